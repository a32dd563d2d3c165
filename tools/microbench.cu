// microbench.cu -- B200 facts the steric kernels are designed around (run under gpurun):
//   fp64 DFMA peak, F2F.F64.F32 and MUFU.RCP64H throughput (alone and mixed with DFMA),
//   accuracy of rcp.approx.ftz.f64 + one cubic refinement, streaming read bandwidth.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

// MODE 0: 8 independent DFMA chains.  1: + one F2F.F64.F32 per 8 DFMA.  2: F2F only.
// 3: MUFU.RCP64H only.  4: 8 DFMA + 1 MUFU.RCP64H.  5: 8 DFMA + 2 F2F (the steric ratio ~ 19:2)
template <int MODE>
__global__ void __launch_bounds__(256) k_pipe(double* out, float seedf, double seed) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x;
  const double m = 1.0000001, c = 1e-9;
  float f = seedf + threadIdx.x;
  double acc = 0.0;
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0 || MODE == 1 || MODE == 4 || MODE == 5) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
    }
    if (MODE == 1 || MODE == 2 || MODE == 5) {
      double d;
      asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(f));
      acc += 0.0;  // keep structure similar
      f = __int_as_float(__float_as_int(f) ^ (int)__double2loint(d));
      if (MODE == 5 || MODE == 2) {
        double d2;
        asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d2) : "f"(f + 1.0f));
        f = __int_as_float(__float_as_int(f) ^ (int)__double2loint(d2));
      }
    }
    if (MODE == 3 || MODE == 4) {
      double r;
      asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a[0]));
      f = __int_as_float(__float_as_int(f) ^ __double2hiint(r));
      if (MODE == 3) a[0] = __hiloint2double(__double2hiint(a[0]) ^ (__double2hiint(r) & 1), __double2loint(a[0]));
    }
  }
  double s = acc + f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run_pipe(const char* name, double ops_per_iter, const char* unit) {
  const int blocks = 148 * 8, threads = 256;
  double* out;
  CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) k_pipe<MODE><<<blocks, threads>>>(out, 1.5f, 1.25);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    k_pipe<MODE><<<blocks, threads>>>(out, 1.5f, 1.25);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double total = ops_per_iter * ITERS * (double)blocks * threads;
  const double rate = total / (best * 1e-3);
  printf("%-34s %8.3f ms  %10.3f %s\n", name, best, rate * 1e-12, unit);
  CK(cudaFree(out));
  return rate;
}

__global__ void k_rcp_acc(const double* d, int n, double* err_seed, double* err_cubic) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = d[i], r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  err_seed[i] = fabs(e);
  double t = fma(e, e, e);
  double r2 = fma(r, t, r);
  // residual of the refined reciprocal, evaluated with an exact fma
  err_cubic[i] = fabs(fma(-x, r2, 1.0));
}

__global__ void __launch_bounds__(256) k_read(const float4* __restrict__ p, size_t n4, float* out) {
  float s = 0.f;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 a = __ldg(p + i), b = __ldg(p + i + stride), c = __ldg(p + i + 2 * stride), d = __ldg(p + i + 3 * stride);
    s += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
  }
  for (; i < n4; i += stride) { float4 a = __ldg(p + i); s += a.x + a.y + a.z + a.w; }
  if (s == 123.456f) out[0] = s;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs=%d clock=%d kHz\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.clockRate);
  double dfma = run_pipe<0>("DFMA x8 chains", 8, "T DFMA/s");
  printf("  -> fp64 peak %.2f TFLOP/s\n", 2 * dfma * 1e-12);
  run_pipe<2>("F2F.F64.F32 only (2/iter)", 2, "T cvt/s");
  run_pipe<1>("8 DFMA + 1 F2F  (DFMA rate)", 8, "T DFMA/s");
  run_pipe<5>("8 DFMA + 2 F2F  (DFMA rate)", 8, "T DFMA/s");
  run_pipe<3>("MUFU.RCP64H only", 1, "T rcp/s");
  run_pipe<4>("8 DFMA + 1 RCP64H (DFMA rate)", 8, "T DFMA/s");

  // reciprocal accuracy over the Wright denominator range and a wide range
  {
    const int n = 1 << 20;
    double* h = (double*)malloc(sizeof(double) * n);
    for (int i = 0; i < n; ++i) {
      double u = (double)rand() / RAND_MAX;
      h[i] = (i & 1) ? 5.0e5 + 3.0e5 * u : exp((u - 0.5) * 600.0);
    }
    double *d, *es, *ec;
    CK(cudaMalloc(&d, sizeof(double) * n)); CK(cudaMalloc(&es, sizeof(double) * n)); CK(cudaMalloc(&ec, sizeof(double) * n));
    CK(cudaMemcpy(d, h, sizeof(double) * n, cudaMemcpyHostToDevice));
    k_rcp_acc<<<n / 256, 256>>>(d, n, es, ec);
    double* hs = (double*)malloc(sizeof(double) * n); double* hc = (double*)malloc(sizeof(double) * n);
    CK(cudaMemcpy(hs, es, sizeof(double) * n, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hc, ec, sizeof(double) * n, cudaMemcpyDeviceToHost));
    double ms = 0, mc = 0;
    for (int i = 0; i < n; ++i) { if (hs[i] > ms) ms = hs[i]; if (hc[i] > mc) mc = hc[i]; }
    printf("rcp.approx.ftz.f64: max |1-d*r0| = %.3e (2^%.1f); after cubic step max |1-d*r| = %.3e\n", ms, log2(ms), mc);
  }
  // streaming read bandwidth
  {
    const size_t bytes = (size_t)8 << 30;
    float4* p; float* o;
    CK(cudaMalloc(&p, bytes)); CK(cudaMalloc(&o, 4));
    CK(cudaMemset(p, 0, bytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int g = 4; g <= 32; g *= 2) {
      float best = 1e30f;
      for (int r = 0; r < 4; ++r) {
        CK(cudaEventRecord(e0));
        k_read<<<148 * g, 256>>>(p, bytes / 16, o);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
      }
      printf("read 8 GiB, grid 148x%-2d: %.3f ms  %.1f GB/s\n", g, best, bytes / (best * 1e-3) * 1e-9);
    }
  }
  return 0;
}

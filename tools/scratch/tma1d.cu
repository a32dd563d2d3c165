// Does a rank-1 tensor map accept a box that starts at an element that is not 16-byte aligned?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__global__ void k(const __grid_constant__ CUtensorMap map, int coord, float* out) {
  __shared__ __align__(128) float buf[256];
  __shared__ uint64_t bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 1024;" ::"r"(b) : "memory");
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];" ::"r"(d), "l"(&map), "r"(b), "r"(coord) : "memory");
  }
  __syncthreads();
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(b) : "memory");
  out[threadIdx.x] = buf[threadIdx.x];
}
typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  Fn fn = (Fn)p;
  const int n = 100000;
  float *d, *o; cudaMalloc(&d, n * 4); cudaMalloc(&o, 1024);
  float* h = new float[n]; for (int i = 0; i < n; ++i) h[i] = (float)i;
  cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
  CUtensorMap m; cuuint64_t dims[1] = {(cuuint64_t)n}; cuuint64_t str[1] = {0}; cuuint32_t box[1] = {256}, es[1] = {1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  for (int coord : {0, 4, 1, 1961, 99900}) {
    k<<<1, 256>>>(m, coord, o);
    cudaError_t e = cudaDeviceSynchronize();
    float out[256]; cudaMemcpy(out, o, 1024, cudaMemcpyDeviceToHost);
    printf("coord %d: %s first=%g last=%g\n", coord, cudaGetErrorString(e), out[0], out[255]);
    if (e != cudaSuccess) break;
  }
}

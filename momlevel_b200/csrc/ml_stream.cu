// ml_stream.cu -- the elementwise operators as streaming kernels over a shared-memory ring (sm_100a).
//
// eos.wright.density / eos.linear.density applied over a field (src/momlevel/derived.py:624-630),
// spice.flament.spice (src/momlevel/spice/flament.py:78-90) and the reference-state pass
// (src/momlevel/reference.py:71-80) read 8-12 bytes and write 8 bytes per point around 18-30 fp64
// instructions.  With plain loads a thread holds its operands in registers while they are in flight, and
// the ~50-60 registers of the evaluation leave room for ~1000 threads per SM: too few bytes in flight to
// cover HBM latency at 6.5 TB/s, so those kernels sat at 0.66-0.82 of the copy bandwidth.  Here a
// persistent CTA streams its tiles through a 4-stage ring filled by 1-D bulk copies
// (cp.async.bulk.shared::cluster.global, one mbarrier per stage): loads cost no registers and no issue
// slots, ~128 KB per SM are in flight whatever the occupancy, the math reads 128-bit words from shared
// memory and results leave as 256-bit streaming stores.
//
// Layout: the operands are rows of `ncol` points (a level of one time step; a flat array is one row); a tile
// is kTile consecutive points of one row.  Rows must start 16-byte aligned in every operand
// (ncol % 4 == 0 and aligned bases); anything else takes the plain kernels in ml_api.cu.
#include "ml_tma_dev.cuh"

#include "ml_stream.cuh"

namespace ml {
namespace stream {

// experiment knobs (tools/ab_stream.sh): ring depth, CTAs per SM of the map kernel, cache hint of the result stores
#ifndef ML_STREAM_STAGES
#define ML_STREAM_STAGES 4
#endif
#ifndef ML_STREAM_CTAS
#define ML_STREAM_CTAS 3
#endif
#ifndef ML_STREAM_ST
#define ML_STREAM_ST 0
#endif
constexpr int kTile = 2048;     // points per tile: 8 KB per fp32 operand
constexpr int kStages = ML_STREAM_STAGES;
constexpr int kThreads = 256;   // two quads of a tile per thread
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tma::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(tma::smem_u32(bar))
               : "memory");
}
// four results of a quad leave as ONE 256-bit store (sm_100: STG.256): a lane writes a whole 32-byte sector and a
// warp 1 KB in a row; two 128-bit stores per lane would each write half of every sector they touch
__device__ __forceinline__ void st4(double* p, double a, double b, double c, double d) {
#if ML_STREAM_ST == 1
  asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#elif ML_STREAM_ST == 2
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#else
  asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d)
               : "memory");
#endif
}
// A warp is done with a stage: true in lane 0 of the warp that is the LAST of the CTA to leave it (that lane
// refills the stage); tma::stage_done orders every warp's reads of the stage before the refill.
__device__ __forceinline__ bool last_to_leave(uint64_t* empty_bar, int* counter, uint32_t parity) {
  __syncwarp();
  if ((threadIdx.x & 31) != 0) return false;
  return tma::stage_done(empty_bar, counter, kWarps, parity);
}

struct Geom {
  i64 ncol;           // points per row
  i64 ntiles;         // rows * tiles_per_row
  int tiles_per_row;
  int nz;             // rows per outer (time) index: row = t * nz + z
};

// Ring of kStages stages of NIN fp32 operand tiles.  Everybody waits on the stage's mbarrier; the warp that is the
// last to leave a stage refills it (no producer warp, no CTA-wide barrier in the loop).
template <int NIN>
struct Ring {
  float* base;
  uint64_t *full, *empty;
  int* released;  // [kStages] warps that have left the stage (mod kWarps)
  __device__ __forceinline__ float* stage(int s, int k) const { return base + ((size_t)s * NIN + k) * kTile; }
  __device__ __forceinline__ void init(unsigned char* smem) {
    base = reinterpret_cast<float*>(smem);
    full = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * NIN * kTile * sizeof(float));
    empty = full + kStages;
    released = reinterpret_cast<int*>(full + 2 * kStages);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < kStages; ++s) {
        tma::mbar_init(full + s, 1);
        tma::mbar_init(empty + s, kWarps);
        released[s] = 0;
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
};
template <int NIN>
constexpr size_t ring_bytes() {
  return (size_t)kStages * NIN * kTile * sizeof(float) + kStages * (2 * sizeof(uint64_t) + sizeof(int)) + 64;
}

// ------------------------------------------------------------------------------ spice / density
// OP 0: Flament spiciness (T, S).  OP 1: density of EOS (T, S, p per row: scalar or per level).
template <int OP, int EOS>
__global__ void __launch_bounds__(kThreads, ML_STREAM_CTAS)
    k_stream_map(const float* __restrict__ T, const float* __restrict__ S, i64 t_stride, i64 s_stride,
                 const double* __restrict__ p, int pmode, Geom g, double* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Ring<2> ring;
  ring.init(smem_raw);
  const int tid = threadIdx.x;
  auto issue = [&](i64 tile, int s) {
    const i64 row = tile / g.tiles_per_row;
    const i64 col0 = (tile - row * g.tiles_per_row) * kTile;
    const i64 t = row / g.nz, z = row - t * g.nz;
    const i64 len = g.ncol - col0 < kTile ? g.ncol - col0 : kTile;
    const uint32_t bytes = (uint32_t)len * 4u;
    tma::mbar_expect_tx(ring.full + s, 2 * bytes);
    bulk_load(ring.stage(s, 0), T + t * t_stride + z * g.ncol + col0, bytes, ring.full + s);
    bulk_load(ring.stage(s, 1), S + t * s_stride + z * g.ncol + col0, bytes, ring.full + s);
  };
  __syncthreads();
  if (tid == 0)
    for (int s = 0; s < kStages; ++s) {
      const i64 tile = (i64)blockIdx.x + (i64)s * gridDim.x;
      if (tile < g.ntiles) issue(tile, s);
    }
  Eos<EOS> eos;
  int it = 0;
  for (i64 tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
    const int s = it % kStages;
    const i64 row = tile / g.tiles_per_row;
    const i64 col0 = (tile - row * g.tiles_per_row) * kTile;
    const i64 len = g.ncol - col0 < kTile ? g.ncol - col0 : kTile;
    if (OP == 1) {
      const i64 z = row % g.nz;
      eos.set_level(pmode == ML_P_SCALAR ? __ldg(p) : (pmode == ML_P_PER_LEVEL ? __ldg(p + z) : 0.0));
    }
    double* o = out + row * g.ncol + col0;
    tma::mbar_wait(ring.full + s, (uint32_t)(it / kStages) & 1u);
    const float4* sT = reinterpret_cast<const float4*>(ring.stage(s, 0));
    const float4* sS = reinterpret_cast<const float4*>(ring.stage(s, 1));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = tid + h * kThreads;
      if (4 * q < len) {
        const float4 a = sT[q], b = sS[q];
        double r0, r1, r2, r3;
        if (OP == 0) {
          r0 = flament_spice((double)a.x, (double)b.x);
          r1 = flament_spice((double)a.y, (double)b.y);
          r2 = flament_spice((double)a.z, (double)b.z);
          r3 = flament_spice((double)a.w, (double)b.w);
        } else {
          r0 = eos.rho_checked((double)a.x, (double)b.x);
          r1 = eos.rho_checked((double)a.y, (double)b.y);
          r2 = eos.rho_checked((double)a.z, (double)b.z);
          r3 = eos.rho_checked((double)a.w, (double)b.w);
        }
        st4(o + 4 * q, r0, r1, r2, r3);
      }
    }
    if (last_to_leave(ring.empty + s, ring.released + s, (uint32_t)(it / kStages) & 1u)) {
      const i64 next = tile + (i64)kStages * gridDim.x;
      if (next < g.ntiles) issue(next, s);
    }
  }
}

// ------------------------------------------------------------------------------ reference state
// rho_ref = rho(T0, S0, p_z) everywhere, volo = nansum(V0), masso = nansum(rho_ref * V0)
// (reference.py:71-80, derived.py:787-789, :435-438).  Block partials: [2][gridDim.x], fixed order.
template <int EOS>
__global__ void __launch_bounds__(kThreads, 2)
    k_stream_refstate(const float* __restrict__ T0, const float* __restrict__ S0, const float* __restrict__ V0,
                      const double* __restrict__ p_level, Geom g, double* __restrict__ rho_ref,
                      double* __restrict__ partials) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double sm[kWarps];
  Ring<3> ring;
  ring.init(smem_raw);
  const int tid = threadIdx.x;
  auto issue = [&](i64 tile, int s) {
    const i64 row = tile / g.tiles_per_row;
    const i64 col0 = (tile - row * g.tiles_per_row) * kTile;
    const i64 len = g.ncol - col0 < kTile ? g.ncol - col0 : kTile;
    const uint32_t bytes = (uint32_t)len * 4u;
    const i64 off = row * g.ncol + col0;
    tma::mbar_expect_tx(ring.full + s, 3 * bytes);
    bulk_load(ring.stage(s, 0), T0 + off, bytes, ring.full + s);
    bulk_load(ring.stage(s, 1), S0 + off, bytes, ring.full + s);
    bulk_load(ring.stage(s, 2), V0 + off, bytes, ring.full + s);
  };
  __syncthreads();
  if (tid == 0)
    for (int s = 0; s < kStages; ++s) {
      const i64 tile = (i64)blockIdx.x + (i64)s * gridDim.x;
      if (tile < g.ntiles) issue(tile, s);
    }
  Eos<EOS> eos;
  double vol = 0.0, mass = 0.0;
  int it = 0;
  for (i64 tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++it) {
    const int s = it % kStages;
    const i64 row = tile / g.tiles_per_row;
    const i64 col0 = (tile - row * g.tiles_per_row) * kTile;
    const i64 len = g.ncol - col0 < kTile ? g.ncol - col0 : kTile;
    eos.set_level(__ldg(p_level + row));
    double* o = rho_ref + row * g.ncol + col0;
    tma::mbar_wait(ring.full + s, (uint32_t)(it / kStages) & 1u);
    const float4* sT = reinterpret_cast<const float4*>(ring.stage(s, 0));
    const float4* sS = reinterpret_cast<const float4*>(ring.stage(s, 1));
    const float4* sV = reinterpret_cast<const float4*>(ring.stage(s, 2));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = tid + h * kThreads;
      if (4 * q < len) {
        const float4 a = sT[q], b = sS[q], v = sV[q];
        const float vv[4] = {v.x, v.y, v.z, v.w};
        double r[4];
        r[0] = eos.rho((double)a.x, (double)b.x);
        r[1] = eos.rho((double)a.y, (double)b.y);
        r[2] = eos.rho((double)a.z, (double)b.z);
        r[3] = eos.rho((double)a.w, (double)b.w);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (!isnan(vv[j])) {
            vol += (double)vv[j];
            const double m = r[j] * (double)vv[j];
            if (!is_nan_q(m)) mass += m;
          }
        st4(o + 4 * q, r[0], r[1], r[2], r[3]);
      }
    }
    if (last_to_leave(ring.empty + s, ring.released + s, (uint32_t)(it / kStages) & 1u)) {
      const i64 next = tile + (i64)kStages * gridDim.x;
      if (next < g.ntiles) issue(next, s);
    }
  }
  vol = block_sum<kWarps>(vol, sm);
  mass = block_sum<kWarps>(mass, sm);
  if (tid == 0) {
    partials[blockIdx.x] = vol;
    partials[(i64)gridDim.x + blockIdx.x] = mass;
  }
}

// ------------------------------------------------------------------------------ density anomaly
// delta_rho[t][z][col] = V_ref notnull ? rho(T, S, p_z) - rho_ref : NaN (steric.py:151-153).  A CTA owns (level,
// column segment) pairs; rho_ref and the volume mask of a pair are read once into registers and reused for every
// time step, whose T and S tiles stream through the ring (the sequence runs on across pair boundaries, so the ring
// never drains).
template <int EOS, typename TV>
__global__ void __launch_bounds__(kThreads, 2)
    k_stream_delta_rho(const float* __restrict__ T, const float* __restrict__ S, i64 t_stride, i64 s_stride,
                       const double* __restrict__ rho_ref, const TV* __restrict__ v_ref,
                       const double* __restrict__ p_level, int nt, Geom g, double* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Ring<2> ring;
  ring.init(smem_raw);
  const int tid = threadIdx.x;
  const i64 npairs = g.ntiles;                         // (level, segment) pairs; g.nz rows
  const i64 mine = npairs > (i64)blockIdx.x ? (npairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const i64 nseq = mine * nt;                          // tiles this CTA streams, time fastest
  auto issue = [&](i64 q, int s) {
    const i64 pair = (i64)blockIdx.x + (q / nt) * gridDim.x;
    const int t = (int)(q % nt);
    const i64 row = pair / g.tiles_per_row;            // level
    const i64 col0 = (pair - row * g.tiles_per_row) * kTile;
    const i64 len = g.ncol - col0 < kTile ? g.ncol - col0 : kTile;
    const uint32_t bytes = (uint32_t)len * 4u;
    const i64 off = row * g.ncol + col0;
    tma::mbar_expect_tx(ring.full + s, 2 * bytes);
    bulk_load(ring.stage(s, 0), T + (i64)t * t_stride + off, bytes, ring.full + s);
    bulk_load(ring.stage(s, 1), S + (i64)t * s_stride + off, bytes, ring.full + s);
  };
  __syncthreads();
  if (tid == 0)
    for (int s = 0; s < kStages && s < nseq; ++s) issue(s, s);
  Eos<EOS> eos;
  const i64 lvl = (i64)g.nz * g.ncol;
  i64 q = 0;
  for (i64 i = 0; i < mine; ++i) {
    const i64 pair = (i64)blockIdx.x + i * gridDim.x;
    const i64 row = pair / g.tiles_per_row;
    const i64 col0 = (pair - row * g.tiles_per_row) * kTile;
    const i64 len = g.ncol - col0 < kTile ? g.ncol - col0 : kTile;
    const i64 off = row * g.ncol + col0;
    eos.set_level(__ldg(p_level + row));
    double ref[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int qd = tid + h * kThreads;
#pragma unroll
      for (int j = 0; j < 4; ++j) ref[h][j] = nan("");
      if (4 * qd < len) {
        const double2 a = __ldg(reinterpret_cast<const double2*>(rho_ref + off + 4 * qd));
        const double2 b = __ldg(reinterpret_cast<const double2*>(rho_ref + off + 4 * qd) + 1);
        const double r4[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
        for (int j = 0; j < 4; ++j) ref[h][j] = isnan(__ldg(v_ref + off + 4 * qd + j)) ? nan("") : r4[j];
      }
    }
    for (int t = 0; t < nt; ++t, ++q) {
      const int s = (int)(q % kStages);
      tma::mbar_wait(ring.full + s, (uint32_t)(q / kStages) & 1u);
      const float4* sT = reinterpret_cast<const float4*>(ring.stage(s, 0));
      const float4* sS = reinterpret_cast<const float4*>(ring.stage(s, 1));
      double* o = out + (i64)t * lvl + off;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int qd = tid + h * kThreads;
        if (4 * qd < len) {
          const float4 a = sT[qd], b = sS[qd];
          st4(o + 4 * qd, eos.rho((double)a.x, (double)b.x) - ref[h][0], eos.rho((double)a.y, (double)b.y) - ref[h][1],
              eos.rho((double)a.z, (double)b.z) - ref[h][2], eos.rho((double)a.w, (double)b.w) - ref[h][3]);
        }
      }
      if (last_to_leave(ring.empty + s, ring.released + s, (uint32_t)(q / kStages) & 1u)) {
        const i64 next = q + kStages;
        if (next < nseq) issue(next, s);
      }
    }
  }
}

// ----------------------------------------------------------------------- host side
static Geom geometry(i64 nrows, int nz, i64 ncol) {
  Geom g;
  g.ncol = ncol;
  g.tiles_per_row = (int)((ncol + kTile - 1) / kTile);
  g.ntiles = nrows * g.tiles_per_row;
  g.nz = nz;
  return g;
}

static unsigned persistent_grid(i64 ntiles, int ctas_per_sm) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const i64 want = (i64)ctas_per_sm * sms;
  return (unsigned)(ntiles < want ? ntiles : want);
}

bool eligible(const void* a, const void* b, const void* c, const void* out, i64 ncol) {
  const uintptr_t bits = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c);
  // fp32 rows start on 16 bytes (bulk copies), fp64 result rows on 32 (256-bit stores)
  return (bits & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 31u) == 0 && ncol % 4 == 0 && ncol >= 4 &&
         ncol / kTile < 0x7fffffff;
}

template <typename K>
static int opt_in(K kern, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  return e == cudaSuccess ? ML_OK : cuda_fail(e, "cudaFuncSetAttribute(k_stream)");
}

int launch_spice(const float* T, const float* S, i64 n, double* out, cudaStream_t st) {
  const Geom g = geometry(1, 1, n);
  auto kern = k_stream_map<0, 0>;
  if (int rc = opt_in(kern, ring_bytes<2>())) return rc;
  kern<<<persistent_grid(g.ntiles, ML_STREAM_CTAS), kThreads, ring_bytes<2>(), st>>>(T, S, 0, 0, nullptr, 0, g, out);
  return launched("k_stream_map(spice)");
}

int launch_density(int eos, const float* T, const float* S, i64 t_stride, i64 s_stride, const double* p, int pmode,
                   i64 nrows, int nz, i64 ncol, double* out, cudaStream_t st) {
  const Geom g = geometry(nrows, nz, ncol);
  const unsigned grid = persistent_grid(g.ntiles, ML_STREAM_CTAS);
  if (eos == ML_EOS_WRIGHT) {
    auto kern = k_stream_map<1, 0>;
    if (int rc = opt_in(kern, ring_bytes<2>())) return rc;
    kern<<<grid, kThreads, ring_bytes<2>(), st>>>(T, S, t_stride, s_stride, p, pmode, g, out);
  } else {
    auto kern = k_stream_map<1, 1>;
    if (int rc = opt_in(kern, ring_bytes<2>())) return rc;
    kern<<<grid, kThreads, ring_bytes<2>(), st>>>(T, S, t_stride, s_stride, p, pmode, g, out);
  }
  return launched("k_stream_map(density)");
}

int refstate_blocks(i64 nz, i64 ncol) {
  const Geom g = geometry(nz, (int)nz, ncol);
  return (int)persistent_grid(g.ntiles, 2);
}

int launch_refstate(int eos, const float* T0, const float* S0, const float* V0, const double* p_level, i64 nz, i64 ncol,
                    double* rho_ref, double* partials, cudaStream_t st) {
  const Geom g = geometry(nz, (int)nz, ncol);
  const unsigned grid = persistent_grid(g.ntiles, 2);
  if (eos == ML_EOS_WRIGHT) {
    auto kern = k_stream_refstate<0>;
    if (int rc = opt_in(kern, ring_bytes<3>())) return rc;
    kern<<<grid, kThreads, ring_bytes<3>(), st>>>(T0, S0, V0, p_level, g, rho_ref, partials);
  } else {
    auto kern = k_stream_refstate<1>;
    if (int rc = opt_in(kern, ring_bytes<3>())) return rc;
    kern<<<grid, kThreads, ring_bytes<3>(), st>>>(T0, S0, V0, p_level, g, rho_ref, partials);
  }
  return launched("k_stream_refstate");
}

int launch_delta_rho(int eos, const float* T, const float* S, i64 t_stride, i64 s_stride, const double* rho_ref,
                     const void* v_ref, int v_f32, const double* p_level, int nt, i64 nz, i64 ncol, double* out,
                     cudaStream_t st) {
  const Geom g = geometry(nz, (int)nz, ncol);
  const unsigned grid = persistent_grid(g.ntiles, 2);
#define ML_STREAM_DRHO(E, TV)                                                                                     \
  do {                                                                                                            \
    auto kern = k_stream_delta_rho<E, TV>;                                                                        \
    if (int rc = opt_in(kern, ring_bytes<2>())) return rc;                                                        \
    kern<<<grid, kThreads, ring_bytes<2>(), st>>>(T, S, t_stride, s_stride, rho_ref, (const TV*)v_ref, p_level, nt, g, out); \
  } while (0)
  if (eos == ML_EOS_WRIGHT) {
    if (v_f32) ML_STREAM_DRHO(0, float); else ML_STREAM_DRHO(0, double);
  } else {
    if (v_f32) ML_STREAM_DRHO(1, float); else ML_STREAM_DRHO(1, double);
  }
#undef ML_STREAM_DRHO
  return launched("k_stream_delta_rho");
}

}  // namespace stream
}  // namespace ml

// ml_strat.cu -- stratification diagnostics that share the vertical sweep of the steric path
// (SURVEY.md section 8f, rank 4): buoyancy frequency, its Chelton adjustment, the stability
// (Turner) angle and the first-mode gravity-wave speed.
//
//   derived.calc_n2              src/momlevel/derived.py:391-411  (cell centres; the `interfaces`
//                                branch needs xgcm and stays out of scope)
//   derived.adjust_negative_n2   src/momlevel/derived.py:30-71
//   derived.calc_stability_angle src/momlevel/derived.py:714-766
//   derived.calc_wave_speed      src/momlevel/derived.py:798-828
//
// One thread owns one water column of one outer (time) slab and walks it top to bottom with a
// three-level window in registers, so T and S cross HBM once (8 B per point in, 8 B out) and the
// vertical derivative -- numpy.gradient(..., edge_order=2) on the uneven z grid, which is what
// DataArray.differentiate evaluates -- never materialises dT/dz, dS/dz, alpha or beta.  Threads
// are adjacent along x, so every load and store of a warp is one contiguous row segment.
#include "ml_common.cuh"
#include "ml_host.cuh"

namespace ml {
namespace {

constexpr int kBlock = 128;
constexpr int kMaxLevels = 1024;  // gradient coefficients of every level live in shared memory

// numpy.gradient's second-order coefficients for an uneven grid (numpy/lib/function_base.py,
// `gradient`: interior a,b,c from dx1 = x[i]-x[i-1], dx2 = x[i+1]-x[i]; one-sided three-point
// formulas at both ends for edge_order=2).  out[i] = a f[i-1] + b f[i] + c f[i+1], with the
// window shifted inwards at the ends.  Explicitly rounded like numpy evaluates them.
__device__ void gradient_coefficients(const double* __restrict__ z, int nz, double* __restrict__ coef /*[nz][3]*/) {
  for (int i = threadIdx.x; i < nz; i += blockDim.x) {
    double a, b, c;
    if (i == 0) {
      const double dx1 = __dsub_rn(z[1], z[0]), dx2 = __dsub_rn(z[2], z[1]);
      const double s = __dadd_rn(dx1, dx2);
      a = -__ddiv_rn(__dadd_rn(__dmul_rn(2.0, dx1), dx2), __dmul_rn(dx1, s));
      b = __ddiv_rn(s, __dmul_rn(dx1, dx2));
      c = -__ddiv_rn(dx1, __dmul_rn(dx2, s));
    } else if (i == nz - 1) {
      const double dx1 = __dsub_rn(z[nz - 2], z[nz - 3]), dx2 = __dsub_rn(z[nz - 1], z[nz - 2]);
      const double s = __dadd_rn(dx1, dx2);
      a = __ddiv_rn(dx2, __dmul_rn(dx1, s));
      b = -__ddiv_rn(__dadd_rn(dx2, dx1), __dmul_rn(dx1, dx2));
      c = __ddiv_rn(__dadd_rn(__dmul_rn(2.0, dx2), dx1), __dmul_rn(dx2, s));
    } else {
      const double dx1 = __dsub_rn(z[i], z[i - 1]), dx2 = __dsub_rn(z[i + 1], z[i]);
      const double s = __dadd_rn(dx1, dx2);
      a = -__ddiv_rn(dx2, __dmul_rn(dx1, s));
      b = __ddiv_rn(__dsub_rn(dx2, dx1), __dmul_rn(dx1, dx2));
      c = __ddiv_rn(dx1, __dmul_rn(dx2, s));
    }
    coef[3 * i + 0] = a;
    coef[3 * i + 1] = b;
    coef[3 * i + 2] = c;
  }
}

// alpha = -(drho/dT)/rho and beta = (drho/dS)/rho (wright.py:122-165) from ONE division:
// rho = pp/den and drho/dX = N_X/den^2 give alpha = -N_T/(den pp), beta = N_S/(den pp).
template <int EOS>
__device__ __forceinline__ void alpha_beta(double T, double S, double p, double& alpha, double& beta) {
  if (EOS == 0) {
    double al0, p0, lam;
    wright_terms(T, S, al0, p0, lam);
    const double pp = p + p0;
    const double den = fma(al0, pp, lam);
    const double dp0_t = fma(wr::b5, S, fma(T, fma(3.0 * wr::b3, T, 2.0 * wr::b2), wr::b1));
    const double dlam_t = fma(wr::c5, S, fma(T, fma(3.0 * wr::c3, T, 2.0 * wr::c2), wr::c1));
    const double n_t = fma(lam, dp0_t, -(pp * fma(pp, wr::a1, dlam_t)));
    const double n_s = fma(lam, fma(wr::b5, T, wr::b4), -(pp * fma(pp, wr::a2, fma(wr::c5, T, wr::c4))));
    const double r = 1.0 / (den * pp);
    alpha = -(n_t * r);
    beta = n_s * r;
  } else {
    const double r = 1.0 / linear_rho(T, S);  // linear.py:113-162
    alpha = -(lin::drho_dt * r);
    beta = lin::drho_ds * r;
  }
}

enum StratOut { kN2 = 0, kN2Adjusted = 1, kTurnerAngle = 2 };

// adjust_negative_n2 for one value (derived.py:56-69).  `fill` = this cell belongs to index 0 of
// the array's FIRST axis, which the reference fills with 1e-8 (time step 0 of a 4-D field, the
// surface level of a 3-D one); `carry` is the last positive value met further up the column.
__device__ __forceinline__ double adjust_step(double n2, bool fill, double& carry) {
  const bool missing = isnan(n2);
  double adj = (missing || n2 <= 0.0) ? nan("") : n2;
  if (fill && isnan(adj)) adj = 1.0e-8;
  if (isnan(adj)) adj = carry;  // ffill along z
  else carry = adj;
  return missing ? nan("") : adj;
}

template <typename TIn, int EOS, int OUT>
__global__ void __launch_bounds__(kBlock) k_strat(const TIn* __restrict__ T, const TIn* __restrict__ S,
                                                  const double* __restrict__ z_l, const double* __restrict__ p_level,
                                                  double gravity, double patm, int fill_mode, int nz, i64 ncol,
                                                  double* __restrict__ out) {
  extern __shared__ double s_coef[];  // [nz][3] gradient coefficients, then [nz] pressure
  double* s_p = s_coef + 3 * nz;
  gradient_coefficients(z_l, nz, s_coef);
  for (int i = threadIdx.x; i < nz; i += blockDim.x)
    s_p[i] = p_level ? __ldg(p_level + i) : fma(__ldg(z_l + i), 1.0e4, patm);  // derived.py:396
  __syncthreads();
  const i64 c = (i64)blockIdx.x * kBlock + threadIdx.x;
  if (c >= ncol) return;
  const i64 slab = (i64)blockIdx.y * nz * ncol + c;
  const TIn* Tc = T + slab;
  const TIn* Sc = S + slab;
  double* oc = out + slab;
  // rolling window over levels z-2 .. z+2 plus one more level in flight (nz >= 3 is checked by the host)
  double t_m2 = 0.0, t_m1 = 0.0, t_0 = ldf(Tc), t_p1 = ldf(Tc + ncol), t_p2 = ldf(Tc + 2 * ncol);
  double s_m2 = 0.0, s_m1 = 0.0, s_0 = ldf(Sc), s_p1 = ldf(Sc + ncol), s_p2 = ldf(Sc + 2 * ncol);
  double t_p3 = 0.0, s_p3 = 0.0;
  if (nz > 3) {
    t_p3 = ldf(Tc + 3 * ncol);
    s_p3 = ldf(Sc + 3 * ncol);
  }
  double carry = nan("");
  const bool fill_slab = fill_mode == 1 && blockIdx.y == 0;  // adjusted[0] of a 4-D field is its first time step
  for (int z = 0; z < nz; ++z) {
    double t_p4 = 0.0, s_p4 = 0.0;
    if (z + 4 < nz) {
      t_p4 = ldf(Tc + (i64)(z + 4) * ncol);
      s_p4 = ldf(Sc + (i64)(z + 4) * ncol);
    }
    const double a = s_coef[3 * z], b = s_coef[3 * z + 1], cc = s_coef[3 * z + 2];
    double dtdz, dsdz;
    if (z == 0) {  // forward three-point formula
      dtdz = fma(cc, t_p2, fma(b, t_p1, a * t_0));
      dsdz = fma(cc, s_p2, fma(b, s_p1, a * s_0));
    } else if (z == nz - 1) {  // backward three-point formula
      dtdz = fma(cc, t_0, fma(b, t_m1, a * t_m2));
      dsdz = fma(cc, s_0, fma(b, s_m1, a * s_m2));
    } else {
      dtdz = fma(cc, t_p1, fma(b, t_0, a * t_m1));
      dsdz = fma(cc, s_p1, fma(b, s_0, a * s_m1));
    }
    double alpha, beta;
    alpha_beta<EOS>(t_0, s_0, s_p[z], alpha, beta);
    double r;
    if (OUT == kTurnerAngle) {
      const double ratio = (beta * dsdz) / (alpha * dtdz);  // derived.py:753
      r = atan((1.0 + ratio) / (1.0 - ratio)) * 57.29577951308232;  // np.degrees: x * (180 / pi)
    } else {
      r = gravity * (alpha * dtdz - beta * dsdz);  // derived.py:401
      if (OUT == kN2Adjusted) r = adjust_step(r, fill_slab || (fill_mode == 0 && z == 0), carry);
    }
    oc[(i64)z * ncol] = r;
    t_m2 = t_m1; t_m1 = t_0; t_0 = t_p1; t_p1 = t_p2; t_p2 = t_p3; t_p3 = t_p4;
    s_m2 = s_m1; s_m1 = s_0; s_0 = s_p1; s_p1 = s_p2; s_p2 = s_p3; s_p3 = s_p4;
  }
}

// adjust_negative_n2 on an existing field (derived.py:30-71)
__global__ void __launch_bounds__(kBlock) k_adjust_n2(const double* __restrict__ n2, int fill_mode, int nz, i64 ncol,
                                                      double* __restrict__ out) {
  const i64 c = (i64)blockIdx.x * kBlock + threadIdx.x;
  if (c >= ncol) return;
  const i64 slab = (i64)blockIdx.y * nz * ncol + c;
  const bool fill_slab = fill_mode == 1 && blockIdx.y == 0;
  double carry = nan("");
  for (int z = 0; z < nz; ++z) {
    const double v = __ldg(n2 + slab + (i64)z * ncol);
    out[slab + (i64)z * ncol] = adjust_step(v, fill_slab || (fill_mode == 0 && z == 0), carry);
  }
}

// calc_wave_speed (derived.py:821): sum_z sqrt(adjusted n2) dz / pi, skipna.  dz is [nz][ncol].
__global__ void __launch_bounds__(kBlock) k_wave_speed(const double* __restrict__ n2, const double* __restrict__ dz,
                                                       int fill_mode, int nz, i64 ncol, double* __restrict__ out) {
  const i64 c = (i64)blockIdx.x * kBlock + threadIdx.x;
  if (c >= ncol) return;
  const i64 slab = (i64)blockIdx.y * nz * ncol + c;
  const bool fill_slab = fill_mode == 1 && blockIdx.y == 0;
  double carry = nan(""), acc = 0.0;
  for (int z = 0; z < nz; ++z) {
    const double v = __ldg(n2 + slab + (i64)z * ncol);
    const double term = sqrt(adjust_step(v, fill_slab || (fill_mode == 0 && z == 0), carry)) * __ldg(dz + (i64)z * ncol + c);
    if (!isnan(term)) acc += term;
  }
  out[(i64)blockIdx.y * ncol + c] = acc / 3.141592653589793;
}

template <int OUT>
int launch_strat(int eos, int dtype, const void* T, const void* S, const double* z_l, const double* p_level,
                 double gravity, double patm, int fill_mode, i64 nouter, int nz, i64 ncol, double* out,
                 cudaStream_t st) {
  const dim3 grid((unsigned)((ncol + kBlock - 1) / kBlock), (unsigned)nouter);
  const size_t smem = (size_t)4 * nz * sizeof(double);
#define ML_STRAT(TIN, E) \
  k_strat<TIN, E, OUT><<<grid, kBlock, smem, st>>>((const TIN*)T, (const TIN*)S, z_l, p_level, gravity, patm, fill_mode, nz, ncol, out)
  if (dtype == ML_F32) {
    if (eos == ML_EOS_WRIGHT) ML_STRAT(float, 0); else ML_STRAT(float, 1);
  } else {
    if (eos == ML_EOS_WRIGHT) ML_STRAT(double, 0); else ML_STRAT(double, 1);
  }
#undef ML_STRAT
  return launched("k_strat");
}

int check_strat(int eos, int dtype, const void* T, const void* S, const double* z_l, const double* out, int64_t nouter,
                int64_t nz, int64_t ncol) {
  if (eos != ML_EOS_WRIGHT && eos != ML_EOS_LINEAR) return fail(ML_ERR_EOS, "unknown equation of state id %d", eos);
  if (dtype != ML_F32 && dtype != ML_F64) return fail(ML_ERR_DTYPE, "unknown dtype id %d", dtype);
  ML_REQUIRE_PTR(T);
  ML_REQUIRE_PTR(S);
  ML_REQUIRE_PTR(z_l);
  ML_REQUIRE_PTR(out);
  // numpy.gradient(edge_order=2) needs three points; the coefficient table lives in shared memory
  if (nouter <= 0 || ncol <= 0 || nz < 3 || nz > kMaxLevels || nouter > 65535)
    return fail(ML_ERR_SHAPE, "bad extents nouter=%lld nz=%lld ncol=%lld (3 <= nz <= %d, nouter <= 65535)",
                (long long)nouter, (long long)nz, (long long)ncol, kMaxLevels);
  ML_REQUIRE_ALIGNED(T, elem_size(dtype));
  ML_REQUIRE_ALIGNED(S, elem_size(dtype));
  return ML_OK;
}

}  // namespace
}  // namespace ml

using namespace ml;

extern "C" {

int ml_calc_n2(int eos, int dtype, const void* T, const void* S, const double* z_l, double gravity, double patm,
               int adjust_negative, int fill_mode, int64_t nouter, int64_t nz, int64_t ncol, double* out,
               void* stream) {
  int rc = check_strat(eos, dtype, T, S, z_l, out, nouter, nz, ncol);
  if (rc) return rc;
  if (fill_mode != 0 && fill_mode != 1) return fail(ML_ERR_MODE, "fill_mode must be 0 or 1, got %d", fill_mode);
  cudaStream_t st = (cudaStream_t)stream;
  if (adjust_negative)
    return launch_strat<kN2Adjusted>(eos, dtype, T, S, z_l, nullptr, gravity, patm, fill_mode, nouter, (int)nz, ncol, out, st);
  return launch_strat<kN2>(eos, dtype, T, S, z_l, nullptr, gravity, patm, fill_mode, nouter, (int)nz, ncol, out, st);
}

int ml_stability_angle(int eos, int dtype, const void* T, const void* S, const double* p_level, const double* z_l,
                       int64_t nouter, int64_t nz, int64_t ncol, double* out, void* stream) {
  int rc = check_strat(eos, dtype, T, S, z_l, out, nouter, nz, ncol);
  if (rc) return rc;
  ML_REQUIRE_PTR(p_level);
  return launch_strat<kTurnerAngle>(eos, dtype, T, S, z_l, p_level, 0.0, 0.0, 0, nouter, (int)nz, ncol, out,
                                    (cudaStream_t)stream);
}

int ml_adjust_negative_n2(const double* n2, int fill_mode, int64_t nouter, int64_t nz, int64_t ncol, double* out,
                          void* stream) {
  ML_REQUIRE_PTR(n2);
  ML_REQUIRE_PTR(out);
  if (fill_mode != 0 && fill_mode != 1) return fail(ML_ERR_MODE, "fill_mode must be 0 or 1, got %d", fill_mode);
  if (nouter <= 0 || nz <= 0 || ncol <= 0 || nouter > 65535 || nz > INT32_MAX)
    return fail(ML_ERR_SHAPE, "bad extents nouter=%lld nz=%lld ncol=%lld", (long long)nouter, (long long)nz, (long long)ncol);
  const dim3 grid((unsigned)((ncol + kBlock - 1) / kBlock), (unsigned)nouter);
  k_adjust_n2<<<grid, kBlock, 0, (cudaStream_t)stream>>>(n2, fill_mode, (int)nz, ncol, out);
  return launched("k_adjust_n2");
}

int ml_wave_speed(const double* n2, const double* dz, int fill_mode, int64_t nouter, int64_t nz, int64_t ncol,
                  double* out, void* stream) {
  ML_REQUIRE_PTR(n2);
  ML_REQUIRE_PTR(dz);
  ML_REQUIRE_PTR(out);
  if (fill_mode != 0 && fill_mode != 1) return fail(ML_ERR_MODE, "fill_mode must be 0 or 1, got %d", fill_mode);
  if (nouter <= 0 || nz <= 0 || ncol <= 0 || nouter > 65535 || nz > INT32_MAX)
    return fail(ML_ERR_SHAPE, "bad extents nouter=%lld nz=%lld ncol=%lld", (long long)nouter, (long long)nz, (long long)ncol);
  const dim3 grid((unsigned)((ncol + kBlock - 1) / kBlock), (unsigned)nouter);
  k_wave_speed<<<grid, kBlock, 0, (cudaStream_t)stream>>>(n2, dz, fill_mode, (int)nz, ncol, out);
  return launched("k_wave_speed");
}

}  // extern "C"

// ml_hostpath.cu -- ml_steric_local_host / ml_steric_global_host: the steric path on HOST buffers.
//
// What momlevel.steric(dset) does for variant="steric", domain="local" when the Dataset
// lives in host memory (src/momlevel/steric.py:84-184 with the reference state taken from
// time step 0, src/momlevel/reference.py:60-80): time steps are streamed through two device
// windows, the host->device copy of window k+1 running on a copy stream while the kernels of
// window k run on the compute stream.  Pinned (page-locked) host buffers make the copies
// truly asynchronous; pageable buffers work but serialise inside the driver.
#include <vector>

#include "ml_host.cuh"

namespace {

// Device staging buffers, streams and events are kept per host thread between calls (a year of
// OM4p25 needs ~5 GB of windows; allocating and freeing them costs tens of milliseconds per call)
// and released by ml_host_release() or at thread exit.
struct Resources {
  static constexpr int kSlots = 16;
  void* buf[kSlots] = {nullptr};
  size_t cap[kSlots] = {0};
  int device = -1;
  cudaStream_t copy = nullptr, comp = nullptr;
  cudaEvent_t copied[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
  void release() {
    for (int i = 0; i < kSlots; ++i) {
      if (buf[i]) cudaFree(buf[i]);
      buf[i] = nullptr;
      cap[i] = 0;
    }
    for (int i = 0; i < 2; ++i) {
      if (copied[i]) cudaEventDestroy(copied[i]);
      if (freed[i]) cudaEventDestroy(freed[i]);
      copied[i] = freed[i] = nullptr;
    }
    if (copy) cudaStreamDestroy(copy);
    if (comp) cudaStreamDestroy(comp);
    copy = comp = nullptr;
    device = -1;
  }
  ~Resources() { release(); }
  cudaError_t prepare() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev != device) {
      release();
      device = dev;
    }
    if (!copy && (e = cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if (!comp && (e = cudaStreamCreateWithFlags(&comp, cudaStreamNonBlocking)) != cudaSuccess) return e;
    for (int b = 0; b < 2; ++b) {
      if (!copied[b] && (e = cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming)) != cudaSuccess) return e;
      if (!freed[b] && (e = cudaEventCreateWithFlags(&freed[b], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  // slot i grows to at least `bytes`
  cudaError_t alloc(int i, void** p, size_t bytes) {
    if (cap[i] < bytes) {
      if (buf[i]) cudaFree(buf[i]);
      buf[i] = nullptr;
      cap[i] = 0;
      cudaError_t e = cudaMalloc(&buf[i], bytes);
      if (e != cudaSuccess) return e;
      cap[i] = bytes;
    }
    *p = buf[i];
    return cudaSuccess;
  }
};

Resources& resources() {
  static thread_local Resources r;
  return r;
}

}  // namespace

extern "C" int ml_host_release(void) {
  resources().release();
  return ML_OK;
}

// eta_thermo / eta_halo: optional extra heights from the same transfer (NULL = steric only)
static int steric_local_host_impl(int eos, int dtype, const void* T, const void* S, const void* v0, const double* z_i,
                                  const double* deptho, const double* p_level, double neg_inv_rhozero, int64_t nt,
                                  int64_t nz, int64_t ncol, int steps_per_window, double* eta, double* eta_thermo,
                                  double* eta_halo, double* rho_ref_out, double* sums_out) {
  using namespace ml;
  if (eos != ML_EOS_WRIGHT && eos != ML_EOS_LINEAR) return fail(ML_ERR_EOS, "unknown equation of state id %d", eos);
  if (dtype != ML_F32 && dtype != ML_F64) return fail(ML_ERR_DTYPE, "unknown dtype id %d", dtype);
  ML_REQUIRE_PTR(T);
  ML_REQUIRE_PTR(S);
  ML_REQUIRE_PTR(v0);
  ML_REQUIRE_PTR(z_i);
  ML_REQUIRE_PTR(deptho);
  ML_REQUIRE_PTR(p_level);
  ML_REQUIRE_PTR(eta);
  if (nt <= 0 || nz <= 0 || ncol <= 0 || steps_per_window < 1)
    return fail(ML_ERR_SHAPE, "bad extents nt=%lld nz=%lld ncol=%lld window=%d", (long long)nt, (long long)nz,
                (long long)ncol, steps_per_window);

  const size_t es = (size_t)elem_size(dtype);
  const size_t lvl = (size_t)nz * (size_t)ncol;
  const int64_t spw = steps_per_window < nt ? steps_per_window : nt;
  const size_t win_bytes = (size_t)spw * lvl * es;
  const size_t ws_bytes = ml_workspace_bytes(2, nz, ncol);

  Resources& r = resources();
  ML_CUDA(r.prepare());
  void *dT[2], *dS[2], *dV, *dRho, *dEta, *dZi, *dDepth, *dP, *dSums, *dWs;
  for (int b = 0; b < 2; ++b) {
    ML_CUDA(r.alloc(2 * b, &dT[b], win_bytes));
    ML_CUDA(r.alloc(2 * b + 1, &dS[b], win_bytes));
  }
  ML_CUDA(r.alloc(4, &dV, lvl * es));
  ML_CUDA(r.alloc(5, &dRho, lvl * sizeof(double)));
  ML_CUDA(r.alloc(6, &dEta, (size_t)nt * ncol * sizeof(double)));
  ML_CUDA(r.alloc(7, &dZi, (size_t)(nz + 1) * sizeof(double)));
  ML_CUDA(r.alloc(8, &dDepth, (size_t)ncol * sizeof(double)));
  ML_CUDA(r.alloc(9, &dP, (size_t)nz * sizeof(double)));
  ML_CUDA(r.alloc(10, &dSums, 2 * sizeof(double)));
  ML_CUDA(r.alloc(11, &dWs, ws_bytes));
  // the other two variants hold one field at the reference slab (steric.py:115-121), so step 0 of T and S
  // stays on the device for the whole call, next to one more height field per variant
  const bool variants = eta_thermo != nullptr || eta_halo != nullptr;
  void *dT0 = nullptr, *dS0 = nullptr, *dEtaT = nullptr, *dEtaH = nullptr;
  if (variants) {
    ML_CUDA(r.alloc(12, &dT0, lvl * es));
    ML_CUDA(r.alloc(13, &dS0, lvl * es));
    if (eta_thermo) ML_CUDA(r.alloc(14, &dEtaT, (size_t)nt * ncol * sizeof(double)));
    if (eta_halo) ML_CUDA(r.alloc(15, &dEtaH, (size_t)nt * ncol * sizeof(double)));
  }

  ML_CUDA(cudaMemcpyAsync(dZi, z_i, (size_t)(nz + 1) * sizeof(double), cudaMemcpyHostToDevice, r.copy));
  ML_CUDA(cudaMemcpyAsync(dDepth, deptho, (size_t)ncol * sizeof(double), cudaMemcpyHostToDevice, r.copy));
  ML_CUDA(cudaMemcpyAsync(dP, p_level, (size_t)nz * sizeof(double), cudaMemcpyHostToDevice, r.copy));
  ML_CUDA(cudaMemcpyAsync(dV, v0, lvl * es, cudaMemcpyHostToDevice, r.copy));

  const int64_t nwin = (nt + spw - 1) / spw;
  for (int64_t w = 0; w < nwin; ++w) {
    const int b = (int)(w & 1);
    const int64_t t_first = w * spw;
    const int64_t nt_w = (t_first + spw <= nt) ? spw : (nt - t_first);
    const size_t off = (size_t)t_first * lvl * es;
    const size_t bytes = (size_t)nt_w * lvl * es;
    if (w >= 2) ML_CUDA(cudaStreamWaitEvent(r.copy, r.freed[b], 0));
    ML_CUDA(cudaMemcpyAsync(dT[b], (const char*)T + off, bytes, cudaMemcpyHostToDevice, r.copy));
    ML_CUDA(cudaMemcpyAsync(dS[b], (const char*)S + off, bytes, cudaMemcpyHostToDevice, r.copy));
    ML_CUDA(cudaEventRecord(r.copied[b], r.copy));

    ML_CUDA(cudaStreamWaitEvent(r.comp, r.copied[b], 0));
    int rc;
    if (w == 0)  // the window that starts at the reference step (reference.py:60-80): fused pass
      rc = ml_steric_local_selfref(eos, dtype, dT[0], dS[0], 0, 0, dV, dtype, (const double*)dZi,
                                   (const double*)dDepth, (const double*)dP, neg_inv_rhozero, nt_w, nz, ncol,
                                   (double*)dEta, (double*)dRho, (double*)dSums, dWs, ws_bytes, r.comp);
    else
      rc = ml_steric_local(eos, dtype, dT[b], dS[b], 0, 0, (const double*)dRho, dV, dtype, (const double*)dZi,
                           (const double*)dDepth, (const double*)dP, neg_inv_rhozero, nt_w, nz, ncol,
                           (double*)dEta + (size_t)t_first * ncol, nullptr, r.comp);
    if (rc) return rc;
    if (variants) {
      if (w == 0) {  // keep the reference slabs before window 0 is recycled
        ML_CUDA(cudaMemcpyAsync(dT0, dT[0], lvl * es, cudaMemcpyDeviceToDevice, r.comp));
        ML_CUDA(cudaMemcpyAsync(dS0, dS[0], lvl * es, cudaMemcpyDeviceToDevice, r.comp));
      }
      // window 0 starts at the reference step: the fused self-reference pass again, so that the step-0
      // heights of these variants are exactly zero as well (it rewrites rho_ref / sums with the same values)
      for (int v = 0; v < 2; ++v) {
        double* out = (double*)(v == 0 ? dEtaT : dEtaH);
        if ((v == 0 ? eta_thermo : eta_halo) == nullptr) continue;
        const void* Tv = v == 0 ? dT[b] : dT0;
        const void* Sv = v == 0 ? dS0 : dS[b];
        if (w == 0)
          rc = ml_steric_local_selfref(eos, dtype, Tv, Sv, v == 1, v == 0, dV, dtype, (const double*)dZi,
                                       (const double*)dDepth, (const double*)dP, neg_inv_rhozero, nt_w, nz, ncol, out,
                                       (double*)dRho, (double*)dSums, dWs, ws_bytes, r.comp);
        else
          rc = ml_steric_local(eos, dtype, Tv, Sv, v == 1, v == 0, (const double*)dRho, dV, dtype, (const double*)dZi,
                               (const double*)dDepth, (const double*)dP, neg_inv_rhozero, nt_w, nz, ncol,
                               out + (size_t)t_first * ncol, nullptr, r.comp);
        if (rc) return rc;
      }
    }
    ML_CUDA(cudaEventRecord(r.freed[b], r.comp));
  }
  ML_CUDA(cudaMemcpyAsync(eta, dEta, (size_t)nt * ncol * sizeof(double), cudaMemcpyDeviceToHost, r.comp));
  if (eta_thermo) ML_CUDA(cudaMemcpyAsync(eta_thermo, dEtaT, (size_t)nt * ncol * sizeof(double), cudaMemcpyDeviceToHost, r.comp));
  if (eta_halo) ML_CUDA(cudaMemcpyAsync(eta_halo, dEtaH, (size_t)nt * ncol * sizeof(double), cudaMemcpyDeviceToHost, r.comp));
  if (sums_out) ML_CUDA(cudaMemcpyAsync(sums_out, dSums, 2 * sizeof(double), cudaMemcpyDeviceToHost, r.comp));
  if (rho_ref_out) ML_CUDA(cudaMemcpyAsync(rho_ref_out, dRho, lvl * sizeof(double), cudaMemcpyDeviceToHost, r.comp));
  ML_CUDA(cudaStreamSynchronize(r.comp));
  ML_CUDA(cudaStreamSynchronize(r.copy));
  return ML_OK;
}

extern "C" int ml_steric_local_host(int eos, int dtype, const void* T, const void* S, const void* v0,
                                    const double* z_i, const double* deptho, const double* p_level,
                                    double neg_inv_rhozero, int64_t nt, int64_t nz, int64_t ncol,
                                    int steps_per_window, double* eta, double* rho_ref_out, double* sums_out) {
  return steric_local_host_impl(eos, dtype, T, S, v0, z_i, deptho, p_level, neg_inv_rhozero, nt, nz, ncol,
                                steps_per_window, eta, nullptr, nullptr, rho_ref_out, sums_out);
}

extern "C" int ml_steric_local_variants_host(int eos, int dtype, const void* T, const void* S, const void* v0,
                                             const double* z_i, const double* deptho, const double* p_level,
                                             double neg_inv_rhozero, int64_t nt, int64_t nz, int64_t ncol,
                                             int steps_per_window, double* eta_steric, double* eta_thermosteric,
                                             double* eta_halosteric, double* rho_ref_out, double* sums_out) {
  return steric_local_host_impl(eos, dtype, T, S, v0, z_i, deptho, p_level, neg_inv_rhozero, nt, nz, ncol,
                                steps_per_window, eta_steric, eta_thermosteric, eta_halosteric, rho_ref_out, sums_out);
}

// The global branch (src/momlevel/steric.py:134-147) on host buffers: per-step masses
// M(t) = sum rho(t) * volcello_ref (derived.py:435-438) with T, S streamed through the same two
// windows; the ln() formula stays with the caller, as for ml_steric_global.
extern "C" int ml_steric_global_host(int eos, int dtype, const void* T, const void* S, const void* v_ref,
                                     const double* p_level, int64_t nt, int64_t nz, int64_t ncol,
                                     int steps_per_window, double* masso) {
  using namespace ml;
  if (eos != ML_EOS_WRIGHT && eos != ML_EOS_LINEAR) return fail(ML_ERR_EOS, "unknown equation of state id %d", eos);
  if (dtype != ML_F32 && dtype != ML_F64) return fail(ML_ERR_DTYPE, "unknown dtype id %d", dtype);
  ML_REQUIRE_PTR(T);
  ML_REQUIRE_PTR(S);
  ML_REQUIRE_PTR(v_ref);
  ML_REQUIRE_PTR(p_level);
  ML_REQUIRE_PTR(masso);
  if (nt <= 0 || nz <= 0 || ncol <= 0 || steps_per_window < 1)
    return fail(ML_ERR_SHAPE, "bad extents nt=%lld nz=%lld ncol=%lld window=%d", (long long)nt, (long long)nz,
                (long long)ncol, steps_per_window);
  const size_t es = (size_t)elem_size(dtype);
  const size_t lvl = (size_t)nz * (size_t)ncol;
  const int64_t spw = steps_per_window < nt ? steps_per_window : nt;
  const size_t win_bytes = (size_t)spw * lvl * es;
  const size_t ws_bytes = ml_workspace_bytes(spw, nz, ncol);

  Resources& r = resources();
  ML_CUDA(r.prepare());
  void *dT[2], *dS[2], *dV, *dM, *dP, *dWs;
  for (int b = 0; b < 2; ++b) {
    ML_CUDA(r.alloc(2 * b, &dT[b], win_bytes));
    ML_CUDA(r.alloc(2 * b + 1, &dS[b], win_bytes));
  }
  ML_CUDA(r.alloc(4, &dV, lvl * es));
  ML_CUDA(r.alloc(6, &dM, (size_t)nt * sizeof(double)));
  ML_CUDA(r.alloc(9, &dP, (size_t)nz * sizeof(double)));
  ML_CUDA(r.alloc(11, &dWs, ws_bytes));
  ML_CUDA(cudaMemcpyAsync(dP, p_level, (size_t)nz * sizeof(double), cudaMemcpyHostToDevice, r.copy));
  ML_CUDA(cudaMemcpyAsync(dV, v_ref, lvl * es, cudaMemcpyHostToDevice, r.copy));

  const int64_t nwin = (nt + spw - 1) / spw;
  for (int64_t w = 0; w < nwin; ++w) {
    const int b = (int)(w & 1);
    const int64_t t_first = w * spw;
    const int64_t nt_w = (t_first + spw <= nt) ? spw : (nt - t_first);
    const size_t off = (size_t)t_first * lvl * es;
    const size_t bytes = (size_t)nt_w * lvl * es;
    if (w >= 2) ML_CUDA(cudaStreamWaitEvent(r.copy, r.freed[b], 0));
    ML_CUDA(cudaMemcpyAsync(dT[b], (const char*)T + off, bytes, cudaMemcpyHostToDevice, r.copy));
    ML_CUDA(cudaMemcpyAsync(dS[b], (const char*)S + off, bytes, cudaMemcpyHostToDevice, r.copy));
    ML_CUDA(cudaEventRecord(r.copied[b], r.copy));
    ML_CUDA(cudaStreamWaitEvent(r.comp, r.copied[b], 0));
    int rc = ml_steric_global(eos, dtype, dT[b], dS[b], 0, 0, dV, dtype, (const double*)dP, nt_w, nz, ncol,
                              (double*)dM + t_first, dWs, ws_bytes, r.comp);
    if (rc) return rc;
    ML_CUDA(cudaEventRecord(r.freed[b], r.comp));
  }
  ML_CUDA(cudaMemcpyAsync(masso, dM, (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost, r.comp));
  ML_CUDA(cudaStreamSynchronize(r.comp));
  ML_CUDA(cudaStreamSynchronize(r.copy));
  return ML_OK;
}

// hostpath_sim.cpp -- the host entry points of libmomlevel_b200 (csrc/ml_hostpath.cu + csrc/ml_pack.cpp) run on a
// box without a GPU, against the simulated runtime of tests/sim/cuda_runtime.h and under ThreadSanitizer.
//
// What is under test is everything the *_host entry points do on the host: the windows, the double buffers and
// their events, the presence index, the packers and the calling thread working towards each other, the staging
// ring, the pinned bounce buffers of pageable operands.  The device kernels are replaced by toy ones with the
// same data dependences (a "density" T + 2 S, used only where the reference volume is present); the real kernels
// are tested on the GPU.  Every combination must give, bit for bit, what the toy kernels give on the caller's
// whole arrays -- whichever rows crossed packed -- and ThreadSanitizer must stay silent.
//
//   g++ -std=c++17 -O1 -g -fsanitize=thread -pthread -I tests/sim -x c++ tests/sim/hostpath_sim.cpp -o hostpath_sim
#define ML_HOSTPATH_TEST_HOOKS
#include <cuda_runtime.h>  // tests/sim/cuda_runtime.h

#include <math.h>
#include <stdio.h>

#include <random>
#include <vector>

#include "../../momlevel_b200/csrc/ml_host.cuh"

namespace ml {
ThreadState& tls() {
  static thread_local ThreadState s = {{0}, 0, 0, 0};
  return s;
}
}  // namespace ml

// ---- toy kernels: run on the stream they are given, like the real ones -----------------------------------------

static inline double toy_rho(float t, float s) { return (double)t + 2.0 * (double)s; }
static inline bool present(float v) { return !std::isnan(v); }

static void toy_local(const float* T, const float* S, int tb, int sb, const double* rho_ref, const float* V, int64_t nt,
                      int64_t nz, int64_t ncol, double* eta) {
  for (int64_t t = 0; t < nt; ++t)
    for (int64_t c = 0; c < ncol; ++c) {
      double acc = 0.0;
      for (int64_t z = 0; z < nz; ++z) {
        const size_t i = (size_t)z * ncol + c;
        if (!present(V[i])) continue;
        const double d = toy_rho(T[(tb ? 0 : (size_t)t * nz * ncol) + i], S[(sb ? 0 : (size_t)t * nz * ncol) + i]) - rho_ref[i];
        if (!std::isnan(d)) acc += d;
      }
      eta[(size_t)t * ncol + c] = acc;
    }
}

static void toy_reference(const float* T, const float* S, const float* V, int64_t nz, int64_t ncol, double* rho_ref,
                          double* sums) {
  double volo = 0.0, masso = 0.0;
  for (size_t i = 0; i < (size_t)nz * ncol; ++i) {
    rho_ref[i] = toy_rho(T[i], S[i]);
    if (!present(V[i])) continue;
    volo += V[i];
    if (!std::isnan(rho_ref[i])) masso += rho_ref[i] * V[i];
  }
  sums[0] = volo;
  sums[1] = masso;
}

static void toy_global(const float* T, const float* S, int tb, int sb, const float* V, int64_t nt, int64_t nz, int64_t ncol,
                       double* masso) {
  for (int64_t t = 0; t < nt; ++t) {
    double m = 0.0;
    for (size_t i = 0; i < (size_t)nz * ncol; ++i) {
      if (!present(V[i])) continue;
      const double x = toy_rho(T[(tb ? 0 : (size_t)t * nz * ncol) + i], S[(sb ? 0 : (size_t)t * nz * ncol) + i]) * V[i];
      if (!std::isnan(x)) m += x;
    }
    masso[t] = m;
  }
}

extern "C" size_t ml_workspace_bytes(int64_t, int64_t, int64_t) { return 64; }

extern "C" int ml_steric_local(int, int, const void* T, const void* S, int tb, int sb, const double* rho_ref,
                               const void* v_ref, int, const double*, const double*, const double*, double, int64_t nt,
                               int64_t nz, int64_t ncol, double* eta, double*, void* stream) {
  ((cudaStream_t)stream)->enqueue([=] {
    toy_local((const float*)T, (const float*)S, tb, sb, rho_ref, (const float*)v_ref, nt, nz, ncol, eta);
  });
  return 0;
}

extern "C" int ml_steric_local_selfref(int, int, const void* T, const void* S, int tb, int sb, const void* v_ref, int,
                                       const double*, const double*, const double*, double, int64_t nt, int64_t nz,
                                       int64_t ncol, double* eta, double* rho_ref, double* sums, void*, size_t,
                                       void* stream) {
  ((cudaStream_t)stream)->enqueue([=] {
    // the reference state is step 0 of whichever operand varies in time (the other one IS the reference slab)
    toy_reference((const float*)T, (const float*)S, (const float*)v_ref, nz, ncol, rho_ref, sums);
    toy_local((const float*)T, (const float*)S, tb, sb, rho_ref, (const float*)v_ref, nt, nz, ncol, eta);
  });
  return 0;
}

extern "C" int ml_steric_global(int, int, const void* T, const void* S, int tb, int sb, const void* v_ref, int, const double*,
                                int64_t nt, int64_t nz, int64_t ncol, double* masso, void*, size_t, void* stream) {
  ((cudaStream_t)stream)->enqueue([=] { toy_global((const float*)T, (const float*)S, tb, sb, (const float*)v_ref, nt, nz, ncol, masso); });
  return 0;
}

extern "C" int ml_reference_state(int, int, const void* T0, const void* S0, const void* V0, const double*, int64_t nz,
                                  int64_t ncol, double* rho_ref, double* sums, void*, size_t, void* stream) {
  ((cudaStream_t)stream)->enqueue([=] { toy_reference((const float*)T0, (const float*)S0, (const float*)V0, nz, ncol, rho_ref, sums); });
  return 0;
}

// the expansion kernel, as a loop (k_unpack_rows itself is tested on the GPU)
static int launch_unpack(cudaStream_t stream, int nrows, int, const float* pT, const float* pS, float* T, float* S,
                         const uint32_t* words, const uint32_t* before, const uint64_t* lvloff, const uint8_t* flags,
                         int nz, int64_t ncol, int64_t ngrp, uint64_t nwet) {
  stream->enqueue([=] {
    for (int row = 0; row < nrows; ++row) {
      if (!flags[row]) continue;
      const int t = row / nz, z = row % nz;
      for (int64_t g = 0; g < ngrp; ++g) {
        const uint32_t m = words[(size_t)z * ngrp + g];
        uint64_t src = (uint64_t)t * nwet + lvloff[z] + before[(size_t)z * ngrp + g];
        for (int lane = 0; lane < 32; ++lane) {
          const int64_t col = g * 32 + lane;
          if (col >= ncol) break;
          const bool here = (m >> lane) & 1u;
          T[(size_t)row * ncol + col] = here ? pT[src] : NAN;
          S[(size_t)row * ncol + col] = here ? pS[src] : NAN;
          src += here;
        }
      }
    }
  });
  ml::tls().launches++;
  return 0;
}

// the index kernels, as the host body they must agree with (k_presence_* themselves are tested on the GPU)
static int launch_presence_index(cudaStream_t stream, const float* v, int64_t nz, int64_t ncol, int64_t, uint32_t* words,
                                 uint32_t* before, uint64_t* count) {
  stream->enqueue([=] { ml_pack_index_rows(v, nz, ncol, words, before, count); });
  ml::tls().launches += 2;
  return 0;
}

#include "../../momlevel_b200/csrc/ml_hostpath.cu"
#include "../../momlevel_b200/csrc/ml_pack.cpp"

// ---- the harness ------------------------------------------------------------------------------------------------

template <class T>
static T* alloc(size_t n, bool pinned) {
  void* p = nullptr;
  if (pinned)
    cudaHostAlloc(&p, n * sizeof(T), 0);
  else
    p = malloc(n * sizeof(T));
  return (T*)p;
}
template <class T>
static void release(T* p, bool pinned) {
  if (pinned)
    cudaFreeHost(p);
  else
    free(p);
}

static bool same(const double* a, const double* b, size_t n) { return memcmp(a, b, n * sizeof(double)) == 0; }

int main() {
  struct Shape {
    int64_t nt, nz, ncol;
    int pattern;  // 0 = levels thin out with depth, 1 = every volume present, 2 = none
  } shapes[] = {{5, 7, 1000, 0}, {4, 3, 33, 0}, {3, 12, 4111, 0},  // 32, 2 and 129 groups per row: 4, 1 and 16 segments
                {1, 1, 5, 0},    {2, 3, 64, 1}, {2, 3, 70, 2}};
  int runs = 0, bad = 0;
  double frac_sum[5][2] = {{0}}, frac_max[5][2] = {{0}};  // pinned sources: share of rows packed by mode, fast / slow copies
  int frac_n[5][2] = {{0}};
  for (const Shape& sh : shapes) {
    const int64_t nt = sh.nt, nz = sh.nz, ncol = sh.ncol;
    const size_t lvl = (size_t)nz * ncol, all = (size_t)nt * lvl;
    for (int pinned = 0; pinned < 2; ++pinned) {
      float* T = alloc<float>(all, pinned);
      float* S = alloc<float>(all, pinned);
      float* V = alloc<float>(lvl, pinned);
      double* eta[3];
      for (auto& e : eta) e = alloc<double>((size_t)nt * ncol, pinned);
      std::mt19937 rng(17 + (unsigned)ncol);
      std::uniform_real_distribution<float> u(0.f, 1.f);
      for (int64_t z = 0; z < nz; ++z) {
        float wet = z == 0 ? 0.97f : 1.0f - (float)z / (float)nz;  // level 0 is never worth packing
        if (sh.pattern) wet = sh.pattern == 1 ? 2.0f : -1.0f;
        for (int64_t c = 0; c < ncol; ++c) V[(size_t)z * ncol + c] = u(rng) < wet ? 1.0f + u(rng) : NAN;
      }
      for (size_t i = 0; i < all; ++i) {
        const bool here = present(V[i % lvl]);
        const float r = u(rng);
        // present cells: data with a few holes; absent cells: NaN or junk the reference never reads
        T[i] = here ? (r < 0.01f ? NAN : 10.f + 5.f * u(rng)) : (r < 0.5f ? NAN : 99.f);
        S[i] = here ? (r > 0.99f ? NAN : 35.f + u(rng)) : (r < 0.5f ? NAN : -7.f);
      }
      // what the toy kernels give on the caller's whole arrays
      std::vector<double> rho_ref(lvl), want[3], want_m(nt);
      double want_sums[2];
      toy_reference(T, S, V, nz, ncol, rho_ref.data(), want_sums);
      for (auto& w : want) w.resize((size_t)nt * ncol);
      toy_local(T, S, 0, 0, rho_ref.data(), V, nt, nz, ncol, want[0].data());
      toy_local(T, S, 0, 1, rho_ref.data(), V, nt, nz, ncol, want[1].data());  // thermosteric: S held at step 0
      toy_local(T, S, 1, 0, rho_ref.data(), V, nt, nz, ncol, want[2].data());  // halosteric: T held at step 0
      toy_global(T, S, 0, 0, V, nt, nz, ncol, want_m.data());
      std::vector<double> z_i(nz + 1, 0.0), depth(ncol, 1.0), p(nz, 0.0), masso(nt), rho_out(lvl);
      int packable = 0;  // levels with less than 90 % of their cells present: what mode 2 packs from pinned memory
      for (int64_t z = 0; z < nz; ++z) {
        int64_t n = 0;
        for (int64_t c = 0; c < ncol; ++c) n += present(V[(size_t)z * ncol + c]);
        packable += (double)n < 0.9 * (double)ncol;
      }
      if (sh.pattern == 2) packable = 0;  // nothing present at all: the library does not bother (every row as it is)
      const int spws[] = {1, 2, (int)nt};
      for (int mode = 0; mode <= 4; ++mode)
        for (int threads : {0, 1, 3, 7})  // 0: the library tunes the number of packing threads window by window
          for (int spw : spws)
            for (int slow = 0; slow < 2; ++slow) {
              if ((mode == 0 && threads != 1) || (mode != 1 && threads == 0)) continue;
              simcuda::copy_ns_per_kib() = slow ? 300 : 0;
              ml_host_set_packing(mode, threads);
              double sums[2] = {0, 0};
              for (auto& e : eta) memset(e, 0, (size_t)nt * ncol * sizeof(double));
              int rc = ml_steric_local_variants_host(0, ML_F32, T, S, V, z_i.data(), depth.data(), p.data(), -1.0, nt, nz,
                                                     ncol, spw, eta[0], eta[1], eta[2], nullptr, sums);
              const double frac = ml_host_last_packed_fraction();
              bool ok = rc == 0 && sums[0] == want_sums[0] && sums[1] == want_sums[1];
              for (int v = 0; v < 3; ++v) ok = ok && same(eta[v], want[v].data(), (size_t)nt * ncol);
              ok = ok && (mode != 0 || frac == 0.0) && (mode == 0 || pinned || frac == (sh.pattern == 2 ? 0.0 : 1.0)) &&
                   (mode != 2 || !pinned || fabs(frac - (double)packable / (double)nz) < 1e-9) &&
                   (mode == 0 || !pinned || frac <= (double)packable / (double)nz + 1e-9);
              if (pinned) {
                frac_sum[mode][slow] += frac;
                frac_max[mode][slow] += (double)packable / (double)nz;
                frac_n[mode][slow]++;
              }
              rc = ml_steric_global_host(0, ML_F32, T, S, V, p.data(), nt, nz, ncol, spw, masso.data());
              ok = ok && rc == 0 && same(masso.data(), want_m.data(), (size_t)nt);
              // a call that wants rho_ref back sends every row as it is
              rc = ml_steric_local_host(0, ML_F32, T, S, V, z_i.data(), depth.data(), p.data(), -1.0, nt, nz, ncol, spw,
                                        eta[0], rho_out.data(), sums);
              ok = ok && rc == 0 && ml_host_last_packed_fraction() == 0.0 && same(eta[0], want[0].data(), (size_t)nt * ncol) &&
                   same(rho_out.data(), rho_ref.data(), lvl);
              ++runs;
              if (!ok) {
                ++bad;
                printf("MISMATCH nt=%lld nz=%lld ncol=%lld pinned=%d mode=%d threads=%d spw=%d slow=%d frac=%.3f\n",
                       (long long)nt, (long long)nz, (long long)ncol, pinned, mode, threads, spw, slow, frac);
              }
            }
      // ---- a tuner that has settled on plain copies: the call skips the index and the packers altogether, counts its
      // windows up to the next retry point and no further, and the call after that one tries the choices again
      if (pinned && sh.pattern != 2) {
        simcuda::copy_ns_per_kib() = 0;
        ml_host_set_packing(1, 0);
        bool ok = ml_steric_global_host(0, ML_F32, T, S, V, p.data(), nt, nz, ncol, 1, masso.data()) == 0;  // the table is this grid's
        PackTuner& tn = resources().tuner;
        ok = ok && tn.nz == nz && tn.ncol == ncol;
        const double table[4] = {5.0, 1.0, 5.0, 5.0};
        for (int i = 0; i < 4; ++i) tn.ms_per_step[i] = table[i];
        tn.windows = PackTuner::kRetry - 3;
        ok = ok && tn.settled_on_none();
        for (auto& e : eta) memset(e, 0, (size_t)nt * ncol * sizeof(double));
        ok = ok && ml_steric_local_variants_host(0, ML_F32, T, S, V, z_i.data(), depth.data(), p.data(), -1.0, nt, nz, ncol, 1,
                                                 eta[0], eta[1], eta[2], nullptr, nullptr) == 0;
        ok = ok && ml_host_last_packed_fraction() == 0.0 && ml_host_last_pack_threads() == 0;
        for (int v = 0; v < 3; ++v) ok = ok && same(eta[v], want[v].data(), (size_t)nt * ncol);
        ok = ok && tn.windows == std::min<int64_t>(PackTuner::kRetry - 3 + nt, PackTuner::kRetry);
        if (nt >= 3) {  // parked at the retry point: this call plans again and walks through the choices
          ok = ok && !tn.settled_on_none();
          ok = ok && ml_steric_global_host(0, ML_F32, T, S, V, p.data(), nt, nz, ncol, 1, masso.data()) == 0;
          ok = ok && same(masso.data(), want_m.data(), (size_t)nt) && tn.windows == PackTuner::kRetry + nt;
        }
        ++runs;
        if (!ok) {
          ++bad;
          printf("SETTLED TUNER MISMATCH nt=%lld nz=%lld ncol=%lld windows=%lld\n", (long long)nt, (long long)nz, (long long)ncol,
                 (long long)tn.windows);
        }
      }
      // ---- the same fields pushed block by block (ml_host_stream_*): blocks of uneven length travel through two
      // scratch buffers, and a buffer is scribbled over as soon as the contract allows it (after the NEXT push returns)
      {
        std::vector<double> want_g[3];
        for (auto& w : want_g) w.resize((size_t)nt);
        toy_global(T, S, 0, 0, V, nt, nz, ncol, want_g[0].data());
        toy_global(T, S, 0, 1, V, nt, nz, ncol, want_g[1].data());
        toy_global(T, S, 1, 0, V, nt, nz, ncol, want_g[2].data());
        const int64_t maxb = 3;
        float* sc[2][2];
        for (int b = 0; b < 2; ++b)
          for (int f = 0; f < 2; ++f) sc[b][f] = alloc<float>((size_t)maxb * lvl, pinned);
        double* out_s[3];
        for (auto& o : out_s) o = alloc<double>((size_t)nt * ncol, pinned);
        for (int mode : {0, 1, 2, 4})
          for (int threads : {0, 1, 3})
            for (int domain = 0; domain < 2; ++domain)
              for (int supplied = 0; supplied < 2; ++supplied)
                for (int slow = 0; slow < 2; ++slow) {
                  if ((mode == 0 && threads != 1) || (mode != 1 && threads == 0)) continue;
                  simcuda::copy_ns_per_kib() = slow ? 300 : 0;
                  ml_host_set_packing(mode, threads);
                  for (auto& o : out_s) memset(o, 0, (size_t)nt * ncol * sizeof(double));
                  void* hs = nullptr;
                  double sums[2] = {0, 0};
                  int rc = ml_host_stream_begin(domain, 0, ML_F32, 7, V, ML_F32, supplied ? T : nullptr, supplied ? S : nullptr,
                                                supplied ? rho_ref.data() : nullptr, z_i.data(), depth.data(), p.data(), -1.0, nz,
                                                ncol, maxb, domain == 1 ? 2 : 0, &hs);
                  bool ok = rc == 0;
                  int64_t t = 0;
                  for (int k = 0; ok && t < nt; ++k) {
                    const int64_t n = std::min<int64_t>(nt - t, 1 + (k % maxb));  // 1, 2, 3, 1, ... steps
                    memcpy(sc[k & 1][0], T + (size_t)t * lvl, (size_t)n * lvl * sizeof(float));
                    memcpy(sc[k & 1][1], S + (size_t)t * lvl, (size_t)n * lvl * sizeof(float));
                    const size_t o = (size_t)t * (domain == 0 ? (size_t)ncol : 1);
                    rc = ml_host_stream_push(hs, sc[k & 1][0], sc[k & 1][1], n, out_s[0] + o, out_s[1] + o, out_s[2] + o);
                    ok = ok && rc == 0;
                    if (k >= 1)  // the previous block's memory is the caller's again
                      for (int f = 0; f < 2; ++f)
                        for (size_t i = 0; i < (size_t)maxb * lvl; ++i) sc[(k - 1) & 1][f][i] = -1234.5f;
                    t += n;
                  }
                  if (ok) {
                    rc = ml_host_stream_finish(hs, nullptr, sums);
                    ok = rc == 0;
                  } else if (hs) {
                    ml_host_stream_abort(hs);
                  }
                  for (int v = 0; v < 3 && ok; ++v)
                    ok = domain == 0 ? same(out_s[v], want[v].data(), (size_t)nt * ncol) : same(out_s[v], want_g[v].data(), (size_t)nt);
                  if (ok && !supplied) ok = sums[0] == want_sums[0] && sums[1] == want_sums[1];
                  ++runs;
                  if (!ok) {
                    ++bad;
                    printf("STREAM MISMATCH nt=%lld nz=%lld ncol=%lld pinned=%d mode=%d threads=%d domain=%d supplied=%d slow=%d rc=%d %s\n",
                           (long long)nt, (long long)nz, (long long)ncol, pinned, mode, threads, domain, supplied, slow, rc,
                           ml::tls().err);
                  }
                }
        // misuse: a second stream on the same thread, a block that is too long, a missing output
        {
          void *a = nullptr, *b2 = nullptr;
          int rc = ml_host_stream_begin(1, 0, ML_F32, 1, V, ML_F32, nullptr, nullptr, nullptr, nullptr, nullptr, p.data(), 0.0, nz,
                                        ncol, 2, 0, &a);
          bool ok = rc == 0 && ml_host_stream_begin(1, 0, ML_F32, 1, V, ML_F32, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                    p.data(), 0.0, nz, ncol, 2, 0, &b2) != 0;
          ok = ok && ml_host_stream_push(a, T, S, 3, out_s[0], nullptr, nullptr) != 0;   // longer than max_block_steps
          ok = ok && ml_host_stream_push(a, T, S, 1, nullptr, nullptr, nullptr) != 0;    // no output
          ok = ok && ml_host_stream_abort(a) == 0;
          ++runs;
          if (!ok) {
            ++bad;
            printf("STREAM MISUSE not rejected\n");
          }
        }
        for (int b = 0; b < 2; ++b)
          for (int f = 0; f < 2; ++f) release(sc[b][f], pinned);
        for (auto& o : out_s) release(o, pinned);
      }
      release(T, pinned);
      release(S, pinned);
      release(V, pinned);
      for (auto& e : eta) release(e, pinned);
    }
  }
  ml_host_release();
  for (int mode = 1; mode <= 4; ++mode)
    for (int slow = 0; slow < 2; ++slow)
      printf("mode %d, %s copies: %.2f of the rows packed on average (%.2f packable)\n", mode, slow ? "slow" : "fast",
             frac_sum[mode][slow] / frac_n[mode][slow], frac_max[mode][slow] / frac_n[mode][slow]);
  printf("hostpath_sim: %d runs, %d mismatches\n", runs, bad);
  return bad ? 1 : 0;
}

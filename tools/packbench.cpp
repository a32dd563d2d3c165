// packbench.cpp -- host-side rates behind the packed transfer of the *_host entry points: the presence index of
// a volcello field, one OM4p25 step of T and S compressed by N threads (csrc/ml_pack.cpp), and a plain N-thread
// memcpy of the same bytes for comparison.  No GPU involved.
//   g++ -O3 -std=c++17 -pthread [-DML_PACK_PLAIN_STORES] -o packbench packbench.cpp ../momlevel_b200/csrc/ml_pack.cpp
//   ./packbench <threads>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <atomic>
#include "../include/momlevel_b200.h"
int main(int argc, char** argv) {
  int nthr = argc > 1 ? atoi(argv[1]) : 8;
  const int64_t nz = 75, ncol = 1440 * 1080;
  const int64_t ngrp = (ncol + 31) / 32;
  std::vector<float> V(nz * ncol), T(nz * ncol), S(nz * ncol);
  std::vector<float> depth(ncol);
  srand(1);
  for (int64_t c = 0; c < ncol; ++c) depth[c] = ((c / 16) % 90 * 7919 % 10 < 3) ? -1.f : (rand() / (float)RAND_MAX);
  for (int64_t z = 0; z < nz; ++z) {
    float zi = powf(z / 75.f, 2.2f);
    for (int64_t c = 0; c < ncol; ++c) {
      bool wet = depth[c] > zi;
      V[z * ncol + c] = wet ? 1.f : NAN;
      T[z * ncol + c] = wet ? 10.f + c % 7 : NAN;
      S[z * ncol + c] = wet ? 35.f : NAN;
    }
  }
  std::vector<uint32_t> words(nz * ngrp), before(nz * ngrp);
  std::vector<uint64_t> cnt(nz);
  auto t0 = std::chrono::steady_clock::now();
  uint64_t total = ml_pack_index_rows(V.data(), nz, ncol, words.data(), before.data(), cnt.data());
  auto t1 = std::chrono::steady_clock::now();
  printf("simd %d index 1 thread: %.1f ms, wet %.3f\n", ml_pack_simd(), std::chrono::duration<double, std::milli>(t1 - t0).count(), total / (double)(nz * ncol));
  std::vector<uint64_t> off(nz + 1, 0);
  for (int z = 0; z < nz; ++z) off[z + 1] = off[z] + cnt[z];
  std::vector<float> Tp(total + 64), Sp(total + 64);
  const int nseg = 24;
  for (int rep = 0; rep < 4; ++rep) {
    std::atomic<int> next{0};
    auto a = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int i = 0; i < nthr; ++i)
      th.emplace_back([&] {
        for (;;) {
          int k = next.fetch_add(1);
          if (k >= nz * nseg) break;
          int z = k / nseg, s = k % nseg;
          int64_t g0 = ngrp * s / nseg, g1 = ngrp * (s + 1) / nseg;
          ml_pack_rows(T.data() + z * ncol, S.data() + z * ncol, words.data() + z * ngrp, before.data() + z * ngrp, g0, g1, ncol, Tp.data() + off[z], Sp.data() + off[z]);
        }
      });
    for (auto& t : th) t.join();
    auto b = std::chrono::steady_clock::now();
    double ms = std::chrono::duration<double, std::milli>(b - a).count();
    printf("pack step (%d thr): %.2f ms  read %.1f GB/s  (read+write %.1f GB/s)\n", nthr, ms, 2 * nz * ncol * 4 / ms / 1e6, (2 * nz * ncol * 4 + 2 * total * 4) / ms / 1e6);
  }
  {  // the same bytes through memcpy
    const size_t n = (size_t)nz * ncol * 4;
    std::vector<float> D(nz * ncol);
    for (int rep = 0; rep < 3; ++rep) {
      auto a = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for (int i = 0; i < nthr; ++i)
        th.emplace_back([&, i] { size_t c = n / nthr; memcpy((char*)D.data() + i * c, (const char*)T.data() + i * c, c); });
      for (auto& t : th) t.join();
      double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
      printf("memcpy (%d thr): %.2f ms  %.1f GB/s copied\n", nthr, ms, n / ms / 1e6);
    }
  }
  // verify
  uint64_t k = 0; int bad = 0;
  for (int64_t i = 0; i < nz * ncol; ++i) if (!std::isnan(V[i])) { if (Tp[k] != T[i] || Sp[k] != S[i]) ++bad; ++k; }
  printf("k=%llu total=%llu bad=%d\n", (unsigned long long)k, (unsigned long long)total, bad);
}

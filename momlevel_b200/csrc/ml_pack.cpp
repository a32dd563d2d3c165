// ml_pack.cpp -- host-side wet-cell packing for the *_host entry points (plain C++, no CUDA).
//
// The host path is bound by PCIe, and roughly half of an ocean grid is land or below the sea
// floor.  The reference discards those cells itself: delta_rho is NaN wherever the reference
// volcello is missing (src/momlevel/steric.py:151-153), the column sum skips NaN (:163) and
// volo / masso skip them too (derived.py:435-438, 787-789).  So the values T and S hold where
// volcello(t=0) is NaN never reach a result, and a level row can cross PCIe as its wet cells
// only.  This file holds the two CPU loops of that scheme:
//
//   ml_pack_index_rows   volcello rows -> one 32-column presence word per group + the
//                        running count of present cells in front of each group
//   ml_pack_rows         compress one segment of a T row and an S row through those words
//
// Both have an AVX-512 body (VCOMPRESSPS in registers, whole lines written with non-temporal stores)
// chosen at run time and a scalar body for anything else.  Threading lives with the caller (ml_hostpath.cu).
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

// The AVX-512 bodies exist on x86-64 hosts only; anything else (an aarch64 Grace host in front of the same GPU)
// builds the scalar bodies alone.
#if defined(__x86_64__) && !defined(ML_PACK_NO_X86)
#define ML_PACK_X86 1
#include <immintrin.h>
#else
#define ML_PACK_X86 0
#endif

#include "../../include/momlevel_b200.h"

namespace {

// ML_PACK_FORCE_SCALAR=1 in the environment selects the scalar bodies on any CPU (tests/test_pack.py)
bool has_avx512() {
#if ML_PACK_X86
  static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                         __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("popcnt") &&
                         !(getenv("ML_PACK_FORCE_SCALAR") && getenv("ML_PACK_FORCE_SCALAR")[0] == '1');
  return ok;
#else
  return false;
#endif
}

inline bool present_f32(uint32_t bits) { return (bits & 0x7fffffffu) <= 0x7f800000u; }  // not NaN

// ---- index -------------------------------------------------------------------------------

uint64_t index_scalar(const float* v, int64_t ncol, uint32_t* words, uint32_t* before) {
  const int64_t ngrp = (ncol + 31) / 32;
  uint64_t run = 0;
  for (int64_t g = 0; g < ngrp; ++g) {
    const int64_t c0 = g * 32;
    const int n = (int)((ncol - c0) < 32 ? (ncol - c0) : 32);
    uint32_t m = 0;
    for (int i = 0; i < n; ++i) {
      uint32_t b;
      memcpy(&b, v + c0 + i, 4);
      m |= (uint32_t)present_f32(b) << i;
    }
    words[g] = m;
    before[g] = (uint32_t)run;
    run += (uint64_t)__builtin_popcount(m);
  }
  return run;
}

#if ML_PACK_X86
__attribute__((target("avx512f,avx512bw,avx512vl,popcnt"))) uint64_t index_avx512(const float* v, int64_t ncol,
                                                                                   uint32_t* words,
                                                                                   uint32_t* before) {
  const int64_t nfull = ncol / 32;
  uint64_t run = 0;
  for (int64_t g = 0; g < nfull; ++g) {
    const __m512 a = _mm512_loadu_ps(v + g * 32);
    const __m512 b = _mm512_loadu_ps(v + g * 32 + 16);
    const uint32_t m = (uint32_t)_mm512_cmp_ps_mask(a, a, _CMP_ORD_Q) |
                       ((uint32_t)_mm512_cmp_ps_mask(b, b, _CMP_ORD_Q) << 16);
    words[g] = m;
    before[g] = (uint32_t)run;
    run += (uint64_t)_mm_popcnt_u32(m);
  }
  if (nfull * 32 < ncol) {  // ragged last group
    uint32_t w, b;
    const uint64_t n = index_scalar(v + nfull * 32, ncol - nfull * 32, &w, &b);
    words[nfull] = w;
    before[nfull] = (uint32_t)run;
    run += n;
  }
  return run;
}

#else
uint64_t index_avx512(const float* v, int64_t ncol, uint32_t* words, uint32_t* before) {  // never selected
  return index_scalar(v, ncol, words, before);
}
#endif

// ---- pack --------------------------------------------------------------------------------

void pack_scalar(const float* t, const float* s, const uint32_t* words, int64_t g0, int64_t g1, float* t_out,
                 float* s_out) {
  for (int64_t g = g0; g < g1; ++g) {
    uint32_t m = words[g];
    const float* tp = t + g * 32;
    const float* sp = s + g * 32;
    while (m) {
      const int i = __builtin_ctz(m);
      *t_out++ = tp[i];
      *s_out++ = sp[i];
      m &= m - 1;
    }
  }
}

#if !ML_PACK_X86
void pack_avx512(const float* t, const float* s, const uint32_t* words, int64_t g0, int64_t g1, int64_t, float* t_out,
                 float* s_out, bool) {  // never selected
  pack_scalar(t, s, words, g0, g1, t_out, s_out);
}
#elif defined(ML_PACK_PLAIN_STORES)  // A/B: masked stores straight into the staging memory (tools/packbench.cpp)
__attribute__((target("avx512f,avx512bw,avx512vl,popcnt,bmi2"))) void pack_avx512(const float* t, const float* s,
                                                                                 const uint32_t* words, int64_t g0,
                                                                                 int64_t g1, int64_t ncol,
                                                                                 float* t_out, float* s_out, bool) {
  // a ragged last group is left to the scalar body (a 16-lane load would run past the row)
  const int64_t gfull = (g1 * 32 <= ncol) ? g1 : g1 - 1;
  for (int64_t g = g0; g < gfull; ++g) {
    const uint32_t m = words[g];
    if (m == 0) continue;
    const float* tp = t + g * 32;
    const float* sp = s + g * 32;
    if (m == 0xffffffffu) {
      _mm512_storeu_ps(t_out, _mm512_loadu_ps(tp));
      _mm512_storeu_ps(t_out + 16, _mm512_loadu_ps(tp + 16));
      _mm512_storeu_ps(s_out, _mm512_loadu_ps(sp));
      _mm512_storeu_ps(s_out + 16, _mm512_loadu_ps(sp + 16));
      t_out += 32;
      s_out += 32;
      continue;
    }
    const __mmask16 lo = (__mmask16)(m & 0xffffu), hi = (__mmask16)(m >> 16);
    const int nlo = _mm_popcnt_u32(m & 0xffffu), nhi = _mm_popcnt_u32(m >> 16);
    const __mmask16 slo = (__mmask16)((1u << nlo) - 1u), shi = (__mmask16)((1u << nhi) - 1u);
    // the compress stays in registers; the store is masked to the count so that a segment never
    // writes past its own share of the row (another thread owns what follows)
    _mm512_mask_storeu_ps(t_out, slo, _mm512_maskz_compress_ps(lo, _mm512_loadu_ps(tp)));
    _mm512_mask_storeu_ps(s_out, slo, _mm512_maskz_compress_ps(lo, _mm512_loadu_ps(sp)));
    t_out += nlo;
    s_out += nlo;
    _mm512_mask_storeu_ps(t_out, shi, _mm512_maskz_compress_ps(hi, _mm512_loadu_ps(tp + 16)));
    _mm512_mask_storeu_ps(s_out, shi, _mm512_maskz_compress_ps(hi, _mm512_loadu_ps(sp + 16)));
    t_out += nhi;
    s_out += nhi;
  }
  if (gfull < g1) pack_scalar(t, s, words, gfull, g1, t_out, s_out);
}

#else
// Compressed vectors are appended to a small cache-resident buffer and leave it as whole, aligned
// 64-byte lines written with non-temporal stores: the staging memory is only ever read by the DMA
// engine, so pulling its lines into the cache first (what an ordinary store does) is wasted traffic,
// and a masked store that straddles two lines costs several cycles more than a full one.  The
// first and the last line of a segment are shared with its neighbours and are written with masks.
struct LineWriter {
  static constexpr int kLines = 16;
  alignas(64) float buf[(kLines + 2) * 16];
  float* line;  // 64-byte aligned destination of buf[0]
  int n;        // floats in buf (the first `head` of them are not ours)
  int head;
  bool cached;  // ordinary stores: the destination is a small ring meant to stay in the last-level cache
  __attribute__((target("avx512f,avx512bw,avx512vl"))) LineWriter(float* dst, bool keep_in_cache) {
    cached = keep_in_cache;
    head = (int)(((uintptr_t)dst & 63u) >> 2);
    line = dst - head;
    n = head;
  }
  __attribute__((target("avx512f,avx512bw,avx512vl"))) inline void put(__m512 v, int count) {
    _mm512_storeu_ps(buf + n, v);
    n += count;
    if (n >= kLines * 16) flush_full();
  }
  __attribute__((target("avx512f,avx512bw,avx512vl"))) inline void flush_full() {
    const int nl = n >> 4;
    int l = 0;
    if (head) {  // the line we share with the segment in front
      _mm512_mask_storeu_ps(line, (__mmask16)(0xffffu << head), _mm512_load_ps(buf));
      head = 0;
      l = 1;
    }
    if (cached)
      for (; l < nl; ++l) _mm512_store_ps(line + l * 16, _mm512_load_ps(buf + l * 16));
    else
      for (; l < nl; ++l) _mm512_stream_ps(line + l * 16, _mm512_load_ps(buf + l * 16));
    _mm512_store_ps(buf, _mm512_load_ps(buf + nl * 16));
    line += nl * 16;
    n &= 15;
  }
  __attribute__((target("avx512f,avx512bw,avx512vl"))) inline void finish() {
    if (n >= 16) flush_full();
    if (n > head) {
      const __mmask16 m = (__mmask16)(((1u << n) - 1u) & (0xffffu << head));
      _mm512_mask_storeu_ps(line, m, _mm512_load_ps(buf));
    }
    _mm_sfence();
  }
};

__attribute__((target("avx512f,avx512bw,avx512vl,popcnt,bmi2"))) void pack_avx512(const float* t, const float* s,
                                                                                 const uint32_t* words, int64_t g0,
                                                                                 int64_t g1, int64_t ncol,
                                                                                 float* t_out, float* s_out,
                                                                                 bool cached) {
  // a ragged last group is left to the scalar body (a 16-lane load would run past the row)
  const int64_t gfull = (g1 * 32 <= ncol) ? g1 : g1 - 1;
  LineWriter wt(t_out, cached), ws(s_out, cached);
  int64_t written = 0;
  for (int64_t g = g0; g < gfull; ++g) {
    const uint32_t m = words[g];
    if (m == 0) continue;
    const float* tp = t + g * 32;
    const float* sp = s + g * 32;
    const __mmask16 lo = (__mmask16)(m & 0xffffu), hi = (__mmask16)(m >> 16);
    const int nlo = _mm_popcnt_u32(m & 0xffffu), nhi = _mm_popcnt_u32(m >> 16);
    wt.put(_mm512_maskz_compress_ps(lo, _mm512_loadu_ps(tp)), nlo);
    ws.put(_mm512_maskz_compress_ps(lo, _mm512_loadu_ps(sp)), nlo);
    wt.put(_mm512_maskz_compress_ps(hi, _mm512_loadu_ps(tp + 16)), nhi);
    ws.put(_mm512_maskz_compress_ps(hi, _mm512_loadu_ps(sp + 16)), nhi);
    written += nlo + nhi;
  }
  wt.finish();
  ws.finish();
  if (gfull < g1) pack_scalar(t, s, words, gfull, g1, t_out + written, s_out + written);
}

#endif

}  // namespace

extern "C" uint64_t ml_pack_index_rows(const float* v, int64_t nrows, int64_t ncol, uint32_t* words,
                                       uint32_t* before, uint64_t* row_count) {
  const int64_t ngrp = (ncol + 31) / 32;
  uint64_t total = 0;
  for (int64_t r = 0; r < nrows; ++r) {
    const uint64_t n = has_avx512() ? index_avx512(v + r * ncol, ncol, words + r * ngrp, before + r * ngrp)
                                    : index_scalar(v + r * ncol, ncol, words + r * ngrp, before + r * ngrp);
    row_count[r] = n;
    total += n;
  }
  return total;
}

static void pack_rows(const float* t_row, const float* s_row, const uint32_t* words, const uint32_t* before, int64_t g0,
                      int64_t g1, int64_t ncol, float* t_out, float* s_out, bool cached) {
  if (g0 >= g1) return;
  t_out += before[g0];
  s_out += before[g0];
  if (has_avx512())
    pack_avx512(t_row, s_row, words, g0, g1, ncol, t_out, s_out, cached);
  else
    pack_scalar(t_row, s_row, words, g0, g1, t_out, s_out);
}

extern "C" void ml_pack_rows(const float* t_row, const float* s_row, const uint32_t* words, const uint32_t* before,
                             int64_t g0, int64_t g1, int64_t ncol, float* t_out, float* s_out) {
  pack_rows(t_row, s_row, words, before, g0, g1, ncol, t_out, s_out, false);
}

extern "C" void ml_pack_rows_cached(const float* t_row, const float* s_row, const uint32_t* words,
                                    const uint32_t* before, int64_t g0, int64_t g1, int64_t ncol, float* t_out,
                                    float* s_out) {
  pack_rows(t_row, s_row, words, before, g0, g1, ncol, t_out, s_out, true);
}

extern "C" int ml_pack_simd(void) { return has_avx512() ? 512 : 0; }

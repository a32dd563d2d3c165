// ml_hostpath.cu -- ml_steric_local_host / ml_steric_global_host: the steric path on HOST buffers.
//
// What momlevel.steric(dset) does for variant="steric", domain="local" when the Dataset
// lives in host memory (src/momlevel/steric.py:84-184 with the reference state taken from
// time step 0, src/momlevel/reference.py:60-80): time steps are streamed through two device
// windows, the host->device copy of window k+1 running on a copy stream while the kernels of
// window k run on the compute stream.  Pinned (page-locked) host buffers make the copies
// truly asynchronous; pageable buffers work but serialise inside the driver.
#include <sched.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "ml_host.cuh"

namespace {

// Worker threads of the packed transfer (below).  start() hands every worker the same job, the caller does
// its own share of the window meanwhile and then wait()s.
class Pool {
 public:
  ~Pool() {
    {
      std::lock_guard<std::mutex> l(m_);
      stop_ = true;
    }
    cv_start_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return (int)th_.size(); }
  void ensure(int n) {
    std::lock_guard<std::mutex> l(m_);
    while ((int)th_.size() < n) {
      const int id = (int)th_.size();
      const uint64_t born = gen_;
      th_.emplace_back([this, id, born] { loop(id, born); });
    }
  }
  void start(int active, std::function<void(int)> job) {
    std::lock_guard<std::mutex> l(m_);
    job_ = std::move(job);
    active_ = std::min(active, (int)th_.size());
    running_ = active_;
    ++gen_;
    cv_start_.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> l(m_);
    cv_done_.wait(l, [this] { return running_ == 0; });
  }

 private:
  void loop(int id, uint64_t seen) {
    for (;;) {
      std::function<void(int)> job;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_start_.wait(l, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
        if (id >= active_) continue;
        job = job_;
      }
      job(id);
      {
        std::lock_guard<std::mutex> l(m_);
        if (--running_ == 0) cv_done_.notify_all();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_start_, cv_done_;
  std::function<void(int)> job_;
  uint64_t gen_ = 0;
  int active_ = 0, running_ = 0;
  bool stop_ = false;
};

// Device staging buffers, pinned host staging, streams, events and the worker threads are kept per host
// thread between calls (a year of OM4p25 needs ~5 GB of windows; allocating and freeing them costs tens of
// milliseconds per call) and released by ml_host_release() or at thread exit.
struct Resources {
  static constexpr int kSlots = 28;
  static constexpr int kHostSlots = 12;
  static constexpr int kRing = 3;  // dense rows in flight on the copy stream
  static constexpr int kStageRing = 6;  // packed rows between the packers and the copy engine (packing mode 3)
  void* buf[kSlots] = {nullptr};
  size_t cap[kSlots] = {0};
  void* hbuf[kHostSlots] = {nullptr};
  size_t hcap[kHostSlots] = {0};
  int device = -1;
  cudaStream_t copy = nullptr, comp = nullptr, back = nullptr;  // host->device, kernels, device->host
  cudaEvent_t copied[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
  cudaEvent_t ring[kRing] = {nullptr};
  cudaEvent_t slot_done[kStageRing] = {nullptr};  // the copies out of a slot of the staging ring have finished
  Pool pool;
  // settings and accounting of the packed transfer (ml_host_set_packing and friends)
  int pack_mode = 1, pack_threads = 0;
  double last_packed_fraction = 0.0;
  double last_ms[4] = {0.0, 0.0, 0.0, 0.0};  // presence index, windows (host side), drain, whole call
  std::atomic<uint64_t> h2d_bytes{0};
  void release() {
    for (int i = 0; i < kSlots; ++i) {
      if (buf[i]) cudaFree(buf[i]);
      buf[i] = nullptr;
      cap[i] = 0;
    }
    for (int i = 0; i < kHostSlots; ++i) {
      if (hbuf[i]) cudaFreeHost(hbuf[i]);
      hbuf[i] = nullptr;
      hcap[i] = 0;
    }
    for (int i = 0; i < 2; ++i) {
      if (copied[i]) cudaEventDestroy(copied[i]);
      if (freed[i]) cudaEventDestroy(freed[i]);
      copied[i] = freed[i] = nullptr;
    }
    for (int i = 0; i < kRing; ++i) {
      if (ring[i]) cudaEventDestroy(ring[i]);
      ring[i] = nullptr;
    }
    for (int i = 0; i < kStageRing; ++i) {
      if (slot_done[i]) cudaEventDestroy(slot_done[i]);
      slot_done[i] = nullptr;
    }
    if (copy) cudaStreamDestroy(copy);
    if (comp) cudaStreamDestroy(comp);
    if (back) cudaStreamDestroy(back);
    copy = comp = back = nullptr;
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
    if (!back && (e = cudaStreamCreateWithFlags(&back, cudaStreamNonBlocking)) != cudaSuccess) return e;
    for (int b = 0; b < 2; ++b) {
      if (!copied[b] && (e = cudaEventCreateWithFlags(&copied[b], cudaEventDisableTiming)) != cudaSuccess) return e;
      if (!freed[b] && (e = cudaEventCreateWithFlags(&freed[b], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    for (int i = 0; i < kRing; ++i)
      if (!ring[i] && (e = cudaEventCreateWithFlags(&ring[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    for (int i = 0; i < kStageRing; ++i)
      if (!slot_done[i] && (e = cudaEventCreateWithFlags(&slot_done[i], cudaEventDisableTiming)) != cudaSuccess)
        return e;
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
  // pinned host slot i grows to at least `bytes`
  cudaError_t halloc(int i, void** p, size_t bytes) {
    if (hcap[i] < bytes) {
      if (hbuf[i]) cudaFreeHost(hbuf[i]);
      hbuf[i] = nullptr;
      hcap[i] = 0;
      cudaError_t e = cudaHostAlloc(&hbuf[i], bytes, cudaHostAllocDefault);
      if (e != cudaSuccess) return e;
      hcap[i] = bytes;
    }
    *p = hbuf[i];
    return cudaSuccess;
  }
};

// true for memory the driver does not know (malloc / numpy): copies from it are staged by the driver
bool is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

Resources& resources() {
  static thread_local Resources r;
  return r;
}

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Ranks that share this host's cores and memory system: one process per GPU under torchrun exports
// LOCAL_WORLD_SIZE; a lone process counts as one.
int local_ranks() {
  const char* v = getenv("LOCAL_WORLD_SIZE");
  if (v == nullptr || *v == 0) return 1;
  const long n = strtol(v, nullptr, 10);
  return (n >= 1 && n <= 1024) ? (int)n : 1;
}

int default_threads() {
  cpu_set_t set;
  int n = 0;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
  if (n <= 0) n = (int)std::thread::hardware_concurrency();
  // half the cores: the packers and the DMA engine share the host's memory bandwidth, and past that point
  // every extra thread slows the copies by as much as it saves (tools/e2e_sweep.py: 4 / 6 / 8 / 10 / 12 / 15
  // threads on a 16-core host gave 156 / 148 / 145 / 148 / 152 / 157 ms for an OM4p25 year).  That half is
  // shared by the ranks of the host: with every rank taking half the affinity mask, four ranks oversubscribed
  // the cores and eight lost to plain copies (SCALE_r01: packed 366 ms against 409 ms dense at four ranks with
  // cores / (2 ranks) threads each, a tie at eight).
  return std::max(1, std::min(n / (2 * local_ranks()), 64));
}

// memcpy by the worker threads (one thread moves ~10 GB/s, the memory system many times that)
void parallel_copy(Resources& r, void* dst, const void* src, size_t bytes, int threads) {
  threads = std::max(1, threads);
  if (bytes < (size_t)(8u << 20) || threads == 1) {
    memcpy(dst, src, bytes);
    return;
  }
  r.pool.ensure(threads);
  const size_t chunk = ((bytes + (size_t)threads - 1) / (size_t)threads + 63) & ~(size_t)63;
  r.pool.start(threads, [=](int id) {
    const size_t a = (size_t)id * chunk;
    if (a < bytes) memcpy((char*)dst + a, (const char*)src + a, std::min(chunk, bytes - a));
  });
  r.pool.wait();
}

// ---------------------------------------------------------------------------------------------------
// Packed transfer.  A window is nt_w * nz level rows of T and of S.  The reference never uses either
// field where the reference volcello is missing (steric.py:151-153, 163; derived.py:435-438), so a row
// may cross PCIe as its present cells only.  Compressing a row costs host memory bandwidth, moving it
// as it is costs PCIe time, and which of the two runs out first depends on the machine and on what
// else it is doing; so the rows of a window are put in order of how much of them is present and the
// two sides work towards each other: the calling thread queues the fullest rows as plain copies from
// the caller's buffer (at most kRing rows ahead of the copy engine), the workers compress the emptiest
// rows into pinned staging, segment by segment, and queue each one as soon as it is whole.  The
// window is done when they meet.  On the device k_unpack_rows spreads the compressed rows back into
// the dense window (NaN where the volume is missing) and the steric kernels run on it unchanged.
// ---------------------------------------------------------------------------------------------------
struct PackPlan {
  bool on = false;
  int mode = 0, threads = 0, nseg = 1;
  int64_t nz = 0, ncol = 0, ngrp = 0;
  uint64_t nwet = 0;            // present cells of one step
  uint32_t *words = nullptr, *before = nullptr;  // pinned host [nz][ngrp]
  uint64_t *lvloff = nullptr;   // pinned host [nz + 1]: present cells in the levels above
  std::vector<int> order;       // levels, fullest first
  int first_packable = 0;       // position in `order` of the first level worth compressing
  uint32_t *d_words = nullptr, *d_before = nullptr;
  uint64_t* d_lvloff = nullptr;
  float *stage[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // pinned [window parity][T,S]: [t][nwet]
  float *d_packed[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  uint8_t *flags[2] = {nullptr, nullptr}, *d_flags[2] = {nullptr, nullptr};  // [t][z]: 1 = row crossed packed
  int64_t rows_total = 0, rows_packed = 0;
  bool all_staged = false;  // pageable source: no row is copied straight from the caller's buffer
  // mode 3: packed rows wait for the copy engine in a ring of kStageRing row slots (T half, S half) that is
  // small enough to stay in the last-level cache, instead of in staging the size of the window
  bool ring = false;
  float* ring_buf = nullptr;
  int64_t ring_next = 0;                                    // rows that have taken a slot so far (under the lock)
  std::atomic<int64_t> slot_queued[Resources::kStageRing];  // rows whose copies out of the slot have been queued
  // mode 4: a row goes as it is only while no packed row is waiting for the copy stream
  bool packed_first = false;
  std::atomic<int> last_slot{-1};  // slot of the packed row queued last
};

constexpr double kPackableBelow = 0.9;  // a level with more of its cells present than this is never compressed
#ifdef ML_HOSTPATH_TEST_HOOKS
constexpr int64_t kSegmentGroups = 8;  // small rows are cut into several segments too
#else
constexpr int64_t kSegmentGroups = 2048;  // 32-column groups per segment of a row (64 K columns, 256 KB per field)
#endif

// tests/sim/hostpath_sim.cpp compiles this file with g++ against a simulated CUDA runtime (streams that run
// behind the host, events, pinned / pageable bookkeeping) to put the threads and the buffer recycling below under
// ThreadSanitizer on a box without a GPU; it brings its own launch_unpack().
#ifndef ML_HOSTPATH_TEST_HOOKS
__global__ void __launch_bounds__(256) k_unpack_rows(const float* __restrict__ pT, const float* __restrict__ pS,
                                                      float* __restrict__ T, float* __restrict__ S,
                                                      const uint32_t* __restrict__ words,
                                                      const uint32_t* __restrict__ before,
                                                      const uint64_t* __restrict__ lvloff,
                                                      const uint8_t* __restrict__ flags, int nz, int64_t ncol,
                                                      int64_t ngrp, uint64_t nwet, int xblocks) {
  const int row = blockIdx.x / xblocks;  // (t, z)
  if (!flags[row]) return;
  const int t = row / nz, z = row % nz;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t src0 = (uint64_t)t * nwet + lvloff[z];
  const size_t dst0 = (size_t)row * (size_t)ncol;
  const float nanf32 = __int_as_float(0x7fc00000);
  const int64_t gfirst = ((int64_t)(blockIdx.x % xblocks) * 8 + warp) * 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t g = gfirst + k;
    if (g >= ngrp) break;
    const uint32_t m = words[(size_t)z * ngrp + g];
    const uint64_t src = src0 + before[(size_t)z * ngrp + g] + __popc(m & ((1u << lane) - 1u));
    const bool here = (m >> lane) & 1u;
    const float a = here ? pT[src] : nanf32;
    const float b = here ? pS[src] : nanf32;
    const int64_t col = g * 32 + lane;
    if (col < ncol) {
      T[dst0 + col] = a;
      S[dst0 + col] = b;
    }
  }
}

int launch_unpack(cudaStream_t stream, int nrows, int xblocks, const float* pT, const float* pS, float* T, float* S,
                  const uint32_t* words, const uint32_t* before, const uint64_t* lvloff, const uint8_t* flags, int nz,
                  int64_t ncol, int64_t ngrp, uint64_t nwet) {
  k_unpack_rows<<<(unsigned)((int64_t)nrows * xblocks), 256, 0, stream>>>(pT, pS, T, S, words, before, lvloff, flags, nz,
                                                                          ncol, ngrp, nwet, xblocks);
  return ml::launched("k_unpack_rows");
}
#endif

// Presence words of the reference volcello, the device copies of them and the staging buffers.
// Leaves plan.on false (every row crosses as it is) when packing is off, cannot help or cannot get
// its pinned memory.
int plan_packing(Resources& r, PackPlan& plan, int dtype, const void* v0, int64_t nz, int64_t ncol, int64_t spw,
                 bool allowed, bool pageable) {
  using namespace ml;
  plan.on = false;
  plan.mode = r.pack_mode;
  plan.all_staged = false;
  if (!allowed || plan.mode == 0 || dtype != ML_F32) return ML_OK;
  plan.nz = nz;
  plan.ncol = ncol;
  plan.ngrp = (ncol + 31) / 32;
  plan.threads = r.pack_threads > 0 ? std::min(r.pack_threads, 64) : default_threads();
  plan.nseg = (int)std::max<int64_t>(1, std::min<int64_t>(64, plan.ngrp / kSegmentGroups));
  const size_t nw = (size_t)nz * (size_t)plan.ngrp;
  void *hw, *hb, *hl;
  if (r.halloc(0, &hw, nw * 4) != cudaSuccess || r.halloc(1, &hb, nw * 4) != cudaSuccess ||
      r.halloc(2, &hl, (size_t)(nz + 1) * 8) != cudaSuccess) {
    cudaGetLastError();
    return ML_OK;
  }
  plan.words = (uint32_t*)hw;
  plan.before = (uint32_t*)hb;
  plan.lvloff = (uint64_t*)hl;
  r.pool.ensure(plan.threads);
  std::vector<uint64_t> cnt((size_t)nz);
  {
    std::atomic<int64_t> next{0};
    const float* v = (const float*)v0;
    r.pool.start(plan.threads, [&](int) {
      for (;;) {
        const int64_t z = next.fetch_add(1);
        if (z >= nz) break;
        ml_pack_index_rows(v + z * ncol, 1, ncol, plan.words + z * plan.ngrp, plan.before + z * plan.ngrp, &cnt[z]);
      }
    });
    r.pool.wait();
  }
  plan.lvloff[0] = 0;
  for (int64_t z = 0; z < nz; ++z) plan.lvloff[z + 1] = plan.lvloff[z] + cnt[z];
  plan.nwet = plan.lvloff[nz];
  plan.order.resize((size_t)nz);
  for (int64_t z = 0; z < nz; ++z) plan.order[z] = (int)z;
  std::stable_sort(plan.order.begin(), plan.order.end(), [&](int a, int b) { return cnt[a] > cnt[b]; });
  plan.first_packable = (int)nz;
  for (int64_t i = 0; i < nz; ++i)
    if ((double)cnt[plan.order[i]] < kPackableBelow * (double)ncol) {
      plan.first_packable = (int)i;
      break;
    }
  if (pageable && plan.nwet != 0) {
    // Pageable source (a plain numpy array): a DMA from it is a synchronous, single-threaded bounce through the
    // driver at a fraction of the PCIe rate, so every row goes through the packers and the pinned staging,
    // full rows included.
    plan.all_staged = true;
    plan.first_packable = 0;
  }
  if (plan.first_packable == (int)nz || plan.nwet == 0) return ML_OK;  // nothing worth compressing
  const size_t stage_bytes = (size_t)spw * plan.nwet * 4;
  const size_t flag_bytes = (size_t)spw * (size_t)nz;
  plan.ring = plan.mode >= 3;
  plan.packed_first = plan.mode == 4;
  plan.last_slot.store(-1);
  plan.ring_next = 0;
  for (auto& q : plan.slot_queued) q.store(0, std::memory_order_relaxed);
  if (plan.ring) {
    void* h;
    if (r.halloc(3, &h, (size_t)Resources::kStageRing * 2 * (size_t)ncol * 4) != cudaSuccess) {
      cudaGetLastError();
      return ML_OK;
    }
    plan.ring_buf = (float*)h;
  }
  for (int b = 0; b < 2; ++b) {
    for (int f = 0; f < 2; ++f) {
      void *h = nullptr, *d;
      if (!plan.ring && r.halloc(3 + 2 * b + f, &h, stage_bytes) != cudaSuccess) {
        cudaGetLastError();
        return ML_OK;
      }
      ML_CUDA(r.alloc(16 + 2 * b + f, &d, stage_bytes));
      plan.stage[b][f] = (float*)h;
      plan.d_packed[b][f] = (float*)d;
    }
    void *hf, *df;
    if (r.halloc(7, &hf, 2 * flag_bytes) != cudaSuccess) {
      cudaGetLastError();
      return ML_OK;
    }
    ML_CUDA(r.alloc(23 + b, &df, flag_bytes));
    plan.flags[b] = (uint8_t*)hf + (size_t)b * flag_bytes;
    plan.d_flags[b] = (uint8_t*)df;
  }
  void *dw, *db, *dl;
  ML_CUDA(r.alloc(20, &dw, nw * 4));
  ML_CUDA(r.alloc(21, &db, nw * 4));
  ML_CUDA(r.alloc(22, &dl, (size_t)(nz + 1) * 8));
  plan.d_words = (uint32_t*)dw;
  plan.d_before = (uint32_t*)db;
  plan.d_lvloff = (uint64_t*)dl;
  ML_CUDA(cudaMemcpyAsync(dw, plan.words, nw * 4, cudaMemcpyHostToDevice, r.copy));
  ML_CUDA(cudaMemcpyAsync(db, plan.before, nw * 4, cudaMemcpyHostToDevice, r.copy));
  ML_CUDA(cudaMemcpyAsync(dl, plan.lvloff, (size_t)(nz + 1) * 8, cudaMemcpyHostToDevice, r.copy));
  r.h2d_bytes += 2 * nw * 4 + (size_t)(nz + 1) * 8;
  plan.on = true;
  return ML_OK;
}

// One window of T and S to the device buffers dT / dS of parity b: on return every copy is queued on the
// copy stream, copied[b] is recorded behind them and the compute stream waits for it (and has expanded
// the rows that crossed packed).  T_w / S_w point at the window's first step in the caller's buffers.
int stage_window(Resources& r, PackPlan& plan, int b, int64_t w, const void* T_w, const void* S_w, int64_t nt_w,
                 size_t es, int64_t nz, int64_t ncol, void* dT, void* dS) {
  using namespace ml;
  const size_t bytes = (size_t)nt_w * (size_t)nz * (size_t)ncol * es;
  if (w >= 2) ML_CUDA(cudaStreamWaitEvent(r.copy, r.freed[b], 0));
  if (!plan.on) {
    ML_CUDA(cudaMemcpyAsync(dT, T_w, bytes, cudaMemcpyHostToDevice, r.copy));
    ML_CUDA(cudaMemcpyAsync(dS, S_w, bytes, cudaMemcpyHostToDevice, r.copy));
    r.h2d_bytes += 2 * bytes;
    ML_CUDA(cudaEventRecord(r.copied[b], r.copy));
    ML_CUDA(cudaStreamWaitEvent(r.comp, r.copied[b], 0));
    return ML_OK;
  }
  // the staging of this parity was last read by the copies of window w - 2
  if (w >= 2) ML_CUDA(cudaEventSynchronize(r.copied[b]));

  const float* Th = (const float*)T_w;
  const float* Sh = (const float*)S_w;
  const int nrows = (int)(nt_w * nz);
  const size_t row_bytes = (size_t)ncol * 4;
  uint8_t* flags = plan.flags[b];
  memset(flags, 0, (size_t)nrows);
  // position i of the window's row order = step i % nt_w of level order[i / nt_w]
  auto row_of = [&](int i, int& t, int& z) {
    z = plan.order[i / (int)nt_w];
    t = i % (int)nt_w;
  };
  struct Shared {
    std::mutex m;
    int lo = 0, hi = 0, hi_min = 0, lo_end = 0;
    int cur = -1, next_seg = 0, cur_slot = 0;
    int packed = 0;
    std::atomic<int> err{(int)cudaSuccess};  // first CUDA error of any thread; set without the lock
  } sh;
  sh.hi = nrows - 1;
  sh.hi_min = plan.first_packable * (int)nt_w;
  sh.lo_end = (plan.mode == 2 || plan.all_staged) ? sh.hi_min : nrows;
  sh.next_seg = plan.nseg;  // no row open yet
  std::vector<std::atomic<int>> done((size_t)nrows);
  for (auto& d : done) d.store(0, std::memory_order_relaxed);
  const int dev = r.device;
  float *stT = plan.stage[b][0], *stS = plan.stage[b][1];
  float *dpT = plan.d_packed[b][0], *dpS = plan.d_packed[b][1];

  r.pool.start(plan.threads, [&, dev](int) {
    cudaSetDevice(dev);
    for (;;) {
      int pos, seg, slot;
      {
        std::lock_guard<std::mutex> l(sh.m);
        if (sh.next_seg >= plan.nseg) {
          if (sh.hi < sh.lo || sh.hi < sh.hi_min || sh.err.load() != (int)cudaSuccess) break;
          sh.cur = sh.hi--;
          sh.next_seg = 0;
          sh.packed++;
          if (plan.ring) {
            // the next slot of the ring, once the copy engine has emptied it.  Waiting here, with the lock held,
            // keeps every packer (they all want this row) no more than kStageRing rows ahead of the copies.
            const int64_t turn = plan.ring_next++;
            sh.cur_slot = (int)(turn % Resources::kStageRing);
            // the slot's previous row must have been queued (its last segment may still be with a thread that
            // lost its core) before the event below speaks for it
            while (plan.slot_queued[sh.cur_slot].load(std::memory_order_acquire) < turn / Resources::kStageRing &&
                   sh.err.load() == (int)cudaSuccess)
              std::this_thread::yield();
            const cudaError_t e = cudaEventSynchronize(r.slot_done[sh.cur_slot]);
            if (e != cudaSuccess) {
              sh.err.store((int)e);
              break;
            }
          }
        }
        pos = sh.cur;
        seg = sh.next_seg++;
        slot = sh.cur_slot;
      }
      int t, z;
      row_of(pos, t, z);
      const size_t src = ((size_t)t * (size_t)nz + (size_t)z) * (size_t)ncol;
      const size_t dst = (size_t)t * plan.nwet + plan.lvloff[z];
      const int64_t g0 = plan.ngrp * seg / plan.nseg, g1 = plan.ngrp * (seg + 1) / plan.nseg;
      float* rowT = plan.ring ? plan.ring_buf + (size_t)slot * 2 * (size_t)ncol : stT + dst;
      float* rowS = plan.ring ? rowT + (size_t)ncol : stS + dst;
      if (plan.ring)
        ml_pack_rows_cached(Th + src, Sh + src, plan.words + (size_t)z * plan.ngrp,
                            plan.before + (size_t)z * plan.ngrp, g0, g1, ncol, rowT, rowS);
      else
        ml_pack_rows(Th + src, Sh + src, plan.words + (size_t)z * plan.ngrp, plan.before + (size_t)z * plan.ngrp, g0,
                     g1, ncol, rowT, rowS);
      const int row = t * (int)nz + z;
      if (done[row].fetch_add(1, std::memory_order_acq_rel) + 1 == plan.nseg) {  // the row is whole: queue it
        const size_t nb = (size_t)(plan.lvloff[z + 1] - plan.lvloff[z]) * 4;
        flags[row] = 1;
        cudaError_t e = cudaSuccess;
        if (nb) {
          e = cudaMemcpyAsync(dpT + dst, rowT, nb, cudaMemcpyHostToDevice, r.copy);
          if (e == cudaSuccess) e = cudaMemcpyAsync(dpS + dst, rowS, nb, cudaMemcpyHostToDevice, r.copy);
          r.h2d_bytes += 2 * nb;
        }
        if (plan.ring && e == cudaSuccess) e = cudaEventRecord(r.slot_done[slot], r.copy);
        if (e != cudaSuccess) sh.err.store((int)e);
        if (plan.ring) {
          plan.last_slot.store(slot, std::memory_order_release);
          plan.slot_queued[slot].fetch_add(1, std::memory_order_release);
        }
      }
    }
  });

  // this thread: the fullest rows as they are, straight from the caller's buffer
  cudaError_t err = cudaSuccess;
  const int in_flight = plan.packed_first ? 1 : Resources::kRing;
  for (int issued = 0;; ++issued) {
    int pos;
    if (plan.packed_first) {
      // the copy stream belongs to the packed rows: wait until the one queued last has crossed (the stream is
      // in order, so then none is waiting) or until the rows have run out
      for (;;) {
        const int last = plan.last_slot.load(std::memory_order_acquire);
        if (last < 0) break;
        const cudaError_t q = cudaEventQuery(r.slot_done[last]);
        if (q == cudaSuccess) break;
        cudaGetLastError();  // cudaErrorNotReady is an answer, not a failure
        if (q != cudaErrorNotReady) {
          err = q;
          break;
        }
        {
          std::lock_guard<std::mutex> l(sh.m);
          if (sh.lo > sh.hi || sh.lo >= sh.lo_end) break;
        }
        std::this_thread::sleep_for(std::chrono::microseconds(20));
      }
      if (err != cudaSuccess) break;
    }
    {
      std::lock_guard<std::mutex> l(sh.m);
      if (sh.lo > sh.hi || sh.lo >= sh.lo_end) break;
      pos = sh.lo++;
    }
    int t, z;
    row_of(pos, t, z);
    const size_t off = ((size_t)t * (size_t)nz + (size_t)z) * (size_t)ncol;
    cudaEvent_t ev = r.ring[issued % in_flight];
    if (issued >= in_flight && (err = cudaEventSynchronize(ev)) != cudaSuccess) break;
    if ((err = cudaMemcpyAsync((float*)dT + off, Th + off, row_bytes, cudaMemcpyHostToDevice, r.copy)) != cudaSuccess) break;
    if ((err = cudaMemcpyAsync((float*)dS + off, Sh + off, row_bytes, cudaMemcpyHostToDevice, r.copy)) != cudaSuccess) break;
    r.h2d_bytes += 2 * row_bytes;
    if ((err = cudaEventRecord(ev, r.copy)) != cudaSuccess) break;
  }
  if (err != cudaSuccess) sh.err.store((int)err);  // stops the workers; wait for them before the locals they captured go
  r.pool.wait();
  if (sh.err.load() != (int)cudaSuccess) return cuda_fail((cudaError_t)sh.err.load(), "packed transfer");
  plan.rows_total += nrows;
  plan.rows_packed += sh.packed;

  if (sh.packed) {
    ML_CUDA(cudaMemcpyAsync(plan.d_flags[b], flags, (size_t)nrows, cudaMemcpyHostToDevice, r.copy));
    r.h2d_bytes += (size_t)nrows;
  }
  ML_CUDA(cudaEventRecord(r.copied[b], r.copy));
  ML_CUDA(cudaStreamWaitEvent(r.comp, r.copied[b], 0));
  if (sh.packed) {
    const int xblocks = (int)((plan.ngrp + 31) / 32);
    if (int rc = launch_unpack(r.comp, nrows, xblocks, dpT, dpS, (float*)dT, (float*)dS, plan.d_words, plan.d_before,
                               plan.d_lvloff, plan.d_flags[b], (int)nz, ncol, plan.ngrp, plan.nwet))
      return rc;
  }
  return ML_OK;
}

}  // namespace

extern "C" int ml_host_release(void) {
  resources().release();
  return ML_OK;
}

extern "C" int ml_host_set_packing(int mode, int threads) {
  if (mode < 0 || mode > 4) return ml::fail(ML_ERR_MODE, "packing mode %d is not 0 ... 4", mode);
  Resources& r = resources();
  r.pack_mode = mode;
  r.pack_threads = threads > 0 ? threads : 0;
  return ML_OK;
}

extern "C" double ml_host_last_packed_fraction(void) { return resources().last_packed_fraction; }

extern "C" uint64_t ml_host_last_h2d_bytes(void) { return resources().h2d_bytes.load(); }

extern "C" int ml_host_last_timings(double* ms4) {
  ML_REQUIRE_PTR(ms4);
  for (int i = 0; i < 4; ++i) ms4[i] = resources().last_ms[i];
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

  r.h2d_bytes = 0;
  r.last_packed_fraction = 0.0;
  const double t_call = now_ms();
  ML_CUDA(cudaMemcpyAsync(dZi, z_i, (size_t)(nz + 1) * sizeof(double), cudaMemcpyHostToDevice, r.copy));
  ML_CUDA(cudaMemcpyAsync(dDepth, deptho, (size_t)ncol * sizeof(double), cudaMemcpyHostToDevice, r.copy));
  ML_CUDA(cudaMemcpyAsync(dP, p_level, (size_t)nz * sizeof(double), cudaMemcpyHostToDevice, r.copy));
  // Pageable memory (plain numpy) on either end of a copy makes it a synchronous bounce through the driver that
  // holds this thread -- and with it the pipeline -- for the length of the copy.  Unless packing is off such
  // operands go through pinned buffers of our own, filled / emptied by the worker threads.
  const int nthreads = r.pack_threads > 0 ? std::min(r.pack_threads, 64) : default_threads();
  const void* v0_src = v0;
  if (r.pack_mode != 0 && is_pageable(v0)) {
    void* hv;
    if (r.halloc(11, &hv, lvl * es) == cudaSuccess) {
      parallel_copy(r, hv, v0, lvl * es, nthreads);
      v0_src = hv;
    } else {
      cudaGetLastError();
    }
  }
  ML_CUDA(cudaMemcpyAsync(dV, v0_src, lvl * es, cudaMemcpyHostToDevice, r.copy));
  r.h2d_bytes += (size_t)(2 * nz + 1 + ncol) * sizeof(double) + lvl * es;
  double* user_eta[3] = {eta, eta_thermo, eta_halo};
  double* host_eta[3] = {eta, eta_thermo, eta_halo};  // where the device->host copies land
  const size_t eta_bytes = (size_t)nt * ncol * sizeof(double);
  for (int v = 0; v < 3; ++v) {
    void* hb;
    if (user_eta[v] == nullptr || r.pack_mode == 0 || !is_pageable(user_eta[v])) continue;
    if (r.halloc(8 + v, &hb, eta_bytes) == cudaSuccess)
      host_eta[v] = (double*)hb;
    else
      cudaGetLastError();
  }
  // rho_ref is defined where the volume is missing too (reference.py:71), so a call that wants it back
  // moves every row as it is
  PackPlan plan;
  if (int rc = plan_packing(r, plan, dtype, v0, nz, ncol, spw, rho_ref_out == nullptr,
                            is_pageable(T) || is_pageable(S))) return rc;
  const double t_plan = now_ms();

  const int64_t nwin = (nt + spw - 1) / spw;
  for (int64_t w = 0; w < nwin; ++w) {
    const int b = (int)(w & 1);
    const int64_t t_first = w * spw;
    const int64_t nt_w = (t_first + spw <= nt) ? spw : (nt - t_first);
    const size_t off = (size_t)t_first * lvl * es;
    int rc = stage_window(r, plan, b, w, (const char*)T + off, (const char*)S + off, nt_w, es, nz, ncol, dT[b], dS[b]);
    if (rc) return rc;
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
    // the heights of this window go home while the next windows come in (PCIe carries both directions)
    const size_t eoff = (size_t)t_first * ncol, ebytes = (size_t)nt_w * ncol * sizeof(double);
    ML_CUDA(cudaStreamWaitEvent(r.back, r.freed[b], 0));
    ML_CUDA(cudaMemcpyAsync(host_eta[0] + eoff, (double*)dEta + eoff, ebytes, cudaMemcpyDeviceToHost, r.back));
    if (eta_thermo) ML_CUDA(cudaMemcpyAsync(host_eta[1] + eoff, (double*)dEtaT + eoff, ebytes, cudaMemcpyDeviceToHost, r.back));
    if (eta_halo) ML_CUDA(cudaMemcpyAsync(host_eta[2] + eoff, (double*)dEtaH + eoff, ebytes, cudaMemcpyDeviceToHost, r.back));
  }
  const double t_windows = now_ms();
  if (sums_out) ML_CUDA(cudaMemcpyAsync(sums_out, dSums, 2 * sizeof(double), cudaMemcpyDeviceToHost, r.comp));
  if (rho_ref_out) ML_CUDA(cudaMemcpyAsync(rho_ref_out, dRho, lvl * sizeof(double), cudaMemcpyDeviceToHost, r.comp));
  ML_CUDA(cudaStreamSynchronize(r.comp));
  ML_CUDA(cudaStreamSynchronize(r.back));
  ML_CUDA(cudaStreamSynchronize(r.copy));
  for (int v = 0; v < 3; ++v)
    if (user_eta[v] != nullptr && host_eta[v] != user_eta[v]) parallel_copy(r, user_eta[v], host_eta[v], eta_bytes, nthreads);
  r.last_packed_fraction = plan.rows_total ? (double)plan.rows_packed / (double)plan.rows_total : 0.0;
  const double t_end = now_ms();
  r.last_ms[0] = t_plan - t_call;
  r.last_ms[1] = t_windows - t_plan;
  r.last_ms[2] = t_end - t_windows;
  r.last_ms[3] = t_end - t_call;
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
  r.h2d_bytes = 0;
  r.last_packed_fraction = 0.0;
  ML_CUDA(cudaMemcpyAsync(dP, p_level, (size_t)nz * sizeof(double), cudaMemcpyHostToDevice, r.copy));
  ML_CUDA(cudaMemcpyAsync(dV, v_ref, lvl * es, cudaMemcpyHostToDevice, r.copy));
  r.h2d_bytes += (size_t)nz * sizeof(double) + lvl * es;
  // rho * volcello is skipped where the reference volume is missing (derived.py:435-438)
  PackPlan plan;
  if (int rc = plan_packing(r, plan, dtype, v_ref, nz, ncol, spw, true, is_pageable(T) || is_pageable(S))) return rc;

  const int64_t nwin = (nt + spw - 1) / spw;
  for (int64_t w = 0; w < nwin; ++w) {
    const int b = (int)(w & 1);
    const int64_t t_first = w * spw;
    const int64_t nt_w = (t_first + spw <= nt) ? spw : (nt - t_first);
    const size_t off = (size_t)t_first * lvl * es;
    int rc = stage_window(r, plan, b, w, (const char*)T + off, (const char*)S + off, nt_w, es, nz, ncol, dT[b], dS[b]);
    if (rc) return rc;
    rc = ml_steric_global(eos, dtype, dT[b], dS[b], 0, 0, dV, dtype, (const double*)dP, nt_w, nz, ncol,
                              (double*)dM + t_first, dWs, ws_bytes, r.comp);
    if (rc) return rc;
    ML_CUDA(cudaEventRecord(r.freed[b], r.comp));
  }
  ML_CUDA(cudaMemcpyAsync(masso, dM, (size_t)nt * sizeof(double), cudaMemcpyDeviceToHost, r.comp));
  ML_CUDA(cudaStreamSynchronize(r.comp));
  ML_CUDA(cudaStreamSynchronize(r.copy));
  r.last_packed_fraction = plan.rows_total ? (double)plan.rows_packed / (double)plan.rows_total : 0.0;
  return ML_OK;
}

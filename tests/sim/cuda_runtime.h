// cuda_runtime.h -- a SIMULATED CUDA runtime for tests/sim/hostpath_sim.cpp (test infrastructure, never shipped).
//
// Just enough of the runtime API for momlevel_b200/csrc/ml_hostpath.cu to compile with g++ and run on a box
// without a GPU, faithful in the one respect that matters to that file: work queued on a stream runs LATER, on the
// stream's own thread, in order; events complete when the stream reaches them; a copy reads its source when it
// runs, not when it is queued.  So a staging buffer recycled too early, a missing event wait or an unsynchronised
// flag shows up as a wrong result or as a ThreadSanitizer report.  "Device" memory is malloc'd host memory.
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <thread>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorNotReady = 600 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes {
  cudaMemoryType type;
};

namespace simcuda {

// nanoseconds a copy sleeps per KiB moved (0 = none); the harness varies it so that the copies are sometimes
// slower and sometimes faster than the packers
inline int& copy_ns_per_kib() {
  static int v = 0;
  return v;
}

struct Stream {
  std::mutex m;
  std::condition_variable cv;
  std::deque<std::function<void()>> q;
  bool stop = false, busy = false;
  std::thread th;
  Stream() : th([this] { run(); }) {}
  ~Stream() {
    {
      std::lock_guard<std::mutex> l(m);
      stop = true;
    }
    cv.notify_all();
    th.join();
  }
  void run() {
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> l(m);
        cv.wait(l, [this] { return stop || !q.empty(); });
        if (q.empty()) return;
        f = std::move(q.front());
        q.pop_front();
        busy = true;
      }
      f();
      {
        std::lock_guard<std::mutex> l(m);
        busy = false;
      }
      cv.notify_all();
    }
  }
  void enqueue(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> l(m);
      q.push_back(std::move(f));
    }
    cv.notify_all();
  }
  void drain() {
    std::unique_lock<std::mutex> l(m);
    cv.wait(l, [this] { return q.empty() && !busy; });
  }
};

struct Event {
  std::mutex m;
  std::condition_variable cv;
  uint64_t recorded = 0, completed = 0;
  std::chrono::steady_clock::time_point when;  // of the last completion (cudaEventElapsedTime)
  void wait_for(uint64_t n) {
    std::unique_lock<std::mutex> l(m);
    cv.wait(l, [&] { return completed >= n; });
  }
};

// allocations the "driver" knows: pinned host memory and device memory
struct Registry {
  std::mutex m;
  std::map<uintptr_t, std::pair<size_t, cudaMemoryType>> blocks;
  void add(void* p, size_t n, cudaMemoryType t) {
    std::lock_guard<std::mutex> l(m);
    blocks[(uintptr_t)p] = {n, t};
  }
  void remove(void* p) {
    std::lock_guard<std::mutex> l(m);
    blocks.erase((uintptr_t)p);
  }
  cudaMemoryType find(const void* p) {
    std::lock_guard<std::mutex> l(m);
    auto it = blocks.upper_bound((uintptr_t)p);
    if (it == blocks.begin()) return cudaMemoryTypeUnregistered;
    --it;
    return ((uintptr_t)p < it->first + it->second.first) ? it->second.second : cudaMemoryTypeUnregistered;
  }
};
inline Registry& registry() {
  static Registry r;
  return r;
}
inline Stream& legacy_stream() {
  static Stream s;
  return s;
}

}  // namespace simcuda

typedef simcuda::Stream* cudaStream_t;
typedef simcuda::Event* cudaEvent_t;

inline cudaError_t cudaGetDevice(int* d) {
  *d = 0;
  return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "simulated error"; }
inline const char* cudaGetErrorName(cudaError_t) { return "simError"; }

inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  *s = new simcuda::Stream();
  return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t s) {
  delete s;
  return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t s) {
  (s ? s : &simcuda::legacy_stream())->drain();
  return cudaSuccess;
}
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
  *e = new simcuda::Event();
  return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
  delete e;
  return cudaSuccess;
}
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
  uint64_t n;
  {
    std::lock_guard<std::mutex> l(e->m);
    n = ++e->recorded;
  }
  (s ? s : &simcuda::legacy_stream())->enqueue([e, n] {
    {
      std::lock_guard<std::mutex> l(e->m);
      if (e->completed < n) e->completed = n;
      e->when = std::chrono::steady_clock::now();
    }
    e->cv.notify_all();
  });
  return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t e) {
  uint64_t n;
  {
    std::lock_guard<std::mutex> l(e->m);
    n = e->recorded;
  }
  e->wait_for(n);
  return cudaSuccess;
}
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  std::chrono::steady_clock::time_point ta, tb;
  {
    std::lock_guard<std::mutex> l(a->m);
    if (a->completed < a->recorded || a->recorded == 0) return cudaErrorNotReady;
    ta = a->when;
  }
  {
    std::lock_guard<std::mutex> l(b->m);
    if (b->completed < b->recorded || b->recorded == 0) return cudaErrorNotReady;
    tb = b->when;
  }
  *ms = std::chrono::duration<float, std::milli>(tb - ta).count();
  return cudaSuccess;
}
inline cudaError_t cudaEventQuery(cudaEvent_t e) {
  std::lock_guard<std::mutex> l(e->m);
  return e->completed >= e->recorded ? cudaSuccess : cudaErrorNotReady;
}
inline cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
  uint64_t n;
  {
    std::lock_guard<std::mutex> l(e->m);
    n = e->recorded;
  }
  (s ? s : &simcuda::legacy_stream())->enqueue([e, n] { e->wait_for(n); });
  return cudaSuccess;
}

inline cudaError_t cudaMalloc(void** p, size_t n) {
  *p = malloc(n ? n : 1);
  if (!*p) return cudaErrorMemoryAllocation;
  memset(*p, 0xA5, n);  // not zero, not NaN: whatever a kernel does not write is noticed
  simcuda::registry().add(*p, n ? n : 1, cudaMemoryTypeDevice);
  return cudaSuccess;
}
inline cudaError_t cudaFree(void* p) {
  simcuda::registry().remove(p);
  free(p);
  return cudaSuccess;
}
inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) {
  if (posix_memalign(p, 4096, n ? n : 1) != 0) return cudaErrorMemoryAllocation;
  simcuda::registry().add(*p, n ? n : 1, cudaMemoryTypeHost);
  return cudaSuccess;
}
inline cudaError_t cudaFreeHost(void* p) {
  simcuda::registry().remove(p);
  free(p);
  return cudaSuccess;
}
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  a->type = simcuda::registry().find(p);
  return cudaSuccess;
}

// A copy that involves pageable memory holds the caller until it is done, like the real one; otherwise it is
// queued and reads its source when the stream gets to it.
inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind kind, cudaStream_t s) {
  simcuda::Stream* st = s ? s : &simcuda::legacy_stream();
  const bool pageable = (kind == cudaMemcpyHostToDevice && simcuda::registry().find(src) == cudaMemoryTypeUnregistered) ||
                        (kind == cudaMemcpyDeviceToHost && simcuda::registry().find(dst) == cudaMemoryTypeUnregistered);
  const int ns = simcuda::copy_ns_per_kib();
  st->enqueue([=] {
    memcpy(dst, src, n);
    if (ns) std::this_thread::sleep_for(std::chrono::nanoseconds((uint64_t)ns * (n >> 10)));
  });
  if (pageable) st->drain();
  return cudaSuccess;
}

// ml_hosttune.h -- how the host path decides how many threads pack (host-only C++; included by ml_hostpath.cu inside its
// anonymous namespace, and with it by tests/sim/hostpath_sim.cpp).
//
//   local_ranks / cores_per_rank / default_threads / max_threads   this rank's share of the host's cores
//   SharedTuning    the table of window timings the ranks of one host share (POSIX shared memory), so that they choose together
//   PackTuner       the choice itself: trial windows, timed intervals, fastest per step for the slowest rank
#pragma once

// Ranks that share this host's cores and memory system: one process per GPU under torchrun exports
// LOCAL_WORLD_SIZE; a lone process counts as one.
int local_ranks() {
  const char* v = getenv("LOCAL_WORLD_SIZE");
  if (v == nullptr || *v == 0) return 1;
  const long n = strtol(v, nullptr, 10);
  return (n >= 1 && n <= 1024) ? (int)n : 1;
}

// Cores of this rank's share of the host: the calling thread's affinity mask divided by the ranks on the host.
int cores_per_rank() {
  cpu_set_t set;
  int n = 0;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
  if (n <= 0) n = (int)std::thread::hardware_concurrency();
  return std::max(1, n / local_ranks());
}

int default_threads() {
  // half the cores: the packers and the DMA engine share the host's memory bandwidth, and past that point
  // every extra thread slows the copies by as much as it saves (tools/e2e_sweep.py: 4 / 6 / 8 / 10 / 12 / 15
  // threads on a 16-core host gave 156 / 148 / 145 / 148 / 152 / 157 ms for an OM4p25 year).  That half is
  // shared by the ranks of the host: with every rank taking half the affinity mask, four ranks oversubscribed
  // the cores and eight lost to plain copies (SCALE_r01: packed 366 ms against 409 ms dense at four ranks with
  // cores / (2 ranks) threads each, a tie at eight).
  return std::max(1, std::min(cores_per_rank() / 2, 64));
}

// The most packers the tuner may try: the rank's share less one core for the calling thread, which queues the plain
// rows and should not have to wait for a core behind its own packers.
int max_threads() { return std::max(1, std::min(cores_per_rank() - 1, 64)); }

// The ranks of one host decide TOGETHER how to move their rows.  They share the host's memory system, so a rank that
// packs slows the plain copies of its neighbours: four ranks that each took what was fastest for themselves ended at
// 227 / 273 / 304 / 253 ms -- two packing, two not -- where all four copying plainly took 230 ms each, and the job
// runs at the pace of its slowest rank.  Every rank therefore publishes its table of timings in a small POSIX
// shared-memory segment named after the job (MASTER_ADDR / MASTER_PORT / TORCHELASTIC_RUN_ID and the user id), and
// every rank picks from the element-wise MAXIMUM over the ranks' tables: the same numbers, hence the same choice.
// Without those variables, or if the segment cannot be had, a rank decides from its own table.
struct SharedTuning {
  static constexpr int kMaxRanks = 64;
  struct Slot {
    int64_t nz, ncol;
    int32_t threads[4];
    double ms[4];
    int32_t seen, pad_;  // ml_host_tuner_share_selftest only
  };
  Slot* slots = nullptr;  // [kMaxRanks], zero-filled by ftruncate
  int me = -1, ranks = 1;
  char name[64] = {0};
  explicit SharedTuning(bool from_env = true) {
#ifndef ML_HOSTPATH_TEST_HOOKS
    if (!from_env) return;
    ranks = local_ranks();
    const char* lr = getenv("LOCAL_RANK");
    const char *addr = getenv("MASTER_ADDR"), *port = getenv("MASTER_PORT"), *run = getenv("TORCHELASTIC_RUN_ID");
    if (ranks <= 1 || ranks > kMaxRanks || lr == nullptr || (port == nullptr && run == nullptr)) return;
    const long r = strtol(lr, nullptr, 10);
    if (r < 0 || r >= ranks) return;
    attach(addr, port, run, (int)r);
#endif
  }
  void attach(const char* addr, const char* port, const char* run, int rank) {
    uint64_t h = 1469598103934665603ull;  // FNV-1a over the job's coordinates
    for (const char* part : {addr, port, run})
      for (const char* c = part ? part : ""; ; ++c) {
        h = (h ^ (uint64_t)(unsigned char)*c) * 1099511628211ull;
        if (*c == 0) break;
      }
    snprintf(name, sizeof(name), "/momlevel_b200_tuner_%u_%016llx", (unsigned)getuid(), (unsigned long long)h);
    const int fd = shm_open(name, O_CREAT | O_RDWR, 0600);
    if (fd < 0) return;
    const size_t bytes = sizeof(Slot) * kMaxRanks;
    void* m = ftruncate(fd, (off_t)bytes) == 0 ? mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0) : MAP_FAILED;
    close(fd);
    if (m == MAP_FAILED) return;
    slots = static_cast<Slot*>(m);
    me = rank;
    clear();
  }
  ~SharedTuning() {
    if (slots == nullptr) return;
    clear();
    munmap(slots, sizeof(Slot) * kMaxRanks);
    if (me == 0) shm_unlink(name);  // the name goes; ranks that still have it mapped keep the memory
  }
  void clear() {
    if (slots == nullptr) return;
    for (int i = 0; i < 4; ++i) __atomic_store_n(reinterpret_cast<int64_t*>(&slots[me].ms[i]), (int64_t)0xbff0000000000000ll, __ATOMIC_RELAXED);  // -1.0
    __atomic_store_n(&slots[me].nz, (int64_t)0, __ATOMIC_RELAXED);
  }
  // this rank's table, for the others to see
  void publish(int64_t nz, int64_t ncol, const int* threads, const double* ms) {
    if (slots == nullptr) return;
    Slot& s = slots[me];
    for (int i = 0; i < 4; ++i) {
      __atomic_store_n(&s.threads[i], (int32_t)threads[i], __ATOMIC_RELAXED);
      int64_t bits;
      memcpy(&bits, &ms[i], 8);
      __atomic_store_n(reinterpret_cast<int64_t*>(&s.ms[i]), bits, __ATOMIC_RELAXED);
    }
    __atomic_store_n(&s.ncol, ncol, __ATOMIC_RELAXED);
    __atomic_store_n(&s.nz, nz, __ATOMIC_RELEASE);
  }
  // element-wise maximum over the ranks that work on the same grid with the same choices (own table included);
  // a choice some rank has no timing for yet stays unknown (-1)
  void combine(int64_t nz, int64_t ncol, const int* threads, const double* mine, double* out) const {
    for (int i = 0; i < 4; ++i) out[i] = mine[i];
    if (slots == nullptr) return;
    for (int r = 0; r < ranks; ++r) {
      if (r == me) continue;
      const Slot& s = slots[r];
      if (__atomic_load_n(&s.nz, __ATOMIC_ACQUIRE) != nz || __atomic_load_n(&s.ncol, __ATOMIC_RELAXED) != ncol) continue;
      bool same = true;
      for (int i = 0; i < 4; ++i) same = same && __atomic_load_n(&s.threads[i], __ATOMIC_RELAXED) == threads[i];
      double theirs[4];
      bool any = false;
      for (int i = 0; i < 4; ++i) {
        const int64_t bits = __atomic_load_n(reinterpret_cast<const int64_t*>(&s.ms[i]), __ATOMIC_RELAXED);
        memcpy(&theirs[i], &bits, 8);
        any = any || theirs[i] >= 0.0;
      }
      if (!same || !any) continue;  // a rank that is not timing its windows (pageable source, fixed thread count) has no say
      for (int i = 0; i < 4; ++i) {
        if (theirs[i] < 0.0 || out[i] < 0.0) out[i] = -1.0;
        else out[i] = std::max(out[i], theirs[i]);
      }
    }
  }
};

SharedTuning& shared_tuning() {
  static SharedTuning t;  // one per process, attached on first use
  return t;
}

// How many threads pack, when the caller leaves it to the library (ml_host_set_packing(1, 0)).  Packing trades host
// memory bandwidth for PCIe bytes, and which of the two runs out first depends on the machine and on who else is
// using it (one rank of four packed SLOWER than plain copies on a box where four ranks still get the full PCIe rate
// each; one or two ranks, or eight, did not).  So the library measures: every window's interval on the copy stream
// (idle gap in front of it included) is timed with events, the first windows try {default, none, twice, half} the
// default thread count for two windows each, and the rest run with whatever was fastest per step -- for the slowest
// rank of the host (SharedTuning above); the table lives with the thread's Resources, so later calls start from it,
// and it is tried afresh every kRetry windows.
struct PackTuner {
  static constexpr int kChoices = 4;
  static constexpr int kRetry = 256;
  int64_t nz = 0, ncol = 0;        // the table belongs to this grid
  int threads[kChoices] = {0, 0, 0, 0};
  double ms_per_step[kChoices] = {-1.0, -1.0, -1.0, -1.0};
  int64_t windows = 0;
  void reset(int64_t nz_, int64_t ncol_, int dflt, int cap) {
    nz = nz_;
    ncol = ncol_;
    threads[0] = dflt;
    threads[1] = 0;
    threads[2] = std::min(std::max(2 * dflt, 1), std::max(cap, 1));
    threads[3] = std::max(dflt / 2, 1);
    for (double& m : ms_per_step) m = -1.0;
    windows = 0;
    shared_tuning().publish(nz, ncol, threads, ms_per_step);
  }
  int choose() {  // which entry of threads[] the next window runs with
    const int64_t w = windows++ % kRetry;
    if (w < 2 * kChoices) return (int)(w / 2);
    return best();
  }
  // Packing has costs the interval of a window on the copy stream does not show (the presence index at the start of
  // every call, cores and memory bandwidth the caller could use), so it has to beat plain copies by kMargin to be chosen.
  static constexpr double kMargin = 0.06;
  int best() const {
    double ms[kChoices];
    shared_tuning().combine(nz, ncol, threads, ms_per_step, ms);
    auto score = [&](int i) { return threads[i] == 0 ? ms[i] * (1.0 - kMargin) : ms[i]; };
    int b = -1;
    for (int i = 0; i < kChoices; ++i)
      if (ms[i] >= 0.0 && (b < 0 || score(i) < score(b))) b = i;
    return b < 0 ? 0 : b;
  }
  // Past the trial windows with plain copies in front: a call that starts now does not even build the presence
  // index.  Its windows are counted by idle_window(), which stops at the next multiple of kRetry, so that the call
  // after that one tries the choices again.
  bool settled_on_none() const {
    const int b = best();
    return windows % kRetry >= 2 * kChoices && threads[b] == 0 && ms_per_step[b] >= 0.0;
  }
  void idle_window() {
    if (windows % kRetry != 0) ++windows;
  }
  void report(int choice, double ms, int64_t steps) {
    if (choice < 0 || steps <= 0 || ms <= 0.0) return;
    const double v = ms / (double)steps;
    ms_per_step[choice] = ms_per_step[choice] < 0.0 ? v : 0.5 * (ms_per_step[choice] + v);
    shared_tuning().publish(nz, ncol, threads, ms_per_step);
  }
};

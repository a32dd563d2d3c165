// Unit checks of csrc/ml_hosttune.h (no CUDA, no threads): the order of the trial windows, the margin plain copies
// enjoy, the settled-on-none shortcut and its retry point, and the cores a rank may use.
//   g++ -std=c++17 -DML_HOSTPATH_TEST_HOOKS -I tests/sim -x c++ tests/sim/tuner_unit.cpp -o tuner_unit && ./tuner_unit
#include <fcntl.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <thread>

namespace {
#include "../../momlevel_b200/csrc/ml_hosttune.h"
}

static int bad = 0;
#define CHECK(cond)                                              \
  do {                                                           \
    if (!(cond)) {                                               \
      printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond);   \
      ++bad;                                                     \
    }                                                            \
  } while (0)

int main() {
  // a rank's share of the cores
  setenv("LOCAL_WORLD_SIZE", "4", 1);
  const int share = cores_per_rank();
  CHECK(local_ranks() == 4 && share >= 1);
  CHECK(default_threads() == std::max(1, std::min(share / 2, 64)));
  CHECK(max_threads() == std::max(1, std::min(share - 1, 64)));
  setenv("LOCAL_WORLD_SIZE", "not a number", 1);
  CHECK(local_ranks() == 1);
  unsetenv("LOCAL_WORLD_SIZE");

  PackTuner t;
  t.reset(75, 1555200, 8, 15);
  CHECK(t.threads[0] == 8 && t.threads[1] == 0 && t.threads[2] == 15 && t.threads[3] == 4);
  // two trial windows per choice, in order
  for (int w = 0; w < 2 * PackTuner::kChoices; ++w) CHECK(t.choose() == w / 2);
  CHECK(!t.settled_on_none());  // nothing timed yet
  // packing 10 % faster than plain copies per step: chosen
  t.report(0, 10.0 * 9.0, 10);
  t.report(1, 10.0 * 10.0, 10);
  t.report(2, 10.0 * 9.5, 10);
  t.report(3, 10.0 * 9.8, 10);
  CHECK(t.best() == 0 && t.choose() == 0 && !t.settled_on_none());
  // only 4 % faster: plain copies keep the job (the margin is 6 %)
  t.ms_per_step[0] = 9.6;
  CHECK(t.best() == 1 && t.settled_on_none());
  // the timings are running means of what is reported
  t.report(0, 8.0, 1);
  CHECK(t.ms_per_step[0] == 0.5 * (9.6 + 8.0));
  t.report(0, -1.0, 1);  // an unreadable span changes nothing
  t.report(-1, 5.0, 1);
  CHECK(t.ms_per_step[0] == 0.5 * (9.6 + 8.0));
  // idle windows stop at the retry point; the window after it starts the trials again
  t.ms_per_step[0] = 9.6;
  t.windows = PackTuner::kRetry - 2;
  t.idle_window();
  CHECK(t.windows == PackTuner::kRetry - 1 && t.settled_on_none());
  t.idle_window();
  t.idle_window();
  CHECK(t.windows == PackTuner::kRetry && !t.settled_on_none());
  CHECK(t.choose() == 0 && t.choose() == 0 && t.choose() == 1);
  // a small default: every choice stays a legal thread count
  t.reset(10, 100, 1, 1);
  CHECK(t.threads[0] == 1 && t.threads[1] == 0 && t.threads[2] == 1 && t.threads[3] == 1);
  printf("tuner_unit: %d failed\n", bad);
  return bad != 0;
}

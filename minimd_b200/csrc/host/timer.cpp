#include "timer.h"

#include <time.h>

static double now() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

Timer::Timer() : t0_(0) { clear(); }
void Timer::clear() {
  for (int i = 0; i < TIME_N; i++) array[i] = 0.0;
}
void Timer::barrier_start(int which) {
  (void)which;
  t0_ = now();
}
void Timer::barrier_stop(int which) { array[which] += now() - t0_; }
double Timer::elapsed_since_start() const { return now() - t0_; }

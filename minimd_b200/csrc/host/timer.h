// Timer buckets of the reference (ref/timer.h:35-40).  TOTAL is wall clock around Integrate::run
// (as in the reference); COMM / FORCE / NEIGH are device times from CUDA events inside mmd_run.
#pragma once

enum { TIME_TOTAL = 0, TIME_COMM, TIME_FORCE, TIME_NEIGH, TIME_TEST, TIME_N };

class Timer {
 public:
  Timer();
  void clear();
  void barrier_start(int which);
  void barrier_stop(int which);
  double elapsed_since_start() const;  // seconds since the last barrier_start
  double array[TIME_N];

 private:
  double t0_;
};

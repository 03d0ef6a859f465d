// Integrate: velocity-Verlet halves and the time loop (ref/integrate.h, ref/integrate.cpp).
#pragma once
#include "atom.h"
#include "comm.h"
#include "force.h"
#include "neighbor.h"
#include "thermo.h"
#include "timer.h"
#include "types.h"

class Integrate {
 public:
  MMD_float dt;
  MMD_float dtforce;
  MMD_int ntimes;
  MMD_int nlocal, nmax;
  MMD_float mass;
  int sort_every;

  // stepwise = 1: the loop body of ref/integrate.cpp:88-205 call by call through the classes
  // (each one C-ABI call).  0 (default): the same sequence fused inside mmd_run (launch-only host loop).
  int stepwise;
  double device_ms;  // CUDA-event time of the last run()
  int steps_done;    // steps already integrated (run() may be called repeatedly)

  Integrate();
  ~Integrate();
  void setup();
  void initialIntegrate();  // ref/integrate.cpp:46-57
  void finalIntegrate();    // ref/integrate.cpp:59-68
  void run(Atom& atom, Force* force, Neighbor& neighbor, Comm& comm, Thermo& thermo, Timer& timer, int nsteps = -1);

 private:
  mmd_ctx* ctx_;
  bool mass_folded_;
};

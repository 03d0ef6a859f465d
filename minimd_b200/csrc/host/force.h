// Force: abstract pair-force interface (ref/force.h:40-70).  Same public members; compute()
// launches the device kernels through the C ABI and, when evflag is set, brings eng_vdwl and
// virial back (this rank's partial sums, exactly what Thermo expects).
#pragma once
#include <vector>

#include "atom.h"
#include "comm.h"
#include "neighbor.h"
#include "types.h"

class Force {
 public:
  MMD_float cutforce;
  MMD_float* cutforcesq;
  MMD_float eng_vdwl;
  MMD_float mass;
  MMD_int evflag;
  MMD_float virial;
  int ntypes;

  explicit Force(int ntypes_) : cutforce(0), eng_vdwl(0), mass(1), evflag(0), virial(0), ntypes(ntypes_), use_sse(0),
                                use_oldcompute(0), reneigh(1), style(FORCELJ), ctx(nullptr),
                                cutsq_((size_t)ntypes_ * ntypes_, 0), eps_((size_t)ntypes_ * ntypes_, 1),
                                s6_((size_t)ntypes_ * ntypes_, 1), sig_((size_t)ntypes_ * ntypes_, 1) {
    cutforcesq = cutsq_.data();
    epsilon = eps_.data();
    sigma6 = s6_.data();
    sigma = sig_.data();
  }
  virtual ~Force() {}
  virtual int setup(Atom& atom) = 0;  // pushes parameters/tables to the device; 0 = ok
  virtual void finalise() {}
  virtual void compute(Atom& atom, Neighbor& neighbor, Comm& comm, int me) = 0;

  int use_sse;
  int use_oldcompute;
  MMD_int reneigh;
  MMD_float *epsilon, *sigma6, *sigma;  // LJ parameters per type pair
  ForceStyle style;

 protected:
  mmd_ctx* ctx;
  std::vector<MMD_float> cutsq_, eps_, s6_, sig_;
};

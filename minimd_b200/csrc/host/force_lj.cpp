#include "force_lj.h"

#include <stdio.h>
#include <stdlib.h>

ForceLJ::ForceLJ(int ntypes_) : Force(ntypes_) {
  cutforce = 0.0;
  use_oldcompute = 0;
  reneigh = 1;
  style = FORCELJ;
}
ForceLJ::~ForceLJ() {}

int ForceLJ::setup(Atom& atom) {
  for (int i = 0; i < ntypes * ntypes; i++) cutforcesq[i] = cutforce * cutforce;  // ref/force_lj.cpp:65-69
  ctx = atom.ctx;
  if (ctx && mmd_force_lj_setup(ctx, cutforcesq, sigma6, epsilon)) {
    fprintf(stderr, "ERROR: mmd_force_lj_setup: %s\n", mmd_last_error());
    return 1;
  }
  return 0;
}

// ref/force_lj.cpp:72-113: the reference picks one of its loop variants here; on the device the
// list style (half / half+ghost_newton / full) and evflag select the kernel instantiation.
// --half_neigh -1 (the legacy serial loop, "not OpenMP safe") is served by the half-list kernel.
void ForceLJ::compute(Atom& atom, Neighbor& neighbor, Comm&, int) {
  eng_vdwl = 0;
  virial = 0;
  const int half = neighbor.halfneigh != 0;
  if (mmd_force_lj_compute(atom.ctx, half, half ? neighbor.ghost_newton : 0, evflag, &eng_vdwl, &virial)) {
    fprintf(stderr, "ERROR: mmd_force_lj_compute: %s\n", mmd_last_error());
    exit(1);
  }
}

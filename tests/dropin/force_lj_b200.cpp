// TEST INFRASTRUCTURE (tests/test_gpu_dropin.py): the per-call seam of INTEGRATION.md section B.
//
// A replacement for ONE translation unit of the reference, ref/force_lj.cpp, compiled against the reference's own
// headers and linked with the reference's other, unmodified objects by oracle/build_dropin.sh.  The reference's time
// loop, Comm and Neighbor stay on the host; ForceLJ::compute (ref/force_lj.cpp:72-113) becomes: copy the positions the
// host holds (local atoms + ghosts) to the device, hand over the host-built neighbor list when it has changed, call
// mmd_force_lj_compute, copy the forces back -- the shape of mpi-spec's raw-pointer hook compute_lj
// (mpi-spec/force_lj_custom.cpp:17-30).  It pays the copies every step and exists to prove the boundary, not to be fast.
#include <cstdio>
#include <cstdlib>

#include "mpi.h"
#include "force_lj.h"
#include "minimd_b200.h"

#define MMD_CHECK(call)                                                             \
  do {                                                                              \
    if ((call) != MMD_OK) {                                                         \
      fprintf(stderr, "minimd_b200: %s failed: %s\n", #call, mmd_last_error());     \
      exit(1);                                                                      \
    }                                                                               \
  } while (0)

ForceLJ::ForceLJ(int ntypes_) {
  cutforce = 0.0;
  use_oldcompute = 0;
  reneigh = 1;
  style = FORCELJ;
  ntypes = ntypes_;
  const int nn = ntypes * ntypes;
  cutforcesq = new MMD_float[nn];
  epsilon = new MMD_float[nn];
  sigma6 = new MMD_float[nn];
  sigma = new MMD_float[nn];
  for (int i = 0; i < nn; i++) {
    cutforcesq[i] = 0.0;
    epsilon[i] = sigma6[i] = sigma[i] = 1.0;
  }
}
ForceLJ::~ForceLJ() {}

void ForceLJ::setup() {
  for (int i = 0; i < ntypes * ntypes; i++) cutforcesq[i] = cutforce * cutforce;
}

void ForceLJ::compute(Atom& atom, Neighbor& neighbor, Comm&, int) {
  // the reference calls this from every OpenMP thread of its parallel regions (ref/ljs.cpp:451-456,
  // ref/integrate.cpp:84-206); one thread drives the device
  #pragma omp master
  {
    static mmd_ctx* ctx = NULL;
    static int list_calls = -1, list_atoms = -1;
    if (!ctx) {
      int device = 0;
      if (const char* lr = getenv("LOCAL_RANK")) device = atoi(lr);
      MMD_CHECK(mmd_ctx_create(device, (int)sizeof(MMD_float), ntypes, NULL, &ctx));
      MMD_CHECK(mmd_force_lj_setup(ctx, cutforcesq, sigma6, epsilon));
    }
    const int nall = atom.nlocal + atom.nghost;
    if (neighbor.ncalls != list_calls || nall != list_atoms) {
      // new ghosts / new list: positions of all atoms the host holds, the last nghost of them ghosts, and the host's rows
      MMD_CHECK(mmd_atom_upload(ctx, atom.x, atom.v, atom.type, nall, PAD));
      MMD_CHECK(mmd_atom_split(ctx, atom.nlocal));
      MMD_CHECK(mmd_neigh_upload(ctx, neighbor.numneigh, neighbor.neighbors, atom.nlocal, neighbor.maxneighs));
      list_calls = neighbor.ncalls;
      list_atoms = nall;
    } else {
      MMD_CHECK(mmd_atom_update(ctx, atom.x, NULL, 0, nall, PAD));
    }
    MMD_CHECK(mmd_force_lj_compute(ctx, neighbor.halfneigh, neighbor.ghost_newton, evflag, &eng_vdwl, &virial));
    if (!evflag) { eng_vdwl = 0; virial = 0; }
    // half lists leave contributions on the ghosts, which the host's Comm::reverse_communicate sends home
    MMD_CHECK(mmd_atom_download(ctx, NULL, NULL, atom.f, NULL, 0, neighbor.halfneigh ? nall : atom.nlocal, PAD));
  }
  #pragma omp barrier
}

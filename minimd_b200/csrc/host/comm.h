// Comm: 3-D spatial decomposition (ref/comm.cpp:60-272) computed on the host; the halo swaps,
// migration and ghost-list rebuild run on the device, remote swaps as NCCL point-to-point
// (one rank per GPU) -- see mmd_comm_* in include/minimd_b200.h.
#pragma once
#include "atom.h"
#include "types.h"
#include "world.h"

class Comm {
 public:
  Comm();
  ~Comm();
  int setup(MMD_float cutneigh, Atom& atom);  // needs `world` set
  void communicate(Atom& atom);               // ref/comm.cpp:276-317
  void reverse_communicate(Atom& atom);       // ref/comm.cpp:321-355
  void exchange(Atom& atom);                  // ref/comm.cpp:364-597
  void borders(Atom& atom);                   // ref/comm.cpp:700-883

  World* world;
  int me;
  int nswap;
  int procneigh[3][2];  // my 6 proc neighbors
  int procgrid[3];      // # of procs in each dim
  int myloc[3];         // my coordinates in the proc grid
  int need[3];          // how many procs away needed in each dim
  mmd_swap_table table;  // sendproc/recvproc, pbc flags, slab bounds per swap

  int check_safeexchange;
  int do_safeexchange;

  // per-swap counts after borders (Comm::sendnum/recvnum/firstrecv)
  int sendnum[MMD_MAX_SWAPS], recvnum[MMD_MAX_SWAPS], firstrecv[MMD_MAX_SWAPS];
  void refresh_counts(Atom& atom);
};

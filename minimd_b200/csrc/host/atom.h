// Atom: owner of the per-atom state (ref/atom.h:47-106).  The public members keep the
// reference's names and meaning; the difference is WHERE the truth lives: once upload() has
// run, positions/velocities/forces/types live in the device context (`ctx`, include/minimd_b200.h)
// and the host arrays x/v/f/type are mirrors refreshed by download().  Every hot-path method
// (pbc, sort, the pack/unpack family used by Comm) is one C-ABI call.
#pragma once
#include <vector>

#include "minimd_b200.h"
#include "types.h"

class Neighbor;

struct Box {
  MMD_float xprd, yprd, zprd;
  MMD_float xlo, xhi;
  MMD_float ylo, yhi;
  MMD_float zlo, zhi;
};

class Atom {
 public:
  int natoms;          // global number of atoms
  int nlocal, nghost;  // owned / ghost atoms of this rank
  int nmax;            // capacity of the host mirrors

  MMD_float* x;  // AoS, stride PAD (ref/types.h:77-81)
  MMD_float* v;
  MMD_float* f;
  int ntypes;
  int* type;

  MMD_float virial, mass;
  Box box;

  mmd_ctx* ctx;  // device context that owns the live copy (not owned by Atom)

  explicit Atom(int ntypes_);
  ~Atom();

  // setup-time (host) construction, ref/atom.cpp:86-104
  void addatom(MMD_float x_in, MMD_float y_in, MMD_float z_in, MMD_float vx_in, MMD_float vy_in, MMD_float vz_in);
  void growarray();

  // host mirrors -> device (drops ghosts, like the state right after create_atoms) and back
  int upload();
  int download(bool with_ghosts = false);
  // nlocal / nghost as the device sees them (after exchange / borders)
  void refresh_counts();

  void pbc();                    // ref/atom.cpp:106-122
  void sort(Neighbor& neighbor);  // ref/atom.cpp:355-421

 private:
  std::vector<MMD_float> xs_, vs_, fs_;
  std::vector<int> types_;
  void bind();
};

// Neighbor: bin geometry + stencil on the host (bit-faithful to ref/neighbor.cpp:318-482, they
// decide which bin every atom lands in), binning and list construction on the device.
// Public members follow ref/neighbor.h:38-92.
#pragma once
#include <vector>

#include "atom.h"
#include "types.h"

class Neighbor {
 public:
  int every;                // re-neighbor every this often
  int nbinx, nbiny, nbinz;  // # of global bins
  MMD_float cutneigh;       // neighbor cutoff
  MMD_float* cutneighsq;    // neighbor cutoff squared, [ntypes*ntypes]
  int ncalls;               // # of times build has been called
  int max_totalneigh;       // largest # of neighbors ever stored

  int* numneigh;   // host mirror, filled by download_lists()
  int* neighbors;  // host mirror, row-major [i*maxneighs + k] (ref/force_lj.cpp:207)
  int maxneighs;   // row capacity (ref/neighbor.cpp:48 starts at 100; grows by the *1.2 protocol)
  int halfneigh;
  MMD_int ghost_newton;

  int* bincount;  // host mirrors, filled by download_bins()
  int* bins;
  int mbins;
  int atoms_per_bin;

  long long total_neigh;  // sum(numneigh) of the last build

  explicit Neighbor(int ntypes_);
  ~Neighbor();
  int setup(Atom& atom);                      // ref/neighbor.cpp:318-452 (+ pushes geometry to the device)
  void build(Atom& atom);                     // ref/neighbor.cpp:79-213
  void binatoms(Atom& atom, int count = -1);  // ref/neighbor.cpp:215-268
  int download_lists(Atom& atom);
  int download_bins(Atom& atom);

  // geometry, readable for reports and tests
  int nstencil;
  std::vector<int> stencil;
  int mbinx, mbiny, mbinz;
  int mbinxlo, mbinylo, mbinzlo;
  MMD_float binsizex, binsizey, binsizez;
  MMD_float bininvx, bininvy, bininvz;

 private:
  MMD_float xprd, yprd, zprd;
  int ntypes;
  std::vector<MMD_float> cutsq_;
  std::vector<int> numneigh_, neighbors_, bincount_, bins_;
  MMD_float bindist(int i, int j, int k);
};

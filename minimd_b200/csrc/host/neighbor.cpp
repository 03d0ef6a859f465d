#include "neighbor.h"

#include <stdio.h>
#include <stdlib.h>

#define FACTOR 0.999
#define SMALL 1.0e-6

static void die(const char* what) {
  fprintf(stderr, "ERROR: %s: %s\n", what, mmd_last_error());
  exit(1);
}

Neighbor::Neighbor(int ntypes_) {
  ncalls = 0;
  ntypes = ntypes_;
  max_totalneigh = 0;
  numneigh = neighbors = nullptr;
  maxneighs = 100;
  bincount = bins = nullptr;
  atoms_per_bin = 8;
  mbins = 0;
  halfneigh = 1;
  ghost_newton = 1;
  every = 20;
  nbinx = nbiny = nbinz = 1;
  cutneigh = 0;
  cutsq_.assign((size_t)ntypes * ntypes, 0);
  cutneighsq = cutsq_.data();
  nstencil = 0;
  total_neigh = 0;
  mbinx = mbiny = mbinz = mbinxlo = mbinylo = mbinzlo = 0;
  binsizex = binsizey = binsizez = bininvx = bininvy = bininvz = 0;
  xprd = yprd = zprd = 0;
}

Neighbor::~Neighbor() {}

// lowest / highest bin index that an atom within cutneigh of [lo,hi] can fall into, plus one
// extra layer on each side for stencil coverage
static void axis_bins(MMD_float lo, MMD_float hi, MMD_float cutneigh, MMD_float prd, MMD_float bininv, int& mlo, int& mhi) {
  MMD_float coord = lo - cutneigh - SMALL * prd;
  mlo = static_cast<int>(coord * bininv);
  if (coord < 0.0) mlo = mlo - 1;
  coord = hi + cutneigh + SMALL * prd;
  mhi = static_cast<int>(coord * bininv);
  mlo = mlo - 1;
  mhi = mhi + 1;
}

int Neighbor::setup(Atom& atom) {
  for (int i = 0; i < ntypes * ntypes; i++) cutneighsq[i] = cutneigh * cutneigh;

  xprd = atom.box.xprd;
  yprd = atom.box.yprd;
  zprd = atom.box.zprd;

  binsizex = xprd / nbinx;
  binsizey = yprd / nbiny;
  binsizez = zprd / nbinz;
  bininvx = 1.0 / binsizex;
  bininvy = 1.0 / binsizey;
  bininvz = 1.0 / binsizez;

  int mbinxhi, mbinyhi, mbinzhi;
  axis_bins(atom.box.xlo, atom.box.xhi, cutneigh, xprd, bininvx, mbinxlo, mbinxhi);
  axis_bins(atom.box.ylo, atom.box.yhi, cutneigh, yprd, bininvy, mbinylo, mbinyhi);
  axis_bins(atom.box.zlo, atom.box.zhi, cutneigh, zprd, bininvz, mbinzlo, mbinzhi);
  mbinx = mbinxhi - mbinxlo + 1;
  mbiny = mbinyhi - mbinylo + 1;
  mbinz = mbinzhi - mbinzlo + 1;

  // how far the stencil can reach; FACTOR < 1 makes the LJ benchmark (3 bins per 5 lattice
  // spacings) come out with the intended size
  int nextx = static_cast<int>(cutneigh * bininvx);
  if (nextx * binsizex < FACTOR * cutneigh) nextx++;
  int nexty = static_cast<int>(cutneigh * bininvy);
  if (nexty * binsizey < FACTOR * cutneigh) nexty++;
  int nextz = static_cast<int>(cutneigh * bininvz);
  if (nextz * binsizez < FACTOR * cutneigh) nextz++;

  stencil.clear();
  const bool upper_only = halfneigh && ghost_newton;
  int kstart = -nextz;
  if (upper_only) {
    kstart = 0;
    stencil.push_back(0);
  }
  for (int k = kstart; k <= nextz; k++)
    for (int j = -nexty; j <= nexty; j++)
      for (int i = -nextx; i <= nextx; i++) {
        if (upper_only && !(k > 0 || j > 0 || (j == 0 && i > 0))) continue;
        if (bindist(i, j, k) < cutneighsq[0]) stencil.push_back(k * mbiny * mbinx + j * mbinx + i);
      }
  nstencil = (int)stencil.size();
  mbins = mbinx * mbiny * mbinz;

  if (atom.ctx) {
    mmd_bin_geometry g;
    g.nbinx = nbinx; g.nbiny = nbiny; g.nbinz = nbinz;
    g.mbinx = mbinx; g.mbiny = mbiny; g.mbinz = mbinz;
    g.mbinxlo = mbinxlo; g.mbinylo = mbinylo; g.mbinzlo = mbinzlo;
    g.bininvx = (double)bininvx; g.bininvy = (double)bininvy; g.bininvz = (double)bininvz;
    if (mmd_neigh_setup(atom.ctx, &g, stencil.data(), nstencil, cutneighsq)) {
      fprintf(stderr, "ERROR: mmd_neigh_setup: %s\n", mmd_last_error());
      return 1;
    }
  }
  return 0;
}

MMD_float Neighbor::bindist(int i, int j, int k) {
  MMD_float delx, dely, delz;
  if (i > 0) delx = (i - 1) * binsizex;
  else if (i == 0) delx = 0.0;
  else delx = (i + 1) * binsizex;
  if (j > 0) dely = (j - 1) * binsizey;
  else if (j == 0) dely = 0.0;
  else dely = (j + 1) * binsizey;
  if (k > 0) delz = (k - 1) * binsizez;
  else if (k == 0) delz = 0.0;
  else delz = (k + 1) * binsizez;
  return (delx * delx + dely * dely + delz * delz);
}

void Neighbor::binatoms(Atom& atom, int count) {
  int mx = 0;
  if (mmd_neigh_binatoms(atom.ctx, count, &atoms_per_bin, &mx)) die("mmd_neigh_binatoms");
}

void Neighbor::build(Atom& atom) {
  ncalls++;
  if (mmd_neigh_build(atom.ctx, halfneigh, ghost_newton, &maxneighs, &total_neigh)) die("mmd_neigh_build");
  if (total_neigh > max_totalneigh) max_totalneigh = (int)total_neigh;
  long long apb = 0;
  if (mmd_query_int(atom.ctx, "atoms_per_bin", &apb) == 0) atoms_per_bin = (int)apb;
}

int Neighbor::download_lists(Atom& atom) {
  atom.refresh_counts();
  numneigh_.assign((size_t)atom.nlocal, 0);
  neighbors_.assign((size_t)atom.nlocal * maxneighs, 0);
  numneigh = numneigh_.data();
  neighbors = neighbors_.data();
  return mmd_neigh_download(atom.ctx, numneigh, neighbors, atom.nlocal, maxneighs);
}

int Neighbor::download_bins(Atom& atom) {
  bincount_.assign((size_t)mbins, 0);
  bins_.assign((size_t)mbins * atoms_per_bin, 0);
  bincount = bincount_.data();
  bins = bins_.data();
  return mmd_neigh_bins_download(atom.ctx, bincount, bins, atoms_per_bin);
}

#include "atom.h"

#include <stdio.h>
#include <stdlib.h>

#include "neighbor.h"

#define DELTA 20000  // growth step of the host mirrors, ref/atom.cpp:39

static void die(const char* what) {
  fprintf(stderr, "ERROR: %s: %s\n", what, mmd_last_error());
  exit(1);
}

Atom::Atom(int ntypes_) {
  natoms = 0;
  nlocal = nghost = 0;
  nmax = 0;
  x = v = f = nullptr;
  type = nullptr;
  ntypes = ntypes_;
  virial = 0;
  mass = 1;
  box = Box();
  ctx = nullptr;
}

Atom::~Atom() {}

void Atom::bind() {
  x = xs_.data();
  v = vs_.data();
  f = fs_.data();
  type = types_.data();
}

void Atom::growarray() {
  nmax += DELTA;
  xs_.resize((size_t)nmax * PAD);
  vs_.resize((size_t)nmax * PAD);
  fs_.resize((size_t)nmax * PAD);
  types_.resize((size_t)nmax);
  bind();
}

void Atom::addatom(MMD_float x_in, MMD_float y_in, MMD_float z_in, MMD_float vx_in, MMD_float vy_in, MMD_float vz_in) {
  if (nlocal == nmax) growarray();
  x[nlocal * PAD + 0] = x_in;
  x[nlocal * PAD + 1] = y_in;
  x[nlocal * PAD + 2] = z_in;
  v[nlocal * PAD + 0] = vx_in;
  v[nlocal * PAD + 1] = vy_in;
  v[nlocal * PAD + 2] = vz_in;
  type[nlocal] = rand() % ntypes;  // ref/atom.cpp:97 (libc stream seeded in main)
  nlocal++;
}

int Atom::upload() {
  const double prd[3] = {(double)box.xprd, (double)box.yprd, (double)box.zprd};
  const double lo[3] = {(double)box.xlo, (double)box.ylo, (double)box.zlo};
  const double hi[3] = {(double)box.xhi, (double)box.yhi, (double)box.zhi};
  if (mmd_atom_set_box(ctx, prd, lo, hi)) return 1;
  if (nmax == 0) growarray();
  if (mmd_atom_upload(ctx, x, v, type, nlocal, PAD)) return 1;
  nghost = 0;
  return 0;
}

void Atom::refresh_counts() {
  int cap = 0;
  if (mmd_atom_counts(ctx, &nlocal, &nghost, &cap)) die("mmd_atom_counts");
}

int Atom::download(bool with_ghosts) {
  refresh_counts();
  const int n = nlocal + (with_ghosts ? nghost : 0);
  while (nmax < nlocal + nghost) growarray();
  if (mmd_atom_download(ctx, x, nullptr, f, type, 0, n, PAD)) return 1;
  if (mmd_atom_download(ctx, nullptr, v, nullptr, nullptr, 0, nlocal, PAD)) return 1;
  return 0;
}

void Atom::pbc() {
  if (mmd_atom_pbc(ctx)) die("mmd_atom_pbc");
}

void Atom::sort(Neighbor&) {
  if (mmd_atom_sort(ctx)) die("mmd_atom_sort");
}

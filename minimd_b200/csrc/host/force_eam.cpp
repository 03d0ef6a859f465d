#include "force_eam.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

ForceEAM::ForceEAM(int ntypes_) : Force(ntypes_) {
  cutforce = 0.0;
  cutmax = 0;
  use_oldcompute = 0;
  nrho = nr = nrho_tot = nr_tot = 0;
  dr = rdr = drho = rdrho = 0;
  rhor_spline = frho_spline = z2r_spline = nullptr;
  style = FORCEEAM;
  potential_file = "Cu_u6.eam";
}
ForceEAM::~ForceEAM() {}

// whitespace-separated reals, any number per line (ForceEAM::grab, ref/force_eam.cpp:800-815)
static bool grab(FILE* fp, int n, MMD_float* dst) {
  char line[1024];
  int got = 0;
  while (got < n) {
    if (!fgets(line, sizeof line, fp)) return false;
    for (char* tok = strtok(line, " \t\n\r\f"); tok && got < n; tok = strtok(nullptr, " \t\n\r\f")) dst[got++] = atof(tok);
  }
  return true;
}

// DYNAMO single-element funcfl file (ref/force_eam.cpp:505-582).  Every rank reads the file itself
// (the reference reads on rank 0 and broadcasts; the bytes are the same).
int ForceEAM::read_file(const char* filename, int me) {
  FILE* fp = fopen(filename, "r");
  if (!fp) {
    if (me == 0) printf("Can't open EAM Potential file: %s\n", filename);
    return 1;
  }
  char line[1024];
  int tmp = 0;
  bool ok = fgets(line, sizeof line, fp) && fgets(line, sizeof line, fp);
  if (ok) sscanf(line, "%d %lg", &tmp, &funcfl.mass);
  ok = ok && fgets(line, sizeof line, fp);
  if (ok) sscanf(line, "%d %lg %d %lg %lg", &funcfl.nrho, &funcfl.drho, &funcfl.nr, &funcfl.dr, &funcfl.cut);
  if (!ok || funcfl.nrho < 4 || funcfl.nr < 4) {
    if (me == 0) printf("Bad EAM Potential file header: %s\n", filename);
    fclose(fp);
    return 1;
  }
  mass = funcfl.mass;
  funcfl.frho.assign(funcfl.nrho + 1, 0);
  funcfl.rhor.assign(funcfl.nr + 1, 0);
  funcfl.zr.assign(funcfl.nr + 1, 0);
  ok = grab(fp, funcfl.nrho, funcfl.frho.data()) && grab(fp, funcfl.nr, funcfl.zr.data()) &&
       grab(fp, funcfl.nr, funcfl.rhor.data());
  fclose(fp);
  if (!ok) {
    if (me == 0) printf("Truncated EAM Potential file: %s\n", filename);
    return 1;
  }
  // tables are addressed 1-based from here on
  for (int i = funcfl.nrho; i > 0; i--) funcfl.frho[i] = funcfl.frho[i - 1];
  for (int i = funcfl.nr; i > 0; i--) funcfl.rhor[i] = funcfl.rhor[i - 1];
  for (int i = funcfl.nr; i > 0; i--) funcfl.zr[i] = funcfl.zr[i - 1];
  return 0;
}

// value of a 1-based table at r by 4-point Lagrange interpolation (ref/force_eam.cpp:630-726)
static double lagrange4(const MMD_float* tab, int ntab, double dtab, double r) {
  const double sixth = 1.0 / 6.0;
  double p = r / dtab + 1.0;
  int k = static_cast<int>(p);
  k = std::min(k, ntab - 2);
  k = std::max(k, 2);
  p -= k;
  p = std::min(p, 2.0);
  const double cof1 = -sixth * p * (p - 1.0) * (p - 2.0);
  const double cof2 = 0.5 * (p * p - 1.0) * (p - 2.0);
  const double cof3 = -0.5 * p * (p + 1.0) * (p - 2.0);
  const double cof4 = sixth * p * (p * p - 1.0);
  return cof1 * tab[k - 1] + cof2 * tab[k] + cof3 * tab[k + 1] + cof4 * tab[k + 2];
}

// re-grid the file tables onto (nr, dr) / (nrho, drho) (ref/force_eam.cpp:589-728)
void ForceEAM::file2array() {
  dr = funcfl.dr;
  drho = funcfl.drho;
  const double rmax = (funcfl.nr - 1) * funcfl.dr;
  const double rhomax = (funcfl.nrho - 1) * funcfl.drho;
  nr = static_cast<int>(rmax / dr + 0.5);
  nrho = static_cast<int>(rhomax / drho + 0.5);

  frho_.assign(nrho + 1, 0);
  for (int m = 1; m <= nrho; m++) frho_[m] = lagrange4(funcfl.frho.data(), funcfl.nrho, funcfl.drho, (m - 1) * drho);
  rhor_.assign(nr + 1, 0);
  for (int m = 1; m <= nr; m++) rhor_[m] = lagrange4(funcfl.rhor.data(), funcfl.nr, funcfl.dr, (m - 1) * dr);
  // z2r = 27.2 * 0.529 * Zi * Zj : Hartree -> eV, Bohr -> Angstrom (ref/force_eam.cpp:725)
  z2r_.assign(nr + 1, 0);
  for (int m = 1; m <= nr; m++) {
    const double r = (m - 1) * dr;
    const double zri = lagrange4(funcfl.zr.data(), funcfl.nr, funcfl.dr, r);
    const double zrj = lagrange4(funcfl.zr.data(), funcfl.nr, funcfl.dr, r);
    z2r_[m] = 27.2 * 0.529 * zri * zrj;
  }
}

// 7 coefficients per knot: [3..6] cubic for the value, [0..2] its derivative (ref/force_eam.cpp:765-793)
void ForceEAM::interpolate(MMD_int n, MMD_float delta, const MMD_float* f, MMD_float* spline) {
  for (int m = 1; m <= n; m++) spline[m * 7 + 6] = f[m];
  spline[1 * 7 + 5] = spline[2 * 7 + 6] - spline[1 * 7 + 6];
  spline[2 * 7 + 5] = 0.5 * (spline[3 * 7 + 6] - spline[1 * 7 + 6]);
  spline[(n - 1) * 7 + 5] = 0.5 * (spline[n * 7 + 6] - spline[(n - 2) * 7 + 6]);
  spline[n * 7 + 5] = spline[n * 7 + 6] - spline[(n - 1) * 7 + 6];
  for (int m = 3; m <= n - 2; m++)
    spline[m * 7 + 5] = ((spline[(m - 2) * 7 + 6] - spline[(m + 2) * 7 + 6]) +
                         8.0 * (spline[(m + 1) * 7 + 6] - spline[(m - 1) * 7 + 6])) / 12.0;
  for (int m = 1; m <= n - 1; m++) {
    spline[m * 7 + 4] = 3.0 * (spline[(m + 1) * 7 + 6] - spline[m * 7 + 6]) - 2.0 * spline[m * 7 + 5] - spline[(m + 1) * 7 + 5];
    spline[m * 7 + 3] = spline[m * 7 + 5] + spline[(m + 1) * 7 + 5] - 2.0 * (spline[(m + 1) * 7 + 6] - spline[m * 7 + 6]);
  }
  spline[n * 7 + 4] = 0.0;
  spline[n * 7 + 3] = 0.0;
  for (int m = 1; m <= n; m++) {
    spline[m * 7 + 2] = spline[m * 7 + 5] / delta;
    spline[m * 7 + 1] = 2.0 * spline[m * 7 + 4] / delta;
    spline[m * 7 + 0] = 3.0 * spline[m * 7 + 3] / delta;
  }
}

// ref/force_eam.cpp:732-761: spline the three tables, then replicate them for every type pair
void ForceEAM::array2spline() {
  rdr = 1.0 / dr;
  rdrho = 1.0 / drho;
  nrho_tot = (nrho + 1) * 7 + 64;
  nr_tot = (nr + 1) * 7 + 64;
  nrho_tot -= nrho_tot % 64;
  nr_tot -= nr_tot % 64;
  const size_t nn = (size_t)ntypes * ntypes;
  frho_sp_.assign(nn * nrho_tot, 0);
  rhor_sp_.assign(nn * nr_tot, 0);
  z2r_sp_.assign(nn * nr_tot, 0);
  frho_spline = frho_sp_.data();
  rhor_spline = rhor_sp_.data();
  z2r_spline = z2r_sp_.data();
  interpolate(nrho, drho, frho_.data(), frho_spline);
  interpolate(nr, dr, rhor_.data(), rhor_spline);
  interpolate(nr, dr, z2r_.data(), z2r_spline);
  for (size_t t = 1; t < nn; t++) {
    memcpy(frho_spline + t * nrho_tot, frho_spline, sizeof(MMD_float) * nrho_tot);
    memcpy(rhor_spline + t * nr_tot, rhor_spline, sizeof(MMD_float) * nr_tot);
    memcpy(z2r_spline + t * nr_tot, z2r_spline, sizeof(MMD_float) * nr_tot);
  }
}

// ForceEAM::setup = coeff + init_style (ref/force_eam.cpp:74-79, 457-487)
int ForceEAM::setup(Atom& atom) {
  if (read_file(potential_file.c_str(), 0)) return 1;
  cutmax = funcfl.cut;
  for (int i = 0; i < ntypes * ntypes; i++) cutforcesq[i] = cutmax * cutmax;
  file2array();
  array2spline();
  ctx = atom.ctx;
  if (ctx && mmd_force_eam_setup(ctx, rhor_spline, z2r_spline, frho_spline, nr, nrho, nr_tot, nrho_tot, (double)rdr,
                                 (double)rdrho, cutforcesq)) {
    fprintf(stderr, "ERROR: mmd_force_eam_setup: %s\n", mmd_last_error());
    return 1;
  }
  return 0;
}

// ref/force_eam.cpp:82-91.  (The reference silently skips the half-list path when more than one
// OpenMP thread is requested, :84-88; that quirk is not reproduced.)
void ForceEAM::compute(Atom& atom, Neighbor& neighbor, Comm&, int) {
  if (mmd_force_eam_compute(atom.ctx, neighbor.halfneigh != 0, evflag, &eng_vdwl, &virial)) {
    fprintf(stderr, "ERROR: mmd_force_eam_compute: %s\n", mmd_last_error());
    exit(1);
  }
}

#include "setup.h"

#include <math.h>
#include <stdio.h>

#include <algorithm>

void create_box(Atom& atom, int nx, int ny, int nz, double rho) {
  const double lattice = pow((4.0 / rho), (1.0 / 3.0));
  atom.box.xprd = nx * lattice;
  atom.box.yprd = ny * lattice;
  atom.box.zprd = nz * lattice;
}

// Park-Miller "minimal standard" generator without masking (ref/setup.cpp:498-517)
double park_miller(int* state) {
  const int ia = 16807, im = 2147483647, iq = 127773, ir = 2836;
  const int k = *state / iq;
  *state = ia * (*state - k * iq) - ir * k;
  if (*state < 0) *state += im;
  return (1.0 / im) * (*state);
}

// FCC lattice: half-lattice points (i,j,k) with i+j+k even, visited in 8x8x8 blocks (blocks and
// points inside a block both x-fastest), kept if inside this rank's sub-box.  The velocity of a
// point comes from a private Park-Miller stream seeded with its global lattice id, 5 discarded
// draws before each component (ref/setup.cpp:315-422).
int create_atoms(Atom& atom, int nx, int ny, int nz, double rho, World& world) {
  atom.natoms = 4 * nx * ny * nz;
  atom.nlocal = 0;

  const double alat = pow((4.0 / rho), (1.0 / 3.0));
  int ilo = static_cast<int>(atom.box.xlo / (0.5 * alat) - 1);
  int ihi = static_cast<int>(atom.box.xhi / (0.5 * alat) + 1);
  int jlo = static_cast<int>(atom.box.ylo / (0.5 * alat) - 1);
  int jhi = static_cast<int>(atom.box.yhi / (0.5 * alat) + 1);
  int klo = static_cast<int>(atom.box.zlo / (0.5 * alat) - 1);
  int khi = static_cast<int>(atom.box.zhi / (0.5 * alat) + 1);
  ilo = std::max(ilo, 0);
  ihi = std::min(ihi, 2 * nx - 1);
  jlo = std::max(jlo, 0);
  jhi = std::min(jhi, 2 * ny - 1);
  klo = std::max(klo, 0);
  khi = std::min(khi, 2 * nz - 1);

  const int B = 8;
  for (int oz = 0; oz * B <= khi; oz++)
    for (int oy = 0; oy * B <= jhi; oy++)
      for (int ox = 0; ox * B <= ihi; ox++)
        for (int sz = 0; sz < B; sz++)
          for (int sy = 0; sy < B; sy++)
            for (int sx = 0; sx < B; sx++) {
              const int i = ox * B + sx, j = oy * B + sy, k = oz * B + sz;
              if ((i + j + k) % 2) continue;
              if (i < ilo || i > ihi || j < jlo || j > jhi || k < klo || k > khi) continue;
              const double xtmp = 0.5 * alat * i, ytmp = 0.5 * alat * j, ztmp = 0.5 * alat * k;
              if (!(xtmp >= atom.box.xlo && xtmp < atom.box.xhi && ytmp >= atom.box.ylo && ytmp < atom.box.yhi &&
                    ztmp >= atom.box.zlo && ztmp < atom.box.zhi))
                continue;
              int n = k * (2 * ny) * (2 * nx) + j * (2 * nx) + i + 1;
              double vel[3];
              for (int c = 0; c < 3; c++) {
                for (int m = 0; m < 5; m++) park_miller(&n);
                vel[c] = park_miller(&n);
              }
              atom.addatom(xtmp, ytmp, ztmp, vel[0], vel[1], vel[2]);
            }

  const long long natoms = world.sum_ll(atom.nlocal);
  if (natoms != atom.natoms) {
    if (world.me == 0) printf("Created incorrect # of atoms\n");
    return 1;
  }
  return 0;
}

// zero the centre-of-mass velocity, then rescale to the requested temperature (ref/setup.cpp:454-494)
void create_velocity(double t_request, Atom& atom, Thermo& thermo, World& world) {
  double vtot[3] = {0.0, 0.0, 0.0};
  for (int i = 0; i < atom.nlocal; i++)
    for (int c = 0; c < 3; c++) vtot[c] += atom.v[i * PAD + c];
  world.sum(vtot, 3);
  for (int c = 0; c < 3; c++) vtot[c] = vtot[c] / atom.natoms;
  for (int i = 0; i < atom.nlocal; i++)
    for (int c = 0; c < 3; c++) atom.v[i * PAD + c] -= vtot[c];

  // Thermo::temperature on the host mirrors (the atoms are not on the device yet)
  MMD_float t = 0.0;
  for (int i = 0; i < atom.nlocal; i++) {
    const MMD_float vx = atom.v[i * PAD + 0], vy = atom.v[i * PAD + 1], vz = atom.v[i * PAD + 2];
    t += (vx * vx + vy * vy + vz * vz) * atom.mass;
  }
  double tsum = (double)t;
  world.sum(&tsum, 1);
  const double temp = (MMD_float)((MMD_float)tsum * thermo.t_scale);
  const double factor = sqrt(t_request / temp);
  for (int i = 0; i < atom.nlocal; i++)
    for (int c = 0; c < 3; c++) atom.v[i * PAD + c] *= factor;
}

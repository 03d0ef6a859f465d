#include "setup.h"

#include "comm.h"
#include "integrate.h"
#include "neighbor.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

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

// ---------------------------------------------------------------------------------------------
// LAMMPS data file
// ---------------------------------------------------------------------------------------------
namespace {

bool blank(const std::string& s) { return s.find_first_not_of(" \t\n\r") == std::string::npos; }

std::string trimmed(const std::string& s) {
  const size_t a = s.find_first_not_of(" \t\n\r");
  if (a == std::string::npos) return "";
  const size_t b = s.find_last_not_of(" \t\n\r");
  return s.substr(a, b - a + 1);
}

bool next_line(FILE* fp, std::string& out) {
  char buf[512];
  if (!fgets(buf, sizeof buf, fp)) return false;
  out = buf;
  return true;
}

}  // namespace

int read_lammps_data(Atom& atom, Comm& comm, Neighbor& neighbor, Integrate& integrate, Thermo& thermo, const char* file,
                     int units, World& world) {
  FILE* fp = fopen(file, "r");
  if (!fp) {
    if (world.me == 0) printf("ERROR: Cannot open file %s\n", file);
    return 1;
  }
  std::string line;
  next_line(fp, line);  // title

  // header: "<n> atoms", "<n> atom types", "<lo> <hi> xlo xhi" ...; anything after '#' is a comment.
  // As in the reference only the box LENGTHS are used (the box is taken to start at the origin).
  std::string section;
  atom.natoms = 0;
  while (next_line(fp, line)) {
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line.erase(hash);
    if (blank(line)) continue;
    double lo = 0, hi = 0;
    if (line.find("atom types") != std::string::npos) continue;
    if (line.find("atoms") != std::string::npos) sscanf(line.c_str(), "%i", &atom.natoms);
    else if (line.find("xlo xhi") != std::string::npos) { sscanf(line.c_str(), "%lg %lg", &lo, &hi); atom.box.xprd = hi - lo; }
    else if (line.find("ylo yhi") != std::string::npos) { sscanf(line.c_str(), "%lg %lg", &lo, &hi); atom.box.yprd = hi - lo; }
    else if (line.find("zlo zhi") != std::string::npos) { sscanf(line.c_str(), "%lg %lg", &lo, &hi); atom.box.zprd = hi - lo; }
    else { section = trimmed(line); break; }
  }
  if (atom.natoms <= 0 || !(atom.box.xprd > 0 && atom.box.yprd > 0 && atom.box.zprd > 0)) {
    if (world.me == 0) printf("ERROR: bad header in data file %s\n", file);
    fclose(fp);
    return 1;
  }

  if (comm.setup(neighbor.cutneigh, atom)) { fclose(fp); return 1; }
  if (neighbor.nbinx < 0) {  // no -b: about 16 atoms per 2x2x2 bins (ref/setup.cpp:226-233)
    const MMD_float volume = atom.box.xprd * atom.box.yprd * atom.box.zprd;
    const MMD_float rho = 1.0 * atom.natoms / volume;
    const MMD_float neigh_bin_size = pow(rho * 16, MMD_float(1.0 / 3.0));
    neighbor.nbinx = atom.box.xprd / neigh_bin_size;
    neighbor.nbiny = atom.box.yprd / neigh_bin_size;
    neighbor.nbinz = atom.box.zprd / neigh_bin_size;
  }
  if (neighbor.nbinx == 0) neighbor.nbinx = 1;
  if (neighbor.nbiny == 0) neighbor.nbiny = 1;
  if (neighbor.nbinz == 0) neighbor.nbinz = 1;
  if (neighbor.setup(atom)) { fclose(fp); return 1; }
  integrate.setup();
  thermo.setup(atom.box.xprd * atom.box.yprd * atom.box.zprd / atom.natoms, integrate, atom, units);

  std::vector<MMD_float> x((size_t)atom.natoms * 3, 0), v((size_t)atom.natoms * 3, 0);
  bool have_atoms = false;
  while (!section.empty()) {
    next_line(fp, line);  // the blank line under the section keyword
    if (section == "Atoms" || section == "Velocities") {
      const bool pos = section == "Atoms";
      if (!pos && !have_atoms && world.me == 0) printf("Must read Atoms before Velocities\n");
      std::vector<MMD_float>& dst = pos ? x : v;
      for (int nread = 0; nread < atom.natoms; nread++) {
        if (!next_line(fp, line)) break;
        int id = 0, type = 0;
        double a = 0, b = 0, c = 0;
        if (pos) sscanf(line.c_str(), "%i %i %lg %lg %lg", &id, &type, &a, &b, &c);
        else sscanf(line.c_str(), "%i %lg %lg %lg", &id, &a, &b, &c);
        if (id < 1 || id > atom.natoms) continue;
        dst[(size_t)(id - 1) * 3 + 0] = a;
        dst[(size_t)(id - 1) * 3 + 1] = b;
        dst[(size_t)(id - 1) * 3 + 2] = c;
      }
      if (pos) have_atoms = true;
    } else if (section == "Masses") {
      if (next_line(fp, line)) {
        int t = 0;
        double m = 0;
        if (sscanf(line.c_str(), "%i %lg", &t, &m) == 2) atom.mass = m;
      }
      while (next_line(fp, line) && !blank(line)) {}  // further types (all atoms share one mass in miniMD)
    }
    // next section keyword = next non-blank line
    section.clear();
    while (next_line(fp, line)) {
      if (!blank(line)) { section = trimmed(line); break; }
    }
  }
  fclose(fp);

  atom.nlocal = 0;
  for (int i = 0; i < atom.natoms; i++) {
    const MMD_float* p = &x[(size_t)i * 3];
    if (p[0] >= atom.box.xlo && p[0] < atom.box.xhi && p[1] >= atom.box.ylo && p[1] < atom.box.yhi &&
        p[2] >= atom.box.zlo && p[2] < atom.box.zhi)
      atom.addatom(p[0], p[1], p[2], v[(size_t)i * 3 + 0], v[(size_t)i * 3 + 1], v[(size_t)i * 3 + 2]);
  }
  const long long natoms = world.sum_ll(atom.nlocal);
  if (natoms != atom.natoms) {
    if (world.me == 0) printf("Created incorrect # of atoms\n");
    return 1;
  }
  return 0;
}

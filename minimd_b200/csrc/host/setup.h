// Synthetic-input construction (ref/setup.cpp:305-517): these functions DEFINE the benchmark's
// inputs, so lattice traversal order, the per-atom Park-Miller velocity streams and the
// centre-of-mass / temperature normalisation follow the reference exactly.  Host code, run once.
#pragma once
#include "atom.h"
#include "thermo.h"
#include "world.h"

void create_box(Atom& atom, int nx, int ny, int nz, double rho);
int create_atoms(Atom& atom, int nx, int ny, int nz, double rho, World& world);
void create_velocity(double t_request, Atom& atom, Thermo& thermo, World& world);
double park_miller(int* state);

class Comm;
class Neighbor;
class Integrate;
// Start from a LAMMPS data file instead of the synthetic lattice (ref/setup.cpp:54-301): header (atoms,
// box bounds), sections Atoms ("id type x y z"), Velocities ("id vx vy vz"), Masses.  Performs the same
// setup sequence the reference does inside this routine (Comm::setup, bin counts from the density,
// Neighbor::setup, Integrate::setup, Thermo::setup) and keeps the atoms of this rank's sub-box, in id order.
int read_lammps_data(Atom& atom, Comm& comm, Neighbor& neighbor, Integrate& integrate, Thermo& thermo, const char* file,
                     int units, World& world);

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

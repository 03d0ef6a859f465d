// The 14-line positional input file as one record (ref/ljs.h:37-51, parsed by input.cpp).
#pragma once
#include <string>

#include "types.h"

struct In {
  int nx, ny, nz;
  MMD_float t_request;
  MMD_float rho;
  int units;
  ForceStyle forcetype;
  MMD_float epsilon, sigma;
  std::string datafile;  // empty = "none"
  int ntimes;
  MMD_float dt;
  int neigh_every;
  MMD_float force_cut;
  MMD_float neigh_cut;  // force_cut + skin after input()
  int thermo_nstat;
};

// returns 0 on success; prints "ERROR: Cannot open <file>" like ref/input.cpp:68-72 otherwise
int input(In& in, const char* filename, int me);

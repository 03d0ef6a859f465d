// Thermo: temperature / potential energy / pressure records (ref/thermo.h, ref/thermo.cpp).
// The per-atom reduction sum m v^2 runs on the device; the scalings and the cross-rank sums
// (the reference's three MPI_Allreduce calls) stay here.
#pragma once
#include <vector>

#include "atom.h"
#include "force.h"
#include "neighbor.h"
#include "timer.h"
#include "types.h"
#include "world.h"

class Integrate;

class Thermo {
 public:
  MMD_int nstat;
  MMD_int mstat;
  MMD_int ntimes;
  std::vector<MMD_int> steparr;
  std::vector<MMD_float> tmparr, engarr, prsarr;

  Thermo();
  ~Thermo();
  void setup(MMD_float rho_in, Integrate& integrate, Atom& atom, MMD_int units);
  // ref/thermo.cpp:74-115.  iflag: 0 = initial record, >0 = step number, -1 = final record.
  void compute(MMD_int iflag, Atom& atom, Neighbor& neighbor, Force* force, Timer& timer, Comm& comm);
  // same record from reductions the fused time loop already produced (this rank's partial sums)
  void record(MMD_int istep, double sum_mv2, double eng_vdwl, double virial, Atom& atom, Neighbor& neighbor, Timer& timer);
  MMD_float temperature(Atom& atom);

  World* world;
  bool quiet;  // no stdout lines (embedding)

  MMD_float t_act, p_act, e_act;
  MMD_float t_scale, e_scale, p_scale, mvv2e, dof_boltz;

 private:
  MMD_float rho;
  void store_and_print(MMD_int istep, MMD_float t, MMD_float eng, MMD_float p, Timer& timer);
};

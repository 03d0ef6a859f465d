#include "thermo.h"

#include <stdio.h>
#include <stdlib.h>

#include "integrate.h"

Thermo::Thermo() {
  nstat = 0;
  mstat = 0;
  ntimes = 0;
  world = nullptr;
  quiet = false;
  t_act = p_act = e_act = 0;
  t_scale = e_scale = p_scale = mvv2e = dof_boltz = 0;
  rho = 0;
}
Thermo::~Thermo() {}

void Thermo::setup(MMD_float rho_in, Integrate& integrate, Atom& atom, MMD_int units) {
  rho = rho_in;
  ntimes = integrate.ntimes;
  steparr.clear();
  tmparr.clear();
  engarr.clear();
  prsarr.clear();

  if (units == LJ) {
    mvv2e = 1.0;
    dof_boltz = (atom.natoms * 3 - 3);
    t_scale = mvv2e / dof_boltz;
    p_scale = 1.0 / 3 / atom.box.xprd / atom.box.yprd / atom.box.zprd;
    e_scale = 0.5;
  } else if (units == METAL) {
    mvv2e = 1.036427e-04;
    dof_boltz = (atom.natoms * 3 - 3) * 8.617343e-05;
    t_scale = mvv2e / dof_boltz;
    p_scale = 1.602176e+06 / 3 / atom.box.xprd / atom.box.yprd / atom.box.zprd;
    e_scale = 524287.985533;  // the reference's literal (ref/thermo.cpp:68); the goldens depend on it
    integrate.dtforce /= mvv2e;
  }
}

MMD_float Thermo::temperature(Atom& atom) {
  double s = 0.0;
  if (mmd_thermo_sum_mv2(atom.ctx, (double)atom.mass, &s)) {
    fprintf(stderr, "ERROR: mmd_thermo_sum_mv2: %s\n", mmd_last_error());
    exit(1);
  }
  t_act = (MMD_float)s;
  if (world) world->sum(&s, 1);
  return (MMD_float)s * t_scale;
}

void Thermo::store_and_print(MMD_int istep, MMD_float t, MMD_float eng, MMD_float p, Timer& timer) {
  if (istep == 0) mstat = 0;
  steparr.resize(mstat + 1);
  tmparr.resize(mstat + 1);
  engarr.resize(mstat + 1);
  prsarr.resize(mstat + 1);
  steparr[mstat] = istep;
  tmparr[mstat] = t;
  engarr[mstat] = eng;
  prsarr[mstat] = p;
  mstat++;
  if (!quiet && (!world || world->me == 0))
    fprintf(stdout, "%i %e %e %e %6.3lf\n", istep, t, eng, p, istep == 0 ? 0.0 : timer.elapsed_since_start());
}

void Thermo::record(MMD_int istep, double sum_mv2, double eng_vdwl, double virial, Atom& atom, Neighbor& neighbor,
                    Timer& timer) {
  // cross-rank sums of the three partial reductions (ref/thermo.cpp:131-133,168-170,188-190)
  e_act = (MMD_float)eng_vdwl;
  if (neighbor.halfneigh) e_act *= 2.0;
  e_act *= e_scale;
  t_act = (MMD_float)sum_mv2;
  p_act = (MMD_float)virial;
  double v[3] = {(double)t_act, (double)e_act, (double)p_act};
  if (world) world->sum(v, 3);
  const MMD_float t = (MMD_float)v[0] * t_scale;
  const MMD_float eng = (MMD_float)v[1] / atom.natoms;
  const MMD_float p = (t * dof_boltz + (MMD_float)v[2]) * p_scale;
  store_and_print(istep, t, eng, p, timer);
}

void Thermo::compute(MMD_int iflag, Atom& atom, Neighbor& neighbor, Force* force, Timer& timer, Comm&) {
  if (iflag > 0 && nstat > 0 && iflag % nstat) return;
  if (iflag > 0 && nstat == 0) return;
  if (iflag == -1 && nstat > 0 && ntimes % nstat == 0) return;
  double s = 0.0;
  if (mmd_thermo_sum_mv2(atom.ctx, (double)atom.mass, &s)) {
    fprintf(stderr, "ERROR: mmd_thermo_sum_mv2: %s\n", mmd_last_error());
    exit(1);
  }
  record(iflag == -1 ? ntimes : iflag, s, (double)force->eng_vdwl, (double)force->virial, atom, neighbor, timer);
}

#include "integrate.h"

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

static void die(const char* what) {
  fprintf(stderr, "ERROR: %s: %s\n", what, mmd_last_error());
  exit(1);
}

Integrate::Integrate() {
  dt = dtforce = 0;
  ntimes = 0;
  nlocal = nmax = 0;
  mass = 1;
  sort_every = 20;
  stepwise = 0;
  device_ms = 0;
  steps_done = 0;
  ctx_ = nullptr;
  mass_folded_ = false;
}
Integrate::~Integrate() {}

void Integrate::setup() { dtforce = 0.5 * dt; }

void Integrate::initialIntegrate() {
  if (mmd_integrate_initial(ctx_, (double)dt, (double)dtforce)) die("mmd_integrate_initial");
}
void Integrate::finalIntegrate() {
  if (mmd_integrate_final(ctx_, (double)dtforce)) die("mmd_integrate_final");
}

void Integrate::run(Atom& atom, Force* force, Neighbor& neighbor, Comm& comm, Thermo& thermo, Timer& timer, int nsteps) {
  ctx_ = atom.ctx;
  mass = atom.mass;
  if (!mass_folded_) {  // ref/integrate.cpp:80-81, once per simulation
    dtforce = dtforce / mass;
    mass_folded_ = true;
  }
  if (nsteps < 0) nsteps = ntimes - steps_done;
  const int first = steps_done, last = steps_done + nsteps;
  const bool reverse = neighbor.halfneigh && neighbor.ghost_newton;
  device_ms = 0;

  if (stepwise) {
    int next_sort = sort_every > 0 ? sort_every : ntimes + 1;
    while (sort_every > 0 && next_sort <= first) next_sort += sort_every;
    for (int n = first; n < last; n++) {
      initialIntegrate();
      if ((n + 1) % neighbor.every) {
        comm.communicate(atom);
      } else {
        comm.exchange(atom);
        if (n + 1 >= next_sort) {
          atom.sort(neighbor);
          next_sort += sort_every;
        }
        comm.borders(atom);
        neighbor.build(atom);
      }
      force->evflag = thermo.nstat > 0 ? ((n + 1) % thermo.nstat == 0) : 0;
      force->compute(atom, neighbor, comm, comm.me);
      if (reverse) comm.reverse_communicate(atom);
      finalIntegrate();
      if (thermo.nstat) thermo.compute(n + 1, atom, neighbor, force, timer, comm);
    }
    steps_done = last;
    atom.refresh_counts();
    return;
  }

  mmd_run_params p;
  p.total_steps = ntimes;
  p.neigh_every = neighbor.every;
  p.sort_every = sort_every;
  p.thermo_nstat = thermo.nstat;
  p.halfneigh = neighbor.halfneigh != 0;
  p.ghost_newton = neighbor.ghost_newton;
  p.force_style = force->style == FORCELJ ? 0 : 1;
  p.dt = (double)dt;
  p.dtforce = (double)dtforce;
  p.mass = (double)mass;
  // one mmd_run per stretch that ends on a thermo step, so records print as the run advances
  int n = first;
  while (n < last) {
    int stop = last;
    if (thermo.nstat > 0) stop = std::min(last, (n / thermo.nstat + 1) * thermo.nstat);
    p.first_step = n;
    p.ntimes = stop - n;
    mmd_thermo_sample sample[2];
    int ns = 0;
    float ms = 0;
    if (mmd_run(ctx_, &p, sample, 2, &ns, &ms)) die("mmd_run");
    device_ms += ms;
    for (int k = 0; k < ns && k < 2; k++)
      thermo.record(sample[k].step, sample[k].sum_mv2, sample[k].eng_vdwl, sample[k].virial, atom, neighbor, timer);
    n = stop;
  }
  steps_done = last;
  atom.refresh_counts();
  long long v = 0;
  if (!mmd_query_int(ctx_, "maxneighs", &v)) neighbor.maxneighs = (int)v;
  if (!mmd_query_int(ctx_, "total_neigh", &v)) neighbor.total_neigh = v;
  if (!mmd_query_int(ctx_, "neigh_builds", &v)) neighbor.ncalls = (int)v;
  comm.refresh_counts(atom);
}

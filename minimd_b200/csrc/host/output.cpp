#include "output.h"

#include <stdarg.h>
#include <stdio.h>
#include <time.h>

#include <string>

#include "sim.h"

namespace {

// writes to the report file and, with --yaml_screen, to stdout as well
struct Sink {
  FILE* fp = nullptr;
  bool screen = false;
  void operator()(const char* fmt, ...) const {
    va_list ap;
    if (fp) {
      va_start(ap, fmt);
      vfprintf(fp, fmt, ap);
      va_end(ap);
    }
    if (screen) {
      va_start(ap, fmt);
      vfprintf(stdout, fmt, ap);
      va_end(ap);
    }
  }
};

// average / extremes / 10-bucket histogram of one value per rank (ref/output.cpp:496-547)
struct Stats {
  double ave, max, min;
  int histo[10];
};
Stats rank_stats(const World& w, double value) {
  Stats s;
  double sum = value, mx = value, mn = -value;
  w.sum(&sum, 1);
  w.max(&mx, 1);
  w.max(&mn, 1);
  mn = -mn;
  s.ave = sum / w.nprocs;
  s.max = mx;
  s.min = mn;
  double h[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const double del = mx - mn;
  int m = del == 0.0 ? 0 : static_cast<int>((value - mn) / del * 10);
  if (m > 9) m = 9;
  h[m] = 1;
  w.sum(h, 10);
  for (int i = 0; i < 10; i++) s.histo[i] = (int)(h[i] + 0.5);
  return s;
}
void print_stats(const Sink& out, const char* label, const Stats& s) {
  out("# %s %g ave %g max %g min\n", label, s.ave, s.max, s.min);
  out("# Histogram:");
  for (int i = 0; i < 10; i++) out(" %d", s.histo[i]);
  out("\n");
}

}  // namespace

void output(Simulation& sim) {
  const World& w = sim.world;
  Atom& atom = *sim.atom;
  Neighbor& neighbor = *sim.neighbor;
  const In& in = sim.in;

  // lost-atom check after wrapping (ref/output.cpp:62-85)
  atom.pbc();
  atom.download(false);
  const long long natoms = w.sum_ll(atom.nlocal);
  long long nlost = 0;
  for (int i = 0; i < atom.nlocal; i++)
    if (atom.x[i * PAD + 0] < 0.0 || atom.x[i * PAD + 0] >= atom.box.xprd || atom.x[i * PAD + 1] < 0.0 ||
        atom.x[i * PAD + 1] >= atom.box.yprd || atom.x[i * PAD + 2] < 0.0 || atom.x[i * PAD + 2] >= atom.box.zprd)
      nlost++;
  nlost = w.sum_ll(nlost);
  if (natoms != atom.natoms || nlost > 0) {
    if (w.me == 0) printf("Atom counts = %lld %lld %d\n", nlost, natoms, atom.natoms);
    if (w.me == 0) printf("ERROR: Incorrect number of atoms\n");
    return;
  }

  Sink out;
  out.screen = w.me == 0 && sim.opt.screen_yaml;
  if (w.me == 0) {
    time_t now = time(nullptr);
    tm lt = *localtime(&now);
    char filename[256];
    snprintf(filename, sizeof filename, "miniMD-%4d-%02d-%02d-%02d-%02d-%02d.yaml", lt.tm_year + 1900, lt.tm_mon + 1,
             lt.tm_mday, lt.tm_hour, lt.tm_min, lt.tm_sec);
    out.fp = fopen(filename, "w");
  }

  out("run_configuration: \n");
  out("  variant: " VARIANT_STRING "\n");
  out("  mpi_processes: %i\n", w.nprocs);
  out("  thread_teams: %i\n", sim.opt.teams);
  out("  threads: %i\n", sim.opt.num_threads);
  out("  datafile: %s\n", in.datafile.empty() ? "None" : in.datafile.c_str());
  out("  units: %s\n", in.units == 0 ? "LJ" : "METAL");
  out("  atoms: %i\n", atom.natoms);
  out("  atom_types: %i\n", atom.ntypes);
  out("  system_size: %2.2lf %2.2lf %2.2lf\n", (double)atom.box.xprd, (double)atom.box.yprd, (double)atom.box.zprd);
  out("  unit_cells: %i %i %i\n", in.nx, in.ny, in.nz);
  out("  density: %lf\n", (double)in.rho);
  out("  force_type: %s\n", in.forcetype == FORCELJ ? "LJ" : "EAM");
  out("  force_cutoff: %lf\n", (double)sim.force->cutforce);
  out("  force_params: %2.2lf %2.2lf\n", (double)sim.force->epsilon[0], (double)sim.force->sigma[0]);
  out("  neighbor_cutoff: %lf\n", (double)neighbor.cutneigh);
  out("  neighbor_type: %i\n", neighbor.halfneigh);
  out("  neighbor_bins: %i %i %i\n", neighbor.nbinx, neighbor.nbiny, neighbor.nbinz);
  out("  neighbor_frequency: %i\n", neighbor.every);
  out("  sort_frequency: %i\n", sim.integrate.sort_every);
  out("  timestep_size: %lf\n", (double)sim.integrate.dt);
  out("  thermo_frequency: %i\n", sim.thermo.nstat);
  out("  ghost_newton: %i\n", neighbor.ghost_newton);
  out("  use_intrinsics: %i\n", sim.force->use_sse);
  out("  safe_exchange: %i\n", sim.comm.do_safeexchange);
  out("  float_size: %i\n\n", (int)sizeof(MMD_float));

  out("\n\nthermodynamic_output:\n");
  const Thermo& th = sim.thermo;
  for (int i = 0; i < th.mstat; i++) {
    const double conserve = (1.5 * th.tmparr[i] + th.engarr[i]) / (1.5 * th.tmparr[0] + th.engarr[0]);
    out("  timestep: %d \n", th.steparr[i]);
    out("      T*:           %15.10g \n", (double)th.tmparr[i]);
    out("      U*:           %15.10g \n", (double)th.engarr[i]);
    out("      P*:           %15.10g \n", (double)th.prsarr[i]);
    out("      Conservation: %15.10g \n", conserve);
  }
  out("\n\n");

  // time split, averaged over ranks
  const double time_total = sim.timer.array[TIME_TOTAL];
  double t[3] = {sim.timer.array[TIME_FORCE], sim.timer.array[TIME_NEIGH], sim.timer.array[TIME_COMM]};
  const double t_other = time_total - (t[0] + t[1] + t[2]);
  const int nsteps = sim.integrate.steps_done;
  out("time:\n");
  out("  total:\n");
  out("    time: %g \n", time_total);
  out("    performance: %10.5e \n", natoms * nsteps / time_total);
  out("    performance_proc: %10.5e \n", natoms * nsteps / time_total / w.nprocs / sim.opt.num_threads);
  double avg[4] = {t[0], t[1], t[2], t_other};
  w.sum(avg, 4);
  for (double& a : avg) a /= w.nprocs;
  out("  force: %g\n", avg[0]);
  out("  neigh: %g\n", avg[1]);
  out("  comm:  %g\n", avg[2]);
  out("  other: %g\n", avg[3]);
  out("\n");

  out("# Timing histograms \n");
  print_stats(out, "Force time:", rank_stats(w, t[0]));
  print_stats(out, "Neigh time:", rank_stats(w, t[1]));
  print_stats(out, "Comm  time:", rank_stats(w, t[2]));
  print_stats(out, "Other time:", rank_stats(w, t_other));
  out("\n");

  print_stats(out, "Nlocal:    ", rank_stats(w, atom.nlocal));
  print_stats(out, "Nghost:    ", rank_stats(w, atom.nghost));
  long long nsw = 0;
  for (int i = 0; i < sim.comm.nswap; i++) nsw += sim.comm.sendnum[i];
  print_stats(out, "Nswaps:    ", rank_stats(w, (double)nsw));
  print_stats(out, "Neighs:    ", rank_stats(w, (double)neighbor.total_neigh));
  double total = (double)neighbor.total_neigh;
  w.sum(&total, 1);
  out("# Total # of neighbors = %g\n", total);
  out("\n");
  if (out.fp) fclose(out.fp);
}

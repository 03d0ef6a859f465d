// TEST INFRASTRUCTURE (tests/test_gpu_dropin.py): INTEGRATION.md section B made real.
//
// A replacement for ONE translation unit of the reference, ref/integrate.cpp, compiled against the reference's own
// headers and linked with the reference's other, unmodified objects (ljs.cpp main, input, setup, atom, neighbor, comm,
// thermo, force_*, output, timer) by oracle/build_dropin.sh.  Integrate::run (ref/integrate.cpp:70-206) hands the time
// loop to the C ABI (include/minimd_b200.h): the reference's main() sets the problem up on the host exactly as before,
// this file uploads it once, lets mmd_run execute the steps on the GPU, and feeds the thermo samples through the
// reference's own Thermo::compute (ref/thermo.cpp:74-115) so that the printed "# Timestep T U P" lines and the YAML
// report come from reference code.  After the loop the atoms are copied back and the host-side ghost / neighbor state
// is rebuilt with the reference's own Comm::exchange / borders and Neighbor::build, so everything main() does after the
// run (ref/ljs.cpp:474-499) works on consistent host arrays.
//
// Nothing here is reference source: only its public class surface is used (plus Neighbor's private bin geometry, reached
// by compiling this one file with `private` opened -- a maintainer would add accessors instead).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mpi.h"
#define private public   // Neighbor's bin geometry / stencil (ref/neighbor.h:72-89) and ForceEAM's tables
#define protected public
#include "neighbor.h"
#include "force_eam.h"
#undef private
#undef protected
#include "integrate.h"
#include "minimd_b200.h"

#define MMD_CHECK(call)                                                             \
  do {                                                                              \
    if ((call) != MMD_OK) {                                                         \
      fprintf(stderr, "minimd_b200: %s failed: %s\n", #call, mmd_last_error());     \
      exit(1);                                                                      \
    }                                                                               \
  } while (0)

Integrate::Integrate() { sort_every = 20; }
Integrate::~Integrate() {}
void Integrate::setup() { dtforce = 0.5 * dt; }

// the two velocity-Verlet halves as separate entry points (ref/integrate.cpp:46-68); unused by run() below, which
// executes whole steps on the device, but part of the class surface
static mmd_ctx* g_ctx = NULL;
void Integrate::initialIntegrate() { if (g_ctx) MMD_CHECK(mmd_integrate_initial(g_ctx, dt, dtforce)); }
void Integrate::finalIntegrate() { if (g_ctx) MMD_CHECK(mmd_integrate_final(g_ctx, dtforce)); }

static void setup_context(Atom& atom, Force* force, Neighbor& neighbor, Comm& comm) {
  int device = 0;
  if (const char* lr = getenv("LOCAL_RANK")) device = atoi(lr);
  MMD_CHECK(mmd_ctx_create(device, (int)sizeof(MMD_float), atom.ntypes, NULL, &g_ctx));
  const double prd[3] = {atom.box.xprd, atom.box.yprd, atom.box.zprd};
  const double lo[3] = {atom.box.xlo, atom.box.ylo, atom.box.zlo};
  const double hi[3] = {atom.box.xhi, atom.box.yhi, atom.box.zhi};
  MMD_CHECK(mmd_atom_set_box(g_ctx, prd, lo, hi));
  mmd_bin_geometry g;
  g.nbinx = neighbor.nbinx; g.nbiny = neighbor.nbiny; g.nbinz = neighbor.nbinz;
  g.mbinx = neighbor.mbinx; g.mbiny = neighbor.mbiny; g.mbinz = neighbor.mbinz;
  g.mbinxlo = neighbor.mbinxlo; g.mbinylo = neighbor.mbinylo; g.mbinzlo = neighbor.mbinzlo;
  g.bininvx = neighbor.bininvx; g.bininvy = neighbor.bininvy; g.bininvz = neighbor.bininvz;
  MMD_CHECK(mmd_neigh_setup(g_ctx, &g, neighbor.stencil, neighbor.nstencil, neighbor.cutneighsq));
  mmd_swap_table t;
  memset(&t, 0, sizeof t);
  t.me = comm.me;
  MPI_Comm_size(MPI_COMM_WORLD, &t.nprocs);
  t.nswap = comm.nswap;
  for (int d = 0; d < 3; d++) {
    t.need[d] = comm.need[d];
    t.procgrid[d] = comm.procgrid[d];
    t.procneigh[d][0] = comm.procneigh[d][0];
    t.procneigh[d][1] = comm.procneigh[d][1];
  }
  for (int w = 0; w < comm.nswap; w++) {
    t.sendproc[w] = comm.sendproc[w]; t.recvproc[w] = comm.recvproc[w];
    t.pbc_any[w] = comm.pbc_any[w];
    t.pbc_flagx[w] = comm.pbc_flagx[w]; t.pbc_flagy[w] = comm.pbc_flagy[w]; t.pbc_flagz[w] = comm.pbc_flagz[w];
    t.slablo[w] = comm.slablo[w]; t.slabhi[w] = comm.slabhi[w];
  }
  MMD_CHECK(mmd_comm_setup(g_ctx, &t));
  if (force->style == FORCELJ) {
    MMD_CHECK(mmd_force_lj_setup(g_ctx, force->cutforcesq, force->sigma6, force->epsilon));
  } else {
    ForceEAM* e = (ForceEAM*)force;
    MMD_CHECK(mmd_force_eam_setup(g_ctx, e->rhor_spline, e->z2r_spline, e->frho_spline, e->nr, e->nrho, e->nr_tot,
                                  e->nrho_tot, e->rdr, e->rdrho, force->cutforcesq));
  }
}

void Integrate::run(Atom& atom, Force* force, Neighbor& neighbor, Comm& comm, Thermo& thermo, Timer& timer) {
  comm.timer = &timer;
  mass = atom.mass;
  dtforce = dtforce / mass;   // as the reference does on entry (ref/integrate.cpp:80-81)
  if (!g_ctx) setup_context(atom, force, neighbor, comm);

  // the state main() has prepared (ref/ljs.cpp:445-459): local atoms; ghosts and lists are rebuilt on the device
  MMD_CHECK(mmd_atom_upload(g_ctx, atom.x, atom.v, atom.type, atom.nlocal, PAD));
  MMD_CHECK(mmd_comm_exchange(g_ctx));
  MMD_CHECK(mmd_comm_borders(g_ctx));
  int maxneighs = neighbor.maxneighs;
  long long total = 0;
  MMD_CHECK(mmd_neigh_build(g_ctx, neighbor.halfneigh, neighbor.ghost_newton, &maxneighs, &total));
  // forces of step 0 (main() computed them on the host; the device needs its own copy before the first half kick)
  if (force->style == FORCELJ) {
    MMD_CHECK(mmd_force_lj_compute(g_ctx, neighbor.halfneigh, neighbor.ghost_newton, 0, NULL, NULL));
    if (neighbor.halfneigh && neighbor.ghost_newton) MMD_CHECK(mmd_comm_reverse_communicate(g_ctx));
  } else {
    MMD_CHECK(mmd_force_eam_compute(g_ctx, neighbor.halfneigh, 0, NULL, NULL));
  }

  mmd_run_params p;
  memset(&p, 0, sizeof p);
  p.total_steps = ntimes;
  p.neigh_every = neighbor.every;
  p.sort_every = sort_every;
  p.thermo_nstat = thermo.nstat;
  p.halfneigh = neighbor.halfneigh;
  p.ghost_newton = neighbor.ghost_newton;
  p.force_style = force->style == FORCELJ ? 0 : 1;
  p.dt = dt;
  p.dtforce = dtforce;
  p.mass = mass;
  int n = 0;
  while (n < ntimes) {
    // up to and including the next thermo step
    int chunk = ntimes - n;
    if (thermo.nstat > 0) chunk = std::min(chunk, thermo.nstat - n % thermo.nstat);
    p.first_step = n;
    p.ntimes = chunk;
    mmd_thermo_sample s[2];
    int ns = 0;
    float ms = 0;
    MMD_CHECK(mmd_run(g_ctx, &p, s, 2, &ns, &ms));
    n += chunk;
    if (thermo.nstat > 0 && n % thermo.nstat == 0 && ns > 0) {
      // Thermo::compute reads Atom::v, Atom::nlocal, Force::eng_vdwl and Force::virial (ref/thermo.cpp:86-91,121-135,183-193)
      int nl = 0, ng = 0;
      MMD_CHECK(mmd_atom_counts(g_ctx, &nl, &ng, NULL));
      while (nl + ng >= atom.nmax) atom.growarray();
      atom.nlocal = nl;
      MMD_CHECK(mmd_atom_download(g_ctx, NULL, atom.v, NULL, NULL, 0, nl, PAD));
      force->eng_vdwl = (MMD_float)s[ns - 1].eng_vdwl;
      force->virial = (MMD_float)s[ns - 1].virial;
      thermo.compute(n, atom, neighbor, force, timer, comm);
    }
  }
  // hand the atoms back and let the reference rebuild its own ghost / list state
  int nl = 0, ng = 0;
  MMD_CHECK(mmd_atom_counts(g_ctx, &nl, &ng, NULL));
  while (nl + ng >= atom.nmax) atom.growarray();
  atom.nlocal = nl;
  atom.nghost = 0;
  MMD_CHECK(mmd_atom_download(g_ctx, atom.x, atom.v, atom.f, atom.type, 0, nl, PAD));
  comm.exchange(atom);
  comm.borders(atom);
  neighbor.build(atom);
}

// extern "C" surface of the host layer (include/minimd_host.h).
#include <stdio.h>
#include <string.h>

#include <string>

#include "minimd_host.h"
#include "force_eam.h"
#include "sim.h"

struct mmd_sim {
  Simulation s;
};

static thread_local std::string g_err;

extern "C" {

int mmd_sim_precision_bytes(void) { return (int)sizeof(MMD_float); }
const char* mmd_sim_last_error(void) { return g_err.c_str(); }

static int create_common(int argc, const char* const* argv, int rank, int nranks, int device, const void* nccl_id128,
                         mmd_reduce_fn reduce, void* user, bool host_only, mmd_sim** out);

int mmd_sim_plan(int argc, const char* const* argv, int rank, int nranks, mmd_reduce_fn reduce, void* user,
                 mmd_sim** out) {
  return create_common(argc, argv, rank, nranks, -1, nullptr, reduce, user, true, out);
}

int mmd_sim_create(int argc, const char* const* argv, int rank, int nranks, int device, const void* nccl_id128,
                   mmd_sim** out) {
  return create_common(argc, argv, rank, nranks, device, nccl_id128, nullptr, nullptr, false, out);
}

static int create_common(int argc, const char* const* argv, int rank, int nranks, int device, const void* nccl_id128,
                         mmd_reduce_fn reduce, void* user, bool host_only, mmd_sim** out) {
  if (!out) {
    g_err = "out is null";
    return 1;
  }
  *out = nullptr;
  Options opt;
  // parse_options skips argv[0]; the embedding API has none
  std::vector<const char*> av;
  av.push_back("minimd_host");
  for (int i = 0; i < argc; i++) av.push_back(argv[i]);
  parse_options((int)av.size(), av.data(), opt);
  if (!opt.unknown.empty()) {
    g_err = "unknown option '" + opt.unknown[0] + "'";
    return 1;
  }
  World w;
  w.me = rank;
  w.nprocs = nranks < 1 ? 1 : nranks;
  w.device = device >= 0 ? device : rank;
  w.reduce_cb = reduce;
  w.reduce_user = user;
  if (host_only) opt.quiet = 1;
  mmd_sim* h = new mmd_sim();
  if (h->s.init(opt, w, (const unsigned char*)nccl_id128, host_only)) {
    g_err = h->s.error.empty() ? std::string(mmd_last_error()) : h->s.error;
    delete h;
    return 1;
  }
  *out = h;
  return 0;
}

int mmd_sim_host_array(mmd_sim* h, const char* name, const void** data, long long* count) {
  if (!h || !name || !data || !count) { g_err = "null argument"; return 1; }
  Simulation& s = h->s;
  const std::string k(name);
  const long long nn = (long long)s.atom->ntypes * s.atom->ntypes;
  ForceEAM* eam = s.force->style == FORCEEAM ? static_cast<ForceEAM*>(s.force) : nullptr;
  if (k == "x") { *data = s.atom->x; *count = (long long)s.atom->nlocal * PAD; }
  else if (k == "v") { *data = s.atom->v; *count = (long long)s.atom->nlocal * PAD; }
  else if (k == "type") { *data = s.atom->type; *count = s.atom->nlocal; }
  else if (k == "stencil") { *data = s.neighbor->stencil.data(); *count = s.neighbor->nstencil; }
  else if (k == "cutneighsq") { *data = s.neighbor->cutneighsq; *count = nn; }
  else if (k == "cutforcesq") { *data = s.force->cutforcesq; *count = nn; }
  else if (k == "epsilon") { *data = s.force->epsilon; *count = nn; }
  else if (k == "sigma6") { *data = s.force->sigma6; *count = nn; }
  else if (eam && k == "rhor_spline") { *data = eam->rhor_spline; *count = nn * eam->nr_tot; }
  else if (eam && k == "z2r_spline") { *data = eam->z2r_spline; *count = nn * eam->nr_tot; }
  else if (eam && k == "frho_spline") { *data = eam->frho_spline; *count = nn * eam->nrho_tot; }
  else { g_err = "unknown array " + k; return 1; }
  return 0;
}

int mmd_sim_bin_geometry(mmd_sim* h, mmd_bin_geometry* g) {
  if (!h || !g) { g_err = "null argument"; return 1; }
  const Neighbor& n = *h->s.neighbor;
  g->nbinx = n.nbinx; g->nbiny = n.nbiny; g->nbinz = n.nbinz;
  g->mbinx = n.mbinx; g->mbiny = n.mbiny; g->mbinz = n.mbinz;
  g->mbinxlo = n.mbinxlo; g->mbinylo = n.mbinylo; g->mbinzlo = n.mbinzlo;
  g->bininvx = (double)n.bininvx; g->bininvy = (double)n.bininvy; g->bininvz = (double)n.bininvz;
  return 0;
}

int mmd_sim_swap_table(mmd_sim* h, mmd_swap_table* t) {
  if (!h || !t) { g_err = "null argument"; return 1; }
  *t = h->s.comm.table;
  return 0;
}

int mmd_sim_run(mmd_sim* h, int nsteps, double* device_ms) {
  if (!h) { g_err = "null simulation"; return 1; }
  if (!h->s.ctx) { g_err = "this simulation was only planned (mmd_sim_plan): no device context"; return 1; }
  h->s.run(nsteps);
  if (device_ms) *device_ms = h->s.integrate.device_ms;
  return 0;
}

int mmd_sim_finish(mmd_sim* h) {
  if (!h) { g_err = "null simulation"; return 1; }
  if (!h->s.ctx) { g_err = "this simulation was only planned (mmd_sim_plan): no device context"; return 1; }
  return h->s.finish();
}

int mmd_sim_destroy(mmd_sim* h) {
  delete h;
  return 0;
}

int mmd_sim_thermo(mmd_sim* h, int max, int* step, double* T, double* U, double* P) {
  if (!h) return 0;
  const Thermo& t = h->s.thermo;
  for (int i = 0; i < t.mstat && i < max; i++) {
    if (step) step[i] = t.steparr[i];
    if (T) T[i] = (double)t.tmparr[i];
    if (U) U[i] = (double)t.engarr[i];
    if (P) P[i] = (double)t.prsarr[i];
  }
  return t.mstat;
}

int mmd_sim_get_int(mmd_sim* h, const char* key, long long* v) {
  if (!h || !key || !v) { g_err = "null argument"; return 1; }
  Simulation& s = h->s;
  const std::string k(key);
  if (k == "natoms") *v = s.atom->natoms;
  else if (k == "nlocal") *v = s.atom->nlocal;
  else if (k == "nghost") *v = s.atom->nghost;
  else if (k == "nswap") *v = s.comm.nswap;
  else if (k == "maxneighs") *v = s.neighbor->maxneighs;
  else if (k == "total_neigh") *v = s.neighbor->total_neigh;
  else if (k == "neigh_builds") *v = s.neighbor->ncalls;
  else if (k == "mbins") *v = s.neighbor->mbins;
  else if (k == "nstencil") *v = s.neighbor->nstencil;
  else if (k == "nbinx") *v = s.neighbor->nbinx;
  else if (k == "nbiny") *v = s.neighbor->nbiny;
  else if (k == "nbinz") *v = s.neighbor->nbinz;
  else if (k == "halfneigh") *v = s.neighbor->halfneigh;
  else if (k == "ghost_newton") *v = s.neighbor->ghost_newton;
  else if (k == "steps_done") *v = s.integrate.steps_done;
  else if (k == "ntimes") *v = s.integrate.ntimes;
  else if (k == "sort_every") *v = s.integrate.sort_every;
  else if (k == "neigh_every") *v = s.neighbor->every;
  else if (k == "thermo_nstat") *v = s.thermo.nstat;
  else if (k == "procgrid0") *v = s.comm.procgrid[0];
  else if (k == "procgrid1") *v = s.comm.procgrid[1];
  else if (k == "procgrid2") *v = s.comm.procgrid[2];
  else if (k == "force_style") *v = s.force->style == FORCELJ ? 0 : 1;
  else if (k == "ntypes") *v = s.atom->ntypes;
  else if (k == "me") *v = s.world.me;
  else if (k == "nprocs") *v = s.world.nprocs;
  else if (k == "myloc0") *v = s.comm.myloc[0];
  else if (k == "myloc1") *v = s.comm.myloc[1];
  else if (k == "myloc2") *v = s.comm.myloc[2];
  else if (k == "nr" && s.force->style == FORCEEAM) *v = static_cast<ForceEAM*>(s.force)->nr;
  else if (k == "nrho" && s.force->style == FORCEEAM) *v = static_cast<ForceEAM*>(s.force)->nrho;
  else if (k == "nr_tot" && s.force->style == FORCEEAM) *v = static_cast<ForceEAM*>(s.force)->nr_tot;
  else if (k == "nrho_tot" && s.force->style == FORCEEAM) *v = static_cast<ForceEAM*>(s.force)->nrho_tot;
  else { g_err = "unknown key " + k; return 1; }
  return 0;
}

int mmd_sim_get_real(mmd_sim* h, const char* key, double* v) {
  if (!h || !key || !v) { g_err = "null argument"; return 1; }
  Simulation& s = h->s;
  const std::string k(key);
  if (k == "dt") *v = s.integrate.dt;
  else if (k == "dtforce") *v = s.integrate.dtforce;
  else if (k == "mass") *v = s.atom->mass;
  else if (k == "t_scale") *v = s.thermo.t_scale;
  else if (k == "e_scale") *v = s.thermo.e_scale;
  else if (k == "p_scale") *v = s.thermo.p_scale;
  else if (k == "dof_boltz") *v = s.thermo.dof_boltz;
  else if (k == "xprd") *v = s.atom->box.xprd;
  else if (k == "yprd") *v = s.atom->box.yprd;
  else if (k == "zprd") *v = s.atom->box.zprd;
  else if (k == "cutneigh") *v = s.neighbor->cutneigh;
  else if (k == "cutforce") *v = s.force->cutforce;
  else if (k == "t_total") *v = s.timer.array[TIME_TOTAL];
  else if (k == "t_force") *v = s.timer.array[TIME_FORCE];
  else if (k == "t_neigh") *v = s.timer.array[TIME_NEIGH];
  else if (k == "t_comm") *v = s.timer.array[TIME_COMM];
  else if (k == "device_ms_total") *v = s.device_ms_total;
  else if (k == "xlo") *v = s.atom->box.xlo;
  else if (k == "xhi") *v = s.atom->box.xhi;
  else if (k == "ylo") *v = s.atom->box.ylo;
  else if (k == "yhi") *v = s.atom->box.yhi;
  else if (k == "zlo") *v = s.atom->box.zlo;
  else if (k == "zhi") *v = s.atom->box.zhi;
  else if (k == "rdr" && s.force->style == FORCEEAM) *v = static_cast<ForceEAM*>(s.force)->rdr;
  else if (k == "rdrho" && s.force->style == FORCEEAM) *v = static_cast<ForceEAM*>(s.force)->rdrho;
  else { g_err = "unknown key " + k; return 1; }
  return 0;
}

mmd_ctx* mmd_sim_ctx(mmd_sim* h) { return h ? h->s.ctx : nullptr; }

int mmd_sim_run_params(mmd_sim* h, int nsteps, mmd_run_params* p) {
  if (!h || !p) { g_err = "null argument"; return 1; }
  Simulation& s = h->s;
  // dtforce is divided by the mass on the first Integrate::run (ref/integrate.cpp:80-81)
  const bool folded = s.integrate.steps_done > 0;
  p->ntimes = nsteps;
  p->first_step = s.integrate.steps_done;
  p->total_steps = s.integrate.ntimes;
  p->neigh_every = s.neighbor->every;
  p->sort_every = s.integrate.sort_every;
  p->thermo_nstat = s.thermo.nstat;
  p->halfneigh = s.neighbor->halfneigh != 0;
  p->ghost_newton = s.neighbor->ghost_newton;
  p->force_style = s.force->style == FORCELJ ? 0 : 1;
  p->dt = (double)s.integrate.dt;
  p->dtforce = (double)(folded ? s.integrate.dtforce : (MMD_float)(s.integrate.dtforce / s.atom->mass));
  p->mass = (double)s.atom->mass;
  return 0;
}

}  // extern "C"

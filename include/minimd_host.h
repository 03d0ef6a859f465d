/* =====================================================================================
 * minimd_host.h -- C ABI of the host layer (minimd_b200/csrc/host): a whole miniMD run as an
 * object, for embedders that want the reference's main() (ref/ljs.cpp:61-504) without a
 * process boundary -- bench.py and the parity tests bind it with ctypes.
 *
 * The library is built twice, libminimd_host_f64.so (PRECISION=2) and libminimd_host_f32.so
 * (PRECISION=1), exactly as the reference is compiled per precision (ref/types.h:61-74).
 * Both link libminimd_b200.so (include/minimd_b200.h), which does all the device work.
 * Status: 0 = ok, nonzero = error (message via mmd_sim_last_error()).
 * ===================================================================================== */
#ifndef MINIMD_HOST_H
#define MINIMD_HOST_H

#include "minimd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mmd_sim mmd_sim;

/* sizeof(MMD_float) of this build: 8 or 4 */
int mmd_sim_precision_bytes(void);
const char* mmd_sim_last_error(void);

/* main() up to and including the step-0 thermo record (ref/ljs.cpp:263-468).
 * argv: the reference's command line WITHOUT argv[0] (-i in.lj.miniMD -s 80 --half_neigh 1 ...).
 * rank/nranks: this process's place in the run (one rank per GPU); nccl_id128: the id from
 * mmd_comm_nccl_unique_id() of rank 0, NULL when nranks == 1.  device < 0: use `rank`. */
int mmd_sim_create(int argc, const char* const* argv, int rank, int nranks, int device, const void* nccl_id128,
                   mmd_sim** out);
/* Host-only planning: the same code path as mmd_sim_create up to create_velocity (input, CLI, box,
 * Comm::setup decomposition + swap table, Neighbor::setup bins + stencil, Force::setup tables,
 * create_atoms, create_velocity) WITHOUT a GPU.  Nothing can be run on a planned simulation; it
 * exists so the host logic can be inspected and tested anywhere.  reduce (may be NULL when
 * nranks == 1) performs the cross-rank sums/maxima (op 0 sum, 1 max) the setup needs. */
typedef void (*mmd_reduce_fn)(double* values, int n, int op, void* user);
int mmd_sim_plan(int argc, const char* const* argv, int rank, int nranks, mmd_reduce_fn reduce, void* user,
                 mmd_sim** out);
/* views into the host-side state (valid until the simulation is destroyed):
 * "x" "v" (MMD_float[nlocal*PAD]) "type" (int[nlocal]) "stencil" (int[nstencil])
 * "cutneighsq" "cutforcesq" "epsilon" "sigma6" (MMD_float[ntypes^2])
 * "rhor_spline" "z2r_spline" "frho_spline" (MMD_float, EAM only).  count = number of elements. */
int mmd_sim_host_array(mmd_sim* sim, const char* name, const void** data, long long* count);
int mmd_sim_bin_geometry(mmd_sim* sim, mmd_bin_geometry* out);
int mmd_sim_swap_table(mmd_sim* sim, mmd_swap_table* out);

/* Integrate::run (ref/integrate.cpp:70-207) for nsteps more steps (<0: all that remain of in.ntimes;
 * running past in.ntimes is allowed and continues the same cadence).  device_ms (may be NULL):
 * CUDA-event time of the loop. */
int mmd_sim_run(mmd_sim* sim, int nsteps, double* device_ms);
/* final force + thermo record + PERF_SUMMARY + optional YAML (ref/ljs.cpp:474-498) */
int mmd_sim_finish(mmd_sim* sim);
int mmd_sim_destroy(mmd_sim* sim);

/* thermo records so far (Thermo::steparr/tmparr/engarr/prsarr): returns the count; copies up to max */
int mmd_sim_thermo(mmd_sim* sim, int max, int* step, double* T, double* U, double* P);
/* named scalars: "natoms" "nlocal" "nghost" "nswap" "maxneighs" "total_neigh" "neigh_builds" "mbins"
 * "nstencil" "nbinx" "nbiny" "nbinz" "halfneigh" "ghost_newton" "steps_done" "ntimes" "sort_every"
 * "neigh_every" "thermo_nstat" "procgrid0..2" / "dt" "dtforce" "mass" "t_scale" "e_scale" "p_scale"
 * "dof_boltz" "xprd" "yprd" "zprd" "cutneigh" "cutforce" "t_total" "t_force" "t_neigh" "t_comm" */
int mmd_sim_get_int(mmd_sim* sim, const char* key, long long* value);
int mmd_sim_get_real(mmd_sim* sim, const char* key, double* value);
/* the device context of this simulation (owned by the simulation) */
mmd_ctx* mmd_sim_ctx(mmd_sim* sim);
/* mmd_run parameters that continue this simulation for nsteps (what mmd_sim_run passes down) */
int mmd_sim_run_params(mmd_sim* sim, int nsteps, mmd_run_params* out);

#ifdef __cplusplus
}
#endif
#endif /* MINIMD_HOST_H */

/* =====================================================================================
 * minimd_b200.h -- C ABI of the B200 (sm_100a) implementation of the miniMD hot path.
 *
 * Plain C: opaque handle, raw pointers, sizes, int status codes.  No C++/torch types.
 * Every entry point names the reference interface it replaces (paths relative to the
 * Mantevo/miniMD tree, `ref/` variant).  The reference has no FFI of its own; its seams are
 * the public members of Atom / Neighbor / Force / Integrate / Comm and
 * mpi-spec's raw-pointer hook  bool compute_lj(const ForceLJ&, half_neigh, ghost_newton)
 * (mpi-spec/force_lj.h:89, body mpi-spec/force_lj_custom.cpp:17-30).
 *
 * Model: an `mmd_ctx` owns ALL device memory of one rank (one GPU) and is the single source
 * of truth between calls; host arrays use the reference's layout (AoS, stride PAD = 3 or 4,
 * MMD_float = float|double chosen per context, ref/types.h:61-81) and cross the boundary only
 * in the explicit upload/download calls.  Device layout is private.
 *
 * Status: 0 = ok; nonzero = error, message via mmd_last_error().  Nothing throws or exits
 * across this boundary (the reference printf()s and exit(0)s, e.g. ref/ljs.cpp:105-108).
 * Threading: one host thread per context (the reference's orphaned-OpenMP calling convention,
 * ref/integrate.cpp:84-206, does not apply: a GPU backend runs the time loop single-threaded,
 * as kokkos/integrate.cpp:85-185 does).
 * ===================================================================================== */
#ifndef MINIMD_B200_H
#define MINIMD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MMD_ABI_VERSION 1
#define MMD_MAX_SWAPS 32

typedef struct mmd_ctx mmd_ctx;

/* ---- status ------------------------------------------------------------------------- */
enum {
  MMD_OK = 0,
  MMD_ERR_ARG = 1,      /* bad argument / call order */
  MMD_ERR_CUDA = 2,     /* CUDA runtime error (message has the CUDA string) */
  MMD_ERR_NODEVICE = 3, /* no CUDA device: the product path never falls back to the CPU */
  MMD_ERR_STATE = 4,    /* required setup call missing */
  MMD_ERR_NCCL = 5
};
const char* mmd_last_error(void);
int mmd_abi_version(void);
/* number of visible CUDA devices (0 on a CPU-only box); never fails */
int mmd_device_count(void);

/* ---- context ------------------------------------------------------------------------ */
/* precision_bytes = sizeof(MMD_float) (ref/types.h:61-74): 8 or 4.
 * stream: a cudaStream_t to launch on (e.g. torch's current stream) or NULL to create one. */
int mmd_ctx_create(int device, int precision_bytes, int ntypes, void* stream, mmd_ctx** out);
int mmd_ctx_destroy(mmd_ctx* ctx);
int mmd_ctx_sync(mmd_ctx* ctx);
/* the cudaStream_t all kernels of this context are launched on */
void* mmd_ctx_stream(mmd_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
long long mmd_ctx_launches(mmd_ctx* ctx);

/* ---- Atom (ref/atom.h:47-106) ------------------------------------------------------- */
/* Box (ref/atom.h:40-45): periodic lengths and this rank's sub-box bounds. */
int mmd_atom_set_box(mmd_ctx* ctx, const double prd[3], const double lo[3], const double hi[3]);
/* Replace the local atoms: x,v are MMD_float[nlocal*pad], type int[nlocal] (Atom::x/v/type).
 * Ghosts are dropped (nghost=0) -- exactly the state after create_atoms (ref/setup.cpp:315). */
int mmd_atom_upload(mmd_ctx* ctx, const void* x, const void* v, const int* type, int nlocal, int pad);
/* Declare the last (uploaded - nlocal) atoms of the preceding mmd_atom_upload to be ghosts: the host keeps Comm::borders
 * (the mpi-spec compute_lj seam, mpi-spec/force_lj_custom.cpp:17-30, hands over x of nlocal + nghost atoms).  Ghosts made
 * this way have no send lists on the device (their swap counts are zero): a forward / reverse halo does not touch them. */
int mmd_atom_split(mmd_ctx* ctx, int nlocal);
/* Overwrite x and/or v of atoms [first, first+count) without touching counts/lists
 * (the per-step H2D of a host-resident Atom). NULL pointers are skipped. */
int mmd_atom_update(mmd_ctx* ctx, const void* x, const void* v, int first, int count, int pad);
/* Copy atoms [first, first+count) back; any of x,v,f,type may be NULL.  first+count may
 * extend over ghosts for x,f,type (v is defined for locals only). */
int mmd_atom_download(mmd_ctx* ctx, void* x, void* v, void* f, int* type, int first, int count, int pad);
int mmd_atom_counts(mmd_ctx* ctx, int* nlocal, int* nghost, int* nmax);
/* Atom::pbc (ref/atom.cpp:106-122) */
int mmd_atom_pbc(mmd_ctx* ctx);
/* Atom::sort (ref/atom.cpp:355-421): reorder local x,v,type by bin, ascending old index
 * inside a bin (the reference's single-thread order). Requires mmd_neigh_setup. */
int mmd_atom_sort(mmd_ctx* ctx);

/* ---- Neighbor (ref/neighbor.h:38-92) ------------------------------------------------ */
/* Bin geometry + stencil as computed by Neighbor::setup (ref/neighbor.cpp:318-452) on the
 * host; values are passed as doubles holding the exact MMD_float values. */
typedef struct {
  int nbinx, nbiny, nbinz;          /* global bins */
  int mbinx, mbiny, mbinz;          /* local bins incl. ghost layers */
  int mbinxlo, mbinylo, mbinzlo;
  double bininvx, bininvy, bininvz; /* 1/binsize, already rounded to MMD_float */
} mmd_bin_geometry;
/* cutneighsq: MMD_float[ntypes*ntypes] (Neighbor::cutneighsq). */
int mmd_neigh_setup(mmd_ctx* ctx, const mmd_bin_geometry* geo, const int* stencil, int nstencil,
                    const void* cutneighsq);
/* Neighbor::binatoms (ref/neighbor.cpp:215-268). count<0 => nlocal+nghost.
 * *atoms_per_bin: in = current row width, out = the width the reference's doubling protocol
 * (:257-261) would have ended with.  max_count (may be NULL) = fullest bin. */
int mmd_neigh_binatoms(mmd_ctx* ctx, int count, int* atoms_per_bin, int* max_count);
/* Neighbor::build (ref/neighbor.cpp:79-213) incl. binatoms and the maxneighs*1.2 resize
 * protocol (:186-208).  *maxneighs in/out (Neighbor::maxneighs); total = sum(numneigh). */
int mmd_neigh_build(mmd_ctx* ctx, int halfneigh, int ghost_newton, int* maxneighs, long long* total);
/* numneigh int[nlocal], neighbors int[nlocal*maxneighs] row-major (ref/force_lj.cpp:207). */
int mmd_neigh_download(mmd_ctx* ctx, int* numneigh, int* neighbors, int nrows, int maxneighs);
/* Upload a host-built list instead (the mpi-spec compute_lj seam passes host lists). */
int mmd_neigh_upload(mmd_ctx* ctx, const int* numneigh, const int* neighbors, int nrows, int maxneighs);
/* Neighbor::bincount int[mbins] and Neighbor::bins int[mbins*atoms_per_bin] in the reference's
 * fixed-width row layout (either may be NULL). */
int mmd_neigh_bins_download(mmd_ctx* ctx, int* bincount, int* bins, int atoms_per_bin);
/* per-atom bin index (coord2bin, ref/neighbor.cpp:274-300, including its +1) of the last binning */
int mmd_neigh_atom_bins_download(mmd_ctx* ctx, int* bin_of_atom, int count);

/* ---- Force (ref/force.h:40-70) ------------------------------------------------------ */
/* ForceLJ parameters (ref/force_lj.cpp:41-69): MMD_float[ntypes*ntypes] each. */
int mmd_force_lj_setup(mmd_ctx* ctx, const void* cutforcesq, const void* sigma6, const void* epsilon);
/* ForceLJ::compute (ref/force_lj.cpp:72-113) -> compute_halfneigh<EV,GN> (:185-263) or
 * compute_fullneigh<EV> (:366-449).  eng_vdwl/virial: MMD_float*, written only if evflag
 * (Force::eng_vdwl, Force::virial); pass NULL to leave them on the device. */
int mmd_force_lj_compute(mmd_ctx* ctx, int halfneigh, int ghost_newton, int evflag, void* eng_vdwl,
                         void* virial);
/* ForceEAM tables after array2spline (ref/force_eam.cpp:732-761): MMD_float arrays of
 * ntypes^2*nr_tot (rhor, z2r) and ntypes^2*nrho_tot (frho); cutforcesq MMD_float[ntypes^2]. */
int mmd_force_eam_setup(mmd_ctx* ctx, const void* rhor_spline, const void* z2r_spline, const void* frho_spline,
                        int nr, int nrho, int nr_tot, int nrho_tot, double rdr, double rdrho,
                        const void* cutforcesq);
/* ForceEAM::compute (ref/force_eam.cpp:82-91) -> compute_halfneigh (:94-270) or
 * compute_fullneigh (:274-449) including the fp halo (ForceEAM::communicate :851-887). */
int mmd_force_eam_compute(mmd_ctx* ctx, int halfneigh, int evflag, void* eng_vdwl, void* virial);

/* ---- Integrate (ref/integrate.h) + Thermo ------------------------------------------- */
/* Integrate::initialIntegrate (ref/integrate.cpp:46-57): v += dtforce*f; x += dt*v. */
int mmd_integrate_initial(mmd_ctx* ctx, double dt, double dtforce);
/* Integrate::finalIntegrate (ref/integrate.cpp:59-68): v += dtforce*f. */
int mmd_integrate_final(mmd_ctx* ctx, double dtforce);
/* The reduction of Thermo::temperature (ref/thermo.cpp:140-166): *sum_mv2 = sum_i (v.v)*mass,
 * returned as double; the caller applies t_scale. */
int mmd_thermo_sum_mv2(mmd_ctx* ctx, double mass, double* sum_mv2);

/* ---- Comm (ref/comm.h) -------------------------------------------------------------- */
/* Swap table from Comm::setup (ref/comm.cpp:60-272). sendproc/recvproc == me => self swap. */
typedef struct {
  int me, nprocs;
  int nswap;
  int need[3];
  int procgrid[3];
  int procneigh[3][2];
  int sendproc[MMD_MAX_SWAPS], recvproc[MMD_MAX_SWAPS];
  int pbc_any[MMD_MAX_SWAPS], pbc_flagx[MMD_MAX_SWAPS], pbc_flagy[MMD_MAX_SWAPS], pbc_flagz[MMD_MAX_SWAPS];
  double slablo[MMD_MAX_SWAPS], slabhi[MMD_MAX_SWAPS];
} mmd_swap_table;
int mmd_comm_setup(mmd_ctx* ctx, const mmd_swap_table* table);
/* Multi-GPU transport: one NCCL communicator over all ranks (one rank per GPU).
 * mmd_comm_nccl_unique_id fills a 128-byte id on rank 0; the launcher broadcasts it. */
int mmd_comm_nccl_unique_id(void* id128);
int mmd_comm_nccl_init(mmd_ctx* ctx, const void* id128, int rank, int nranks);
/* Comm::exchange (ref/comm.cpp:364-597): pbc + migrate atoms that left the sub-box. */
int mmd_comm_exchange(mmd_ctx* ctx);
/* Comm::borders (ref/comm.cpp:700-883): rebuild ghost atoms and per-swap send lists. */
int mmd_comm_borders(mmd_ctx* ctx);
/* Comm::communicate (ref/comm.cpp:276-317): forward halo of x. */
int mmd_comm_communicate(mmd_ctx* ctx);
/* Comm::reverse_communicate (ref/comm.cpp:321-355): reverse halo of f. */
int mmd_comm_reverse_communicate(mmd_ctx* ctx);
/* Comm::sendnum/recvnum/firstrecv after borders (int[nswap] each; any may be NULL). */
int mmd_comm_swap_counts(mmd_ctx* ctx, int* sendnum, int* recvnum, int* firstrecv);
/* Comm::sendlist[iswap] (int[sendnum[iswap]]). */
int mmd_comm_sendlist_download(mmd_ctx* ctx, int iswap, int* list, int count);
/* sum / max over ranks of n doubles (Thermo's MPI_Allreduce, ref/thermo.cpp:131,168,188);
 * identity on one rank. op: 0 sum, 1 max; n <= 16. */
int mmd_comm_allreduce(mmd_ctx* ctx, double* values, int n, int op);

/* ---- fused time loop (Integrate::run, ref/integrate.cpp:70-207) ---------------------- */
typedef struct {
  int ntimes;          /* steps to run in this call */
  int first_step;      /* global index of the first step (n in ref's loop), for cadence */
  int total_steps;     /* Integrate::ntimes of the whole run (next_sort init, :86) */
  int neigh_every;     /* Neighbor::every */
  int sort_every;      /* Integrate::sort_every (0 = never) */
  int thermo_nstat;    /* Thermo::nstat (0 = never inside the loop) */
  int halfneigh, ghost_newton;
  int force_style;     /* 0 LJ, 1 EAM */
  double dt, dtforce;  /* dtforce already divided by mass (:81) */
  double mass;
} mmd_run_params;
/* One record per thermo step inside the loop: raw reductions, the caller applies
 * Thermo's scalings (ref/thermo.cpp:119-194). */
typedef struct {
  int step;
  double sum_mv2, eng_vdwl, virial;
} mmd_thermo_sample;
/* Runs params->ntimes steps entirely on the device (launch-only host loop).  samples[max_samples] receives the thermo steps.
 * elapsed_ms (may be NULL): CUDA-event time of the loop on the context's stream. */
int mmd_run(mmd_ctx* ctx, const mmd_run_params* params, mmd_thermo_sample* samples, int max_samples,
            int* nsamples, float* elapsed_ms);

/* Device-time split of mmd_run, the analogue of the reference's Timer buckets (ref/timer.h:35-40:
 * TIME_COMM / TIME_FORCE / TIME_NEIGH; "integrate" is what the reference reports as t_other).
 * Enabled with mmd_set_option(ctx, "phase_timing", 1); measured with CUDA events on the context's
 * stream.  ms[MMD_NPHASE] / calls[MMD_NPHASE] accumulate over mmd_run calls until reset != 0. */
enum { MMD_PHASE_INTEGRATE = 0, MMD_PHASE_COMM = 1, MMD_PHASE_NEIGH = 2, MMD_PHASE_FORCE = 3, MMD_PHASE_OTHER = 4,
       MMD_NPHASE = 5 };
int mmd_run_phase_times(mmd_ctx* ctx, double* ms, long long* calls, int reset);

/* ---- introspection ------------------------------------------------------------------- */
/* Named integer/real queries ("nlocal", "nghost", "maxneighs", "mbins", "total_neigh",
 * "neigh_builds", "nswap", ...); returns MMD_ERR_ARG for unknown keys. */
int mmd_query_int(mmd_ctx* ctx, const char* key, long long* value);
/* Run-time switches (all ranks must use the same values):
 *   "tile_lists" (1)   neighbor lists as 16-bit tile-local rows + shared-memory force kernels (LJ); 0 = classic rows of
 *                      global ids.  Takes effect at the next mmd_neigh_build.  mmd_neigh_download returns the reference's
 *                      rows in either case.
 *   "tile_dealt" (1)   LJ on tile lists: the force kernel walks a bank-dealt copy of the rows (a quarter warp per atom,
 *                      halo window staged by bulk asynchronous copies from a slot-ordered position mirror); 0 = one lane
 *                      pair per row, window staged by an indexed gather.  Takes effect at the next mmd_neigh_build.
 *   "tile_xsort" (1)   tile lists: every bin of the windows' private slot map sorted by x + interval build; 0 = windows in
 *                      CSR order + per-bin candidate-table build
 *   "tile_build2" (1)  0 = tile rows from the warp-per-bin build (no shared-memory window; testing / odd geometries)
 *   "tile_pair_build" (1)  interval build: two atoms of a bin share one sweep over the union of their candidate intervals
 *   "tile_lane_build" (0)  interval build with one lane per atom instead of one warp per atom (measured slower)
 *   "tile_eam" (0)     tile-resident lists for the EAM force too (measured slower than the classic kernels)
 *   "force_nonuniform" (0)  testing: use the per-type parameter-table kernels even when all type pairs are equal
 *   "fuse_integrate" (1)  mmd_run: finalIntegrate(n) + initialIntegrate(n+1) in one kernel
 *   "fuse_force" (1)      mmd_run, tile lists: ... and both inside the force kernel's epilogue
 *   "fuse_halo" (1)       one rank: forward halo in one launch (ghosts resolved to their local source)
 *   "fuse_ghosts" (0)     one rank, dealt LJ lists: the fused force kernel also writes every atom's periodic images, so
 *                         the forward halo of the next step needs no launch at all (measured slower: off by default)
 *   "graph_steps" (1)     mmd_run on one rank: two consecutive plain steps (forward halo + fused force/Verlet kernel, no rebuild,
 *                         no thermo) are captured once per neighbor list as a CUDA graph and replayed for the steps up to the
 *                         next rebuild (ref/integrate.cpp:88-205 runs the same sequence every step).  0: launch every step.
 *   "p2p_halo" (1)        several ranks: forward halo over CUDA-IPC peer windows; 0 = NCCL send/recv
 *   "split_force" (1)     several ranks, dealt LJ lists: tiles without ghosts in their halo window run on a second stream
 *                         while the forward halo of the step is in flight; boundary tiles follow the halo
 *   "lj_threads_per_atom" (0 = auto), "eam_threads_per_atom" (8): lanes per atom of the classic kernels
 *   "phase_timing" (0)    per-phase CUDA-event timing of mmd_run
 *   "kernel_profile" (0)  diagnostic builds (-DMMD_KERNEL_PROFILE) only: the LJ tile force kernels sum clock64 differences
 *                         (staging, CTA life, CTA count), read back with the queries "stage_clocks", "cta_clocks", "cta_count"
 * Queries added by these paths: "list_tile", "list_dealt", "list_xsorted", "tile_ok", "tile_builds", "tile_fallbacks",
 * "tile_max_halo", "tile_max_full", "tile_row_capacity", "tile_dealt_capacity", "tile_count", "tile_interior",
 * "tile_boundary", "split_steps", "fused_halo_steps", "graph_captures", "graph_replays", "p2p_active", "p2p_calls". */
int mmd_set_option(mmd_ctx* ctx, const char* key, long long value);

#ifdef __cplusplus
}
#endif
#endif /* MINIMD_B200_H */

// C-ABI implementation (include/minimd_b200.h): the device context that owns all HBM state of
// one rank and the launch logic behind every entry point.  sm_100a only; no CPU fallback --
// without a CUDA device mmd_ctx_create fails with MMD_ERR_NODEVICE.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <set>
#include <string>
#include <vector>

#include "../../include/minimd_b200.h"
#include "common.cuh"
#include "force_eam_kernels.cuh"
#include "force_lj_kernels.cuh"
#include "integrate_comm_kernels.cuh"
#include "neighbor_kernels.cuh"
#include "scan.cuh"
#include "tile_kernels.cuh"
#include "tile_dealt_kernels.cuh"
#include "tile_build_lane.cuh"
#include "tile_eam_dealt.cuh"

#ifdef MMD_WITH_NCCL
#include <nccl.h>
#endif

using namespace mmd;

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return set_err(MMD_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                                          cudaGetErrorString(e_));                                 \
  } while (0)
#define MM(call)                  \
  do {                            \
    int r_ = (call);              \
    if (r_ != MMD_OK) return r_;  \
  } while (0)
// kernel launch with accounting + immediate launch-error check (template kernels are passed
// parenthesised: LAUNCH(c, (kernel<T, 1>), grid, block, args...))
struct mmd_ctx;
static int launch_fail(int line, cudaError_t e);
template <class C, class K, class... Args>
static inline int launch_impl(int line, C* ctx, K kern, int grid, int block, Args... args) {
  if (grid <= 0) return MMD_OK;
  kern<<<grid, block, 0, ctx->stream>>>(args...);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return launch_fail(line, e);
  return MMD_OK;
}
#define LAUNCH(ctx, kern, grid, block, ...) MM(launch_impl(__LINE__, ctx, kern, grid, block, __VA_ARGS__))
template <class C, class K, class... Args>
static inline int launch_smem_impl(int line, C* ctx, K kern, int grid, int block, size_t smem, Args... args) {
  if (grid <= 0) return MMD_OK;
  kern<<<grid, block, smem, ctx->stream>>>(args...);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return launch_fail(line, e);
  return MMD_OK;
}
#define LAUNCH_SMEM(ctx, kern, grid, block, smem, ...) MM(launch_smem_impl(__LINE__, ctx, kern, grid, block, smem, __VA_ARGS__))
template <class C, class K, class... Args>
static inline int launch_on_impl(int line, C* ctx, cudaStream_t st, K kern, int grid, int block, size_t smem, Args... args) {
  if (grid <= 0) return MMD_OK;
  kern<<<grid, block, smem, st>>>(args...);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return launch_fail(line, e);
  return MMD_OK;
}
#define LAUNCH_ON(ctx, st, kern, grid, block, smem, ...) MM(launch_on_impl(__LINE__, ctx, st, kern, grid, block, smem, __VA_ARGS__))

#ifdef MMD_WITH_NCCL
#define NC(call)                                                                                  \
  do {                                                                                            \
    ncclResult_t r_ = (call);                                                                     \
    if (r_ != ncclSuccess) return set_err(MMD_ERR_NCCL, "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                                          ncclGetErrorString(r_));                                \
  } while (0)
#endif

// ---------------------------------------------------------------------------------------
// device buffer with geometric growth (optionally preserving contents)
// ---------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int reserve(size_t need, cudaStream_t s, size_t keep_bytes = 0, double slack = 1.0) {
    if (need <= bytes) return MMD_OK;
    size_t nb = (size_t)(need * slack) + 256;
    void* q = nullptr;
    CU(cudaMalloc(&q, nb));
    if (p && keep_bytes) CU(cudaMemcpyAsync(q, p, std::min(keep_bytes, bytes), cudaMemcpyDeviceToDevice, s));
    if (p) {
      CU(cudaStreamSynchronize(s));
      CU(cudaFree(p));
    }
    p = q;
    bytes = nb;
    return MMD_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class U> U* as() const { return reinterpret_cast<U*>(p); }
};

struct SwapState {
  DevBuf list;  // Comm::sendlist[iswap]
  int sendnum = 0, recvnum = 0, firstrecv = 0;
};

struct mmd_ctx {
  int device = 0;
  int prec = 8;  // sizeof(MMD_float)
  int ntypes = 1;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  long long launches = 0;

  // Atom
  int nlocal = 0, nghost = 0;
  int cap = 0;  // atoms the per-atom arrays can hold
  DevBuf x, v, f, x_alt, v_alt;
  double prd[3] = {0, 0, 0}, lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  bool have_box = false;
  DevBuf stage_a, stage_b, stage_c, stage_i;  // host<->device AoS staging

  // Neighbor
  bool have_geo = false;
  mmd_bin_geometry geo;
  int mbins = 0;
  std::vector<int> stencil;
  std::vector<double> h_cutneighsq;
  int nruns = 0;
  DevBuf runs, cutneighsq;
  DevBuf atom_bin, bin_atoms, bincount, bin_start, cursor;
  DevBuf tile_sums;
  int binned_count = 0;  // atoms covered by the last binning
  int max_bin_seen = 0;  // running max bin occupancy (drives the reference's atoms_per_bin)
  DevBuf numneigh, neighbors;
  int maxneighs = 100;  // Neighbor::maxneighs (ref/neighbor.cpp:48)
  int neigh_stride = 104;
  int neigh_rows = 0;
  long long total_neigh = 0;
  int neigh_builds = 0, neigh_resizes = 0;
  int list_half = -1, list_gn = -1;

  // tile-resident lists (tile_kernels.cuh): 16-bit tile-local rows + shared-memory force kernels
  bool tile_enable = true;  // option "tile_lists"
  bool tile_eam = true;     // option "tile_eam": EAM on tile-resident, bank-dealt lists (tile_eam_dealt.cuh); 0 = classic rows
  bool tile_ok = false;     // bin grid / stencil admit the tiling (decided by mmd_neigh_setup)
  bool tile_build2 = true;  // option "tile_build2": CTA-per-tile build from shared memory (0: warp-per-bin build)
  bool list_tile = false;   // format of the current list
  TileGeo tgeo;
  int nsruns = 0;           // runs of the symmetric (full) stencil
  DevBuf sruns, tile_runs, tile_center, tile_info, tile_slots, tile_oslot, trows, tnum;
  bool tile_xsort = true;     // option "tile_xsort": x-sorted windows + interval build
  bool tile_pair_build = true;  // option "tile_pair_build": the interval build sweeps two atoms of a bin at a time
  bool tile_lane_build = false;  // option "tile_lane_build": interval build with one lane per atom (default 0: one warp per atom, measured faster)
  bool list_xsorted = false;  // the current list was built on x-sorted windows
  int tcap = 0;             // row capacity (16-bit entries, multiple of 8)
  int tcap_floor = 0;       // raised when a build overflowed its rows
  int tile_max_h = 0, tile_max_full = 0;
  int tile_builds = 0, tile_fallbacks = 0;
  // bank-dealt copy of the rows (tile_dealt_kernels.cuh): what the LJ force kernel walks
  bool tile_dealt = true;   // option "tile_dealt"
  bool list_dealt = false;  // trowsq holds the current list
  DevBuf trowsq;
  int tcapq = 0;            // dealt row capacity (16-bit entries, multiple of 32)
  int tile_max_rows = 0;    // most rows (atoms of the own bins) of any tile: sizes the force kernel's stash
  // slot-ordered mirror of the positions (xs_mirror.cuh), double-buffered like x / x_alt
  DevBuf xs_rec[2], xs_z[2], slot_of, xs_types;
  int xs_cur = 0;
  bool xs_valid = false;    // the current mirror buffer agrees with x[] for every binned atom
  // several ranks: interior tiles (no ghost in the halo window) run on a second stream while the forward halo is in flight
  bool split_enable = true;   // option "split_force"
  bool split_ready = false;   // tile_split holds the lists of the current neighbor list
  bool split_active = false;  // work may be pending on stream2 (joined before anything else touches the atoms)
  bool ev_int_valid = false;
  DevBuf send_flag;           // [nlocal] 1: the atom is in a send list (its tile counts as boundary)
  DevBuf tile_split;          // [3 * ntiles] interior tiles | boundary tiles | all tiles that own local atoms
  int n_interior = 0, n_boundary = 0, n_active = 0;
  long long split_steps = 0;
  // option "kernel_profile": the LJ tile force kernels add up, per CTA, the clocks until the halo window is staged and
  // the clocks of the whole CTA (queries "stage_clocks", "cta_clocks", "cta_count")
  bool kernel_profile = false;
  unsigned long long* d_prof = nullptr;
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_int = nullptr, ev_bnd = nullptr;
  std::set<const void*> smem_optin;  // kernels of this context's device already opted in to > 48 KB dynamic shared memory

  // Force
  bool have_lj = false, lj_uniform = true;
  DevBuf lj_cut, lj_s6, lj_eps;
  double lj_cut0 = 0, lj_s60 = 0, lj_eps0 = 0;
  int lj_tpa = 0;  // lanes per atom in the LJ kernels; 0 = auto (half list 2, full list 4: profiles/r1_ncu_force_neigh_lj80.md)
  bool have_eam = false, eam_uniform = true;
  DevBuf eam_rho_val, eam_rho_der, eam_z2_val, eam_z2_der, eam_frho_val, eam_frho_der, eam_cut;
  double eam_cut0 = 0, eam_rdr = 0, eam_rdrho = 0;
  int eam_nr = 0, eam_nrho = 0;
  int eam_tpa = 8;
  DevBuf rho, fp;
  DevBuf eam_blob1, eam_blob2, fp_s;  // pair-split spline tables of the dealt kernels, slot-ordered fp mirror
  int eam_nkp = 0;

  // Comm
  bool have_comm = false;
  mmd_swap_table swaps;
  SwapState sw[MMD_MAX_SWAPS];
  DevBuf border_tiles;  // 2 arrays of tile counts
  DevBuf ghost_src, ghost_shift;  // single rank: every ghost resolved to its local source + periodic shift
  bool ghosts_resolved = false;
  // ... inverted: the ghost images of every local atom (xs_mirror.cuh), written by the fused force kernel's epilogue
  bool fuse_ghosts = false;       // option "fuse_ghosts" (measured slower than the separate halo launch at every size: off)
  bool images_ready = false;      // img_* describe the current ghosts
  bool ghosts_fresh = false;      // the last launch already wrote the ghosts of the next step
  DevBuf img_start, img_cursor, img_list;
  long long fused_halo_steps = 0;
  // CUDA graph of two consecutive plain steps (forward halo + fused force/Verlet, twice: the position buffers and the
  // mirror end where they started), replayed for the steps between two rebuilds.  Option "graph_steps": 0 off,
  // 1 (default) on one rank
  int graph_steps = 1;
  cudaGraphExec_t pair_exec = nullptr;
  bool pair_valid = false;        // pair_exec matches the current lists, counts and buffer orientation
  bool capturing = false;         // no phase marks while the stream is in capture mode
  long long pair_launches = 0;    // kernels per replay
  long long graph_replays = 0, graph_captures = 0;
  bool fuse_halo = true;          // option "fuse_halo"
  DevBuf sendbuf, recvbuf;
  DevBuf exch_flag, exch_pos, exch_holes;
  long long exch_sent = 0, exch_received = 0;  // atoms migrated so far (introspection)
#ifdef MMD_WITH_NCCL
  ncclComm_t nccl = nullptr;
#endif
  int nranks = 1, rank = 0;
  // forward halo over peer memory (integrate_comm_kernels.cuh): own receive window + the neighbors' mapped windows
  bool p2p_enable = true;   // option "p2p_halo"
  bool p2p_on = false;
  unsigned char* win = nullptr;
  std::vector<unsigned char*> peer_win;
  unsigned* d_done = nullptr;
  unsigned long long p2p_epoch[MMD_MAX_SWAPS] = {0};
  long long p2p_calls = 0;

  // device scalars + pinned mirror
  // d_scal ints : [0] status, [1] max_n, [2] max_bin, [3] border total 0, [4] border total 1, [5] scan total,
  //               [6..9] exchange/border counts, [10] max full row, [11] max halo window, [12] tile status,
  //               [13] tile row counter, [14] max rows of a tile, [16] interior tiles, [17] boundary tiles, [18] all
  // d_ev doubles: [0] eng, [1] virial, [2] sum m v^2, [3] embed energy
  int* d_scal = nullptr;
  unsigned long long* d_total = nullptr;
  double* d_ev = nullptr;
  int* h_scal = nullptr;
  unsigned long long* h_total = nullptr;
  double* h_ev = nullptr;

  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  // per-phase device timing of mmd_run (option "phase_timing"): marks[k] closes an interval of phase
  // mark_phase[k]; intervals are summed after the loop's final synchronisation.
  bool fuse_integrate = true;  // mmd_run: finalIntegrate(n) + initialIntegrate(n+1) in one kernel
  bool fuse_force = true;      // mmd_run, tile-resident lists: ... and both inside the force kernel's epilogue
  bool phase_timing = false;
  std::vector<cudaEvent_t> marks;
  std::vector<int> mark_phase, mark_calls;
  int nmarks = 0;
  double phase_ms[MMD_NPHASE] = {0, 0, 0, 0, 0};
  long long phase_calls[MMD_NPHASE] = {0, 0, 0, 0, 0};
};

// close the interval since the previous mark and attribute it to `phase` (-1: just open an interval)
static int phase_mark(mmd_ctx* c, int phase, int calls = 1 /* force launches etc. the interval stands for */) {
  if (!c->phase_timing || c->capturing) return MMD_OK;
  if (c->nmarks == (int)c->marks.size()) {
    cudaEvent_t e;
    CU(cudaEventCreate(&e));
    c->marks.push_back(e);
    c->mark_phase.push_back(-1);
    c->mark_calls.push_back(1);
  }
  CU(cudaEventRecord(c->marks[c->nmarks], c->stream));
  c->mark_phase[c->nmarks] = phase;
  c->mark_calls[c->nmarks] = calls;
  c->nmarks++;
  return MMD_OK;
}
static int phase_collect(mmd_ctx* c) {
  if (!c->phase_timing || c->nmarks == 0) return MMD_OK;
  CU(cudaEventSynchronize(c->marks[c->nmarks - 1]));
  for (int k = 1; k < c->nmarks; k++) {
    const int ph = c->mark_phase[k];
    if (ph < 0) continue;
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->marks[k - 1], c->marks[k]));
    c->phase_ms[ph] += ms;
    c->phase_calls[ph] += c->mark_calls[k];
  }
  c->nmarks = 0;
  return MMD_OK;
}

static const int TPB = 256;
// > 48 KB of dynamic shared memory is a per-device, per-kernel opt-in: remembered per context (one context = one device)
template <class K> static int smem_optin(mmd_ctx* c, K kern) {
  const void* key = reinterpret_cast<const void*>(kern);
  if (c->smem_optin.count(key)) return MMD_OK;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
  c->smem_optin.insert(key);
  return MMD_OK;
}
static int launch_fail(int line, cudaError_t e) {
  return set_err(MMD_ERR_CUDA, "mmd_device.cu:%d kernel launch: %s", line, cudaGetErrorString(e));
}

// ---------------------------------------------------------------------------------------
// AoS(pad) <-> Vec4 conversion kernels (the only place the reference's host layout appears)
// ---------------------------------------------------------------------------------------
template <class T>
__global__ void pack_x_kernel(const T* __restrict__ src, const int* __restrict__ type, int n, int pad,
                              Vec4<T>* __restrict__ dst, int keep_type) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Vec4<T> p;
  p.x = src[(size_t)i * pad + 0];
  p.y = src[(size_t)i * pad + 1];
  p.z = src[(size_t)i * pad + 2];
  p.w = keep_type ? dst[i].w : type_to_lane<T>(type ? type[i] : 0);
  dst[i] = p;
}
template <class T>
__global__ void unpack_kernel(const Vec4<T>* __restrict__ src, int n, int pad, T* __restrict__ dst, int* __restrict__ type) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Vec4<T> p = src[i];
  if (dst) {
    dst[(size_t)i * pad + 0] = p.x;
    dst[(size_t)i * pad + 1] = p.y;
    dst[(size_t)i * pad + 2] = p.z;
  }
  if (type) type[i] = lane_to_type(p.w);
}
__global__ void rows_restride_kernel(const int* __restrict__ src, int rows, int src_stride, int dst_stride, int ncopy,
                                     int* __restrict__ dst) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * ncopy) return;
  const int r = (int)(idx / ncopy), k = (int)(idx % ncopy);
  dst[(size_t)r * dst_stride + k] = src[(size_t)r * src_stride + k];
}

// ---------------------------------------------------------------------------------------
// receive window of the peer-memory forward halo: 6 flag words (256 B apart), then [parity][swap] regions
// ---------------------------------------------------------------------------------------
static const int P2P_SWAPS = 6;
static const size_t P2P_REGION_ATOMS = 262144;
static const size_t P2P_FLAG_BYTES = 4096;
static const size_t P2P_REGION_BYTES = P2P_REGION_ATOMS * 3 * sizeof(double);
static const size_t P2P_WIN_BYTES = P2P_FLAG_BYTES + 2 * P2P_SWAPS * P2P_REGION_BYTES;
static inline unsigned long long* p2p_flag(unsigned char* win, int w) { return reinterpret_cast<unsigned long long*>(win + (size_t)w * 256); }
static inline unsigned char* p2p_region(unsigned char* win, int parity, int w) {
  return win + P2P_FLAG_BYTES + ((size_t)parity * P2P_SWAPS + w) * P2P_REGION_BYTES;
}

// ---------------------------------------------------------------------------------------
// typed implementation
// ---------------------------------------------------------------------------------------
template <class T> struct Impl {
  typedef Vec4<T> V;

  static int reserve_atoms(mmd_ctx* c, int n) {
    if (n <= c->cap) return MMD_OK;
    int ncap = std::max((int)(n * 1.25) + 1024, c->cap);
    const size_t keep = (size_t)(c->nlocal + c->nghost) * sizeof(V);
    MM(c->x.reserve((size_t)ncap * sizeof(V), c->stream, keep));
    MM(c->v.reserve((size_t)ncap * sizeof(V), c->stream, (size_t)c->nlocal * sizeof(V)));
    MM(c->f.reserve((size_t)ncap * sizeof(V), c->stream, keep));
    MM(c->x_alt.reserve((size_t)ncap * sizeof(V), c->stream));
    MM(c->v_alt.reserve((size_t)ncap * sizeof(V), c->stream));
    MM(c->atom_bin.reserve((size_t)ncap * sizeof(int), c->stream));
    MM(c->bin_atoms.reserve((size_t)ncap * sizeof(int), c->stream));
    if (c->have_eam) {
      MM(c->rho.reserve((size_t)ncap * sizeof(T), c->stream));
      MM(c->fp.reserve((size_t)ncap * sizeof(T), c->stream));
    }
    c->cap = ncap;
    return MMD_OK;
  }

  // ---- Atom ----------------------------------------------------------------------------
  static int upload(mmd_ctx* c, const void* x, const void* v, const int* type, int nlocal, int pad) {
    c->nlocal = 0;
    c->nghost = 0;
    // ghost head-room estimate from the box: shell of thickness ~3 sigma around the sub-box
    int guess = nlocal;
    if (c->have_box) {
      double vol = 1, shell = 1;
      for (int d = 0; d < 3; d++) {
        const double len = c->hi[d] - c->lo[d];
        vol *= len;
        shell *= len + 2 * 3.2 * (c->have_geo ? 1.0 / c->geo.bininvx : 1.0);
      }
      if (vol > 0) guess = (int)std::min(8.0 * nlocal + 4096, nlocal * (shell / vol));
    }
    MM(reserve_atoms(c, std::max(guess, nlocal)));
    const size_t nb = (size_t)nlocal * pad * sizeof(T);
    MM(c->stage_a.reserve(nb, c->stream));
    MM(c->stage_b.reserve(nb, c->stream));
    MM(c->stage_i.reserve((size_t)nlocal * sizeof(int), c->stream));
    CU(cudaMemcpyAsync(c->stage_a.p, x, nb, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->stage_b.p, v, nb, cudaMemcpyHostToDevice, c->stream));
    if (type) CU(cudaMemcpyAsync(c->stage_i.p, type, (size_t)nlocal * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    const int g = div_up(nlocal, TPB);
    LAUNCH(c, pack_x_kernel<T>, g, TPB, c->stage_a.as<T>(), type ? c->stage_i.as<int>() : nullptr, nlocal, pad,
           c->x.as<V>(), 0);
    LAUNCH(c, pack_x_kernel<T>, g, TPB, c->stage_b.as<T>(), nullptr, nlocal, pad, c->v.as<V>(), 0);
    CU(cudaMemsetAsync(c->f.p, 0, (size_t)c->cap * sizeof(V), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->nlocal = nlocal;
    c->xs_valid = false;
    c->neigh_rows = -1;  // lists and ghost tables refer to the previous atoms
    c->ghosts_resolved = false;
    for (int w = 0; w < MMD_MAX_SWAPS; w++) c->sw[w].sendnum = c->sw[w].recvnum = c->sw[w].firstrecv = 0;
    return MMD_OK;
  }

  static int update(mmd_ctx* c, const void* x, const void* v, int first, int count, int pad) {
    if (first < 0 || count < 0 || first + count > c->nlocal + c->nghost) return set_err(MMD_ERR_ARG, "update: range");
    const size_t nb = (size_t)count * pad * sizeof(T);
    const int g = div_up(count, TPB);
    if (x) {
      c->xs_valid = false;  // positions change behind the mirror's back: refilled before the next dealt force launch
      MM(c->stage_a.reserve(nb, c->stream));
      CU(cudaMemcpyAsync(c->stage_a.p, x, nb, cudaMemcpyHostToDevice, c->stream));
      LAUNCH(c, pack_x_kernel<T>, g, TPB, c->stage_a.as<T>(), nullptr, count, pad, c->x.as<V>() + first, 1);
    }
    if (v) {
      if (first + count > c->nlocal) return set_err(MMD_ERR_ARG, "update: v is defined for local atoms only");
      MM(c->stage_b.reserve(nb, c->stream));
      CU(cudaMemcpyAsync(c->stage_b.p, v, nb, cudaMemcpyHostToDevice, c->stream));
      LAUNCH(c, pack_x_kernel<T>, g, TPB, c->stage_b.as<T>(), nullptr, count, pad, c->v.as<V>() + first, 0);
    }
    return MMD_OK;
  }

  static int download(mmd_ctx* c, void* x, void* v, void* f, int* type, int first, int count, int pad) {
    if (first < 0 || count < 0 || first + count > c->nlocal + c->nghost) return set_err(MMD_ERR_ARG, "download: range");
    if (v && first + count > c->nlocal) return set_err(MMD_ERR_ARG, "download: v is defined for local atoms only");
    if (count == 0) return MMD_OK;
    const size_t nb = (size_t)count * pad * sizeof(T);
    const int g = div_up(count, TPB);
    if (x || type) {
      MM(c->stage_a.reserve(nb, c->stream));
      MM(c->stage_i.reserve((size_t)count * sizeof(int), c->stream));
      if (pad > 3 && x) CU(cudaMemsetAsync(c->stage_a.p, 0, nb, c->stream));
      LAUNCH(c, unpack_kernel<T>, g, TPB, c->x.as<V>() + first, count, pad, x ? c->stage_a.as<T>() : nullptr,
             type ? c->stage_i.as<int>() : nullptr);
      if (x) CU(cudaMemcpyAsync(x, c->stage_a.p, nb, cudaMemcpyDeviceToHost, c->stream));
      if (type) CU(cudaMemcpyAsync(type, c->stage_i.p, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    }
    if (v) {
      MM(c->stage_b.reserve(nb, c->stream));
      if (pad > 3) CU(cudaMemsetAsync(c->stage_b.p, 0, nb, c->stream));
      LAUNCH(c, unpack_kernel<T>, g, TPB, c->v.as<V>() + first, count, pad, c->stage_b.as<T>(), nullptr);
      CU(cudaMemcpyAsync(v, c->stage_b.p, nb, cudaMemcpyDeviceToHost, c->stream));
    }
    if (f) {
      MM(c->stage_c.reserve(nb, c->stream));
      if (pad > 3) CU(cudaMemsetAsync(c->stage_c.p, 0, nb, c->stream));
      LAUNCH(c, unpack_kernel<T>, g, TPB, c->f.as<V>() + first, count, pad, c->stage_c.as<T>(), nullptr);
      CU(cudaMemcpyAsync(f, c->stage_c.p, nb, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return MMD_OK;
  }

  static int pbc(mmd_ctx* c) {
    if (!c->have_box) return set_err(MMD_ERR_STATE, "pbc: mmd_atom_set_box not called");
    c->xs_valid = false;
    LAUNCH(c, pbc_kernel<T>, div_up(c->nlocal, TPB), TPB, c->x.as<V>(), c->nlocal, (T)c->prd[0], (T)c->prd[1],
           (T)c->prd[2]);
    return MMD_OK;
  }

  // ---- binning -------------------------------------------------------------------------
  static BinGeo<T> geo(mmd_ctx* c) {
    BinGeo<T> g;
    for (int d = 0; d < 3; d++) g.prd[d] = (T)c->prd[d];
    g.bininv[0] = (T)c->geo.bininvx; g.bininv[1] = (T)c->geo.bininvy; g.bininv[2] = (T)c->geo.bininvz;
    g.nbin[0] = c->geo.nbinx; g.nbin[1] = c->geo.nbiny; g.nbin[2] = c->geo.nbinz;
    g.mbin[0] = c->geo.mbinx; g.mbin[1] = c->geo.mbiny; g.mbin[2] = c->geo.mbinz;
    g.mbinlo[0] = c->geo.mbinxlo; g.mbinlo[1] = c->geo.mbinylo; g.mbinlo[2] = c->geo.mbinzlo;
    g.mbins = c->mbins;
    return g;
  }

  // count -> exclusive scan -> fill -> per-bin sort.  All asynchronous.
  static int binatoms_async(mmd_ctx* c, int n) {
    if (!c->have_geo || !c->have_box) return set_err(MMD_ERR_STATE, "binatoms: mmd_neigh_setup / mmd_atom_set_box missing");
    const int mb = c->mbins;
    CU(cudaMemsetAsync(c->bincount.p, 0, (size_t)mb * sizeof(int), c->stream));
    CU(cudaMemsetAsync(c->cursor.p, 0, (size_t)mb * sizeof(int), c->stream));
    LAUNCH(c, bin_count_kernel<T>, div_up(n, TPB), TPB, c->x.as<V>(), n, geo(c), c->atom_bin.as<int>(),
           c->bincount.as<int>(), c->d_scal + 0);
    const int ntiles = div_up(mb, SCAN_TILE);
    MM(c->tile_sums.reserve((size_t)ntiles * sizeof(int), c->stream));
    LAUNCH(c, scan_tile_sums_kernel, ntiles, SCAN_THREADS, c->bincount.as<int>(), mb, c->tile_sums.as<int>(),
           c->d_scal + 2);
    LAUNCH(c, scan_spine_kernel, 1, 1024, c->tile_sums.as<int>(), ntiles, c->bin_start.as<int>() + mb, 0);
    LAUNCH(c, scan_apply_kernel, ntiles, SCAN_THREADS, c->bincount.as<int>(), mb, c->tile_sums.as<int>(),
           c->bin_start.as<int>());
    LAUNCH(c, bin_fill_kernel, div_up(n, TPB), TPB, c->atom_bin.as<int>(), n, c->bin_start.as<int>(),
           c->cursor.as<int>(), c->bin_atoms.as<int>());
    LAUNCH(c, bin_sort_kernel, div_up(mb, TPB), TPB, c->bin_start.as<int>(), mb, c->bin_atoms.as<int>());
    c->binned_count = n;
    return MMD_OK;
  }

  static int read_status(mmd_ctx* c) {  // status + running max bin occupancy
    CU(cudaMemcpyAsync(c->h_scal, c->d_scal, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_scal + 10, c->d_scal + 10, 5 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->max_bin_seen = std::max(c->max_bin_seen, c->h_scal[2]);
    if (c->h_scal[0] & 5) {
      // report once: the bits describe the state that was just examined, not whatever the caller uploads next
      const int st = c->h_scal[0];
      CU(cudaMemsetAsync(c->d_scal + 0, 0, sizeof(int), c->stream));
      if (st & 1) return set_err(MMD_ERR_STATE, "atom outside the bin grid (lost atom / bad coordinates)");
      return set_err(MMD_ERR_STATE, "forward halo: a neighbor rank's message did not arrive (peer-memory path timed out)");
    }
    return MMD_OK;
  }

  static int sort(mmd_ctx* c) {
    MM(binatoms_async(c, c->nlocal));
    LAUNCH(c, permute_atoms_kernel<T>, div_up(c->nlocal, TPB), TPB, c->bin_atoms.as<int>(), c->nlocal, c->x.as<V>(),
           c->v.as<V>(), c->x_alt.as<V>(), c->v_alt.as<V>());
    std::swap(c->x, c->x_alt);  // pointer swap, ref/atom.cpp:409-418
    std::swap(c->v, c->v_alt);
    c->xs_valid = false;
    return MMD_OK;
  }

  // ---- neighbor build ------------------------------------------------------------------
  // EAM runs on tile lists only through the dealt kernels, which need uniform tables (tile_eam_dealt.cuh)
  static bool want_tile(mmd_ctx* c) {
    return c->tile_enable && c->tile_ok && (!c->have_eam || (c->tile_eam && c->tile_dealt && c->eam_uniform));
  }

  // ---- slot-ordered mirror of the positions (xs_mirror.cuh) -----------------------------------
  static XsMirror<T> mirror(mmd_ctx* c, int which) {
    XsMirror<T> M;
    M.rec = c->xs_rec[which].as<QRec<T>>();
    M.z = c->xs_z[which].as<T>();
    M.slot_of = c->slot_of.as<int>();
    return M;
  }
  // the mirror the atom-moving kernels have to keep current (none unless a dealt list with a valid mirror exists)
  static XsMirror<T> live_mirror(mmd_ctx* c) {
    if (c->list_tile && c->list_dealt && c->xs_valid && c->neigh_rows == c->nlocal) return mirror(c, c->xs_cur);
    XsMirror<T> M;
    M.rec = nullptr; M.z = nullptr; M.slot_of = nullptr;
    return M;
  }
  // (re)build the current mirror buffer from x[] for every binned atom
  static int xs_fill(mmd_ctx* c) {
    const int nall = c->nlocal + c->nghost;
    const size_t cap = (size_t)std::max(c->cap, nall) + 64;
    for (int b = 0; b < 2; b++) {
      MM(c->xs_rec[b].reserve(cap * 16, c->stream));
      if (sizeof(T) == 8) MM(c->xs_z[b].reserve(cap * 8, c->stream));
    }
    MM(c->slot_of.reserve(cap * sizeof(int), c->stream));
    MM(c->xs_types.reserve(cap, c->stream));
    LAUNCH(c, xs_fill_kernel<T>, div_up(nall, TPB), TPB, c->x.as<V>(), c->tile_slots.as<int>(), nall, mirror(c, c->xs_cur),
           c->slot_of.as<int>(), c->xs_types.as<unsigned char>());
    c->xs_valid = true;
    return MMD_OK;
  }
  // local atoms were moved by an unfused integrator: copy them into the mirror
  static int xs_refresh_locals(mmd_ctx* c) {
    const XsMirror<T> M = live_mirror(c);
    if (M.rec) LAUNCH(c, xs_refresh_kernel<T>, div_up(c->nlocal, TPB), TPB, c->x.as<V>(), 0, c->nlocal, M);
    return MMD_OK;
  }

  // tile-resident build (tile_kernels.cuh).  *done = false when the tiling does not fit this state (halo window
  // larger than the shared-memory capacity, stencil leaving the grid): the caller builds classic rows instead.
  static int build_tile(mmd_ctx* c, int halfneigh, int gn, bool* done) {
    *done = false;
    const int nall = c->nlocal + c->nghost;
    TileGeo& g = c->tgeo;
    // tile origin: the first bin that can own a local atom opens a tile (no half-empty tiles along the low faces)
    {
      const double binv[3] = {c->geo.bininvx, c->geo.bininvy, c->geo.bininvz};
      const int mlo[3] = {c->geo.mbinxlo, c->geo.mbinylo, c->geo.mbinzlo};
      const int tb[3] = {TBX, TBY, TBZ};
      int o[3];
      for (int d = 0; d < 3; d++) {
        const int first = std::max(0, (int)(c->lo[d] * binv[d]) - mlo[d]);
        o[d] = (tb[d] - first % tb[d]) % tb[d];
      }
      g.ox = o[0]; g.oy = o[1]; g.oz = o[2];
      g.ntx = (g.mbx + g.ox + TBX - 1) / TBX; g.nty = (g.mby + g.oy + TBY - 1) / TBY; g.ntz = (g.mbz + g.oz + TBZ - 1) / TBZ;
      g.ntiles = g.ntx * g.nty * g.ntz;
    }
    MM(c->tile_runs.reserve((size_t)g.ntiles * g.nrun * sizeof(int2), c->stream));
    MM(c->tile_center.reserve((size_t)g.ntiles * TILE_NCENTER * sizeof(int4), c->stream));
    MM(c->tile_info.reserve((size_t)g.ntiles * sizeof(int2), c->stream));
    MM(c->tile_slots.reserve((size_t)std::max(c->cap, nall) * sizeof(int), c->stream));
    // one row per binned atom, numbered tile by tile (the rows of ghost atoms are never touched)
    MM(c->tnum.reserve((size_t)std::max(nall, 1) * sizeof(int2), c->stream, 0, 1.1));
    CU(cudaMemsetAsync(c->tnum.p, 0xff, (size_t)std::max(nall, 1) * sizeof(int2), c->stream));
    CU(cudaMemsetAsync(c->d_scal + 10, 0, 5 * sizeof(int), c->stream));
    LAUNCH(c, tile_table_kernel, div_up(g.ntiles, 4), 128, g, c->bin_start.as<int>(), c->bin_atoms.as<int>(), c->mbins,
           c->nlocal, c->tile_runs.as<int2>(), c->tile_center.as<int4>(), c->tile_info.as<int2>(), c->d_scal + 11,
           c->d_scal + 13, c->d_scal + 14);
    // the windows' slot -> atom map: a private copy of the CSR bins, every bin sorted by x when the interval build is used
    bool xsorted = c->tile_xsort && c->tile_build2 && c->nsruns <= 32;
    if (xsorted) {
      MM(c->tile_oslot.reserve((size_t)std::max(c->cap, nall) * sizeof(int), c->stream));
      LAUNCH(c, bin_xsort_kernel<T>, div_up(c->mbins, TPB), TPB, c->x.as<V>(), c->bin_start.as<int>(), c->bin_atoms.as<int>(),
             c->mbins, c->tile_slots.as<int>(), c->tile_oslot.as<int>(), c->d_scal + 12);
    } else {
      CU(cudaMemcpyAsync(c->tile_slots.p, c->bin_atoms.p, (size_t)nall * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    }
    const int mode = halfneigh ? (gn ? 1 : 2) : 0;
    // the largest halo window sizes the shared memory of the build and force kernels
    MM(read_status(c));
    if (xsorted && (c->h_scal[12] & 8)) {  // a bin too full for the thread-local sort: windows stay in CSR order
      xsorted = false;
      CU(cudaMemsetAsync(c->d_scal + 12, 0, sizeof(int), c->stream));
      CU(cudaMemcpyAsync(c->tile_slots.p, c->bin_atoms.p, (size_t)nall * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    }
    c->tile_max_h = c->h_scal[11];
    c->tile_max_rows = c->h_scal[14];
    const int hcap_limit = (int)((227 * 1024 - 4096) / (3 * sizeof(T) + 1)) & ~63;
    if (c->tile_max_h > hcap_limit) {
      c->tile_fallbacks++;
      return MMD_OK;  // *done stays false
    }
    g.hcap = std::max(64, (c->tile_max_h + 8 + 63) & ~63);  // + 8: sentinel slots of the builds / the dealt force kernel
    Build2Params<T> B;
    {
      const std::vector<double>& cuts = c->h_cutneighsq;  // host copy kept by mmd_neigh_setup (values exact in T)
      B.cut0 = (T)cuts[0];
      B.uniform_cut = 1;
      double cmax = cuts[0];
      for (double v : cuts) { if (v != cuts[0]) B.uniform_cut = 0; cmax = std::max(cmax, v); }
      B.band = (float)(1e-4 * cmax);
      // culling radius: cutneigh + 0.1 % + 2 % of the smallest bin edge (covers the FP32 images and coord2bin's rounding)
      B.rcull = (float)(std::sqrt(cmax) * 1.001 + 0.02 / std::max(std::max(c->geo.bininvx, c->geo.bininvy), c->geo.bininvz));
      B.binsize[0] = 1.0 / c->geo.bininvx; B.binsize[1] = 1.0 / c->geo.bininvy; B.binsize[2] = 1.0 / c->geo.bininvz;
      B.mbinlo[0] = c->geo.mbinxlo; B.mbinlo[1] = c->geo.mbinylo; B.mbinlo[2] = c->geo.mbinzlo;
    }
    // the interval build's footprint can exceed the force kernel's (FP32: ~13 B per window atom + tables): windows then
    // stay in CSR order and the table build / classic rows take over
    if (xsorted && build3_smem_bytes<T>(g, g.hcap, !B.uniform_cut) > (size_t)(227 * 1024 - 2048)) {
      xsorted = false;
      CU(cudaMemcpyAsync(c->tile_slots.p, c->bin_atoms.p, (size_t)nall * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
    }
    const size_t b2_smem = build2_smem_bytes<T>(g, g.hcap, !B.uniform_cut);
    const bool use_b2 = c->tile_build2 && c->nsruns <= TB2_MAXSR && b2_smem <= (size_t)(227 * 1024 - 2048) &&
                        c->tile_max_h <= TB2_MAXH;
    for (;;) {
      // row stride: twice the reference's maxneighs until a build has shown the longest full row, then that + 15 %
      if (c->tile_max_full > 0) c->tcap = std::max(c->tcap_floor, ((int)(c->tile_max_full * 1.15) + 8) & ~7);
      else c->tcap = std::max(c->tcap_floor, (((halfneigh ? 2 : 1) * c->maxneighs) + 7) & ~7);
      MM(c->trows.reserve((size_t)std::max(nall, 1) * c->tcap * sizeof(unsigned short), c->stream, 0, 1.05));
      CU(cudaMemsetAsync(c->d_scal + 1, 0, sizeof(int), c->stream));
      CU(cudaMemsetAsync(c->d_scal + 10, 0, sizeof(int), c->stream));
      CU(cudaMemsetAsync(c->d_total, 0, sizeof(unsigned long long), c->stream));
      const int grid = div_up(c->mbins, NB_WARPS);
#define NBT_ARGS                                                                                                           \
  c->x.as<V>(), c->nlocal, c->bin_start.as<int>(), c->bin_atoms.as<int>(), c->mbins, c->sruns.as<StencilRun>(), c->nsruns, \
      c->cutneighsq.as<T>(), c->ntypes, g, c->tile_runs.as<int2>(), c->tile_center.as<int4>(),                             \
      c->trows.as<unsigned short>(), c->tcap, c->numneigh.as<int>(), c->tnum.as<int2>(), c->d_scal + 12, c->d_scal + 1,    \
      c->d_scal + 10, c->d_total
#define NB2_ARGS                                                                                                          \
  c->x.as<V>(), c->nlocal, c->bin_start.as<int>(), c->bin_atoms.as<int>(), c->mbins, c->sruns.as<StencilRun>(), c->nsruns, \
      c->cutneighsq.as<T>(), c->ntypes, g, B, c->tile_runs.as<int2>(), c->tile_center.as<int4>(), c->tile_info.as<int2>(),  \
      c->trows.as<unsigned short>(), c->tcap, c->numneigh.as<int>(), c->tnum.as<int2>(), c->d_scal + 12, c->d_scal + 1,    \
      c->d_scal + 10, c->d_total
      if (xsorted && c->tile_lane_build && buildl_smem_bytes<T>(g, g.hcap, !B.uniform_cut) <= (size_t)(227 * 1024 - 2048)) {
        const size_t bl_smem = buildl_smem_bytes<T>(g, g.hcap, !B.uniform_cut);
#define NBL_ARGS                                                                                                          \
  c->x.as<V>(), c->nlocal, c->bin_start.as<int>(), c->tile_slots.as<int>(), c->mbins, c->sruns.as<StencilRun>(), c->nsruns, \
      c->cutneighsq.as<T>(), c->ntypes, g, B, c->tile_runs.as<int2>(), c->tile_center.as<int4>(), c->tile_info.as<int2>(),  \
      c->trows.as<unsigned short>(), c->tcap, c->numneigh.as<int>(), c->tnum.as<int2>(), c->d_scal + 12, c->d_scal + 1,    \
      c->d_scal + 10, c->d_total
#define NBL_LAUNCH(M, U)                                                                       \
  do {                                                                                         \
    MM(smem_optin(c, neigh_build_lane_kernel<T, M, U>));                                       \
    LAUNCH_SMEM(c, (neigh_build_lane_kernel<T, M, U>), g.ntiles, TBL_THREADS, bl_smem, NBL_ARGS); \
  } while (0)
        if (B.uniform_cut) {
          if (mode == 0) NBL_LAUNCH(0, 1);
          if (mode == 1) NBL_LAUNCH(1, 1);
          if (mode == 2) NBL_LAUNCH(2, 1);
        } else {
          if (mode == 0) NBL_LAUNCH(0, 0);
          if (mode == 1) NBL_LAUNCH(1, 0);
          if (mode == 2) NBL_LAUNCH(2, 0);
        }
#undef NBL_LAUNCH
#undef NBL_ARGS
      } else if (xsorted) {
        const size_t b3_smem = build3_smem_bytes<T>(g, g.hcap, !B.uniform_cut);
        MM(smem_optin(c, neigh_build_tile3_kernel<T, 0, 1>)); MM(smem_optin(c, neigh_build_tile3_kernel<T, 1, 1>));
        MM(smem_optin(c, neigh_build_tile3_kernel<T, 2, 1>)); MM(smem_optin(c, neigh_build_tile3_kernel<T, 0, 0>));
        MM(smem_optin(c, neigh_build_tile3_kernel<T, 1, 0>)); MM(smem_optin(c, neigh_build_tile3_kernel<T, 2, 0>));
#define NB3_ARGS                                                                                                          \
  c->x.as<V>(), c->nlocal, c->bin_start.as<int>(), c->tile_slots.as<int>(), c->mbins, c->sruns.as<StencilRun>(), c->nsruns, \
      c->cutneighsq.as<T>(), c->ntypes, g, B, c->tile_runs.as<int2>(), c->tile_center.as<int4>(), c->tile_info.as<int2>(),  \
      c->trows.as<unsigned short>(), c->tcap, c->numneigh.as<int>(), c->tnum.as<int2>(), c->d_scal + 12, c->d_scal + 1,    \
      c->d_scal + 10, c->d_total, (int)c->tile_pair_build
        if (B.uniform_cut) {
          if (mode == 0) LAUNCH_SMEM(c, (neigh_build_tile3_kernel<T, 0, 1>), g.ntiles, TB2_THREADS, b3_smem, NB3_ARGS);
          if (mode == 1) LAUNCH_SMEM(c, (neigh_build_tile3_kernel<T, 1, 1>), g.ntiles, TB2_THREADS, b3_smem, NB3_ARGS);
          if (mode == 2) LAUNCH_SMEM(c, (neigh_build_tile3_kernel<T, 2, 1>), g.ntiles, TB2_THREADS, b3_smem, NB3_ARGS);
        } else {
          if (mode == 0) LAUNCH_SMEM(c, (neigh_build_tile3_kernel<T, 0, 0>), g.ntiles, TB2_THREADS, b3_smem, NB3_ARGS);
          if (mode == 1) LAUNCH_SMEM(c, (neigh_build_tile3_kernel<T, 1, 0>), g.ntiles, TB2_THREADS, b3_smem, NB3_ARGS);
          if (mode == 2) LAUNCH_SMEM(c, (neigh_build_tile3_kernel<T, 2, 0>), g.ntiles, TB2_THREADS, b3_smem, NB3_ARGS);
        }
#undef NB3_ARGS
      } else if (use_b2) {
        MM(smem_optin(c, neigh_build_tile2_kernel<T, 0, 1>)); MM(smem_optin(c, neigh_build_tile2_kernel<T, 1, 1>));
        MM(smem_optin(c, neigh_build_tile2_kernel<T, 2, 1>)); MM(smem_optin(c, neigh_build_tile2_kernel<T, 0, 0>));
        MM(smem_optin(c, neigh_build_tile2_kernel<T, 1, 0>)); MM(smem_optin(c, neigh_build_tile2_kernel<T, 2, 0>));
        if (B.uniform_cut) {
          if (mode == 0) LAUNCH_SMEM(c, (neigh_build_tile2_kernel<T, 0, 1>), g.ntiles, TB2_THREADS, b2_smem, NB2_ARGS);
          if (mode == 1) LAUNCH_SMEM(c, (neigh_build_tile2_kernel<T, 1, 1>), g.ntiles, TB2_THREADS, b2_smem, NB2_ARGS);
          if (mode == 2) LAUNCH_SMEM(c, (neigh_build_tile2_kernel<T, 2, 1>), g.ntiles, TB2_THREADS, b2_smem, NB2_ARGS);
        } else {
          if (mode == 0) LAUNCH_SMEM(c, (neigh_build_tile2_kernel<T, 0, 0>), g.ntiles, TB2_THREADS, b2_smem, NB2_ARGS);
          if (mode == 1) LAUNCH_SMEM(c, (neigh_build_tile2_kernel<T, 1, 0>), g.ntiles, TB2_THREADS, b2_smem, NB2_ARGS);
          if (mode == 2) LAUNCH_SMEM(c, (neigh_build_tile2_kernel<T, 2, 0>), g.ntiles, TB2_THREADS, b2_smem, NB2_ARGS);
        }
      } else {
        if (mode == 0) LAUNCH(c, (neigh_build_tile_kernel<T, 0>), grid, NB_WARPS * 32, NBT_ARGS);
        if (mode == 1) LAUNCH(c, (neigh_build_tile_kernel<T, 1>), grid, NB_WARPS * 32, NBT_ARGS);
        if (mode == 2) LAUNCH(c, (neigh_build_tile_kernel<T, 2>), grid, NB_WARPS * 32, NBT_ARGS);
      }
#undef NB2_ARGS
#undef NBT_ARGS
      CU(cudaMemcpyAsync(c->h_total, c->d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
      MM(read_status(c));
      if (c->h_scal[12] != 0) {  // a stencil range left the bin grid: no tile mapping for this state
        c->tile_fallbacks++;
        return MMD_OK;  // *done stays false
      }
      const int max_half = c->h_scal[1], max_full = c->h_scal[10];
      bool again = false;
      if (max_half >= c->maxneighs) {  // ref/neighbor.cpp:186-208, on the reference's own row lengths
        c->maxneighs = (int)(max_half * 1.2);
        c->neigh_resizes++;
        again = true;
      }
      if (max_full > c->tcap) {
        c->tcap_floor = ((int)(max_full * 1.2) + 7) & ~7;
        again = true;
      }
      c->tile_max_full = std::max(c->tile_max_full, max_full);
      if (!again) break;
    }
    c->neigh_stride = (c->maxneighs + 7) & ~7;
    c->list_tile = true;
    c->list_xsorted = xsorted;
    c->tile_builds++;
    // bank-dealt copy of the rows for the LJ force kernel (tile_dealt_kernels.cuh)
    c->list_dealt = false;
    if (c->tile_dealt) {
      const int nrows = std::max(nall, 1);
      c->tcapq = dealt_capacity(std::min(c->tile_max_full, c->tcap));
      const size_t dsm = deal_smem_bytes(c->tcap, c->tcapq);
      if (dsm <= (size_t)(227 * 1024 - 2048) && c->tcapq <= (DEAL_MAXROW & ~31)) {
        MM(c->trowsq.reserve((size_t)nrows * c->tcapq * sizeof(unsigned short) + 256, c->stream, 0, 1.05));
        MM(smem_optin(c, tile_rows_deal_kernel));
        LAUNCH_SMEM(c, tile_rows_deal_kernel, div_up(nrows, DEAL_THREADS), DEAL_THREADS, dsm, c->trows.as<unsigned short>(),
                    c->tnum.as<int2>(), nrows, c->tcap, c->nlocal, c->trowsq.as<unsigned short>(), c->tcapq, g.hcap - 8);
        c->list_dealt = true;
        c->xs_valid = false;
        MM(xs_fill(c));
        // lists of the tiles that own local atoms (what the force kernels launch over): interior (no ghost in the halo
        // window), boundary, and all of them
        {
          MM(c->tile_split.reserve((size_t)3 * g.ntiles * sizeof(int), c->stream));
          CU(cudaMemsetAsync(c->d_scal + 16, 0, 3 * sizeof(int), c->stream));
          const bool want_split = c->nranks > 1 && c->split_enable && c->stream2 != nullptr;
          if (want_split) {  // atoms the forward halo reads: their tiles go to the boundary list
            MM(c->send_flag.reserve((size_t)std::max(c->nlocal, 1), c->stream, 0, 1.2));
            CU(cudaMemsetAsync(c->send_flag.p, 0, (size_t)std::max(c->nlocal, 1), c->stream));
            for (int ws = 0; ws < c->swaps.nswap; ws++)
              LAUNCH(c, mark_send_atoms_kernel, div_up(c->sw[ws].sendnum, TPB), TPB, c->sw[ws].list.as<int>(), c->sw[ws].sendnum,
                     c->nlocal, c->send_flag.as<unsigned char>());
          }
          LAUNCH(c, tile_classify_kernel, div_up(g.ntiles, 4), 128, g, c->tile_runs.as<int2>(), c->tile_center.as<int4>(),
                 c->tile_info.as<int2>(), c->tile_slots.as<int>(), c->nlocal,
                 want_split ? c->send_flag.as<unsigned char>() : (const unsigned char*)nullptr, c->tile_split.as<int>(), c->d_scal + 16);
          CU(cudaMemcpyAsync(c->h_scal + 16, c->d_scal + 16, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
          CU(cudaStreamSynchronize(c->stream));
          c->n_interior = c->h_scal[16];
          c->n_boundary = c->h_scal[17];
          c->n_active = c->h_scal[18];
          c->split_ready = c->nranks > 1 && c->split_enable && c->stream2 != nullptr;
        }
      }
    }
    *done = true;
    return MMD_OK;
  }

  // reference-format rows (global ids, half subset for half lists) of a tile-resident list, into c->neighbors
  static int export_rows(mmd_ctx* c) {
    if (!c->list_tile) return MMD_OK;
    const TileGeo& g = c->tgeo;
    MM(c->neighbors.reserve((size_t)std::max(c->nlocal, 1) * c->neigh_stride * sizeof(int), c->stream, 0, 1.05));
    CU(cudaMemsetAsync(c->d_scal + 12, 0, sizeof(int), c->stream));
    LAUNCH(c, tile_rows_export_kernel, g.ntiles, TILE_THREADS, g, c->tile_runs.as<int2>(), c->tile_center.as<int4>(),
           c->tile_info.as<int2>(), c->tile_slots.as<int>(), c->list_xsorted ? c->tile_oslot.as<int>() : (const int*)nullptr,
           c->trows.as<unsigned short>(), c->tnum.as<int2>(), c->tcap, c->nlocal, c->list_half ? 1 : 0,
           c->neighbors.as<int>(), c->neigh_stride, c->neigh_stride, c->d_scal + 12);
    CU(cudaMemcpyAsync(c->h_scal + 12, c->d_scal + 12, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->h_scal[12] & 16) return set_err(MMD_ERR_STATE, "neighbor export: a row is longer than the sorting export handles");
    return MMD_OK;
  }
  // switch the current list to the classic format (kernels without a tile variant)
  static int ensure_classic(mmd_ctx* c) {
    if (!c->list_tile) return MMD_OK;
    MM(export_rows(c));
    c->list_tile = false;
    return MMD_OK;
  }

  static int build(mmd_ctx* c, int halfneigh, int gn, int* maxneighs_io, long long* total_out) {
    if (maxneighs_io && *maxneighs_io > 0) c->maxneighs = *maxneighs_io;
    const int nall = c->nlocal + c->nghost;
    MM(binatoms_async(c, nall));
    MM(c->numneigh.reserve((size_t)std::max(c->nlocal, 1) * sizeof(int), c->stream, 0, 1.1));
    c->list_tile = false;
    if (want_tile(c)) {
      bool done = false;
      MM(build_tile(c, halfneigh, gn, &done));
      if (done) {
        c->total_neigh = (long long)*c->h_total;
        c->neigh_rows = c->nlocal;
        c->neigh_builds++;
        c->list_half = halfneigh;
        c->list_gn = gn;
        if (maxneighs_io) *maxneighs_io = c->maxneighs;
        if (total_out) *total_out = c->total_neigh;
        return MMD_OK;
      }
    }
    const int mode = halfneigh ? (gn ? 1 : 2) : 0;
    for (;;) {
      c->neigh_stride = (c->maxneighs + 7) & ~7;
      MM(c->neighbors.reserve((size_t)std::max(c->nlocal, 1) * c->neigh_stride * sizeof(int), c->stream, 0, 1.05));
      CU(cudaMemsetAsync(c->d_scal + 1, 0, sizeof(int), c->stream));
      CU(cudaMemsetAsync(c->d_total, 0, sizeof(unsigned long long), c->stream));
      const int grid = div_up(c->mbins, NB_WARPS);
#define NB_ARGS                                                                                                  \
  c->x.as<V>(), c->nlocal, c->bin_start.as<int>(), c->bin_atoms.as<int>(), c->mbins, c->runs.as<int2>(), c->nruns, \
      c->cutneighsq.as<T>(), c->ntypes, c->neighbors.as<int>(), c->numneigh.as<int>(), c->neigh_stride,           \
      c->d_scal + 1, c->d_total, c->maxneighs
      if (mode == 0) LAUNCH(c, (neigh_build_kernel<T, 0>), grid, NB_WARPS * 32, NB_ARGS);
      if (mode == 1) LAUNCH(c, (neigh_build_kernel<T, 1>), grid, NB_WARPS * 32, NB_ARGS);
      if (mode == 2) LAUNCH(c, (neigh_build_kernel<T, 2>), grid, NB_WARPS * 32, NB_ARGS);
#undef NB_ARGS
      CU(cudaMemcpyAsync(c->h_total, c->d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
      MM(read_status(c));
      const int max_n = c->h_scal[1];
      if (max_n >= c->maxneighs) {  // ref/neighbor.cpp:186-208
        c->maxneighs = (int)(max_n * 1.2);
        c->neigh_resizes++;
        continue;
      }
      break;
    }
    c->total_neigh = (long long)*c->h_total;
    c->neigh_rows = c->nlocal;
    c->neigh_builds++;
    c->list_half = halfneigh;
    c->list_gn = gn;
    if (maxneighs_io) *maxneighs_io = c->maxneighs;
    if (total_out) *total_out = c->total_neigh;
    return MMD_OK;
  }

  // ---- LJ ------------------------------------------------------------------------------
  template <int TPA, int HALF, int GN, int EV>
  static int lj_launch(mmd_ctx* c) {
    LJParams<T> P;
    P.cutforcesq = (T)c->lj_cut0; P.sigma6 = (T)c->lj_s60; P.epsilon = (T)c->lj_eps0;
    P.cutforcesq_tab = c->lj_cut.as<T>(); P.sigma6_tab = c->lj_s6.as<T>(); P.epsilon_tab = c->lj_eps.as<T>();
    P.ntypes = c->ntypes;
    const int grid = div_up((long long)c->nlocal * TPA, LJ_BLOCK);
    if (c->lj_uniform)
      LAUNCH(c, (force_lj_kernel<T, TPA, HALF, GN, EV, 1>), grid, LJ_BLOCK, c->x.as<V>(), c->f.as<V>(),
             c->neighbors.as<int>(), c->numneigh.as<int>(), c->neigh_stride, c->nlocal, P, c->d_ev);
    else
      LAUNCH(c, (force_lj_kernel<T, TPA, HALF, GN, EV, 0>), grid, LJ_BLOCK, c->x.as<V>(), c->f.as<V>(),
             c->neighbors.as<int>(), c->numneigh.as<int>(), c->neigh_stride, c->nlocal, P, c->d_ev);
    return MMD_OK;
  }
  template <int TPA> static int lj_dispatch(mmd_ctx* c, int half, int gn, int ev) {
    if (half) {
      if (gn) return ev ? lj_launch<TPA, 1, 1, 1>(c) : lj_launch<TPA, 1, 1, 0>(c);
      return ev ? lj_launch<TPA, 1, 0, 1>(c) : lj_launch<TPA, 1, 0, 0>(c);
    }
    return ev ? lj_launch<TPA, 0, 0, 1>(c) : lj_launch<TPA, 0, 0, 0>(c);
  }
  // tile-resident list: owner-computes shared-memory kernel (tile_kernels.cuh)
  // part: 0 = every tile on the context's stream; 1 = interior tiles on stream2; 2 = boundary tiles on the context's
  // stream (dealt lists of several ranks only, see run())
  template <int EV, int UNI, int INTEG> static int lj_tile_launch(mmd_ctx* c, int half, const VerletParams<T>& VP, int part = 0) {
    LJTileParams<T> P;
    P.cutforcesq = (T)c->lj_cut0; P.sigma6 = (T)c->lj_s60; P.epsilon = (T)c->lj_eps0;
    P.cutforcesq_tab = c->lj_cut.as<T>(); P.sigma6_tab = c->lj_s6.as<T>(); P.epsilon_tab = c->lj_eps.as<T>();
    P.ntypes = c->ntypes;
    // every pair is visited from both ends: half-list semantics sum each pair once (ref/force_lj.cpp:246-248),
    // full-list semantics report twice the pair energy and half the double-counted virial (:441-442)
    P.e_scale = half ? 0.5 : 1.0;
    P.v_scale = 0.5;
    const TileGeo& g = c->tgeo;
    const int scap = (c->tile_max_rows + 7) & ~7;
    if (c->list_dealt && qwin_smem_bytes<T>(g.hcap, !UNI, scap) <= (size_t)(227 * 1024 - 2048)) {
      if (!c->xs_valid) MM(xs_fill(c));
      LJDealtParams<T> Q;
      Q.cutforcesq = P.cutforcesq; Q.sigma6 = P.sigma6; Q.epsilon = P.epsilon;
      Q.k48 = (T)48 * P.epsilon * P.sigma6;
      Q.kA = Q.k48 * P.sigma6; Q.kB = (T)-0.5 * Q.k48;
      Q.cutforcesq_tab = P.cutforcesq_tab; Q.sigma6_tab = P.sigma6_tab; Q.epsilon_tab = P.epsilon_tab;
      Q.ntypes = P.ntypes; Q.e_scale = P.e_scale; Q.v_scale = P.v_scale;
      MM(smem_optin(c, force_lj_dealt_kernel<T, EV, UNI, INTEG>));
      GhostImages<T> GI;
      memset(&GI, 0, sizeof GI);
      if (INTEG && part == 0 && c->images_ready && c->ghosts_resolved && c->fuse_ghosts && c->fuse_halo) {
        GI.start = c->img_start.as<int>();
        GI.list = c->img_list.as<int2>();
        GI.prd[0] = (T)c->prd[0]; GI.prd[1] = (T)c->prd[1]; GI.prd[2] = (T)c->prd[2];
        GI.nlocal = c->nlocal;
      }
      const int* list = c->tile_split.as<int>() + (part == 0 ? 2 * g.ntiles : (part == 2 ? g.ntiles : 0));
      const int grid = part == 0 ? c->n_active : (part == 1 ? c->n_interior : c->n_boundary);
      LAUNCH_ON(c, part == 1 ? c->stream2 : c->stream, (force_lj_dealt_kernel<T, EV, UNI, INTEG>), grid, TILE_THREADS,
                qwin_smem_bytes<T>(g.hcap, !UNI, scap), c->x.as<V>(), c->f.as<V>(), g, c->tile_runs.as<int2>(),
                c->tile_center.as<int4>(), c->tile_info.as<int2>(), mirror(c, c->xs_cur), c->xs_types.as<unsigned char>(),
                c->trowsq.as<unsigned long long>(), c->tnum.as<int2>(), c->tcapq, c->nlocal, scap, Q, VP,
                mirror(c, c->xs_cur ^ 1), c->d_ev, list, GI, c->kernel_profile ? c->d_prof : (unsigned long long*)nullptr);
      if (INTEG && GI.start) { c->ghosts_fresh = true; c->fused_halo_steps++; }
      // the epilogue wrote the local atoms' new positions into the other mirror buffer (ghosts follow with the next
      // forward halo, as in x_alt); with a split launch the buffers flip once, after the second part
      if (INTEG && part != 1) c->xs_cur ^= 1;
      return MMD_OK;
    }
    const size_t smem = tile_smem_bytes<T>(g.hcap, !UNI);
    MM(smem_optin(c, force_lj_tile_kernel<T, EV, UNI, INTEG>));
    LAUNCH_SMEM(c, (force_lj_tile_kernel<T, EV, UNI, INTEG>), g.ntiles, TILE_THREADS, smem, c->x.as<V>(), c->f.as<V>(), g,
                c->tile_runs.as<int2>(), c->tile_center.as<int4>(), c->tile_info.as<int2>(), c->tile_slots.as<int>(),
                c->trows.as<unsigned short>(), c->tnum.as<int2>(), c->tcap, c->nlocal, P, VP, c->d_ev,
                c->kernel_profile ? c->d_prof : (unsigned long long*)nullptr);
    return MMD_OK;
  }
  // force of step n + finalIntegrate(n) + initialIntegrate(n+1) in one launch (tile-resident lists only); the new
  // positions land in x_alt, which then becomes x
  static int lj_tile_verlet(mmd_ctx* c, int half, int ev, double dt, double dtforce, double mass) {
    VerletParams<T> VP;
    VP.v = c->v.as<V>(); VP.x_out = c->x_alt.as<V>();
    VP.dt = (T)dt; VP.dtforce = (T)dtforce; VP.mass = (T)mass;
    if (ev) CU(cudaMemsetAsync(c->d_ev, 0, 3 * sizeof(double), c->stream));
    if (c->lj_uniform) MM(ev ? (lj_tile_launch<1, 1, 1>(c, half, VP)) : (lj_tile_launch<0, 1, 1>(c, half, VP)));
    else MM(ev ? (lj_tile_launch<1, 0, 1>(c, half, VP)) : (lj_tile_launch<0, 0, 1>(c, half, VP)));
    std::swap(c->x, c->x_alt);
    return MMD_OK;
  }
  static bool split_usable(mmd_ctx* c) {
    return c->split_enable && c->nranks > 1 && c->split_ready && c->list_tile && c->list_dealt && c->n_interior > 0 &&
           c->stream2 != nullptr;
  }
  // anything that is not a split step waits for the interior kernels still running on stream2
  static int split_join(mmd_ctx* c) {
    if (c->split_active) {
      if (c->ev_int_valid) CU(cudaStreamWaitEvent(c->stream, c->ev_int, 0));
      c->split_active = false;
      c->ev_int_valid = false;
    }
    return MMD_OK;
  }
  // One fused step of several ranks with the forward halo hidden behind the interior tiles:
  //   stream2: [boundary(n-1) done]  interior(n)
  //   stream :  forward halo(n)      [interior(n-1) done]  boundary(n)
  // Both parts read the position buffers of step n and write those of step n+1; the halo writes ghost slots only, which
  // no interior window contains.  (No energies on this path: thermo steps run the single launch.)
  static int lj_split_step(mmd_ctx* c, int half, double dt, double dtforce, double mass) {
    VerletParams<T> VP;
    VP.v = c->v.as<V>(); VP.x_out = c->x_alt.as<V>();
    VP.dt = (T)dt; VP.dtforce = (T)dtforce; VP.mass = (T)mass;
    if (!c->xs_valid) MM(xs_fill(c));
    if (!c->split_active) {  // first split step after other work: stream2 starts behind everything queued so far
      CU(cudaEventRecord(c->ev_bnd, c->stream));
      c->split_active = true;
      c->ev_int_valid = false;
    }
    // the halo reads send-list atoms only, all of them owned by boundary tiles (tile_classify_kernel): it follows the
    // previous boundary kernel in stream order and may overlap the tail of the previous interior kernel
    MM(communicate(c, false));
    MM(phase_mark(c, MMD_PHASE_COMM));
    CU(cudaStreamWaitEvent(c->stream2, c->ev_bnd, 0));
    if (c->lj_uniform) MM((lj_tile_launch<0, 1, 1>(c, half, VP, 1))); else MM((lj_tile_launch<0, 0, 1>(c, half, VP, 1)));
    // boundary windows hold atoms the previous interior kernel moved
    if (c->ev_int_valid) CU(cudaStreamWaitEvent(c->stream, c->ev_int, 0));
    CU(cudaEventRecord(c->ev_int, c->stream2));
    c->ev_int_valid = true;
    if (c->lj_uniform) MM((lj_tile_launch<0, 1, 1>(c, half, VP, 2))); else MM((lj_tile_launch<0, 0, 1>(c, half, VP, 2)));
    CU(cudaEventRecord(c->ev_bnd, c->stream));
    std::swap(c->x, c->x_alt);
    c->split_steps++;
    return MMD_OK;
  }

  // clear_f: zero f[0,nall) first (half lists; skipped by mmd_run when a fused prologue did it)
  static int lj_async(mmd_ctx* c, int half, int gn, int ev, bool clear_f) {
    if (!c->have_lj) return set_err(MMD_ERR_STATE, "force_lj: mmd_force_lj_setup missing");
    if (c->neigh_rows != c->nlocal) return set_err(MMD_ERR_STATE, "force_lj: neighbor list is stale (build first)");
    if (c->list_tile) {
      if ((half != 0) != (c->list_half != 0)) return set_err(MMD_ERR_STATE, "force_lj: list was built for the other neighbor style");
      // ghosts receive no force in this scheme; keep their f at zero so that a following reverse halo is a no-op
      if (half && clear_f && c->nghost > 0)
        CU(cudaMemsetAsync(c->f.as<V>() + c->nlocal, 0, (size_t)c->nghost * sizeof(V), c->stream));
      if (ev) CU(cudaMemsetAsync(c->d_ev, 0, 2 * sizeof(double), c->stream));
      VerletParams<T> none;
      memset(&none, 0, sizeof none);
      if (c->lj_uniform) return ev ? lj_tile_launch<1, 1, 0>(c, half, none) : lj_tile_launch<0, 1, 0>(c, half, none);
      return ev ? lj_tile_launch<1, 0, 0>(c, half, none) : lj_tile_launch<0, 0, 0>(c, half, none);
    }
    if (half && clear_f) CU(cudaMemsetAsync(c->f.p, 0, (size_t)(c->nlocal + c->nghost) * sizeof(V), c->stream));
    if (ev) CU(cudaMemsetAsync(c->d_ev, 0, 2 * sizeof(double), c->stream));
    switch (c->lj_tpa ? c->lj_tpa : (half ? 2 : 4)) {
      case 1: return lj_dispatch<1>(c, half, gn, ev);
      case 2: return lj_dispatch<2>(c, half, gn, ev);
      case 4: return lj_dispatch<4>(c, half, gn, ev);
      case 8: return lj_dispatch<8>(c, half, gn, ev);
      case 16: return lj_dispatch<16>(c, half, gn, ev);
      default: return lj_dispatch<32>(c, half, gn, ev);
    }
  }
  static int read_ev(mmd_ctx* c, int n) {
    CU(cudaMemcpyAsync(c->h_ev, c->d_ev, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return MMD_OK;
  }

  // ---- EAM -----------------------------------------------------------------------------
  static EAMTables<T> eam_tables(mmd_ctx* c) {
    EAMTables<T> E;
    E.rho_val = c->eam_rho_val.as<V>(); E.rho_der = c->eam_rho_der.as<V>();
    E.z2_val = c->eam_z2_val.as<V>(); E.z2_der = c->eam_z2_der.as<V>();
    E.frho_val = c->eam_frho_val.as<V>(); E.frho_der = c->eam_frho_der.as<V>();
    E.cutforcesq_tab = c->eam_cut.as<T>();
    E.cutforcesq = (T)c->eam_cut0;
    E.rdr = (T)c->eam_rdr; E.rdrho = (T)c->eam_rdrho;
    E.nr = c->eam_nr; E.nrho = c->eam_nrho; E.ntypes = c->ntypes;
    return E;
  }
  static int eam_setup(mmd_ctx* c, const void* rhor, const void* z2r, const void* frho, int nr, int nrho, int nr_tot,
                       int nrho_tot, double rdr, double rdrho, const void* cutsq) {
    const int nn = c->ntypes * c->ntypes;
    const T* hr = (const T*)rhor; const T* hz = (const T*)z2r; const T* hf = (const T*)frho; const T* hc = (const T*)cutsq;
    bool uni = true;
    for (int t = 1; t < nn && uni; t++) {
      uni = uni && !memcmp(hr, hr + (size_t)t * nr_tot, sizeof(T) * (size_t)(nr + 1) * 7) &&
            !memcmp(hz, hz + (size_t)t * nr_tot, sizeof(T) * (size_t)(nr + 1) * 7) &&
            !memcmp(hf, hf + (size_t)t * nrho_tot, sizeof(T) * (size_t)(nrho + 1) * 7) && hc[t] == hc[0];
    }
    c->eam_uniform = uni;
    auto repack = [&](const T* src, int n, int tot, int off, int cnt, DevBuf& dst) -> int {
      std::vector<V> h((size_t)nn * (n + 1));
      for (int t = 0; t < nn; t++)
        for (int m = 0; m <= n; m++) {
          V r; r.x = r.y = r.z = r.w = (T)0;
          const T* s = src + (size_t)t * tot + (size_t)m * 7 + off;
          r.x = s[0]; r.y = s[1]; r.z = s[2];
          if (cnt == 4) r.w = s[3];
          h[(size_t)t * (n + 1) + m] = r;
        }
      MM(dst.reserve(h.size() * sizeof(V), c->stream));
      CU(cudaMemcpyAsync(dst.p, h.data(), h.size() * sizeof(V), cudaMemcpyHostToDevice, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      return MMD_OK;
    };
    MM(repack(hr, nr, nr_tot, 3, 4, c->eam_rho_val));
    MM(repack(hr, nr, nr_tot, 0, 3, c->eam_rho_der));
    MM(repack(hz, nr, nr_tot, 3, 4, c->eam_z2_val));
    MM(repack(hz, nr, nr_tot, 0, 3, c->eam_z2_der));
    MM(repack(hf, nrho, nrho_tot, 3, 4, c->eam_frho_val));
    MM(repack(hf, nrho, nrho_tot, 0, 3, c->eam_frho_der));
    MM(c->eam_cut.reserve((size_t)nn * sizeof(T), c->stream));
    CU(cudaMemcpy(c->eam_cut.p, hc, (size_t)nn * sizeof(T), cudaMemcpyHostToDevice));
    {  // pair-split tables of the dealt kernels (type pair 0; used when all pairs share it)
      const int nk = nr + 1, nkp = (nk + 3) & ~3;
      c->eam_nkp = nkp;
      std::vector<T> b1((size_t)nkp * 4, (T)0), b2((size_t)nkp * 7, (T)0);
      for (int m = 0; m < nk; m++) {
        const T* r7 = hr + (size_t)m * 7;
        const T* z7 = hz + (size_t)m * 7;
        b1[(size_t)2 * m + 0] = r7[3]; b1[(size_t)2 * m + 1] = r7[4];                                  // rhoA
        b1[(size_t)2 * nkp + 2 * m + 0] = r7[5]; b1[(size_t)2 * nkp + 2 * m + 1] = r7[6];              // rhoB
        b2[(size_t)2 * m + 0] = r7[0]; b2[(size_t)2 * m + 1] = r7[1];                                  // rdA
        b2[(size_t)2 * nkp + 2 * m + 0] = z7[3]; b2[(size_t)2 * nkp + 2 * m + 1] = z7[4];              // z2A
        b2[(size_t)4 * nkp + 2 * m + 0] = z7[5]; b2[(size_t)4 * nkp + 2 * m + 1] = z7[6];              // z2B
        b2[(size_t)6 * nkp + m] = r7[2];                                                               // rdB
      }
      MM(c->eam_blob1.reserve(b1.size() * sizeof(T), c->stream));
      MM(c->eam_blob2.reserve(b2.size() * sizeof(T), c->stream));
      CU(cudaMemcpy(c->eam_blob1.p, b1.data(), b1.size() * sizeof(T), cudaMemcpyHostToDevice));
      CU(cudaMemcpy(c->eam_blob2.p, b2.data(), b2.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
    c->eam_cut0 = (double)hc[0];
    c->eam_rdr = rdr; c->eam_rdrho = rdrho; c->eam_nr = nr; c->eam_nrho = nrho;
    c->have_eam = true;
    if (c->cap > 0) {
      MM(c->rho.reserve((size_t)c->cap * sizeof(T), c->stream));
      MM(c->fp.reserve((size_t)c->cap * sizeof(T), c->stream));
    }
    return MMD_OK;
  }

  template <int TPA, int EV, int UNI> static int eam_launch(mmd_ctx* c, int half) {
    const EAMTables<T> E = eam_tables(c);
    const int grid = div_up((long long)c->nlocal * TPA, EAM_BLOCK);
    const int nall = c->nlocal + c->nghost;
#define EAM_RHO_ARGS c->x.as<V>(), c->neighbors.as<int>(), c->numneigh.as<int>(), c->neigh_stride, c->nlocal, E, \
                     c->rho.as<T>(), c->fp.as<T>(), c->d_ev + 3
#define EAM_PAIR_ARGS c->x.as<V>(), c->f.as<V>(), c->neighbors.as<int>(), c->numneigh.as<int>(), c->neigh_stride, \
                      c->nlocal, E, c->fp.as<T>(), c->d_ev
    if (half) {
      CU(cudaMemsetAsync(c->f.p, 0, (size_t)nall * sizeof(V), c->stream));
      CU(cudaMemsetAsync(c->rho.p, 0, (size_t)c->nlocal * sizeof(T), c->stream));
      LAUNCH(c, (eam_rho_kernel<T, TPA, 0, EV, UNI>), grid, EAM_BLOCK, EAM_RHO_ARGS);
      LAUNCH(c, (eam_embed_kernel<T, EV, UNI>), div_up(c->nlocal, TPB), TPB, c->x.as<V>(), c->nlocal, E, c->rho.as<T>(),
             c->fp.as<T>(), c->d_ev + 3);
    } else {
      LAUNCH(c, (eam_rho_kernel<T, TPA, 1, EV, UNI>), grid, EAM_BLOCK, EAM_RHO_ARGS);
    }
    MM(forward_scalar(c, c->fp.as<T>()));
    if (half) LAUNCH(c, (eam_pair_kernel<T, TPA, 1, EV, UNI>), grid, EAM_BLOCK, EAM_PAIR_ARGS);
    else LAUNCH(c, (eam_pair_kernel<T, TPA, 0, EV, UNI>), grid, EAM_BLOCK, EAM_PAIR_ARGS);
#undef EAM_RHO_ARGS
#undef EAM_PAIR_ARGS
    return MMD_OK;
  }
  template <int TPA> static int eam_dispatch(mmd_ctx* c, int half, int ev) {
    if (c->eam_uniform) return ev ? eam_launch<TPA, 1, 1>(c, half) : eam_launch<TPA, 0, 1>(c, half);
    return ev ? eam_launch<TPA, 1, 0>(c, half) : eam_launch<TPA, 0, 0>(c, half);
  }
  // tile-resident, bank-dealt list: owner-computes shared-memory passes (tile_eam_dealt.cuh)
  static EAMDealtTabs<T> eam_dealt_tabs(mmd_ctx* c) {
    EAMDealtTabs<T> D;
    D.blob1 = c->eam_blob1.as<unsigned char>();
    D.blob2 = c->eam_blob2.as<unsigned char>();
    D.nkp = c->eam_nkp;
    return D;
  }
  static bool eam_dealt_fits(mmd_ctx* c) {
    const int scap = (c->tile_max_rows + 7) & ~7;
    const EAMDealtTabs<T> D = eam_dealt_tabs(c);
    return c->list_tile && c->list_dealt && c->eam_uniform &&
           eam_dealt_smem_bytes<T>(c->tgeo.hcap, scap, 2, D) <= (size_t)(227 * 1024 - 2048);
  }
  // VP == nullptr: forces to f[]; else the velocity-Verlet halves ride in the pair pass's epilogue
  template <int EV> static int eam_dealt_launch(mmd_ctx* c, int half, const VerletParams<T>* VP) {
    const EAMTables<T> E = eam_tables(c);
    const EAMDealtTabs<T> D = eam_dealt_tabs(c);
    const TileGeo& g = c->tgeo;
    const int scap = (c->tile_max_rows + 7) & ~7;
    const size_t sm1 = eam_dealt_smem_bytes<T>(g.hcap, scap, 1, D), sm2 = eam_dealt_smem_bytes<T>(g.hcap, scap, 2, D);
    if (!c->xs_valid) MM(xs_fill(c));
    MM(c->fp_s.reserve(((size_t)std::max(c->cap, c->nlocal + c->nghost) + 64) * sizeof(T), c->stream));
    // ghosts receive no force in this scheme; keep their f at zero so that a following reverse halo is a no-op
    if (half && !VP && c->nghost > 0) CU(cudaMemsetAsync(c->f.as<V>() + c->nlocal, 0, (size_t)c->nghost * sizeof(V), c->stream));
    VerletParams<T> none;
    memset(&none, 0, sizeof none);
    const XsMirror<T> in = mirror(c, c->xs_cur), out = mirror(c, c->xs_cur ^ 1);
#define EAMD_ARGS(vp)                                                                                                    \
  c->x.as<V>(), c->f.as<V>(), g, c->tile_runs.as<int2>(), c->tile_center.as<int4>(), c->tile_info.as<int2>(), in,         \
      c->trowsq.as<unsigned long long>(), c->tnum.as<int2>(), c->tcapq, c->nlocal, scap, E, D, c->fp.as<T>(),             \
      c->fp_s.as<T>(), vp, out, c->d_ev
    // 512 threads when two CTAs fit an SM, else one CTA of 1024 (the same 32 warps per SM either way)
    if (2 * (sm1 + 1024) <= (size_t)227 * 1024) {
      MM(smem_optin(c, eam_dealt_kernel<T, 1, EV, 0, 512>));
      LAUNCH_SMEM(c, (eam_dealt_kernel<T, 1, EV, 0, 512>), g.ntiles, 512, sm1, EAMD_ARGS(none));
    } else {
      MM(smem_optin(c, eam_dealt_kernel<T, 1, EV, 0, 1024>));
      LAUNCH_SMEM(c, (eam_dealt_kernel<T, 1, EV, 0, 1024>), g.ntiles, 1024, sm1, EAMD_ARGS(none));
    }
    MM(forward_scalar(c, c->fp.as<T>()));
    if (c->nghost > 0)
      LAUNCH(c, fp_mirror_ghosts_kernel<T>, div_up(c->nghost, TPB), TPB, c->fp.as<T>(), c->nlocal, c->nghost,
             c->slot_of.as<int>(), c->fp_s.as<T>());
    if (VP) {
      MM(smem_optin(c, eam_dealt_kernel<T, 2, EV, 1, 1024>));
      LAUNCH_SMEM(c, (eam_dealt_kernel<T, 2, EV, 1, 1024>), g.ntiles, 1024, sm2, EAMD_ARGS(*VP));
      c->xs_cur ^= 1;
    } else {
      MM(smem_optin(c, eam_dealt_kernel<T, 2, EV, 0, 1024>));
      LAUNCH_SMEM(c, (eam_dealt_kernel<T, 2, EV, 0, 1024>), g.ntiles, 1024, sm2, EAMD_ARGS(none));
    }
#undef EAMD_ARGS
    return MMD_OK;
  }
  // EAM force of step n + finalIntegrate(n) + initialIntegrate(n+1); the new positions land in x_alt, which becomes x
  static int eam_dealt_verlet(mmd_ctx* c, int half, int ev, double dt, double dtforce, double mass) {
    VerletParams<T> VP;
    VP.v = c->v.as<V>(); VP.x_out = c->x_alt.as<V>();
    VP.dt = (T)dt; VP.dtforce = (T)dtforce; VP.mass = (T)mass;
    MM(c->rho.reserve((size_t)c->cap * sizeof(T), c->stream));
    MM(c->fp.reserve((size_t)c->cap * sizeof(T), c->stream));
    if (ev) CU(cudaMemsetAsync(c->d_ev, 0, 4 * sizeof(double), c->stream));
    MM(ev ? eam_dealt_launch<1>(c, half, &VP) : eam_dealt_launch<0>(c, half, &VP));
    std::swap(c->x, c->x_alt);
    return MMD_OK;
  }

  static int eam_async(mmd_ctx* c, int half, int ev) {
    if (!c->have_eam) return set_err(MMD_ERR_STATE, "force_eam: mmd_force_eam_setup missing");
    if (c->neigh_rows != c->nlocal) return set_err(MMD_ERR_STATE, "force_eam: neighbor list is stale (build first)");
    MM(c->rho.reserve((size_t)c->cap * sizeof(T), c->stream));
    MM(c->fp.reserve((size_t)c->cap * sizeof(T), c->stream));
    if (ev) CU(cudaMemsetAsync(c->d_ev, 0, 4 * sizeof(double), c->stream));
    if (c->list_tile) {
      if ((half != 0) != (c->list_half != 0)) return set_err(MMD_ERR_STATE, "force_eam: list was built for the other neighbor style");
      if (eam_dealt_fits(c)) return ev ? eam_dealt_launch<1>(c, half, nullptr) : eam_dealt_launch<0>(c, half, nullptr);
      MM(ensure_classic(c));  // window too large for the pair pass / per-type tables: export once, continue on classic rows
    }
    switch (c->eam_tpa) {
      case 1: return eam_dispatch<1>(c, half, ev);
      case 2: return eam_dispatch<2>(c, half, ev);
      case 4: return eam_dispatch<4>(c, half, ev);
      case 8: return eam_dispatch<8>(c, half, ev);
      case 16: return eam_dispatch<16>(c, half, ev);
      default: return eam_dispatch<32>(c, half, ev);
    }
  }
  // eng_vdwl from the device partial sums (see eam_pair_kernel): full = 2*(embed + sum 0.5 phi),
  // half = embed + sum phi
  static double eam_energy(mmd_ctx* c, int half) { return half ? c->h_ev[3] + c->h_ev[0] : 2.0 * (c->h_ev[3] + c->h_ev[0]); }

  // ---- integrate -----------------------------------------------------------------------
  static int initial(mmd_ctx* c, double dt, double dtforce, bool zero_f) {
    const int g = div_up(c->nlocal, TPB);
    if (zero_f) LAUNCH(c, (initial_integrate_kernel<T, 1>), g, TPB, c->x.as<V>(), c->v.as<V>(), c->f.as<V>(), c->nlocal, (T)dt, (T)dtforce);
    else LAUNCH(c, (initial_integrate_kernel<T, 0>), g, TPB, c->x.as<V>(), c->v.as<V>(), c->f.as<V>(), c->nlocal, (T)dt, (T)dtforce);
    return xs_refresh_locals(c);
  }
  static int final_(mmd_ctx* c, double dtforce, bool ke, double mass) {
    const int g = div_up(c->nlocal, TPB);
    if (ke) {
      CU(cudaMemsetAsync(c->d_ev + 2, 0, sizeof(double), c->stream));
      LAUNCH(c, (final_integrate_kernel<T, 1>), g, TPB, c->v.as<V>(), c->f.as<V>(), c->nlocal, (T)dtforce, (T)mass, c->d_ev + 2);
    } else {
      LAUNCH(c, (final_integrate_kernel<T, 0>), g, TPB, c->v.as<V>(), c->f.as<V>(), c->nlocal, (T)dtforce, (T)mass, c->d_ev + 2);
    }
    return MMD_OK;
  }
  static int final_initial(mmd_ctx* c, double dt, double dtforce, bool ke, double mass) {
    const int g = div_up(c->nlocal, TPB);
    if (ke) {
      CU(cudaMemsetAsync(c->d_ev + 2, 0, sizeof(double), c->stream));
      LAUNCH(c, (final_initial_integrate_kernel<T, 1>), g, TPB, c->x.as<V>(), c->v.as<V>(), c->f.as<V>(), c->nlocal, (T)dt, (T)dtforce, (T)mass, c->d_ev + 2);
    } else {
      LAUNCH(c, (final_initial_integrate_kernel<T, 0>), g, TPB, c->x.as<V>(), c->v.as<V>(), c->f.as<V>(), c->nlocal, (T)dt, (T)dtforce, (T)mass, c->d_ev + 2);
    }
    return xs_refresh_locals(c);
  }
  static int sum_mv2(mmd_ctx* c, double mass, double* out) {
    CU(cudaMemsetAsync(c->d_ev + 2, 0, sizeof(double), c->stream));
    const int g = std::max(1, std::min(div_up(c->nlocal, TPB), 148 * 8));
    LAUNCH(c, sum_mv2_kernel<T>, g, TPB, c->v.as<V>(), c->nlocal, (T)mass, c->d_ev + 2);
    MM(read_ev(c, 3));
    *out = c->h_ev[2];
    return MMD_OK;
  }

  // ---- Comm ----------------------------------------------------------------------------
  static SwapPairDev pair_desc(mmd_ctx* c, int w0, int nsw) {
    SwapPairDev sp;
    memset(&sp, 0, sizeof sp);
    for (int s = 0; s < nsw; s++) {
      const int w = w0 + s;
      sp.list[s] = c->sw[w].list.as<int>();
      sp.count[s] = c->sw[w].sendnum;
      sp.first[s] = c->sw[w].firstrecv;
      sp.any[s] = c->swaps.pbc_any[w];
      sp.flag[s][0] = c->swaps.pbc_flagx[w];
      sp.flag[s][1] = c->swaps.pbc_flagy[w];
      sp.flag[s][2] = c->swaps.pbc_flagz[w];
    }
    return sp;
  }
  static bool is_self(mmd_ctx* c, int w) { return c->swaps.sendproc[w] == c->swaps.me; }

#ifdef MMD_WITH_NCCL
  static ncclDataType_t nccl_real() { return sizeof(T) == 8 ? ncclDouble : ncclFloat; }
  // one grouped send+recv (MPI_Sendrecv of ref/comm.cpp:304-306)
  static int sendrecv(mmd_ctx* c, const void* sbuf, size_t scount, int dst, void* rbuf, size_t rcount, int src,
                      ncclDataType_t dt) {
    if (!c->nccl) return set_err(MMD_ERR_STATE, "remote swap requested but mmd_comm_nccl_init was not called");
    NC(ncclGroupStart());
    if (scount) NC(ncclSend(sbuf, scount, dt, dst, c->nccl, c->stream));
    if (rcount) NC(ncclRecv(rbuf, rcount, dt, src, c->nccl, c->stream));
    NC(ncclGroupEnd());
    return MMD_OK;
  }
  // both swaps of a dimension layer in ONE group: 2 sends + 2 receives, buffers [part 0 | part 1]
  static int sendrecv_pair2(mmd_ctx* c, const T* sbuf, size_t s0, size_t s1, int dst0, int dst1, T* r0, T* r1, size_t n0,
                            size_t n1, int src0, int src1) {
    if (!c->nccl) return set_err(MMD_ERR_STATE, "remote swap requested but mmd_comm_nccl_init was not called");
    NC(ncclGroupStart());
    if (s0) NC(ncclSend(sbuf, s0, nccl_real(), dst0, c->nccl, c->stream));
    if (n0) NC(ncclRecv(r0, n0, nccl_real(), src0, c->nccl, c->stream));
    if (s1) NC(ncclSend(sbuf + s0, s1, nccl_real(), dst1, c->nccl, c->stream));
    if (n1) NC(ncclRecv(r1, n1, nccl_real(), src1, c->nccl, c->stream));
    NC(ncclGroupEnd());
    return MMD_OK;
  }
  static int sendrecv_pair(mmd_ctx* c, const T* sbuf, size_t s0, size_t s1, int dst0, int dst1, T* rbuf, size_t n0, size_t n1,
                           int src0, int src1) {
    return sendrecv_pair2(c, sbuf, s0, s1, dst0, dst1, rbuf, rbuf + n0, n0, n1, src0, src1);
  }
#endif

  static int communicate(mmd_ctx* c, bool zero_ghost_f) {
    if (!c->have_comm) return set_err(MMD_ERR_STATE, "communicate: mmd_comm_setup missing");
    const T px = (T)c->prd[0], py = (T)c->prd[1], pz = (T)c->prd[2];
    if (c->ghosts_resolved && !zero_ghost_f) {
      LAUNCH(c, halo_forward_resolved_kernel<T>, div_up(c->nghost, TPB), TPB, c->x.as<V>(), c->nlocal, c->nghost,
             c->ghost_src.as<int>(), c->ghost_shift.as<int>(), px, py, pz, live_mirror(c));
      return MMD_OK;
    }
    for (int w = 0; w < c->swaps.nswap; w += 2) {
      const int nsw = std::min(2, c->swaps.nswap - w);
      if (is_self(c, w) && (nsw < 2 || is_self(c, w + 1))) {
        const SwapPairDev sp = pair_desc(c, w, nsw);
        const int n = sp.count[0] + sp.count[1];
        if (zero_ghost_f) LAUNCH(c, (halo_forward_self_kernel<T, 1>), div_up(n, TPB), TPB, c->x.as<V>(), c->f.as<V>(), sp, px, py, pz, live_mirror(c));
        else LAUNCH(c, (halo_forward_self_kernel<T, 0>), div_up(n, TPB), TPB, c->x.as<V>(), c->f.as<V>(), sp, px, py, pz, live_mirror(c));
      } else {
#ifdef MMD_WITH_NCCL
        if (nsw != 2) return set_err(MMD_ERR_STATE, "communicate: odd swap count");
        const SwapPairDev sp = pair_desc(c, w, 2);
        const int ns0 = c->sw[w].sendnum, ns1 = c->sw[w + 1].sendnum, nr0 = c->sw[w].recvnum, nr1 = c->sw[w + 1].recvnum;
        if (c->p2p_on && c->swaps.nswap <= P2P_SWAPS) {
          // peer-memory path: pack straight into the neighbors' windows, then wait for theirs and unpack
          if ((size_t)std::max(std::max(ns0, ns1), std::max(nr0, nr1)) > P2P_REGION_ATOMS)
            return set_err(MMD_ERR_STATE, "communicate: halo message exceeds the peer window (set option p2p_halo to 0)");
          const unsigned long long epoch = ++c->p2p_epoch[w];
          const int par = (int)(epoch & 1ull);
          unsigned char* pw0 = c->peer_win[c->swaps.sendproc[w]];
          unsigned char* pw1 = c->peer_win[c->swaps.sendproc[w + 1]];
          LAUNCH(c, halo_p2p_send_kernel<T>, std::max(1, div_up(ns0 + ns1, TPB)), TPB, c->x.as<V>(), sp, px, py, pz,
                 reinterpret_cast<T*>(p2p_region(pw0, par, w)), reinterpret_cast<T*>(p2p_region(pw1, par, w + 1)),
                 p2p_flag(pw0, w), p2p_flag(pw1, w + 1), epoch, c->d_done + w / 2);
          const T* s0 = reinterpret_cast<const T*>(p2p_region(c->win, par, w));
          const T* s1 = reinterpret_cast<const T*>(p2p_region(c->win, par, w + 1));
          const int ug = std::max(1, div_up(nr0 + nr1, TPB));
          if (zero_ghost_f) LAUNCH(c, (halo_p2p_unpack_kernel<T, 1>), ug, TPB, c->x.as<V>(), c->f.as<V>(), c->sw[w].firstrecv, nr0, c->sw[w + 1].firstrecv, nr1, s0, s1, p2p_flag(c->win, w), p2p_flag(c->win, w + 1), epoch, c->d_scal + 0, live_mirror(c));
          else LAUNCH(c, (halo_p2p_unpack_kernel<T, 0>), ug, TPB, c->x.as<V>(), c->f.as<V>(), c->sw[w].firstrecv, nr0, c->sw[w + 1].firstrecv, nr1, s0, s1, p2p_flag(c->win, w), p2p_flag(c->win, w + 1), epoch, c->d_scal + 0, live_mirror(c));
          c->p2p_calls++;
          continue;
        }
        MM(c->sendbuf.reserve((size_t)3 * (ns0 + ns1) * sizeof(T), c->stream, 0, 1.5));
        MM(c->recvbuf.reserve((size_t)3 * (nr0 + nr1) * sizeof(T), c->stream, 0, 1.5));
        LAUNCH(c, halo_pack_x_pair_kernel<T>, div_up(ns0 + ns1, TPB), TPB, c->x.as<V>(), sp, px, py, pz, c->sendbuf.as<T>());
        MM(sendrecv_pair(c, c->sendbuf.as<T>(), (size_t)3 * ns0, (size_t)3 * ns1, c->swaps.sendproc[w], c->swaps.sendproc[w + 1],
                         c->recvbuf.as<T>(), (size_t)3 * nr0, (size_t)3 * nr1, c->swaps.recvproc[w], c->swaps.recvproc[w + 1]));
        if (zero_ghost_f) LAUNCH(c, (halo_unpack_x_pair_kernel<T, 1>), div_up(nr0 + nr1, TPB), TPB, c->x.as<V>(), c->f.as<V>(), c->sw[w].firstrecv, nr0, c->sw[w + 1].firstrecv, nr1, c->recvbuf.as<T>(), live_mirror(c));
        else LAUNCH(c, (halo_unpack_x_pair_kernel<T, 0>), div_up(nr0 + nr1, TPB), TPB, c->x.as<V>(), c->f.as<V>(), c->sw[w].firstrecv, nr0, c->sw[w + 1].firstrecv, nr1, c->recvbuf.as<T>(), live_mirror(c));
#else
        return set_err(MMD_ERR_STATE, "remote swap but library built without NCCL");
#endif
      }
    }
    return MMD_OK;
  }

  static int reverse(mmd_ctx* c) {
    if (!c->have_comm) return set_err(MMD_ERR_STATE, "reverse_communicate: mmd_comm_setup missing");
    const int npairs = (c->swaps.nswap + 1) / 2;
    for (int p = npairs - 1; p >= 0; p--) {
      const int w = 2 * p;
      const int nsw = std::min(2, c->swaps.nswap - w);
      if (is_self(c, w) && (nsw < 2 || is_self(c, w + 1))) {
        const SwapPairDev sp = pair_desc(c, w, nsw);
        LAUNCH(c, halo_reverse_self_kernel<T>, div_up(sp.count[0] + sp.count[1], TPB), TPB, c->f.as<V>(), sp);
      } else {
#ifdef MMD_WITH_NCCL
        if (nsw != 2) return set_err(MMD_ERR_STATE, "reverse_communicate: odd swap count");
        const SwapPairDev sp = pair_desc(c, w, 2);
        const int ns0 = c->sw[w].sendnum, ns1 = c->sw[w + 1].sendnum, nr0 = c->sw[w].recvnum, nr1 = c->sw[w + 1].recvnum;
        MM(c->sendbuf.reserve((size_t)3 * (nr0 + nr1) * sizeof(T), c->stream, 0, 1.5));
        MM(c->recvbuf.reserve((size_t)3 * (ns0 + ns1) * sizeof(T), c->stream, 0, 1.5));
        LAUNCH(c, halo_pack_f_pair_kernel<T>, div_up(nr0 + nr1, TPB), TPB, c->f.as<V>(), c->sw[w].firstrecv, nr0,
               c->sw[w + 1].firstrecv, nr1, c->sendbuf.as<T>());
        // ghost forces travel back along the swap: to the rank the ghosts came from, from the rank mine went to
        MM(sendrecv_pair(c, c->sendbuf.as<T>(), (size_t)3 * nr0, (size_t)3 * nr1, c->swaps.recvproc[w], c->swaps.recvproc[w + 1],
                         c->recvbuf.as<T>(), (size_t)3 * ns0, (size_t)3 * ns1, c->swaps.sendproc[w], c->swaps.sendproc[w + 1]));
        LAUNCH(c, halo_unpack_f_pair_kernel<T>, div_up(ns0 + ns1, TPB), TPB, c->f.as<V>(), sp, c->recvbuf.as<T>());
#else
        return set_err(MMD_ERR_STATE, "remote swap but library built without NCCL");
#endif
      }
    }
    return MMD_OK;
  }

  // forward halo of one scalar per atom (EAM fp)
  static int forward_scalar(mmd_ctx* c, T* a) {
    if (!c->have_comm) return set_err(MMD_ERR_STATE, "mmd_comm_setup missing");
    for (int w = 0; w < c->swaps.nswap; w += 2) {
      const int nsw = std::min(2, c->swaps.nswap - w);
      if (is_self(c, w) && (nsw < 2 || is_self(c, w + 1))) {
        const SwapPairDev sp = pair_desc(c, w, nsw);
        LAUNCH(c, halo_forward_scalar_self_kernel<T>, div_up(sp.count[0] + sp.count[1], TPB), TPB, a, sp);
      } else {
#ifdef MMD_WITH_NCCL
        if (nsw != 2) return set_err(MMD_ERR_STATE, "forward_scalar: odd swap count");
        const SwapPairDev sp = pair_desc(c, w, 2);
        const int ns0 = c->sw[w].sendnum, ns1 = c->sw[w + 1].sendnum, nr0 = c->sw[w].recvnum, nr1 = c->sw[w + 1].recvnum;
        MM(c->sendbuf.reserve((size_t)(ns0 + ns1) * sizeof(T), c->stream, 0, 1.5));
        LAUNCH(c, gather_scalar_pair_kernel<T>, div_up(ns0 + ns1, TPB), TPB, a, sp, c->sendbuf.as<T>());
        // ghosts of one swap are contiguous from firstrecv: receive straight into the array
        MM(sendrecv_pair2(c, c->sendbuf.as<T>(), (size_t)ns0, (size_t)ns1, c->swaps.sendproc[w], c->swaps.sendproc[w + 1],
                          a + c->sw[w].firstrecv, a + c->sw[w + 1].firstrecv, (size_t)nr0, (size_t)nr1, c->swaps.recvproc[w],
                          c->swaps.recvproc[w + 1]));
#else
        return set_err(MMD_ERR_STATE, "remote swap but library built without NCCL");
#endif
      }
    }
    return MMD_OK;
  }

  // Comm::borders.  Per swap pair: count -> spine -> (host reads the two totals, grows arrays)
  // -> ordered scatter that writes the send lists and, for self swaps, the ghosts themselves.
  static int borders(mmd_ctx* c) {
    if (!c->have_comm || !c->have_box) return set_err(MMD_ERR_STATE, "borders: mmd_comm_setup / mmd_atom_set_box missing");
    c->nghost = 0;
    c->xs_valid = false;
    const T px = (T)c->prd[0], py = (T)c->prd[1], pz = (T)c->prd[2];
    int w = 0;
    for (int dim = 0; dim < 3; dim++) {
      int nfirst = 0, nlast = 0;
      for (int layer = 0; layer < c->swaps.need[dim]; layer++, w += 2) {
        nfirst = nlast;
        nlast = c->nlocal + c->nghost;
        const int nscan = nlast - nfirst;
        const int ntiles = std::max(1, div_up(nscan, BORDER_THREADS));
        MM(c->border_tiles.reserve((size_t)2 * ntiles * sizeof(int), c->stream, 0, 1.2));
        int* t0 = c->border_tiles.as<int>();
        int* t1 = t0 + ntiles;
        const T lo0 = (T)c->swaps.slablo[w], hi0 = (T)c->swaps.slabhi[w];
        const T lo1 = (T)c->swaps.slablo[w + 1], hi1 = (T)c->swaps.slabhi[w + 1];
        LAUNCH(c, border_count_kernel<T>, ntiles, BORDER_THREADS, c->x.as<V>(), nfirst, nlast, dim, lo0, hi0, lo1, hi1, t0, t1);
        LAUNCH(c, scan_spine_kernel, 2, 1024, t0, ntiles, c->d_scal + 3, ntiles);
        const bool self0 = is_self(c, w), self1 = is_self(c, w + 1);
        int* d_cnt = c->d_scal + 6;  // [8],[9] counts received from the partners of the two swaps
        if (!self0 || !self1) {
#ifdef MMD_WITH_NCCL
          // the send counts travel straight from the device scalars the scan wrote (ref/comm.cpp:822-824): one host
          // round trip below reads them together with the counts received
          if (!c->nccl) return set_err(MMD_ERR_STATE, "remote swap requested but mmd_comm_nccl_init was not called");
          CU(cudaMemsetAsync(d_cnt + 2, 0, 2 * sizeof(int), c->stream));
          NC(ncclGroupStart());
          if (!self0) { NC(ncclSend(c->d_scal + 3, 1, ncclInt, c->swaps.sendproc[w], c->nccl, c->stream));
                        NC(ncclRecv(d_cnt + 2, 1, ncclInt, c->swaps.recvproc[w], c->nccl, c->stream)); }
          if (!self1) { NC(ncclSend(c->d_scal + 4, 1, ncclInt, c->swaps.sendproc[w + 1], c->nccl, c->stream));
                        NC(ncclRecv(d_cnt + 3, 1, ncclInt, c->swaps.recvproc[w + 1], c->nccl, c->stream)); }
          NC(ncclGroupEnd());
#else
          return set_err(MMD_ERR_STATE, "remote swap but library built without NCCL");
#endif
        }
        CU(cudaMemcpyAsync(c->h_scal + 3, c->d_scal + 3, 7 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));  // [3..9]
        CU(cudaStreamSynchronize(c->stream));
        const int ns0 = c->h_scal[3], ns1 = c->h_scal[4];
        MM(c->sw[w].list.reserve((size_t)std::max(ns0, 1) * sizeof(int), c->stream, 0, 1.5));
        MM(c->sw[w + 1].list.reserve((size_t)std::max(ns1, 1) * sizeof(int), c->stream, 0, 1.5));
        const int nr0 = self0 ? ns0 : c->h_scal[8], nr1 = self1 ? ns1 : c->h_scal[9];
        T *b0 = nullptr, *b1 = nullptr;
        if (!self0 || !self1) {
          MM(c->sendbuf.reserve((size_t)4 * (ns0 + ns1) * sizeof(T), c->stream, 0, 1.5));
          b0 = c->sendbuf.as<T>();
          b1 = b0 + (size_t)4 * ns0;
        }
        const int first0 = c->nlocal + c->nghost, first1 = first0 + nr0;
        MM(reserve_atoms(c, first1 + nr1));
        SwapPairDev sp = pair_desc(c, w, 2);
        LAUNCH(c, border_scatter_kernel<T>, ntiles, BORDER_THREADS, c->x.as<V>(), nfirst, nlast, dim, lo0, hi0, lo1, hi1,
               t0, t1, c->sw[w].list.as<int>(), c->sw[w + 1].list.as<int>(), (int)self0, (int)self1, first0, first1, sp,
               px, py, pz, b0, b1);
#ifdef MMD_WITH_NCCL
        if (!self0 || !self1) {
          MM(c->recvbuf.reserve((size_t)4 * (nr0 + nr1) * sizeof(T), c->stream, 0, 1.5));
          T* r0 = c->recvbuf.as<T>();
          T* r1 = r0 + (size_t)4 * nr0;
          NC(ncclGroupStart());
          if (!self0) { if (ns0) NC(ncclSend(b0, (size_t)4 * ns0, nccl_real(), c->swaps.sendproc[w], c->nccl, c->stream));
                        if (nr0) NC(ncclRecv(r0, (size_t)4 * nr0, nccl_real(), c->swaps.recvproc[w], c->nccl, c->stream)); }
          if (!self1) { if (ns1) NC(ncclSend(b1, (size_t)4 * ns1, nccl_real(), c->swaps.sendproc[w + 1], c->nccl, c->stream));
                        if (nr1) NC(ncclRecv(r1, (size_t)4 * nr1, nccl_real(), c->swaps.recvproc[w + 1], c->nccl, c->stream)); }
          NC(ncclGroupEnd());
          if (!self0) LAUNCH(c, border_unpack_kernel<T>, div_up(nr0, TPB), TPB, c->x.as<V>(), first0, nr0, r0);
          if (!self1) LAUNCH(c, border_unpack_kernel<T>, div_up(nr1, TPB), TPB, c->x.as<V>(), first1, nr1, r1);
        }
#endif
        c->sw[w].sendnum = ns0; c->sw[w].recvnum = nr0; c->sw[w].firstrecv = first0;
        c->sw[w + 1].sendnum = ns1; c->sw[w + 1].recvnum = nr1; c->sw[w + 1].firstrecv = first1;
        c->nghost += nr0 + nr1;
      }
    }
    c->neigh_rows = -1;  // lists are stale until the next build
    // single rank: resolve every ghost to its local source so that the per-step forward halo is one launch
    c->ghosts_resolved = false;
    c->images_ready = false;
    c->ghosts_fresh = false;
    {
      bool all_self = c->fuse_halo && c->nghost > 0;
      for (int ws = 0; ws < c->swaps.nswap; ws++) all_self = all_self && is_self(c, ws);
      // one layer of swaps per dimension only: a second layer (box edge < cutneigh) composes shifts of +-2, which the
      // packed 2-bit fields cannot hold and which x + 2*prd would not reproduce bit for bit
      for (int d = 0; d < 3; d++) all_self = all_self && c->swaps.need[d] <= 1;
      if (all_self) {
        MM(c->ghost_src.reserve((size_t)c->nghost * sizeof(int), c->stream, 0, 1.3));
        MM(c->ghost_shift.reserve((size_t)c->nghost * sizeof(int), c->stream, 0, 1.3));
        for (int ws = 0; ws < c->swaps.nswap; ws++) {
          const int any = c->swaps.pbc_any[ws];
          LAUNCH(c, ghost_resolve_kernel, div_up(c->sw[ws].sendnum, TPB), TPB, c->sw[ws].list.as<int>(), c->sw[ws].sendnum,
                 c->sw[ws].firstrecv, c->nlocal, any ? c->swaps.pbc_flagx[ws] : 0, any ? c->swaps.pbc_flagy[ws] : 0,
                 any ? c->swaps.pbc_flagz[ws] : 0, c->ghost_src.as<int>(), c->ghost_shift.as<int>());
        }
        c->ghosts_resolved = true;
        c->images_ready = false;
        if (c->fuse_ghosts && c->nlocal > 0) {
          const int n1 = c->nlocal + 1;
          MM(c->img_start.reserve((size_t)(n1 + 1) * sizeof(int), c->stream, 0, 1.2));
          MM(c->img_cursor.reserve((size_t)(n1 + 1) * sizeof(int), c->stream, 0, 1.2));
          MM(c->img_list.reserve((size_t)c->nghost * sizeof(int2), c->stream, 0, 1.3));
          CU(cudaMemsetAsync(c->img_cursor.p, 0, (size_t)(n1 + 1) * sizeof(int), c->stream));
          LAUNCH(c, ghost_image_count_kernel, div_up(c->nghost, TPB), TPB, c->ghost_src.as<int>(), c->nghost, c->img_cursor.as<int>());
          // exclusive scan of the counts (start[nlocal] = nghost)
          const int ntl = std::max(1, div_up(n1, SCAN_TILE));
          MM(c->tile_sums.reserve((size_t)ntl * sizeof(int), c->stream));
          LAUNCH(c, scan_tile_sums_kernel, ntl, SCAN_THREADS, c->img_cursor.as<int>(), n1, c->tile_sums.as<int>(), (int*)nullptr);
          LAUNCH(c, scan_spine_kernel, 1, 1024, c->tile_sums.as<int>(), ntl, c->img_start.as<int>() + n1, 0);
          LAUNCH(c, scan_apply_kernel, ntl, SCAN_THREADS, c->img_cursor.as<int>(), n1, c->tile_sums.as<int>(), c->img_start.as<int>());
          CU(cudaMemsetAsync(c->img_cursor.p, 0, (size_t)(n1 + 1) * sizeof(int), c->stream));
          LAUNCH(c, ghost_image_fill_kernel, div_up(c->nghost, TPB), TPB, c->ghost_src.as<int>(), c->ghost_shift.as<int>(), c->nghost,
                 c->img_start.as<int>(), c->img_cursor.as<int>(), c->img_list.as<int2>());
          c->images_ready = true;
        }
      }
    }
    // the second position buffer (Atom::sort scratch, target of the fused force+Verlet kernel) gets the ghost records
    // too: remote halo unpacks only refresh x,y,z and rely on the type lane being in place
    if (c->nghost > 0)
      CU(cudaMemcpyAsync(c->x_alt.as<V>() + c->nlocal, c->x.as<V>() + c->nlocal, (size_t)c->nghost * sizeof(V),
                         cudaMemcpyDeviceToDevice, c->stream));
    return MMD_OK;
  }

  // ---- CUDA graph of two plain steps ------------------------------------------------------
  static bool graph_usable(mmd_ctx* c, const mmd_run_params* p) {
    if (!c->graph_steps || c->nranks != 1 || c->fuse_ghosts || c->kernel_profile) return false;
    if (!(c->fuse_force && c->fuse_integrate && c->list_tile && c->neigh_rows == c->nlocal)) return false;
    if (p->force_style != 0 && !eam_dealt_fits(c)) return false;
    if (c->list_dealt && !c->xs_valid) return false;
    for (int w = 0; w < c->swaps.nswap; w++)
      if (!is_self(c, w)) return false;
    return true;
  }
  // the body of a plain step, as the eager loop runs it
  static int plain_step_body(mmd_ctx* c, const mmd_run_params* p) {
    if (!c->ghosts_fresh) MM(communicate(c, false));
    c->ghosts_fresh = false;
    if (p->force_style == 0) MM(lj_tile_verlet(c, p->halfneigh, 0, p->dt, p->dtforce, p->mass));
    else MM(eam_dealt_verlet(c, p->halfneigh, 0, p->dt, p->dtforce, p->mass));
    return MMD_OK;
  }
  static int capture_pair(mmd_ctx* c, const mmd_run_params* p) {
    const long long l0 = c->launches;
    CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
    c->capturing = true;
    int rc = plain_step_body(c, p);
    if (rc == MMD_OK) rc = plain_step_body(c, p);
    c->capturing = false;
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    c->pair_launches = c->launches - l0;
    c->launches = l0;
    if (rc != MMD_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess || !g) return set_err(MMD_ERR_CUDA, "graph capture of two steps: %s", cudaGetErrorString(e));
    bool have = false;
    if (c->pair_exec) {  // same topology as the previous neighbor list's graph: update in place
      cudaGraphExecUpdateResultInfo info;
      have = cudaGraphExecUpdate(c->pair_exec, g, &info) == cudaSuccess;
      if (!have) { cudaGetLastError(); cudaGraphExecDestroy(c->pair_exec); c->pair_exec = nullptr; }
    }
    if (!have) {
      const cudaError_t ei = cudaGraphInstantiate(&c->pair_exec, g, 0);
      if (ei != cudaSuccess) { cudaGraphDestroy(g); c->pair_exec = nullptr; return set_err(MMD_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ei)); }
    }
    cudaGraphDestroy(g);
    c->pair_valid = true;
    c->graph_captures++;
    return MMD_OK;
  }

  // ---- fused time loop -------------------------------------------------------------------
  static int run(mmd_ctx* c, const mmd_run_params* p, mmd_thermo_sample* samples, int max_samples, int* nsamples,
                 float* elapsed_ms) {
    int ns = 0;
    int next_sort = p->sort_every > 0 ? p->sort_every : p->total_steps + 1;
    while (p->sort_every > 0 && next_sort <= p->first_step) next_sort += p->sort_every;
    const bool reverse_needed = p->halfneigh && p->ghost_newton;
    if (elapsed_ms) CU(cudaEventRecord(c->ev0, c->stream));
    MM(phase_mark(c, -1));
    const int last = p->first_step + p->ntimes - 1;
    c->pair_valid = false;  // (dt, masses and list style are the caller's: never reuse a graph across calls)
    // a step with nothing but the forward halo and the fused force / Verlet kernel
    auto plain_step = [&](int m) {
      const int evm = p->thermo_nstat > 0 ? ((m + 1) % p->thermo_nstat == 0) : 0;
      return ((m + 1) % p->neigh_every) != 0 && m != p->first_step && m < last && !evm;
    };
    for (int n = p->first_step; n < p->first_step + p->ntimes; n++) {
      // (from the second step on, initialIntegrate already ran fused with the previous finalIntegrate)
      if (n == p->first_step || !c->fuse_integrate) {
        MM(initial(c, p->dt, p->dtforce, false));
        MM(phase_mark(c, MMD_PHASE_INTEGRATE));
      }
      const int ev = p->thermo_nstat > 0 ? ((n + 1) % p->thermo_nstat == 0) : 0;
      // several ranks, dealt LJ lists, no energies wanted: the forward halo of this step runs behind the interior tiles
      const bool do_split = ((n + 1) % p->neigh_every) != 0 && n != p->first_step && n < last && !ev && p->force_style == 0 &&
                            c->fuse_force && c->fuse_integrate && c->neigh_rows == c->nlocal && split_usable(c);
      if (do_split) {
        MM(lj_split_step(c, p->halfneigh, p->dt, p->dtforce, p->mass));
        MM(phase_mark(c, MMD_PHASE_FORCE));
        continue;
      }
      MM(split_join(c));
      // two plain steps in a row on one rank: one graph launch (captured once per neighbor list)
      if (plain_step(n) && plain_step(n + 1) && graph_usable(c, p)) {
        if (!c->pair_valid) MM(capture_pair(c, p));
        CU(cudaGraphLaunch(c->pair_exec, c->stream));
        c->launches += c->pair_launches;
        c->graph_replays++;
        c->ghosts_fresh = false;
        MM(phase_mark(c, MMD_PHASE_FORCE, 2));  // (halo, force and Verlet halves of both steps)
        n++;
        continue;
      }
      c->pair_valid = false;  // an eager step changes the buffer orientation or the lists
      if ((n + 1) % p->neigh_every) {
        // (the fused force kernel of the previous step may have written this step's ghosts already)
        if (!c->ghosts_fresh) MM(communicate(c, false));
        c->ghosts_fresh = false;
        MM(phase_mark(c, MMD_PHASE_COMM));
      } else {
        c->ghosts_fresh = false;
        MM(exchange(c));
        if (n + 1 >= next_sort) {
          MM(sort(c));
          next_sort += p->sort_every;
        }
        MM(borders(c));
        MM(phase_mark(c, MMD_PHASE_COMM));
        int mx = c->maxneighs;
        MM(build(c, p->halfneigh, p->ghost_newton, &mx, nullptr));
        MM(phase_mark(c, MMD_PHASE_NEIGH));
      }
      // tile-resident lists: every atom's force is complete after the kernel -- nothing to clear, nothing to send back,
      // and the two velocity-Verlet halves that follow ride in the kernel's epilogue
      const bool verlet_fused = c->fuse_force && c->fuse_integrate && c->list_tile && n < last &&
                                (p->force_style == 0 || eam_dealt_fits(c));
      if (verlet_fused) {
        if (c->neigh_rows != c->nlocal) return set_err(MMD_ERR_STATE, "run: neighbor list is stale (build first)");
        if (p->force_style == 0) MM(lj_tile_verlet(c, p->halfneigh, ev, p->dt, p->dtforce, p->mass));
        else MM(eam_dealt_verlet(c, p->halfneigh, ev, p->dt, p->dtforce, p->mass));
      } else if (p->force_style == 0) MM(lj_async(c, p->halfneigh, p->ghost_newton, ev, !c->list_tile));
      else MM(eam_async(c, p->halfneigh, ev));
      MM(phase_mark(c, MMD_PHASE_FORCE));
      if (reverse_needed && !c->list_tile) {
        MM(reverse(c));
        MM(phase_mark(c, MMD_PHASE_COMM));
      }
      if (verlet_fused) { /* done inside the force kernel */ }
      else if (n < last && c->fuse_integrate) MM(final_initial(c, p->dt, p->dtforce, ev != 0, p->mass));
      else MM(final_(c, p->dtforce, ev != 0, p->mass));
      MM(phase_mark(c, MMD_PHASE_INTEGRATE));
      if (ev) {
        MM(read_ev(c, 4));
        if (ns < max_samples && samples) {
          samples[ns].step = n + 1;
          samples[ns].sum_mv2 = c->h_ev[2];
          samples[ns].eng_vdwl = p->force_style == 0 ? c->h_ev[0] : eam_energy(c, p->halfneigh);
          samples[ns].virial = c->h_ev[1];
        }
        ns++;
      }
    }
    MM(split_join(c));
    c->ghosts_fresh = false;
    if (elapsed_ms) {
      CU(cudaEventRecord(c->ev1, c->stream));
      CU(cudaEventSynchronize(c->ev1));
      CU(cudaEventElapsedTime(elapsed_ms, c->ev0, c->ev1));
    }
    MM(phase_collect(c));
    if (nsamples) *nsamples = ns;
    return MMD_OK;
  }

  // exclusive scan of flags[0,n) into pos[0,n] (pos[n] = total); the total also lands in h_scal[5]
  // total == nullptr: no host round trip, the total stays on the device in pos[n]
  static int scan_flags(mmd_ctx* c, const int* flags, int n, int* pos, int* total) {
    const int ntiles = std::max(1, div_up(n, SCAN_TILE));
    MM(c->tile_sums.reserve((size_t)ntiles * sizeof(int), c->stream));
    LAUNCH(c, scan_tile_sums_kernel, ntiles, SCAN_THREADS, flags, n, c->tile_sums.as<int>(), (int*)nullptr);
    LAUNCH(c, scan_spine_kernel, 1, 1024, c->tile_sums.as<int>(), ntiles, pos + n, 0);
    LAUNCH(c, scan_apply_kernel, ntiles, SCAN_THREADS, flags, n, c->tile_sums.as<int>(), pos);
    if (!total) return MMD_OK;
    CU(cudaMemcpyAsync(c->h_scal + 5, pos + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *total = c->h_scal[5];
    return MMD_OK;
  }

  // Comm::exchange (ref/comm.cpp:364-597): wrap, then per decomposed dimension send every atom that left
  // my sub-box to both neighbours and keep the arrivals that belong to me.
  static int exchange(mmd_ctx* c) {
    MM(pbc(c));
    c->nghost = 0;  // ghosts are rebuilt by borders(); nlocal may change below
    c->ghosts_resolved = false;
    for (int d = 0; d < 3; d++) {
      if (c->swaps.procgrid[d] <= 1) continue;
#ifdef MMD_WITH_NCCL
      if (!c->nccl) return set_err(MMD_ERR_STATE, "exchange across ranks requested but mmd_comm_nccl_init was not called");
      const T lo = (T)c->lo[d], hi = (T)c->hi[d];
      const int n = c->nlocal;
      const int lower = c->swaps.procneigh[d][0], upper = c->swaps.procneigh[d][1];
      const bool both = c->swaps.procgrid[d] > 2;
      MM(c->exch_flag.reserve((size_t)(n + 2) * sizeof(int), c->stream, 0, 1.2));
      MM(c->exch_pos.reserve((size_t)(n + 2) * sizeof(int), c->stream, 0, 1.2));
      // my count stays on the device (exch_pos[n], written by the scan) and travels from there; ONE host round trip
      // then reads it together with the counts the neighbours sent (ref/comm.cpp:521-530)
      int* d_cnt = c->d_scal + 6;  // [6] my count, [8] from upper, [9] from lower
      CU(cudaMemsetAsync(d_cnt, 0, 4 * sizeof(int), c->stream));
      if (n > 0) {
        LAUNCH(c, exch_flag_kernel<T>, div_up(n, TPB), TPB, c->x.as<V>(), n, d, lo, hi, c->exch_flag.as<int>());
        MM(scan_flags(c, c->exch_flag.as<int>(), n, c->exch_pos.as<int>(), nullptr));
        CU(cudaMemcpyAsync(d_cnt, c->exch_pos.as<int>() + n, sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
      }
      NC(ncclGroupStart());
      NC(ncclSend(d_cnt, 1, ncclInt, lower, c->nccl, c->stream));
      NC(ncclRecv(d_cnt + 2, 1, ncclInt, upper, c->nccl, c->stream));
      if (both) {
        NC(ncclSend(d_cnt, 1, ncclInt, upper, c->nccl, c->stream));
        NC(ncclRecv(d_cnt + 3, 1, ncclInt, lower, c->nccl, c->stream));
      }
      NC(ncclGroupEnd());
      CU(cudaMemcpyAsync(c->h_scal + 6, d_cnt, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      const int nsend = c->h_scal[6];
      const int nkeep = n - nsend;
      if (nsend > 0) {
        MM(c->sendbuf.reserve((size_t)7 * nsend * sizeof(T), c->stream, 0, 1.5));
        MM(c->exch_holes.reserve((size_t)nsend * sizeof(int), c->stream, 0, 1.5));
        LAUNCH(c, exch_pack_kernel<T>, div_up(n, TPB), TPB, c->x.as<V>(), c->v.as<V>(), n, c->exch_flag.as<int>(),
               c->exch_pos.as<int>(), nkeep, c->sendbuf.as<T>(), c->exch_holes.as<int>());
        if (nkeep < n)
          LAUNCH(c, exch_fill_kernel<T>, div_up(n - nkeep, TPB), TPB, c->x.as<V>(), c->v.as<V>(), n, c->exch_flag.as<int>(),
                 c->exch_pos.as<int>(), nkeep, c->exch_holes.as<int>());
      }
      c->nlocal = nkeep;
      c->exch_sent += nsend;
      const int nrecv1 = c->h_scal[8], nrecv2 = both ? c->h_scal[9] : 0;
      const int nrecv = nrecv1 + nrecv2;
      MM(c->recvbuf.reserve((size_t)7 * std::max(nrecv, 1) * sizeof(T), c->stream, 0, 1.5));
      T* r1 = c->recvbuf.as<T>();
      T* r2 = r1 + (size_t)7 * nrecv1;
      NC(ncclGroupStart());
      if (nsend) NC(ncclSend(c->sendbuf.p, (size_t)7 * nsend, nccl_real(), lower, c->nccl, c->stream));
      if (nrecv1) NC(ncclRecv(r1, (size_t)7 * nrecv1, nccl_real(), upper, c->nccl, c->stream));
      if (both) {
        if (nsend) NC(ncclSend(c->sendbuf.p, (size_t)7 * nsend, nccl_real(), upper, c->nccl, c->stream));
        if (nrecv2) NC(ncclRecv(r2, (size_t)7 * nrecv2, nccl_real(), lower, c->nccl, c->stream));
      }
      NC(ncclGroupEnd());
      if (nrecv > 0) {
        MM(c->exch_flag.reserve((size_t)(nrecv + 2) * sizeof(int), c->stream, 0, 1.2));
        MM(c->exch_pos.reserve((size_t)(nrecv + 2) * sizeof(int), c->stream, 0, 1.2));
        LAUNCH(c, exch_recv_flag_kernel<T>, div_up(nrecv, TPB), TPB, c->recvbuf.as<T>(), nrecv, d, lo, hi,
               c->exch_flag.as<int>());
        int nmine = 0;
        MM(scan_flags(c, c->exch_flag.as<int>(), nrecv, c->exch_pos.as<int>(), &nmine));
        if (nmine > 0) {
          MM(reserve_atoms(c, c->nlocal + nmine));
          LAUNCH(c, exch_unpack_kernel<T>, div_up(nrecv, TPB), TPB, c->recvbuf.as<T>(), nrecv, c->exch_flag.as<int>(),
                 c->exch_pos.as<int>(), c->nlocal, c->x.as<V>(), c->v.as<V>());
          c->nlocal += nmine;
          c->exch_received += nmine;
        }
      }
#else
      return set_err(MMD_ERR_STATE, "exchange across ranks requested but the library was built without NCCL");
#endif
    }
    c->neigh_rows = -1;
    return MMD_OK;
  }
};

// ---------------------------------------------------------------------------------------
// extern "C" surface
// ---------------------------------------------------------------------------------------
#define DISPATCH(ctx, expr_d, expr_f) ((ctx)->prec == 8 ? (expr_d) : (expr_f))
#define CHECK_CTX(ctx)                                           \
  do {                                                           \
    if (!(ctx)) return set_err(MMD_ERR_ARG, "null context");     \
    cudaError_t e_ = cudaSetDevice((ctx)->device);               \
    if (e_ != cudaSuccess) return set_err(MMD_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e_)); \
  } while (0)

template <class T> static bool all_equal(const T* a, int n) {
  for (int i = 1; i < n; i++)
    if (a[i] != a[0]) return false;
  return true;
}

#ifdef MMD_WITH_NCCL
// Map every rank's receive window into every other rank (CUDA IPC, one node).  All-or-nothing: if any rank cannot
// export or import a window, all ranks keep the NCCL path (p2p_on stays false, query "p2p_active").
static int p2p_setup(mmd_ctx* c) {
  const int n = c->nranks;
  struct Slot { cudaIpcMemHandle_t h; int ok; int pad[15]; };  // 128 bytes
  static_assert(sizeof(Slot) == 128, "slot size");
  Slot mine;
  memset(&mine, 0, sizeof mine);
  bool ok = cudaMalloc(&c->win, P2P_WIN_BYTES) == cudaSuccess;
  if (ok) ok = cudaMemset(c->win, 0, P2P_WIN_BYTES) == cudaSuccess;
  if (ok) ok = cudaMalloc(&c->d_done, 16 * sizeof(unsigned)) == cudaSuccess && cudaMemset(c->d_done, 0, 16 * sizeof(unsigned)) == cudaSuccess;
  if (ok) ok = cudaIpcGetMemHandle(&mine.h, c->win) == cudaSuccess;
  cudaGetLastError();
  mine.ok = ok ? 1 : 0;
  Slot* d_slots = nullptr;
  CU(cudaMalloc(&d_slots, (size_t)n * sizeof(Slot)));
  CU(cudaMemcpy(d_slots + c->rank, &mine, sizeof(Slot), cudaMemcpyHostToDevice));
  NC(ncclAllGather(d_slots + c->rank, d_slots, sizeof(Slot), ncclChar, c->nccl, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  std::vector<Slot> all(n);
  CU(cudaMemcpy(all.data(), d_slots, (size_t)n * sizeof(Slot), cudaMemcpyDeviceToHost));
  for (int r = 0; r < n; r++) ok = ok && all[r].ok == 1;
  c->peer_win.assign(n, nullptr);
  if (ok) {
    for (int r = 0; r < n && ok; r++) {
      if (r == c->rank) { c->peer_win[r] = c->win; continue; }
      void* q = nullptr;
      if (cudaIpcOpenMemHandle(&q, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); }
      c->peer_win[r] = (unsigned char*)q;
    }
  }
  // second round: did everybody import everything?
  int* d_flag = reinterpret_cast<int*>(d_slots);
  const int v = ok ? 1 : 0;
  CU(cudaMemcpy(d_flag, &v, sizeof(int), cudaMemcpyHostToDevice));
  NC(ncclAllReduce(d_flag, d_flag, 1, ncclInt, ncclMin, c->nccl, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  int all_ok = 0;
  CU(cudaMemcpy(&all_ok, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
  CU(cudaFree(d_slots));
  c->p2p_on = all_ok == 1;
  return MMD_OK;
}
#endif

extern "C" {

const char* mmd_last_error(void) { return g_err; }
int mmd_abi_version(void) { return MMD_ABI_VERSION; }
int mmd_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int mmd_ctx_create(int device, int precision_bytes, int ntypes, void* stream, mmd_ctx** out) {
  if (!out) return set_err(MMD_ERR_ARG, "out is null");
  *out = nullptr;
  if (precision_bytes != 4 && precision_bytes != 8) return set_err(MMD_ERR_ARG, "precision_bytes must be 4 or 8");
  if (ntypes < 1) return set_err(MMD_ERR_ARG, "ntypes must be >= 1");
  int ndev = mmd_device_count();
  if (ndev <= 0) return set_err(MMD_ERR_NODEVICE, "no CUDA device visible: minimd_b200 has no CPU path");
  if (device < 0 || device >= ndev) return set_err(MMD_ERR_ARG, "device %d out of range (%d visible)", device, ndev);
  CU(cudaSetDevice(device));
  mmd_ctx* c = new mmd_ctx();
  c->device = device;
  c->prec = precision_bytes;
  c->ntypes = ntypes;
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    // highest priority: with several ranks the halo and boundary kernels launched here must get SM slots ahead of the
    // interior kernels queued on the low-priority second stream (CTA-granular scheduling, nothing is preempted)
    int pr_least = 0, pr_greatest = 0;
    CU(cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest));
    CU(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, pr_greatest));
    c->own_stream = true;
  }
  CU(cudaMalloc(&c->d_scal, 32 * sizeof(int)));
  CU(cudaMalloc(&c->d_total, sizeof(unsigned long long)));
  CU(cudaMalloc(&c->d_ev, 32 * sizeof(double)));
  CU(cudaMemset(c->d_scal, 0, 32 * sizeof(int)));
  CU(cudaMemset(c->d_total, 0, sizeof(unsigned long long)));
  CU(cudaMemset(c->d_ev, 0, 32 * sizeof(double)));
  CU(cudaMallocHost(&c->h_scal, 32 * sizeof(int)));
  CU(cudaMallocHost(&c->h_total, sizeof(unsigned long long)));
  CU(cudaMallocHost(&c->h_ev, 4 * sizeof(double)));
  memset(c->h_scal, 0, 32 * sizeof(int));
  CU(cudaEventCreate(&c->ev0));
  CU(cudaEventCreate(&c->ev1));
  memset(&c->swaps, 0, sizeof c->swaps);
  *out = c;
  return MMD_OK;
}

int mmd_ctx_destroy(mmd_ctx* c) {
  if (!c) return MMD_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  DevBuf* bufs[] = {&c->x, &c->v, &c->f, &c->x_alt, &c->v_alt, &c->stage_a, &c->stage_b, &c->stage_c, &c->stage_i,
                    &c->runs, &c->cutneighsq, &c->atom_bin, &c->bin_atoms, &c->bincount, &c->bin_start, &c->cursor,
                    &c->tile_sums, &c->numneigh, &c->neighbors, &c->lj_cut, &c->lj_s6, &c->lj_eps, &c->eam_rho_val,
                    &c->eam_rho_der, &c->eam_z2_val, &c->eam_z2_der, &c->eam_frho_val, &c->eam_frho_der, &c->eam_cut,
                    &c->rho, &c->fp, &c->border_tiles, &c->sendbuf, &c->recvbuf, &c->exch_flag, &c->exch_pos,
                    &c->exch_holes, &c->ghost_src, &c->ghost_shift, &c->sruns, &c->tile_runs, &c->tile_center, &c->tile_info, &c->tile_slots, &c->tile_oslot, &c->trows,
                    &c->tnum, &c->trowsq, &c->xs_rec[0], &c->xs_rec[1], &c->xs_z[0], &c->xs_z[1], &c->slot_of, &c->xs_types, &c->eam_blob1, &c->eam_blob2, &c->fp_s, &c->tile_split, &c->send_flag, &c->img_start, &c->img_cursor, &c->img_list};
  if (c->pair_exec) cudaGraphExecDestroy(c->pair_exec);
  for (DevBuf* b : bufs) b->release();
  for (int w = 0; w < MMD_MAX_SWAPS; w++) c->sw[w].list.release();
  for (int r = 0; r < (int)c->peer_win.size(); r++)
    if (c->peer_win[r] && c->peer_win[r] != c->win) cudaIpcCloseMemHandle(c->peer_win[r]);
  if (c->win) cudaFree(c->win);
  if (c->d_done) cudaFree(c->d_done);
  if (c->d_prof) cudaFree(c->d_prof);
#ifdef MMD_WITH_NCCL
  if (c->nccl) ncclCommDestroy(c->nccl);
#endif
  cudaFree(c->d_scal); cudaFree(c->d_total); cudaFree(c->d_ev);
  cudaFreeHost(c->h_scal); cudaFreeHost(c->h_total); cudaFreeHost(c->h_ev);
  cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
  if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
  if (c->ev_int) cudaEventDestroy(c->ev_int);
  if (c->ev_bnd) cudaEventDestroy(c->ev_bnd);
  for (cudaEvent_t e : c->marks) cudaEventDestroy(e);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return MMD_OK;
}

int mmd_ctx_sync(mmd_ctx* c) {
  CHECK_CTX(c);
  CU(cudaStreamSynchronize(c->stream));
  return MMD_OK;
}
void* mmd_ctx_stream(mmd_ctx* c) { return c ? (void*)c->stream : nullptr; }
long long mmd_ctx_launches(mmd_ctx* c) { return c ? c->launches : 0; }

// ---- Atom ------------------------------------------------------------------------------
int mmd_atom_set_box(mmd_ctx* c, const double prd[3], const double lo[3], const double hi[3]) {
  CHECK_CTX(c);
  for (int d = 0; d < 3; d++) {
    if (!(prd[d] > 0)) return set_err(MMD_ERR_ARG, "box length must be positive");
    c->prd[d] = prd[d]; c->lo[d] = lo[d]; c->hi[d] = hi[d];
  }
  c->have_box = true;
  return MMD_OK;
}
int mmd_atom_upload(mmd_ctx* c, const void* x, const void* v, const int* type, int nlocal, int pad) {
  CHECK_CTX(c);
  if (!x || !v || nlocal < 0 || (pad != 3 && pad != 4)) return set_err(MMD_ERR_ARG, "atom_upload: bad arguments");
  return DISPATCH(c, Impl<double>::upload(c, x, v, type, nlocal, pad), Impl<float>::upload(c, x, v, type, nlocal, pad));
}
int mmd_atom_split(mmd_ctx* c, int nlocal) {
  CHECK_CTX(c);
  if (nlocal < 0 || nlocal > c->nlocal || c->nghost != 0) return set_err(MMD_ERR_ARG, "atom_split: call right after mmd_atom_upload with nlocal <= uploaded atoms");
  c->nghost = c->nlocal - nlocal;
  c->nlocal = nlocal;
  return MMD_OK;
}
int mmd_atom_update(mmd_ctx* c, const void* x, const void* v, int first, int count, int pad) {
  CHECK_CTX(c);
  if (pad != 3 && pad != 4) return set_err(MMD_ERR_ARG, "pad must be 3 or 4");
  return DISPATCH(c, Impl<double>::update(c, x, v, first, count, pad), Impl<float>::update(c, x, v, first, count, pad));
}
int mmd_atom_download(mmd_ctx* c, void* x, void* v, void* f, int* type, int first, int count, int pad) {
  CHECK_CTX(c);
  if (pad != 3 && pad != 4) return set_err(MMD_ERR_ARG, "pad must be 3 or 4");
  return DISPATCH(c, Impl<double>::download(c, x, v, f, type, first, count, pad),
                  Impl<float>::download(c, x, v, f, type, first, count, pad));
}
int mmd_atom_counts(mmd_ctx* c, int* nlocal, int* nghost, int* nmax) {
  if (!c) return set_err(MMD_ERR_ARG, "null context");
  if (nlocal) *nlocal = c->nlocal;
  if (nghost) *nghost = c->nghost;
  if (nmax) *nmax = c->cap;
  return MMD_OK;
}
int mmd_atom_pbc(mmd_ctx* c) {
  CHECK_CTX(c);
  return DISPATCH(c, Impl<double>::pbc(c), Impl<float>::pbc(c));
}
int mmd_atom_sort(mmd_ctx* c) {
  CHECK_CTX(c);
  MM(DISPATCH(c, Impl<double>::sort(c), Impl<float>::sort(c)));
  return DISPATCH(c, Impl<double>::read_status(c), Impl<float>::read_status(c));
}

// ---- Neighbor --------------------------------------------------------------------------
int mmd_neigh_setup(mmd_ctx* c, const mmd_bin_geometry* g, const int* stencil, int nstencil, const void* cutneighsq) {
  CHECK_CTX(c);
  if (!g || !stencil || nstencil <= 0 || !cutneighsq) return set_err(MMD_ERR_ARG, "neigh_setup: bad arguments");
  if (g->mbinx <= 0 || g->mbiny <= 0 || g->mbinz <= 0) return set_err(MMD_ERR_ARG, "neigh_setup: bad bin counts");
  const long long mb = (long long)g->mbinx * g->mbiny * g->mbinz;
  if (mb > 0x7fffffff - 2) return set_err(MMD_ERR_ARG, "neigh_setup: too many bins");
  c->geo = *g;
  c->mbins = (int)mb;
  c->stencil.assign(stencil, stencil + nstencil);
  // decompose the stencil into runs of consecutive bin offsets, preserving its order
  std::vector<int2> runs;
  for (int k = 0; k < nstencil; k++) {
    if (!runs.empty() && runs.back().x + runs.back().y == stencil[k]) runs.back().y++;
    else runs.push_back(make_int2(stencil[k], 1));
  }
  c->nruns = (int)runs.size();
  MM(c->runs.reserve(runs.size() * sizeof(int2), c->stream));
  CU(cudaMemcpy(c->runs.p, runs.data(), runs.size() * sizeof(int2), cudaMemcpyHostToDevice));
  const size_t nn = (size_t)c->ntypes * c->ntypes * c->prec;
  MM(c->cutneighsq.reserve(nn, c->stream));
  CU(cudaMemcpy(c->cutneighsq.p, cutneighsq, nn, cudaMemcpyHostToDevice));
  c->h_cutneighsq.resize((size_t)c->ntypes * c->ntypes);
  for (size_t k = 0; k < c->h_cutneighsq.size(); k++)
    c->h_cutneighsq[k] = c->prec == 8 ? ((const double*)cutneighsq)[k] : (double)((const float*)cutneighsq)[k];
  MM(c->bincount.reserve((size_t)(mb + 1) * sizeof(int), c->stream));
  MM(c->bin_start.reserve((size_t)(mb + 1) * sizeof(int), c->stream));
  MM(c->cursor.reserve((size_t)(mb + 1) * sizeof(int), c->stream));
  c->have_geo = true;
  c->neigh_rows = -1;
  c->list_tile = false;
  // ---- tile-resident lists: symmetric closure of the stencil (a half stencil holds only the "upper" bins,
  // ref/neighbor.cpp:424-440), decoded into (dz,dy,dx) and merged into runs along x.  Linear offsets order
  // exactly like the reference's (k,j,i) loops because |dx| < mbinx/2 and |dy| < mbiny/2.
  c->tile_ok = false;
  {
    std::vector<int> full(stencil, stencil + nstencil);
    for (int k = 0; k < nstencil; k++) full.push_back(-stencil[k]);
    std::sort(full.begin(), full.end());
    full.erase(std::unique(full.begin(), full.end()), full.end());
    const int mx = g->mbinx, my = g->mbiny;
    auto rnd_div = [](int a, int b) { return (a >= 0 ? a + b / 2 : a - b / 2) / b; };
    int sx = 0, sy = 0, sz = 0;
    bool ok = true;
    std::vector<StencilRun> sr;
    for (int o : full) {
      const int dz = rnd_div(o, mx * my);
      const int rem = o - dz * mx * my;
      const int dy = rnd_div(rem, mx);
      const int dx = rem - dy * mx;
      if (2 * std::abs(dx) >= mx || 2 * std::abs(dy) >= my) ok = false;
      sx = std::max(sx, std::abs(dx)); sy = std::max(sy, std::abs(dy)); sz = std::max(sz, std::abs(dz));
      if (!sr.empty() && sr.back().off + sr.back().len == o && sr.back().dy == dy && sr.back().dz == dz) sr.back().len++;
      else sr.push_back(StencilRun{o, 1, dy, dz});
    }
    TileGeo t;
    t.mbx = g->mbinx; t.mby = g->mbiny; t.mbz = g->mbinz;
    t.ox = t.oy = t.oz = 0;  // origin and tile counts are fixed at build time (they depend on the sub-box)
    t.ntx = (t.mbx + TBX - 1) / TBX; t.nty = (t.mby + TBY - 1) / TBY; t.ntz = (t.mbz + TBZ - 1) / TBZ;
    t.sx = sx; t.sy = sy; t.sz = sz;
    t.nry = TBY + 2 * sy; t.nrz = TBZ + 2 * sz; t.nrun = t.nry * t.nrz;
    t.ntiles = t.ntx * t.nty * t.ntz;
    t.hcap = 0;  // sized at every build from the largest halo window
    if (t.nrun > TILE_MAXRUN) ok = false;
    if (ok) {
      MM(c->sruns.reserve(sr.size() * sizeof(StencilRun), c->stream));
      CU(cudaMemcpy(c->sruns.p, sr.data(), sr.size() * sizeof(StencilRun), cudaMemcpyHostToDevice));
      c->nsruns = (int)sr.size();
      c->tgeo = t;
      c->tile_ok = true;
    }
  }
  return MMD_OK;
}

static int ref_atoms_per_bin(int current, int max_count) {  // ref/neighbor.cpp:257-261
  int apb = current > 0 ? current : 8;
  while (max_count > apb) apb *= 2;
  return apb;
}

int mmd_neigh_binatoms(mmd_ctx* c, int count, int* atoms_per_bin, int* max_count) {
  CHECK_CTX(c);
  const int n = count < 0 ? c->nlocal + c->nghost : count;
  if (n > c->nlocal + c->nghost) return set_err(MMD_ERR_ARG, "binatoms: count exceeds atoms");
  MM(DISPATCH(c, Impl<double>::binatoms_async(c, n), Impl<float>::binatoms_async(c, n)));
  MM(DISPATCH(c, Impl<double>::read_status(c), Impl<float>::read_status(c)));
  if (max_count) *max_count = c->max_bin_seen;
  if (atoms_per_bin) *atoms_per_bin = ref_atoms_per_bin(*atoms_per_bin, c->max_bin_seen);
  return MMD_OK;
}

int mmd_neigh_build(mmd_ctx* c, int halfneigh, int ghost_newton, int* maxneighs, long long* total) {
  CHECK_CTX(c);
  if (!c->have_geo) return set_err(MMD_ERR_STATE, "neigh_build: mmd_neigh_setup missing");
  return DISPATCH(c, Impl<double>::build(c, halfneigh, ghost_newton, maxneighs, total),
                  Impl<float>::build(c, halfneigh, ghost_newton, maxneighs, total));
}

int mmd_neigh_download(mmd_ctx* c, int* numneigh, int* neighbors, int nrows, int maxneighs) {
  CHECK_CTX(c);
  if (nrows < 0 || nrows > c->neigh_rows) return set_err(MMD_ERR_ARG, "neigh_download: nrows exceeds built rows");
  if (numneigh) CU(cudaMemcpyAsync(numneigh, c->numneigh.p, (size_t)nrows * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (neighbors) {
    if (maxneighs < 1) return set_err(MMD_ERR_ARG, "neigh_download: maxneighs");
    MM(DISPATCH(c, Impl<double>::export_rows(c), Impl<float>::export_rows(c)));
    const int ncopy = std::min(maxneighs, c->maxneighs);
    MM(c->stage_i.reserve((size_t)nrows * maxneighs * sizeof(int), c->stream));
    CU(cudaMemsetAsync(c->stage_i.p, 0xff, (size_t)nrows * maxneighs * sizeof(int), c->stream));
    const long long tot = (long long)nrows * ncopy;
    LAUNCH(c, rows_restride_kernel, div_up(tot, TPB), TPB, c->neighbors.as<int>(), nrows, c->neigh_stride, maxneighs,
           ncopy, c->stage_i.as<int>());
    CU(cudaMemcpyAsync(neighbors, c->stage_i.p, (size_t)nrows * maxneighs * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  return MMD_OK;
}

int mmd_neigh_upload(mmd_ctx* c, const int* numneigh, const int* neighbors, int nrows, int maxneighs) {
  CHECK_CTX(c);
  if (!numneigh || !neighbors || nrows != c->nlocal || maxneighs < 1) return set_err(MMD_ERR_ARG, "neigh_upload: bad arguments (nrows must equal nlocal)");
  c->maxneighs = maxneighs;
  c->neigh_stride = (maxneighs + 7) & ~7;
  c->list_tile = false;
  MM(c->numneigh.reserve((size_t)std::max(nrows, 1) * sizeof(int), c->stream));
  MM(c->neighbors.reserve((size_t)std::max(nrows, 1) * c->neigh_stride * sizeof(int), c->stream));
  MM(c->stage_i.reserve((size_t)nrows * maxneighs * sizeof(int), c->stream));
  CU(cudaMemcpyAsync(c->numneigh.p, numneigh, (size_t)nrows * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->stage_i.p, neighbors, (size_t)nrows * maxneighs * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  const long long tot = (long long)nrows * maxneighs;
  LAUNCH(c, rows_restride_kernel, div_up(tot, TPB), TPB, c->stage_i.as<int>(), nrows, maxneighs, c->neigh_stride, maxneighs,
         c->neighbors.as<int>());
  CU(cudaStreamSynchronize(c->stream));
  c->neigh_rows = nrows;
  return MMD_OK;
}

int mmd_neigh_bins_download(mmd_ctx* c, int* bincount, int* bins, int atoms_per_bin) {
  CHECK_CTX(c);
  if (!c->have_geo) return set_err(MMD_ERR_STATE, "bins_download: mmd_neigh_setup missing");
  const int mb = c->mbins;
  if (bincount) {
    MM(c->stage_i.reserve((size_t)mb * sizeof(int), c->stream));
    LAUNCH(c, bin_counts_from_start_kernel, div_up(mb, TPB), TPB, c->bin_start.as<int>(), mb, c->stage_i.as<int>());
    CU(cudaMemcpyAsync(bincount, c->stage_i.p, (size_t)mb * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  if (bins) {
    if (atoms_per_bin < 1) return set_err(MMD_ERR_ARG, "bins_download: atoms_per_bin");
    MM(c->stage_i.reserve((size_t)mb * atoms_per_bin * sizeof(int), c->stream));
    LAUNCH(c, bins_to_rows_kernel, div_up(mb, TPB), TPB, c->bin_start.as<int>(), c->bin_atoms.as<int>(), mb, atoms_per_bin,
           c->stage_i.as<int>());
    CU(cudaMemcpyAsync(bins, c->stage_i.p, (size_t)mb * atoms_per_bin * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return MMD_OK;
}

int mmd_neigh_atom_bins_download(mmd_ctx* c, int* bin_of_atom, int count) {
  CHECK_CTX(c);
  if (!bin_of_atom || count < 0 || count > c->binned_count) return set_err(MMD_ERR_ARG, "atom_bins_download: count exceeds last binning");
  CU(cudaMemcpyAsync(bin_of_atom, c->atom_bin.p, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return MMD_OK;
}

// ---- Force -----------------------------------------------------------------------------
int mmd_force_lj_setup(mmd_ctx* c, const void* cutforcesq, const void* sigma6, const void* epsilon) {
  CHECK_CTX(c);
  if (!cutforcesq || !sigma6 || !epsilon) return set_err(MMD_ERR_ARG, "force_lj_setup: null table");
  const int nn = c->ntypes * c->ntypes;
  const size_t nb = (size_t)nn * c->prec;
  MM(c->lj_cut.reserve(nb, c->stream));
  MM(c->lj_s6.reserve(nb, c->stream));
  MM(c->lj_eps.reserve(nb, c->stream));
  CU(cudaMemcpy(c->lj_cut.p, cutforcesq, nb, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(c->lj_s6.p, sigma6, nb, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(c->lj_eps.p, epsilon, nb, cudaMemcpyHostToDevice));
  if (c->prec == 8) {
    const double *a = (const double*)cutforcesq, *b = (const double*)sigma6, *e = (const double*)epsilon;
    c->lj_uniform = all_equal(a, nn) && all_equal(b, nn) && all_equal(e, nn);
    c->lj_cut0 = a[0]; c->lj_s60 = b[0]; c->lj_eps0 = e[0];
  } else {
    const float *a = (const float*)cutforcesq, *b = (const float*)sigma6, *e = (const float*)epsilon;
    c->lj_uniform = all_equal(a, nn) && all_equal(b, nn) && all_equal(e, nn);
    c->lj_cut0 = a[0]; c->lj_s60 = b[0]; c->lj_eps0 = e[0];
  }
  c->have_lj = true;
  return MMD_OK;
}

static void store_real(mmd_ctx* c, void* dst, double v) {
  if (!dst) return;
  if (c->prec == 8) *(double*)dst = v;
  else *(float*)dst = (float)v;
}

int mmd_force_lj_compute(mmd_ctx* c, int halfneigh, int ghost_newton, int evflag, void* eng_vdwl, void* virial) {
  CHECK_CTX(c);
  MM(DISPATCH(c, Impl<double>::lj_async(c, halfneigh, ghost_newton, evflag, true),
              Impl<float>::lj_async(c, halfneigh, ghost_newton, evflag, true)));
  if (evflag && (eng_vdwl || virial)) {
    MM(Impl<double>::read_ev(c, 2));
    store_real(c, eng_vdwl, c->h_ev[0]);
    store_real(c, virial, c->h_ev[1]);
  }
  return MMD_OK;
}

int mmd_force_eam_setup(mmd_ctx* c, const void* rhor_spline, const void* z2r_spline, const void* frho_spline, int nr,
                        int nrho, int nr_tot, int nrho_tot, double rdr, double rdrho, const void* cutforcesq) {
  CHECK_CTX(c);
  if (!rhor_spline || !z2r_spline || !frho_spline || !cutforcesq) return set_err(MMD_ERR_ARG, "force_eam_setup: null table");
  if (nr < 2 || nrho < 2 || nr_tot < (nr + 1) * 7 || nrho_tot < (nrho + 1) * 7) return set_err(MMD_ERR_ARG, "force_eam_setup: table sizes");
  return DISPATCH(c, Impl<double>::eam_setup(c, rhor_spline, z2r_spline, frho_spline, nr, nrho, nr_tot, nrho_tot, rdr, rdrho, cutforcesq),
                  Impl<float>::eam_setup(c, rhor_spline, z2r_spline, frho_spline, nr, nrho, nr_tot, nrho_tot, rdr, rdrho, cutforcesq));
}

int mmd_force_eam_compute(mmd_ctx* c, int halfneigh, int evflag, void* eng_vdwl, void* virial) {
  CHECK_CTX(c);
  MM(DISPATCH(c, Impl<double>::eam_async(c, halfneigh, evflag), Impl<float>::eam_async(c, halfneigh, evflag)));
  if (evflag && (eng_vdwl || virial)) {
    MM(Impl<double>::read_ev(c, 4));
    store_real(c, eng_vdwl, Impl<double>::eam_energy(c, halfneigh));
    store_real(c, virial, c->h_ev[1]);
  }
  return MMD_OK;
}

// ---- Integrate / Thermo ----------------------------------------------------------------
int mmd_integrate_initial(mmd_ctx* c, double dt, double dtforce) {
  CHECK_CTX(c);
  return DISPATCH(c, Impl<double>::initial(c, dt, dtforce, false), Impl<float>::initial(c, dt, dtforce, false));
}
int mmd_integrate_final(mmd_ctx* c, double dtforce) {
  CHECK_CTX(c);
  return DISPATCH(c, Impl<double>::final_(c, dtforce, false, 1.0), Impl<float>::final_(c, dtforce, false, 1.0));
}
int mmd_thermo_sum_mv2(mmd_ctx* c, double mass, double* sum_mv2) {
  CHECK_CTX(c);
  if (!sum_mv2) return set_err(MMD_ERR_ARG, "sum_mv2: null output");
  return DISPATCH(c, Impl<double>::sum_mv2(c, mass, sum_mv2), Impl<float>::sum_mv2(c, mass, sum_mv2));
}

// ---- Comm ------------------------------------------------------------------------------
int mmd_comm_setup(mmd_ctx* c, const mmd_swap_table* t) {
  CHECK_CTX(c);
  if (!t || t->nswap < 0 || t->nswap > MMD_MAX_SWAPS || (t->nswap % 2)) return set_err(MMD_ERR_ARG, "comm_setup: bad swap table");
  if (t->nswap != 2 * (t->need[0] + t->need[1] + t->need[2])) return set_err(MMD_ERR_ARG, "comm_setup: nswap != 2*sum(need)");
  c->swaps = *t;
  for (int w = 0; w < MMD_MAX_SWAPS; w++) c->sw[w].sendnum = c->sw[w].recvnum = c->sw[w].firstrecv = 0;
  c->have_comm = true;
  return MMD_OK;
}

int mmd_comm_nccl_unique_id(void* id128) {
#ifdef MMD_WITH_NCCL
  if (!id128) return set_err(MMD_ERR_ARG, "null id");
  ncclUniqueId id;
  NC(ncclGetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(id128, &id, 128);
  return MMD_OK;
#else
  (void)id128;
  return set_err(MMD_ERR_STATE, "library built without NCCL");
#endif
}
int mmd_comm_nccl_init(mmd_ctx* c, const void* id128, int rank, int nranks) {
  CHECK_CTX(c);
#ifdef MMD_WITH_NCCL
  if (!id128 || rank < 0 || rank >= nranks) return set_err(MMD_ERR_ARG, "nccl_init: bad arguments");
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  NC(ncclCommInitRank(&c->nccl, nranks, id, rank));
  c->rank = rank;
  c->nranks = nranks;
  if (nranks > 1 && !c->stream2) {
    int pr_least = 0, pr_greatest = 0;
    CU(cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest));
    CU(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, pr_least));
    CU(cudaEventCreateWithFlags(&c->ev_int, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_bnd, cudaEventDisableTiming));
  }
  if (nranks > 1 && c->p2p_enable) MM(p2p_setup(c));
  return MMD_OK;
#else
  (void)id128; (void)rank; (void)nranks;
  return set_err(MMD_ERR_STATE, "library built without NCCL");
#endif
}

int mmd_comm_exchange(mmd_ctx* c) {
  CHECK_CTX(c);
  if (!c->have_comm) return set_err(MMD_ERR_STATE, "exchange: mmd_comm_setup missing");
  return DISPATCH(c, Impl<double>::exchange(c), Impl<float>::exchange(c));
}
int mmd_comm_borders(mmd_ctx* c) {
  CHECK_CTX(c);
  return DISPATCH(c, Impl<double>::borders(c), Impl<float>::borders(c));
}
int mmd_comm_communicate(mmd_ctx* c) {
  CHECK_CTX(c);
  return DISPATCH(c, Impl<double>::communicate(c, false), Impl<float>::communicate(c, false));
}
int mmd_comm_reverse_communicate(mmd_ctx* c) {
  CHECK_CTX(c);
  return DISPATCH(c, Impl<double>::reverse(c), Impl<float>::reverse(c));
}
int mmd_comm_swap_counts(mmd_ctx* c, int* sendnum, int* recvnum, int* firstrecv) {
  if (!c) return set_err(MMD_ERR_ARG, "null context");
  for (int w = 0; w < c->swaps.nswap; w++) {
    if (sendnum) sendnum[w] = c->sw[w].sendnum;
    if (recvnum) recvnum[w] = c->sw[w].recvnum;
    if (firstrecv) firstrecv[w] = c->sw[w].firstrecv;
  }
  return MMD_OK;
}
int mmd_comm_sendlist_download(mmd_ctx* c, int iswap, int* list, int count) {
  CHECK_CTX(c);
  if (iswap < 0 || iswap >= c->swaps.nswap || !list || count < 0 || count > c->sw[iswap].sendnum)
    return set_err(MMD_ERR_ARG, "sendlist_download: bad arguments");
  CU(cudaMemcpyAsync(list, c->sw[iswap].list.p, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return MMD_OK;
}
int mmd_comm_allreduce(mmd_ctx* c, double* values, int n, int op) {
  CHECK_CTX(c);
  if (!values || n < 0 || n > 16) return set_err(MMD_ERR_ARG, "allreduce: n must be 0..16");
  if (c->nranks <= 1) return MMD_OK;
#ifdef MMD_WITH_NCCL
  double* d = c->d_ev + 8;  // upper part of the scalar block (the lower 8 belong to the kernels)
  CU(cudaMemcpyAsync(d, values, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  NC(ncclAllReduce(d, d, n, ncclDouble, op == 1 ? ncclMax : ncclSum, c->nccl, c->stream));
  CU(cudaMemcpyAsync(values, d, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return MMD_OK;
#else
  return set_err(MMD_ERR_STATE, "library built without NCCL");
#endif
}

// ---- time loop ---------------------------------------------------------------------------
int mmd_run(mmd_ctx* c, const mmd_run_params* p, mmd_thermo_sample* samples, int max_samples, int* nsamples,
            float* elapsed_ms) {
  CHECK_CTX(c);
  if (!p || p->ntimes < 0 || p->neigh_every < 1) return set_err(MMD_ERR_ARG, "run: bad parameters");
  return DISPATCH(c, Impl<double>::run(c, p, samples, max_samples, nsamples, elapsed_ms),
                  Impl<float>::run(c, p, samples, max_samples, nsamples, elapsed_ms));
}

int mmd_run_phase_times(mmd_ctx* c, double* ms, long long* calls, int reset) {
  if (!c) return set_err(MMD_ERR_ARG, "null context");
  for (int k = 0; k < MMD_NPHASE; k++) {
    if (ms) ms[k] = c->phase_ms[k];
    if (calls) calls[k] = c->phase_calls[k];
    if (reset) { c->phase_ms[k] = 0; c->phase_calls[k] = 0; }
  }
  return MMD_OK;
}

// ---- introspection -------------------------------------------------------------------------
int mmd_query_int(mmd_ctx* c, const char* key, long long* value) {
  if (!c || !key || !value) return set_err(MMD_ERR_ARG, "query: null argument");
  std::string k(key);
  if (k == "nlocal") *value = c->nlocal;
  else if (k == "nghost") *value = c->nghost;
  else if (k == "nmax") *value = c->cap;
  else if (k == "maxneighs") *value = c->maxneighs;
  else if (k == "neigh_stride") *value = c->neigh_stride;
  else if (k == "mbins") *value = c->mbins;
  else if (k == "nstencil") *value = (long long)c->stencil.size();
  else if (k == "nruns") *value = c->nruns;
  else if (k == "total_neigh") *value = c->total_neigh;
  else if (k == "neigh_builds") *value = c->neigh_builds;
  else if (k == "neigh_resizes") *value = c->neigh_resizes;
  else if (k == "max_bin_count") *value = c->max_bin_seen;
  else if (k == "atoms_per_bin") *value = ref_atoms_per_bin(8, c->max_bin_seen);
  else if (k == "nswap") *value = c->swaps.nswap;
  else if (k == "precision") *value = c->prec;
  else if (k == "ntypes") *value = c->ntypes;
  else if (k == "lj_uniform") *value = c->lj_uniform;
  else if (k == "eam_uniform") *value = c->eam_uniform;
  else if (k == "lj_threads_per_atom") *value = c->lj_tpa;
  else if (k == "eam_threads_per_atom") *value = c->eam_tpa;
  else if (k == "launches") *value = c->launches;
  else if (k == "exchange_sent") *value = c->exch_sent;
  else if (k == "exchange_received") *value = c->exch_received;
  else if (k == "nranks") *value = c->nranks;
  else if (k == "p2p_active") *value = c->p2p_on;
  else if (k == "p2p_calls") *value = c->p2p_calls;
  else if (k == "list_xsorted") *value = c->list_xsorted;
  else if (k == "tile_lists") *value = c->tile_enable;
  else if (k == "fuse_force") *value = c->fuse_force && c->fuse_integrate;
  else if (k == "list_tile") *value = c->list_tile;
  else if (k == "list_dealt") *value = c->list_tile && c->list_dealt;
  else if (k == "tile_dealt_capacity") *value = c->tcapq;
  else if (k == "split_steps") *value = c->split_steps;
  else if (k == "fused_halo_steps") *value = c->fused_halo_steps;
  else if (k == "graph_replays") *value = c->graph_replays;
  else if (k == "graph_captures") *value = c->graph_captures;
  else if (k == "stage_clocks" || k == "cta_clocks" || k == "cta_count") {
    unsigned long long h[4] = {0, 0, 0, 0};
    if (c->d_prof) {
      cudaStreamSynchronize(c->stream);
      cudaMemcpy(h, c->d_prof, sizeof h, cudaMemcpyDeviceToHost);
    }
    *value = (long long)h[k == "stage_clocks" ? 0 : (k == "cta_clocks" ? 1 : 2)];
  }
  else if (k == "tile_interior") *value = c->split_ready ? c->n_interior : 0;
  else if (k == "tile_boundary") *value = c->split_ready ? c->n_boundary : 0;
  else if (k == "tile_ok") *value = c->tile_ok;
  else if (k == "tile_builds") *value = c->tile_builds;
  else if (k == "tile_fallbacks") *value = c->tile_fallbacks;
  else if (k == "tile_row_capacity") *value = c->tcap;
  else if (k == "tile_max_halo") *value = c->tile_max_h;
  else if (k == "tile_max_full") *value = c->tile_max_full;
  else if (k == "tile_count") *value = c->tile_ok ? c->tgeo.ntiles : 0;
  else return set_err(MMD_ERR_ARG, "query: unknown key '%s'", key);
  return MMD_OK;
}

int mmd_set_option(mmd_ctx* c, const char* key, long long value) {
  if (!c || !key) return set_err(MMD_ERR_ARG, "set_option: null argument");
  std::string k(key);
  auto pow2 = [](long long v) { return v >= 1 && v <= 32 && (v & (v - 1)) == 0; };
  if (k == "lj_threads_per_atom") {
    if (value != 0 && !pow2(value)) return set_err(MMD_ERR_ARG, "lj_threads_per_atom must be 0 (auto),1,2,4,8,16 or 32");
    c->lj_tpa = (int)value;
  } else if (k == "eam_threads_per_atom") {
    if (!pow2(value)) return set_err(MMD_ERR_ARG, "eam_threads_per_atom must be 1,2,4,8,16 or 32");
    c->eam_tpa = (int)value;
  } else if (k == "tile_lists") {  // 1: tile-resident lists + shared-memory force kernels where they apply; 0: classic rows
    c->tile_enable = value != 0;  // takes effect at the next neighbor build
  } else if (k == "fuse_halo") {
    c->fuse_halo = value != 0;
    if (!c->fuse_halo) c->ghosts_resolved = false;
  } else if (k == "tile_dealt") {  // 1: the LJ force kernel walks bank-dealt rows (quarter warp per atom); 0: one lane pair per row
    c->tile_dealt = value != 0;     // takes effect at the next neighbor build
  } else if (k == "fuse_ghosts") {  // one rank: the fused force kernel also writes the periodic images (no halo launch)
    c->fuse_ghosts = value != 0;
  } else if (k == "kernel_profile") {
    if (value && !c->d_prof) {
      if (cudaMalloc(&c->d_prof, 4 * sizeof(unsigned long long)) != cudaSuccess) return set_err(MMD_ERR_CUDA, "kernel_profile: cudaMalloc");
    }
    if (c->d_prof) cudaMemset(c->d_prof, 0, 4 * sizeof(unsigned long long));
    c->kernel_profile = value != 0;
  } else if (k == "graph_steps") {  // 0: launch every step; 1: one rank replays pairs of plain steps from a CUDA graph
    c->graph_steps = value != 0;
    c->pair_valid = false;
  } else if (k == "split_force") {  // several ranks: interior tiles on a second stream behind the forward halo
    c->split_enable = value != 0;
    if (!c->split_enable) c->split_ready = false;
  } else if (k == "tile_pair_build") {
    c->tile_pair_build = value != 0;
  } else if (k == "tile_lane_build") {
    c->tile_lane_build = value != 0;
  } else if (k == "tile_xsort") {
    c->tile_xsort = value != 0;
  } else if (k == "tile_eam") {
    c->tile_eam = value != 0;
  } else if (k == "p2p_halo") {  // 0: forward halo through NCCL send/recv even when peer windows are mapped
    c->p2p_enable = value != 0;
    if (!c->p2p_enable) c->p2p_on = false;
  } else if (k == "tile_build2") {
    c->tile_build2 = value != 0;
  } else if (k == "fuse_integrate") {
    c->fuse_integrate = value != 0;
  } else if (k == "fuse_force") {
    c->fuse_force = value != 0;
  } else if (k == "phase_timing") {
    c->phase_timing = value != 0;
  } else if (k == "force_nonuniform") {  // testing: exercise the per-type table path
    if (value) { c->lj_uniform = false; c->eam_uniform = false; }
  } else {
    return set_err(MMD_ERR_ARG, "set_option: unknown key '%s'", key);
  }
  return MMD_OK;
}

}  // extern "C"

// Tile-resident neighbor lists and the shared-memory pair-force kernels that consume them.
//
// Why: profiles/r1_ncu_force_neigh_lj80.md shows both classic force kernels bound by L1TEX
// wavefronts -- one 32-byte sector per neighbor gather (plus one per RED for half lists) -- at
// 10-20 % of HBM bandwidth.  tools/microbench/tile_gather_bench.cu measured the remedy on B200: keep
// the positions a CTA needs in shared memory and address them with 16-bit indices (1.5x fewer
// pipeline cycles per pair than global gathers, no reductions at all, neighbor rows half the bytes).
//
// Layout.  The bin grid (Neighbor::setup, ref/neighbor.cpp:318-452) is cut into tiles of
// TBX x TBY x TBZ bins.  A tile's *halo window* is the tile grown by the stencil extent in every
// direction; because atoms are held in bin order (CSR: bin_start/bin_atoms), each (y,z) pencil of the
// window is ONE contiguous range of CSR slots -- a "run".  A tile-local index is the position of a
// slot in the concatenation of the tile's runs.  For every local atom the list stores, in the
// reference's row order (stencil order, then bin order: ref/neighbor.cpp:141-183), ALL neighbors
// within cutneigh as 16-bit tile-local indices; bit 15 marks the entries that belong to the
// reference's half list (ref/neighbor.cpp:154-171), so the reference's rows, numneigh and totals are
// recovered exactly (tile_rows_export_kernel) while the force kernels evaluate every atom's complete
// neighborhood ("owner computes"): no scatter to f[j], no reverse halo, no clearing of f.
// The pair set is identical to the reference's (each half-list pair appears in both owners' rows), so
// forces, energies and virials agree with ForceLJ::compute_halfneigh / _fullneigh up to summation
// order (ref/force_lj.cpp:185-263, :366-449).
#pragma once
#include "common.cuh"
#include "neighbor_kernels.cuh"

namespace mmd {

constexpr int TBX = 4, TBY = 4, TBZ = 4;   // bins per tile edge
constexpr int TILE_MAXRUN = 100;           // (TBY+2sy)*(TBZ+2sz) must not exceed this (stencil extent <= 3 bins)
constexpr int TILE_THREADS = 512;          // one warp per centre pencil (TBY*TBZ = 16)
constexpr int TILE_NCENTER = TBY * TBZ;
constexpr unsigned short TILE_HALF_BIT = 0x8000u;

struct TileGeo {
  int mbx, mby, mbz;   // bin grid
  int ox, oy, oz;      // tile (tx,ty,tz) starts at bin (tx*TBX - ox, ...): the first bin that owns local atoms opens a tile
  int ntx, nty, ntz;   // tiles per axis
  int sx, sy, sz;      // stencil half-extent in bins
  int nry, nrz, nrun;  // runs of a halo window: (TBY+2sy) x (TBZ+2sz)
  int ntiles;
  int hcap;            // atoms a halo window may hold (shared-memory capacity of the force kernels)
};

// reference bin id of bin (x,y,z): Neighbor::coord2bin adds 1 (ref/neighbor.cpp:299)
__device__ __forceinline__ int tile_bin_id(const TileGeo& g, int x, int y, int z) { return (z * g.mby + y) * g.mbx + x + 1; }

// ---------------------------------------------------------------------------------------
// Per-tile tables, rebuilt with the bins at every neighbor build.  One warp per tile.
//   runs[t*nrun + p]   = {first CSR slot of run p, tile-local index of that slot}
//   center[t*16 + c]   = {lo, hi, q0, -}: tile-local index range of the tile's own atoms in centre pencil c and
//                        the row number of the first of them.  Rows are numbered tile by tile ("q order"), so
//                        the rows a warp walks are contiguous in memory and addressable without the atom id.
//   info[t]            = {atoms in the halo window, 1 if the tile owns at least one local atom}
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
tile_table_kernel(TileGeo g, const int* __restrict__ bin_start, const int* __restrict__ bin_atoms, int mbins, int nlocal,
                  int2* __restrict__ runs, int4* __restrict__ center, int2* __restrict__ info, int* __restrict__ max_h,
                  int* __restrict__ row_counter) {
  __shared__ int s_start[4][TILE_MAXRUN + 1];
  __shared__ int s_off[4][TILE_MAXRUN + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int t = blockIdx.x * 4 + w;
  if (t >= g.ntiles) return;
  const int tx = t % g.ntx, ty = (t / g.ntx) % g.nty, tz = t / (g.ntx * g.nty);
  const int bx0 = tx * TBX - g.ox, by0 = ty * TBY - g.oy, bz0 = tz * TBZ - g.oz;
  const int xlo = max(0, bx0 - g.sx), xhi = min(g.mbx, bx0 + TBX + g.sx);
  int carry = 0;
  for (int p0 = 0; p0 < g.nrun; p0 += 32) {
    const int p = p0 + lane;
    int start = 0, len = 0;
    if (p < g.nrun) {
      const int y = by0 - g.sy + p % g.nry, z = bz0 - g.sz + p / g.nry;
      if (y >= 0 && y < g.mby && z >= 0 && z < g.mbz && xhi > xlo) {
        start = bin_start[min(tile_bin_id(g, xlo, y, z), mbins)];
        len = bin_start[min(tile_bin_id(g, xhi - 1, y, z) + 1, mbins)] - start;
      }
    }
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (p < g.nrun) {
      s_start[w][p] = start;
      s_off[w][p] = carry + incl - len;
      runs[(size_t)t * g.nrun + p] = make_int2(start, carry + incl - len);
    }
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  // centre pencils: the tile's own bins
  const int cx0 = max(0, bx0), cx1 = min(g.mbx, bx0 + TBX);
  bool owns = false;
  int lo = 0, hi = 0;
  if (lane < TILE_NCENTER) {
    const int cy = lane % TBY, cz = lane / TBY;
    const int y = by0 + cy, z = bz0 + cz;
    if (y >= 0 && y < g.mby && z >= 0 && z < g.mbz && cx1 > cx0) {
      const int p = (cy + g.sy) + (cz + g.sz) * g.nry;
      const int s0 = bin_start[min(tile_bin_id(g, cx0, y, z), mbins)];
      const int s1 = bin_start[min(tile_bin_id(g, cx1 - 1, y, z) + 1, mbins)];
      lo = s_off[w][p] + (s0 - s_start[w][p]);
      hi = lo + (s1 - s0);
      // ids ascend inside a bin: a bin owns a local atom iff its first id is local
      for (int x = cx0; x < cx1; x++) {
        const int b = min(tile_bin_id(g, x, y, z), mbins);
        const int a = bin_start[b], e = bin_start[min(b + 1, mbins)];
        if (e > a && bin_atoms[a] < nlocal) owns = true;
      }
    }
  }
  const bool any = __any_sync(0xffffffffu, owns);
  // rows of this tile: exclusive scan of the centre pencil lengths, base from a global counter
  const int len_c = (lane < TILE_NCENTER) ? hi - lo : 0;
  int incl = len_c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const int tile_rows = __shfl_sync(0xffffffffu, incl, 31);
  int base = 0;
  if (lane == 0 && any) base = atomicAdd(row_counter, tile_rows);
  base = __shfl_sync(0xffffffffu, base, 0);
  if (lane < TILE_NCENTER) center[(size_t)t * TILE_NCENTER + lane] = make_int4(lo, hi, any ? base + incl - len_c : 0, 0);
  if (lane == 0) {
    info[t] = make_int2(carry, any ? 1 : 0);
    if (any) atomicMax(max_h, carry);
  }
}

// stencil run with its pencil decoded: bins [off, off+len) relative to the atom's bin, all in pencil (dy,dz)
struct StencilRun { int off, len, dy, dz; };

// ---------------------------------------------------------------------------------------
// Neighbor build into tile-local rows.  Same decomposition as neigh_build_kernel (one warp per
// bin, candidates staged once per bin and tested against every local atom of the bin, ballot/popc
// compaction keeps the reference's row order); every candidate within cutneigh is stored, the
// reference's half-list filter only sets bit 15 and the half counters.
//   MODE 0: full list requested      (every entry counts)
//   MODE 1: half list, ghost_newton  (own bin: j>i and ghost not "below"; other bins: upper stencil only)
//   MODE 2: half list, no ghost_newton (j > i)
// status |= 2 when a stencil run of a bin that owns local atoms leaves the bin grid (no tile mapping).
// ---------------------------------------------------------------------------------------
template <class T, int MODE>
__global__ void __launch_bounds__(NB_WARPS * 32)
neigh_build_tile_kernel(const Vec4<T>* __restrict__ x, int nlocal, const int* __restrict__ bin_start,
                        const int* __restrict__ bin_atoms, int mbins, const StencilRun* __restrict__ sruns, int nruns,
                        const T* __restrict__ cutneighsq, int ntypes, TileGeo g, const int2* __restrict__ tile_runs,
                        const int4* __restrict__ tile_center, unsigned short* __restrict__ rows, int tcap,
                        int* __restrict__ numneigh_half, int2* __restrict__ row_atom, int* __restrict__ status,
                        int* __restrict__ max_half, int* __restrict__ max_full, unsigned long long* __restrict__ total_half) {
  __shared__ Vec4<T> s_xi[NB_WARPS][NB_CHUNK];
  __shared__ int s_id[NB_WARPS][NB_CHUNK];
  __shared__ int s_q[NB_WARPS][NB_CHUNK];
  __shared__ unsigned s_cand[NB_WARPS][NB_MAXC];         // global id | bit 31: candidate sits in the warp's own bin
  __shared__ unsigned short s_loc[NB_WARPS][NB_MAXC];    // tile-local index | bit 15: bin in the upper half stencil
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.x * NB_WARPS + w;
  if (b >= mbins || b < 1) return;
  const int s0 = bin_start[b], s1 = bin_start[b + 1];
  if (s1 == s0) return;
  const unsigned lt_mask = (1u << lane) - 1u;
  // bin coordinates and tile
  const int lin = b - 1;
  const int bx = lin % g.mbx, by = (lin / g.mbx) % g.mby, bz = lin / (g.mbx * g.mby);
  const int tx = (bx + g.ox) / TBX, ty = (by + g.oy) / TBY, tz = (bz + g.oz) / TBZ;
  const int tile = (tz * g.nty + ty) * g.ntx + tx;
  const int2* trun = tile_runs + (size_t)tile * g.nrun;
  const int cy = by - (ty * TBY - g.oy), cz = bz - (tz * TBZ - g.oz);
  const int ry0 = cy + g.sy, rz0 = cz + g.sz;
  // row number of the atom in CSR slot c of this bin: q = q0 + (tile-local index of c) - lo
  const int4 cen = tile_center[(size_t)tile * TILE_NCENTER + cy + cz * TBY];
  const int2 trc = trun[ry0 + rz0 * g.nry];
  const int q_of_slot = cen.z - cen.x + trc.y - trc.x;

  for (int chunk0 = s0; chunk0 < s1; chunk0 += NB_CHUNK) {
    const int ci = chunk0 + lane;
    int my_id = (ci < s1) ? bin_atoms[ci] : 0x7fffffff;
    const bool is_local = my_id < nlocal;
    const int nloc = __popc(__ballot_sync(0xffffffffu, is_local));
    if (nloc == 0) break;
    if (is_local) {
      s_xi[w][lane] = x[my_id];
      s_id[w][lane] = my_id;
      s_q[w][lane] = q_of_slot + ci;
    }
    __syncwarp();
    int my_n = 0, my_h = 0;  // full / half row length of i-atom `lane`

    int r = 0, off = 0;
    while (r < nruns) {
      int filled = 0;
      while (r < nruns && filled < NB_MAXC) {
        const StencilRun run = sruns[r];
        const int dxlo = run.off - (run.dz * g.mby + run.dy) * g.mbx;
        const int yy = by + run.dy, zz = bz + run.dz;
        if (yy < 0 || yy >= g.mby || zz < 0 || zz >= g.mbz || bx + dxlo < 0 || bx + dxlo + run.len > g.mbx ||
            b + run.off + run.len > mbins) {
          if (lane == 0) atomicOr(status, 2);
          r++; off = 0;
          continue;
        }
        const int blo = b + run.off, bhi = blo + run.len;
        const int2 tr = trun[(ry0 + run.dy) + (rz0 + run.dz) * g.nry];
        const int c_begin = bin_start[blo] + off, c_end = bin_start[bhi];
        const int take = min(c_end - c_begin, NB_MAXC - filled);
        const bool upper_pencil = run.dz > 0 || (run.dz == 0 && run.dy > 0);
        const bool own_pencil = run.dz == 0 && run.dy == 0;
        for (int k = lane; k < take; k += 32) {
          const int c = c_begin + k;
          const bool own = c >= s0 && c < s1;
          const bool upper = upper_pencil || (own_pencil && c >= s1);
          s_cand[w][filled + k] = (unsigned)bin_atoms[c] | (own ? 0x80000000u : 0u);
          s_loc[w][filled + k] = (unsigned short)(((tr.y + (c - tr.x)) & 0x7fff) | (upper ? 0x8000 : 0));
        }
        filled += take;
        if (c_begin + take == c_end) { r++; off = 0; }
        else off += take;
      }
      __syncwarp();
      for (int c0 = 0; c0 < filled; c0 += 32) {
        const int c = c0 + lane;
        const bool valid = c < filled;
        int j = 0;
        bool own_bin = false, upper = false;
        unsigned short loc = 0;
        Vec4<T> xj;
        xj.x = xj.y = xj.z = xj.w = (T)0;
        if (valid) {
          const unsigned e = s_cand[w][c];
          const unsigned short l = s_loc[w][c];
          own_bin = (e & 0x80000000u) != 0u;
          upper = (l & 0x8000) != 0;
          loc = l & 0x7fff;
          j = (int)(e & 0x7fffffffu);
          xj = x[j];
        }
        const int tj = lane_to_type(xj.w);
        for (int t = 0; t < nloc; t++) {
          const Vec4<T> xi = s_xi[w][t];
          const int i = s_id[w][t];
          bool ok = valid && !(own_bin && j == i);
          bool half = true;
          if (MODE == 1) {
            if (own_bin) {
              half = j > i;
              if (j >= nlocal) {
                const bool below = (xj.z < xi.z) || (xj.z == xi.z && xj.y < xi.y) ||
                                   (xj.z == xi.z && xj.y == xi.y && xj.x < xi.x);
                half = half && !below;
              }
            } else {
              half = upper;
            }
          }
          if (MODE == 2) half = j > i;
          const T rsq = rsq_unfused(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z);
          const int ti = lane_to_type(xi.w);
          ok = ok && (rsq <= __ldg(&cutneighsq[ti * ntypes + tj]));
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          const unsigned mh = MODE == 0 ? m : __ballot_sync(0xffffffffu, ok && half);
          const int base = __shfl_sync(0xffffffffu, my_n, t);
          if (ok) {
            const int pos = base + __popc(m & lt_mask);
            if (pos < tcap) rows[(size_t)s_q[w][t] * tcap + pos] = (unsigned short)(loc | ((MODE != 0 && half) ? TILE_HALF_BIT : 0));
          }
          if (lane == t) { my_n += __popc(m); my_h += __popc(mh); }
        }
      }
      __syncwarp();
    }
    if (is_local) {
      numneigh_half[my_id] = my_h;
      row_atom[q_of_slot + ci] = make_int2(my_id, my_n);
    }
    int mxh = is_local ? my_h : 0, mxf = is_local ? my_n : 0;
    unsigned long long sum = is_local ? (unsigned long long)my_h : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mxh = max(mxh, __shfl_xor_sync(0xffffffffu, mxh, o));
      mxf = max(mxf, __shfl_xor_sync(0xffffffffu, mxf, o));
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if (lane == 0) {
      atomicMax(max_half, mxh);
      atomicMax(max_full, mxf);
      atomicAdd(total_half, sum);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------
// Shared-memory image of a tile: run tables + SoA positions of the halo window.
// ---------------------------------------------------------------------------------------
template <class T> struct TileSmem {
  int* run_start;   // [nrun]
  int* run_off;     // [nrun + 1]
  T *sx, *sy, *sz;  // [hcap]
  unsigned char* st;  // [hcap] atom types (non-uniform parameter path only)
  __device__ __forceinline__ void carve(unsigned char* base, int hcap, bool with_types) {
    sx = reinterpret_cast<T*>(base);
    sy = sx + hcap;
    sz = sy + hcap;
    unsigned char* q = reinterpret_cast<unsigned char*>(sz + hcap);
    run_start = reinterpret_cast<int*>(q);
    run_off = run_start + TILE_MAXRUN;
    st = reinterpret_cast<unsigned char*>(run_off + TILE_MAXRUN + 1);
    (void)with_types;
  }
};
template <class T> __host__ __device__ inline size_t tile_smem_bytes(int hcap, bool with_types) {
  return (size_t)hcap * 3 * sizeof(T) + (2 * TILE_MAXRUN + 1) * sizeof(int) + (with_types ? (size_t)hcap : 0) + 16;
}

// load run tables and stage the positions of the whole halo window (coalesced over CSR slots)
template <class T, bool TYPES>
__device__ __forceinline__ int tile_stage(TileSmem<T>& S, const TileGeo& g, int t, int h, const int2* __restrict__ tile_runs,
                                          const int* __restrict__ slots, const Vec4<T>* __restrict__ x) {
  const int2* tr = tile_runs + (size_t)t * g.nrun;
  for (int p = threadIdx.x; p < g.nrun; p += blockDim.x) {
    const int2 r = tr[p];
    S.run_start[p] = r.x;
    S.run_off[p] = r.y;
  }
  if (threadIdx.x == 0) S.run_off[g.nrun] = h;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int p = w; p < g.nrun; p += nw) {
    const int start = S.run_start[p], off = S.run_off[p], len = S.run_off[p + 1] - off;
    for (int k = lane; k < len; k += 32) {
      const int id = __ldg(slots + start + k);
      const Vec4<T> v = ldg4(x + id);
      S.sx[off + k] = v.x;
      S.sy[off + k] = v.y;
      S.sz[off + k] = v.z;
      if (TYPES) S.st[off + k] = (unsigned char)lane_to_type(v.w);
    }
  }
  __syncthreads();
  return h;
}

// ---------------------------------------------------------------------------------------
// LJ pair force over tile-local rows: one CTA per tile, one warp per centre pencil, one lane per atom.
// Each lane walks its own row (8 indices per 128-bit load), gathers x_j from shared memory, keeps
// F_i in registers and writes it once (plain store).  Energy / virial follow the reference's
// conventions through the two scale factors (half list: sum over pairs; full list: eng_vdwl is twice
// the pair energy, ref/force_lj.cpp:441-442).
// ---------------------------------------------------------------------------------------
template <class T> struct LJTileParams {
  T cutforcesq, sigma6, epsilon;
  const T* cutforcesq_tab;
  const T* sigma6_tab;
  const T* epsilon_tab;
  int ntypes;
  double e_scale, v_scale;
};

// reciprocal without the IEEE-division slow path (no branches inside the unrolled pair loop):
// FP64: MUFU.RCP64H seed (>= 20 bits) + two Newton steps -> <= 1 ulp; FP32: the hardware reciprocal.
__device__ __forceinline__ double tile_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
__device__ __forceinline__ float tile_rcp(float x) { return __frcp_rn(x); }

constexpr int LJT_TPA = 2;  // lanes per atom: the two lanes of an atom take alternate 8-entry chunks of its row

// 8 row entries; the L2::128B hint pulls the rest of the 128-byte line into L2 (rows are walked front to back)
__device__ __forceinline__ uint4 ldg_row8(const unsigned short* p) {
  uint4 v;
  asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

template <class T, int EV, int UNIFORM>
__global__ void __launch_bounds__(TILE_THREADS, 2)
force_lj_tile_kernel(const Vec4<T>* __restrict__ x, Vec4<T>* __restrict__ f, TileGeo g, const int2* __restrict__ tile_runs,
                     const int4* __restrict__ tile_center, const int2* __restrict__ tile_info, const int* __restrict__ slots,
                     const unsigned short* __restrict__ rows, const int2* __restrict__ row_atom, int tcap, int nlocal,
                     LJTileParams<T> P, double* __restrict__ ev_out) {
  extern __shared__ __align__(16) unsigned char tile_smem_raw[];
  const int t = blockIdx.x;
  const int2 inf = tile_info[t];
  if (inf.y == 0) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int sub = lane & (LJT_TPA - 1);
  constexpr int APP = 32 / LJT_TPA;  // atoms per pass of a warp

  // A "pass" = APP consecutive atoms of a centre pencil.  {atom id, row length} and the first 8 entries of
  // every row are fetched one pass ahead; the very first fetch is issued before the halo window is staged.
  int cr = w;
  int4 ce = make_int4(0, 0, 0, 0);
  if (cr < TILE_NCENTER) ce = tile_center[(size_t)t * TILE_NCENTER + cr];
  int2 ta_n = make_int2(-1, 0);
  uint4 c_n = make_uint4(0, 0, 0, 0);
  {
    const int a = ce.x + lane / LJT_TPA;
    if (a < ce.y) {
      const size_t q = (size_t)(ce.z + (a - ce.x));
      ta_n = __ldg(row_atom + q);
      c_n = ldg_row8(rows + q * tcap + sub * 8);
    }
  }

  TileSmem<T> S;
  S.carve(tile_smem_raw, g.hcap, !UNIFORM);
  tile_stage<T, !UNIFORM>(S, g, t, inf.x, tile_runs, slots, x);

  double eng = 0.0, vir = 0.0;
  for (; cr < TILE_NCENTER; cr += nw) {
    if (cr != w) {  // only when the block has fewer warps than centre pencils
      ce = tile_center[(size_t)t * TILE_NCENTER + cr];
      ta_n = make_int2(-1, 0);
      const int a = ce.x + lane / LJT_TPA;
      if (a < ce.y) {
        const size_t q = (size_t)(ce.z + (a - ce.x));
        ta_n = __ldg(row_atom + q);
        c_n = ldg_row8(rows + q * tcap + sub * 8);
      }
    }
    for (int a0 = ce.x; a0 < ce.y; a0 += APP) {
      const int a = a0 + lane / LJT_TPA;
      const int2 ta = ta_n;
      uint4 nxt = c_n;
      const int id = (a < ce.y && ta.y > 0) ? ta.x : -1;
      const int cnt = id >= 0 ? ta.y : 0;
      const int aa = id >= 0 ? a : ce.x;
      const unsigned short* __restrict__ row = rows + (size_t)(ce.z + (aa - ce.x)) * tcap;
      {  // next pass
        const int an = a + APP;
        ta_n = make_int2(-1, 0);
        if (an < ce.y) {
          const size_t q = (size_t)(ce.z + (an - ce.x));
          ta_n = __ldg(row_atom + q);
          c_n = ldg_row8(rows + q * tcap + sub * 8);
        }
      }
      const T xi = S.sx[aa], yi = S.sy[aa], zi = S.sz[aa];
      const int ti = UNIFORM ? 0 : (int)S.st[aa];
      T fx = 0, fy = 0, fz = 0;
      const int nch = (cnt + 7) >> 3;
      const int mine = nch > sub ? (nch - sub + LJT_TPA - 1) / LJT_TPA : 0;
      const int iters = __reduce_max_sync(0xffffffffu, mine);
      int k0 = sub * 8;
      for (int it = 0; it < iters; it++, k0 += 8 * LJT_TPA) {
        const uint4 pk = nxt;
        if (k0 + 8 * LJT_TPA < cnt) nxt = ldg_row8(row + k0 + 8 * LJT_TPA);
        const unsigned wds[4] = {pk.x, pk.y, pk.z, pk.w};
        // branch-free: entries past the row end point at the atom itself and are masked out; four independent
        // pair evaluations are in flight per group so the FP64 dependency chains overlap
#pragma unroll
        for (int h = 0; h < 2; h++) {
          int lj[4];
          bool in[4];
          T dx[4], dy[4], dz[4], rsq[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int ee = 4 * h + e;
            in[e] = k0 + ee < cnt;
            lj[e] = in[e] ? (int)((wds[ee >> 1] >> ((ee & 1) * 16)) & 0x7fff) : aa;
          }
#pragma unroll
          for (int e = 0; e < 4; e++) {
            dx[e] = xi - S.sx[lj[e]];
            dy[e] = yi - S.sy[lj[e]];
            dz[e] = zi - S.sz[lj[e]];
          }
#pragma unroll
          for (int e = 0; e < 4; e++) rsq[e] = dx[e] * dx[e] + dy[e] * dy[e] + dz[e] * dz[e];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            T cut, s6, eps;
            if (UNIFORM) {
              cut = P.cutforcesq; s6 = P.sigma6; eps = P.epsilon;
            } else {
              const int tij = ti * P.ntypes + S.st[lj[e]];
              cut = __ldg(P.cutforcesq_tab + tij); s6 = __ldg(P.sigma6_tab + tij); eps = __ldg(P.epsilon_tab + tij);
            }
            const bool hit = in[e] && rsq[e] < cut;
            const T r = hit ? rsq[e] : (T)1;
            const T sr2 = tile_rcp(r);
            const T sr6 = sr2 * sr2 * sr2 * s6;
            T force = (T)48 * sr6 * (sr6 - (T)0.5) * sr2 * eps;
            force = hit ? force : (T)0;
            fx += dx[e] * force;
            fy += dy[e] * force;
            fz += dz[e] * force;
            if (EV) {
              eng += hit ? (double)((T)4 * sr6 * (sr6 - (T)1) * eps) : 0.0;
              vir += (double)(rsq[e] * force);
            }
          }
        }
      }
      if (LJT_TPA > 1) {
        fx = group_sum<LJT_TPA>(fx);
        fy = group_sum<LJT_TPA>(fy);
        fz = group_sum<LJT_TPA>(fz);
      }
      if (id >= 0 && sub == 0) {
        Vec4<T> out;
        out.x = fx; out.y = fy; out.z = fz; out.w = (T)0;
        f[id] = out;
      }
    }
  }
  if (EV) {
    const double v2[2] = {eng * P.e_scale, vir * P.v_scale};
    block_accumulate<2>(v2, ev_out);
  }
}

// ---------------------------------------------------------------------------------------
// Export to the reference's row format (Neighbor::neighbors / numneigh): global atom ids, the
// half-list subset when the list was built for half neighbor semantics.  One CTA per tile.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_THREADS)
tile_rows_export_kernel(TileGeo g, const int2* __restrict__ tile_runs, const int4* __restrict__ tile_center,
                        const int2* __restrict__ tile_info, const int* __restrict__ slots,
                        const unsigned short* __restrict__ rows, const int2* __restrict__ row_atom, int tcap, int nlocal,
                        int half_only, int* __restrict__ out, int out_stride, int out_cap) {
  __shared__ int s_start[TILE_MAXRUN];
  __shared__ int s_off[TILE_MAXRUN + 1];
  const int t = blockIdx.x;
  const int2 inf = tile_info[t];
  if (inf.y == 0) return;
  for (int p = threadIdx.x; p < g.nrun; p += blockDim.x) {
    const int2 r = tile_runs[(size_t)t * g.nrun + p];
    s_start[p] = r.x;
    s_off[p] = r.y;
  }
  if (threadIdx.x == 0) s_off[g.nrun] = inf.x;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int cr = w; cr < TILE_NCENTER; cr += nw) {
    const int4 ce = tile_center[(size_t)t * TILE_NCENTER + cr];
    for (int a = ce.x + lane; a < ce.y; a += 32) {
      const size_t q = (size_t)(ce.z + (a - ce.x));
      const int2 ta = row_atom[q];
      if (ta.y <= 0 || ta.x < 0 || ta.x >= nlocal) continue;
      const int id = ta.x, cnt = ta.y;
      const unsigned short* row = rows + q * tcap;
      int n = 0, p = 0;
      for (int k = 0; k < cnt; k++) {
        const unsigned short e = row[k];
        if (half_only && !(e & TILE_HALF_BIT)) continue;
        const int loc = e & 0x7fff;
        // rows are not monotone across stencil pencils: restart the run search when needed
        if (loc < s_off[p]) p = 0;
        while (p + 1 < g.nrun && loc >= s_off[p + 1]) p++;
        const int j = slots[s_start[p] + (loc - s_off[p])];
        if (n < out_cap) out[(size_t)id * out_stride + n] = j;
        n++;
      }
    }
  }
}

}  // namespace mmd

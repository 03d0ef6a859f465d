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
// slot in the concatenation of the tile's runs (the slot -> atom map is a private copy of the CSR bins,
// each bin sorted by x for the interval build below).  For every local atom the list stores ALL neighbors
// within cutneigh as 16-bit tile-local indices, in ascending order; bit 15 marks the entries that belong
// to the reference's half list (ref/neighbor.cpp:154-171).  The reference's rows, numneigh and totals are
// recovered exactly (tile_rows_export_kernel orders a row by the entries' original CSR position, which is
// the reference's stencil-then-bin order, ref/neighbor.cpp:141-183) while the force kernels evaluate every
// atom's complete neighborhood ("owner computes"): no scatter to f[j], no reverse halo, no clearing of f.
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
                  int* __restrict__ row_counter, int* __restrict__ max_rows) {
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
  if (lane == 0 && any) {
    base = atomicAdd(row_counter, tile_rows);
    atomicMax(max_rows, tile_rows);
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  // .w: atoms of the halo window (-1: the tile owns no local atom)
  if (lane < TILE_NCENTER) center[(size_t)t * TILE_NCENTER + lane] = make_int4(lo, hi, any ? base + incl - len_c : 0, any ? carry : -1);
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
// Neighbor build, tile-resident version: one CTA per tile, one warp per centre pencil.
//
// The halo window is staged into shared memory as FP32 coordinates -- for an FP64 build relative to the window
// origin, so their error is <= 2^-24 * window extent -- together with the tile-local start index of every bin of the
// window.  The candidates of a bin are then ~25 contiguous ranges of tile-local indices, flattened once per bin into
// a per-warp table; for every local atom of the bin the warp sweeps that table 32 candidates at a time (conflict-
// free shared-memory reads, the atom's own coordinates in registers, row length in a warp-uniform register).
// FP64 build: the FP32 distance only decides candidates farther than `band` from cutneighsq (band = 1e-4 * cutneighsq,
// > 20x the worst-case FP32 error).  A sweep that met a candidate inside the band is rolled back and
// repeated by the general loop, which re-evaluates such candidates from the FP64 positions with unfused arithmetic --
// so every accept/reject decision is exactly the reference's (rsq <= cutneighsq, ref/neighbor.cpp:165,179).
// FP32 build: absolute coordinates, unfused FP32 arithmetic, exact.  The general loop also serves the cases that need
// atom ids or FP64 coordinates: ghosts in the atom's own bin (MODE 1) and the j>i filter of MODE 2.
// Row order, half-list flag and counters are those of neigh_build_tile_kernel.
// ---------------------------------------------------------------------------------------
constexpr int TB2_MAXSR = 64;    // stencil runs of the symmetric stencil (25 for the usual 5x5x5 stencil)
constexpr int TB2_NC = 768;      // candidates per table block (a 5x5x5 stencil holds ~570)
constexpr int TB2_MAXH = 8190;   // tile-local indices must fit 13 bits (+ sentinel)
constexpr unsigned short TB2_OWN = 0x2000, TB2_UP = 0x4000;   // candidate class in bits 13-14 (0: lower stencil half)
constexpr int TB2_THREADS = 256, TB2_WARPS = TB2_THREADS / 32;  // each warp takes two centre pencils

template <class T> struct Build2Params {
  T cut0;         // cutneighsq when all type pairs share it
  float band;     // FP64 build: |rsq32 - cut| <= band -> exact FP64 test
  float rcull;    // sqrt(max cutneighsq) plus a safety margin: the half-width of an atom's candidate interval derives from it
  double binsize[3];
  int mbinlo[3];
  int uniform_cut;
};

template <class T> __host__ __device__ inline size_t build2_smem_bytes(const TileGeo& g, int hcap, bool with_types) {
  return (size_t)hcap * 3 * sizeof(float) + (with_types ? (size_t)hcap : 0) + (size_t)g.nrun * (TBX + 2 * g.sx + 1) * sizeof(int) +
         (size_t)TB2_WARPS * TB2_MAXSR * 3 * sizeof(int) + (2 * TILE_MAXRUN + 1) * sizeof(int) +
         (size_t)TB2_WARPS * TB2_NC * sizeof(unsigned short) + 64;
}

// rare paths of the build, kept out of line so that the hot loop carries no predicated FP64 code
template <class T>
__device__ __noinline__ bool build2_exact_within(const Vec4<T>* __restrict__ x, int id_i, int id_j, T cut) {
  const Vec4<T> xi = x[id_i];
  const Vec4<T> xj = x[id_j];
  return rsq_unfused(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z) <= cut;
}
// ref/neighbor.cpp:154-157: a ghost in the atom's own bin is skipped when it lies lexicographically (z,y,x) below
template <class T>
__device__ __noinline__ bool build2_ghost_below(const Vec4<T>* __restrict__ x, int id_i, int id_j) {
  const Vec4<T> xi = x[id_i];
  const Vec4<T> xj = x[id_j];
  return (xj.z < xi.z) || (xj.z == xi.z && xj.y < xi.y) || (xj.z == xi.z && xj.y == xi.y && xj.x < xi.x);
}

template <class T, int MODE, int UC>
__global__ void __launch_bounds__(TB2_THREADS, 4)
neigh_build_tile2_kernel(const Vec4<T>* __restrict__ x, int nlocal, const int* __restrict__ bin_start,
                         const int* __restrict__ bin_atoms, int mbins, const StencilRun* __restrict__ sruns, int nsr,
                         const T* __restrict__ cutneighsq, int ntypes, TileGeo g, Build2Params<T> B,
                         const int2* __restrict__ tile_runs, const int4* __restrict__ tile_center,
                         const int2* __restrict__ tile_info, unsigned short* __restrict__ rows, int tcap,
                         int* __restrict__ numneigh_half, int2* __restrict__ row_atom, int* __restrict__ status,
                         int* __restrict__ max_half, int* __restrict__ max_full, unsigned long long* __restrict__ total_half) {
  extern __shared__ __align__(16) unsigned char b2_smem[];
  const int t = blockIdx.x;
  const int2 inf = tile_info[t];
  if (inf.y == 0) return;
  const int H = inf.x;
  const int WX1 = TBX + 2 * g.sx + 1;
  float* sx = reinterpret_cast<float*>(b2_smem);
  float* sy = sx + g.hcap;
  float* sz = sy + g.hcap;
  int* s_binoff = reinterpret_cast<int*>(sz + g.hcap);           // [nrun][WX1] tile-local start index of each window bin
  int* s_rstart = s_binoff + g.nrun * WX1;                        // [16][TB2_MAXSR] per warp: first index of a stencil range
  int* s_rpref = s_rstart + TB2_WARPS * TB2_MAXSR;                // [warps][TB2_MAXSR] exclusive prefix of range lengths
  int* s_rinfo = s_rpref + TB2_WARPS * TB2_MAXSR;                 // [warps][TB2_MAXSR] slot0 of the pencil | flags in bits 30,31
  int* s_run_start = s_rinfo + TB2_WARPS * TB2_MAXSR;             // [nrun]
  int* s_run_off = s_run_start + TILE_MAXRUN;                     // [nrun+1]
  unsigned short* s_ctab = reinterpret_cast<unsigned short*>(s_run_off + TILE_MAXRUN + 1);  // [16][TB2_NC] candidate table
  unsigned char* st = reinterpret_cast<unsigned char*>(s_ctab + TB2_WARPS * TB2_NC);       // [hcap] types

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int tx = t % g.ntx, ty = (t / g.ntx) % g.nty, tz = t / (g.ntx * g.nty);
  const int bx0 = tx * TBX - g.ox, by0 = ty * TBY - g.oy, bz0 = tz * TBZ - g.oz;
  const int xlo = max(0, bx0 - g.sx), xhi = min(g.mbx, bx0 + TBX + g.sx);
  // window origin (only used to keep the FP32 images small; any point near the tile would do)
  const T org_x = sizeof(T) == 8 ? (T)((bx0 - g.sx + B.mbinlo[0]) * B.binsize[0]) : (T)0;
  const T org_y = sizeof(T) == 8 ? (T)((by0 - g.sy + B.mbinlo[1]) * B.binsize[1]) : (T)0;
  const T org_z = sizeof(T) == 8 ? (T)((bz0 - g.sz + B.mbinlo[2]) * B.binsize[2]) : (T)0;

  const int2* tr = tile_runs + (size_t)t * g.nrun;
  for (int p = threadIdx.x; p < g.nrun; p += blockDim.x) {
    const int2 r = tr[p];
    s_run_start[p] = r.x;
    s_run_off[p] = r.y;
  }
  if (threadIdx.x == 0) {
    s_run_off[g.nrun] = H;
    sx[H] = sy[H] = sz[H] = 1.0e18f;  // sentinel candidate: fails every distance test
    if (!UC) st[H] = 0;
  }
  __syncthreads();
  // bin offsets of the window
  for (int e = threadIdx.x; e < g.nrun * WX1; e += blockDim.x) {
    const int p = e / WX1, wx = e - p * WX1;
    const int y = by0 - g.sy + p % g.nry, z = bz0 - g.sz + p / g.nry;
    int v = s_run_off[p + 1];
    if (y >= 0 && y < g.mby && z >= 0 && z < g.mbz && xlo + wx < xhi)
      v = s_run_off[p] + (bin_start[min(tile_bin_id(g, xlo + wx, y, z), mbins)] - s_run_start[p]);
    s_binoff[e] = v;
  }
  // FP32 images of the positions (+ types)
  for (int p = w; p < g.nrun; p += nw) {
    const int start = s_run_start[p], off = s_run_off[p], len = s_run_off[p + 1] - off;
    for (int k = lane; k < len; k += 32) {
      const int id = __ldg(bin_atoms + start + k);
      const Vec4<T> v = ldg4(x + id);
      sx[off + k] = (float)(v.x - org_x);
      sy[off + k] = (float)(v.y - org_y);
      sz[off + k] = (float)(v.z - org_z);
      if (!UC) st[off + k] = (unsigned char)lane_to_type(v.w);
    }
  }
  __syncthreads();

  const unsigned lt_mask = (1u << lane) - 1u;
  int* rstart = s_rstart + w * TB2_MAXSR;
  int* rpref = s_rpref + w * TB2_MAXSR;
  int* rinfo = s_rinfo + w * TB2_MAXSR;
  unsigned short* ctab = s_ctab + w * TB2_NC;
  const float fcut0 = (float)B.cut0, band = B.band;
  int warp_max_h = 0, warp_max_f = 0;
  unsigned long long warp_total = 0ull;

  for (int cr = w; cr < TILE_NCENTER; cr += nw) {
    const int cy = cr % TBY, cz = cr / TBY;
    const int by = by0 + cy, bz = bz0 + cz;
    if (by < 0 || by >= g.mby || bz < 0 || bz >= g.mbz) continue;
    const int4 ce = tile_center[(size_t)t * TILE_NCENTER + cr];
    const int pc = (cy + g.sy) + (cz + g.sz) * g.nry;
    const int slot0_c = s_run_start[pc] - s_run_off[pc];
    for (int cx = 0; cx < TBX; cx++) {
      const int bx = bx0 + cx;
      if (bx < 0 || bx >= g.mbx) continue;
      const int xw = bx - xlo;
      const int own_lo = s_binoff[pc * WX1 + xw], own_hi = s_binoff[pc * WX1 + xw + 1];
      if (own_hi == own_lo) continue;
      // does the bin own local atoms?  (ids ascend inside a bin: locals first)
      if (__ldg(bin_atoms + slot0_c + own_lo) >= nlocal) continue;

      // ---- candidate ranges of this bin, in stencil order ----
      int ncand = 0;
      for (int r0 = 0; r0 < nsr; r0 += 32) {
        const int r = r0 + lane;
        int start = 0, len = 0, info = 0;
        if (r < nsr) {
          const StencilRun run = sruns[r];
          const int dxlo = run.off - (run.dz * g.mby + run.dy) * g.mbx;
          const int yy = by + run.dy, zz = bz + run.dz;
          const int wlo = xw + dxlo, whi = wlo + run.len;
          if (yy < 0 || yy >= g.mby || zz < 0 || zz >= g.mbz || wlo < 0 || xlo + whi > xhi) {
            atomicOr(status, 2);
          } else {
            const int p = (cy + g.sy + run.dy) + (cz + g.sz + run.dz) * g.nry;
            start = s_binoff[p * WX1 + wlo];
            len = s_binoff[p * WX1 + whi] - start;
            const bool upper = run.dz > 0 || (run.dz == 0 && run.dy > 0);
            const bool ownp = run.dz == 0 && run.dy == 0;
            info = ((s_run_start[p] - s_run_off[p]) & 0x3fffffff) | (upper ? 0x80000000 : 0) | (ownp ? 0x40000000 : 0);
          }
        }
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        if (r < nsr) {
          rstart[r] = start;
          rpref[r] = ncand + incl - len;
          rinfo[r] = info;
        }
        ncand += __shfl_sync(0xffffffffu, incl, 31);
      }
      __syncwarp();

      // first tile-local index of the bin that is NOT a local atom (ids ascend inside a bin: ghosts come last)
      int ghost_lo = own_hi;
      for (int k0 = own_lo; k0 < own_hi; k0 += 32) {
        const int k = k0 + lane;
        const bool gh = k < own_hi && __ldg(bin_atoms + slot0_c + k) >= nlocal;
        const unsigned mg = __ballot_sync(0xffffffffu, gh);
        if (mg) { ghost_lo = k0 + __ffs(mg) - 1; break; }
      }

      // ---- local atoms of the bin, 32 at a time; lane tt keeps the row lengths of atom i0+tt ----
      for (int i0 = own_lo; i0 < ghost_lo; i0 += 32) {
        const int nloc = min(32, ghost_lo - i0);
        const int q_i0 = ce.z + (i0 - ce.x);  // rows of consecutive atoms of a pencil are consecutive
        int my_n = 0, my_h = 0;

        for (int cb0 = 0; cb0 < ncand; cb0 += TB2_NC) {
          const int nblk = min(TB2_NC, ncand - cb0);
          const int nblk32 = (nblk + 31) & ~31;
          // ---- flatten this block of candidates: tile-local index | own-bin bit | upper-stencil bit ----
          bool slow_blk = (MODE == 2);
          __syncwarp();
          // one stencil range after the other (warp-uniform loop), 32 entries per step
          for (int r = 0; r < nsr; r++) {
            const int pref = rpref[r];
            const int len = (r + 1 < nsr ? rpref[r + 1] : ncand) - pref;
            if (len == 0 || pref + len <= cb0 || pref >= cb0 + nblk) continue;
            const int start = rstart[r], info = rinfo[r];
            for (int j = lane; j < len; j += 32) {
              const int k = pref + j - cb0;
              if (k >= 0 && k < nblk) {
                const int lc = start + j;
                const bool own_bin = (info & 0x40000000) && lc >= own_lo && lc < own_hi;
                const bool upper = (info & 0x80000000) != 0 || ((info & 0x40000000) && lc >= own_hi);
                ctab[k] = (unsigned short)(lc | (own_bin ? TB2_OWN : 0) | (upper ? TB2_UP : 0));
                if (MODE == 1 && own_bin && lc >= ghost_lo) slow_blk = true;
              }
            }
          }
          if (nblk + lane < nblk32) ctab[nblk + lane] = (unsigned short)H;  // padding: the sentinel
          slow_blk = __any_sync(0xffffffffu, slow_blk);
          __syncwarp();

          // atoms are swept in pairs: the 32 candidates of a sweep are fetched once and tested against both
          for (int tt = 0; tt < nloc; tt += 2) {
            const int na = min(2, nloc - tt);
            const int liA = i0 + tt, liB = i0 + min(tt + 1, nloc - 1);
            const float xA = sx[liA], yA = sy[liA], zA = sz[liA], xB = sx[liB], yB = sy[liB], zB = sz[liB];
            const int tA = UC ? 0 : (int)st[liA], tB = UC ? 0 : (int)st[liB];
            unsigned short* rowA = rows + (size_t)(q_i0 + tt) * tcap;
            unsigned short* rowB = rowA + tcap;
            const int nA_s = __shfl_sync(0xffffffffu, my_n, tt), hA_s = __shfl_sync(0xffffffffu, my_h, tt);
            const int nB_s = __shfl_sync(0xffffffffu, my_n, (tt + 1) & 31), hB_s = __shfl_sync(0xffffffffu, my_h, (tt + 1) & 31);
            int nA = nA_s, hA = hA_s, nB = nB_s, hB = hB_s;
            bool redo = slow_blk;
            if (!slow_blk) {
              // ---- lean loop: FP32 decision, no ids.  key = index | class bits; with thr = li | OWN:
              //      key == thr <=> the atom itself;  key > thr <=> upper stencil half, or own bin and j > i ----
              bool closeany = false;
              const int thrA = liA | TB2_OWN, thrB = liB | TB2_OWN;
              const bool two = na == 2;
              for (int k0 = 0; k0 < nblk32; k0 += 32) {
                const int key = ctab[k0 + lane] & 0x7fff;
                const int lc = key & 0x1fff;
                const float cxj = sx[lc], cyj = sy[lc], czj = sz[lc];
                float fcA = fcut0, fcB = fcut0;
                if (!UC) {
                  const int tj = (int)st[lc];
                  fcA = (float)__ldg(&cutneighsq[tA * ntypes + tj]);
                  fcB = (float)__ldg(&cutneighsq[tB * ntypes + tj]);
                }
                bool okA, okB;
                {
                  const float dx = xA - cxj, dy = yA - cyj, dz = zA - czj;
                  if (sizeof(T) == 4) {
                    okA = rsq_unfused(dx, dy, dz) <= fcA;
                  } else {
                    const float d = (dx * dx + dy * dy + dz * dz) - fcA;
                    okA = d < -band;
                    closeany = closeany || (fabsf(d) <= band);
                  }
                }
                {
                  const float dx = xB - cxj, dy = yB - cyj, dz = zB - czj;
                  if (sizeof(T) == 4) {
                    okB = rsq_unfused(dx, dy, dz) <= fcB;
                  } else {
                    const float d = (dx * dx + dy * dy + dz * dz) - fcB;
                    okB = d < -band;
                    closeany = closeany || (two && fabsf(d) <= band);
                  }
                }
                okA = okA && key != thrA;
                okB = okB && key != thrB && two;
                const unsigned mA = __ballot_sync(0xffffffffu, okA);
                const unsigned mB = __ballot_sync(0xffffffffu, okB);
                const bool halfA = MODE == 1 ? key > thrA : true, halfB = MODE == 1 ? key > thrB : true;
                const unsigned mhA = MODE == 0 ? mA : __ballot_sync(0xffffffffu, okA && halfA);
                const unsigned mhB = MODE == 0 ? mB : __ballot_sync(0xffffffffu, okB && halfB);
                if (okA) {
                  const int pos = nA + __popc(mA & lt_mask);
                  if (pos < tcap) rowA[pos] = (unsigned short)(lc | ((MODE != 0 && halfA) ? TILE_HALF_BIT : 0));
                }
                if (okB) {
                  const int pos = nB + __popc(mB & lt_mask);
                  if (pos < tcap) rowB[pos] = (unsigned short)(lc | ((MODE != 0 && halfB) ? TILE_HALF_BIT : 0));
                }
                nA += __popc(mA); hA += __popc(mhA);
                nB += __popc(mB); hB += __popc(mhB);
              }
              redo = sizeof(T) == 8 && __any_sync(0xffffffffu, closeany);
            }
            if (redo) {
              nA = nA_s; hA = hA_s; nB = nB_s; hB = hB_s;
            }
            for (int a = 0; redo && a < na; a++) {
              const int li = a ? liB : liA;
              const float xi = a ? xB : xA, yi = a ? yB : yA, zi = a ? zB : zA;
              const int ti = a ? tB : tA;
              unsigned short* rowp = a ? rowB : rowA;
              const int n_s = a ? nB_s : nA_s, h_s = a ? hB_s : hA_s;
              int n_t = n_s, h_t = h_s;
              {
              // ---- general loop: same sweep with ids at hand (exact FP64 test inside the band, ghost and j>i filters) ----
              n_t = n_s;
              h_t = h_s;
              for (int k0 = 0; k0 < nblk; k0 += 32) {
                const int k = k0 + lane;
                const bool valid = k < nblk;
                const unsigned e = ctab[valid ? k : 0];
                const int lc = e & 0x1fff;
                int slot_j = 0;
                {
                  const int n = cb0 + (valid ? k : 0);
                  int lo = 0, hi = nsr - 1;
                  while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (rpref[mid] <= n) lo = mid; else hi = mid - 1;
                  }
                  slot_j = (rinfo[lo] << 2) >> 2;  // low 30 bits, sign-extended: CSR slot of tile-local index 0 of that pencil
                }
                const float dx = xi - sx[lc], dy = yi - sy[lc], dz = zi - sz[lc];
                T cut = B.cut0;
                if (!UC) cut = __ldg(&cutneighsq[ti * ntypes + (int)st[lc]]);
                bool ok;
                if (sizeof(T) == 4) {
                  ok = rsq_unfused(dx, dy, dz) <= (float)cut;
                } else {
                  const float d = (dx * dx + dy * dy + dz * dz) - (float)cut;
                  ok = d < -band;
                  if (valid && fabsf(d) <= band && lc != li)
                    ok = build2_exact_within<T>(x, __ldg(bin_atoms + slot0_c + li), __ldg(bin_atoms + slot_j + lc), cut);
                }
                ok = ok && valid && lc != li;
                bool half = true;
                if (MODE == 1) {
                  half = (e & TB2_UP) || ((e & TB2_OWN) && lc > li);
                  if (ok && half && (e & TB2_OWN) && lc >= ghost_lo)
                    half = !build2_ghost_below<T>(x, __ldg(bin_atoms + slot0_c + li), __ldg(bin_atoms + slot0_c + lc));
                }
                if (MODE == 2) {
                  half = false;
                  if (ok) half = __ldg(bin_atoms + slot_j + lc) > __ldg(bin_atoms + slot0_c + li);
                }
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                const unsigned mh = MODE == 0 ? m : __ballot_sync(0xffffffffu, ok && half);
                if (ok) {
                  const int pos = n_t + __popc(m & lt_mask);
                  if (pos < tcap) rowp[pos] = (unsigned short)(lc | ((MODE != 0 && half) ? TILE_HALF_BIT : 0));
                }
                n_t += __popc(m);
                h_t += __popc(mh);
              }
                          }
              if (a) { nB = n_t; hB = h_t; } else { nA = n_t; hA = h_t; }
            }
            if (lane == tt) { my_n = nA; my_h = hA; }
            if (na == 2 && lane == tt + 1) { my_n = nB; my_h = hB; }
          }
        }
        if (lane < nloc) {
          const int my_id = __ldg(bin_atoms + slot0_c + i0 + lane);
          numneigh_half[my_id] = my_h;
          row_atom[q_i0 + lane] = make_int2(my_id, my_n);
          warp_max_h = max(warp_max_h, my_h);
          warp_max_f = max(warp_max_f, my_n);
          warp_total += (unsigned long long)my_h;
        }
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    warp_max_h = max(warp_max_h, __shfl_xor_sync(0xffffffffu, warp_max_h, o));
    warp_max_f = max(warp_max_f, __shfl_xor_sync(0xffffffffu, warp_max_f, o));
    warp_total += __shfl_xor_sync(0xffffffffu, warp_total, o);
  }
  if (lane == 0) {
    atomicMax(max_half, warp_max_h);
    atomicMax(max_full, warp_max_f);
    atomicAdd(total_half, warp_total);
  }
}

// ---------------------------------------------------------------------------------------
// x-sorted windows.  The slot -> atom map the tile kernels stage from is a private copy of the CSR bins; inside every
// bin it may be put in any order, because tile-local indices only have to agree between the build and the force
// kernels.  Sorting each bin by x makes every pencil run of a halo window ascending in x (bins ascend along the run),
// so the candidates of an atom in a pencil are ONE interval [x_i - xr, x_i + xr] found by two binary searches:
// the build then tests ~140 candidates per atom instead of the ~575 of the bin stencil.  oslot keeps the original
// CSR position of every entry: the export sorts rows by it, which restores the reference's row order.
// status |= 8 when a bin is too full for the thread-local sort (the caller then builds without x-sorted windows).
// ---------------------------------------------------------------------------------------
template <bool V> struct TileTag { static constexpr bool value = V; };
constexpr int XSORT_MAX = 48;
template <class T>
__global__ void bin_xsort_kernel(const Vec4<T>* __restrict__ x, const int* __restrict__ bin_start,
                                 const int* __restrict__ bin_atoms, int mbins, int* __restrict__ slots,
                                 int* __restrict__ oslot, int* __restrict__ status) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= mbins) return;
  const int s0 = bin_start[b], n = bin_start[b + 1] - s0;
  if (n <= 0) return;
  if (n > XSORT_MAX) {
    atomicOr(status, 8);
    for (int k = 0; k < n; k++) { slots[s0 + k] = bin_atoms[s0 + k]; oslot[s0 + k] = s0 + k; }
    return;
  }
  T key[XSORT_MAX];
  int id[XSORT_MAX], os[XSORT_MAX];
  for (int k = 0; k < n; k++) {
    const int a = bin_atoms[s0 + k];
    const T xa = x[a].x;
    int j = k - 1;
    while (j >= 0 && key[j] > xa) {  // stable: equal keys keep their id order
      key[j + 1] = key[j]; id[j + 1] = id[j]; os[j + 1] = os[j];
      j--;
    }
    key[j + 1] = xa; id[j + 1] = a; os[j + 1] = s0 + k;
  }
  for (int k = 0; k < n; k++) { slots[s0 + k] = id[k]; oslot[s0 + k] = os[k]; }
}

// ---------------------------------------------------------------------------------------
// Neighbor build on x-sorted windows: one CTA per tile, one warp per centre pencil, one atom at a time.
// Lane r owns stencil run r (<= 32 runs): it turns the run's bins into the index interval of this atom's candidates
// (slab distance in y/z leaves an x half-width xr; two binary searches over the run's staged x).  The non-empty
// intervals are compacted and their prefix sums kept in shared memory; a sweep of 32 candidates finds its interval
// with one REDUX.OR + POPC.  Distance test, guard band, exact FP64 re-test, half-list flag, counters and row layout
// are those of neigh_build_tile2_kernel; rows come out in candidate order (ascending tile-local index).
// ---------------------------------------------------------------------------------------
// 16-bit store to global memory (the pointer went through an opaque asm, so the compiler no longer knows its space)
__device__ __forceinline__ void stg_u16(unsigned short* p, unsigned short v) {
  asm volatile("st.global.u16 [%0], %1;" ::"l"(p), "h"(v) : "memory");
}

template <class T> __host__ __device__ inline size_t build3_smem_bytes(const TileGeo& g, int hcap, bool with_types) {
  return (size_t)hcap * 3 * sizeof(float) + (with_types ? (size_t)hcap : 0) + (size_t)g.nrun * (TBX + 2 * g.sx + 1) * sizeof(int) +
         (size_t)TB2_WARPS * 32 * sizeof(int4) + (2 * TILE_MAXRUN + 1) * sizeof(int) + 96;
}

template <class T, int MODE, int UC>
__global__ void __launch_bounds__(TB2_THREADS, 4)
neigh_build_tile3_kernel(const Vec4<T>* __restrict__ x, int nlocal, const int* __restrict__ bin_start,
                         const int* __restrict__ slots, int mbins, const StencilRun* __restrict__ sruns, int nsr,
                         const T* __restrict__ cutneighsq, int ntypes, TileGeo g, Build2Params<T> B,
                         const int2* __restrict__ tile_runs, const int4* __restrict__ tile_center,
                         const int2* __restrict__ tile_info, unsigned short* __restrict__ rows, int tcap,
                         int* __restrict__ numneigh_half, int2* __restrict__ row_atom, int* __restrict__ status,
                         int* __restrict__ max_half, int* __restrict__ max_full, unsigned long long* __restrict__ total_half,
                         int pair /* 1: two atoms of a bin per sweep */) {
  extern __shared__ __align__(16) unsigned char b3_smem[];
  const int t = blockIdx.x;
  const int2 inf = tile_info[t];
  if (inf.y == 0) return;
  const int H = inf.x;
  const int WX1 = TBX + 2 * g.sx + 1;
  float* sx = reinterpret_cast<float*>(b3_smem);
  float* sy = sx + g.hcap;
  float* sz = sy + g.hcap;
  // [warps][32] dense intervals of the current atom: {first index - prefix, prefix of lengths, class | slot0*4, -}
  // (directly behind the coordinate arrays: hcap is a multiple of 64, so the records are 16-byte aligned)
  int4* s_dense = reinterpret_cast<int4*>(sz + g.hcap);
  int* s_binoff = reinterpret_cast<int*>(s_dense + TB2_WARPS * 32);   // [nrun][WX1] tile-local start index of each window bin
  int* s_run_start = s_binoff + g.nrun * WX1;                         // [nrun]
  int* s_run_off = s_run_start + TILE_MAXRUN;             // [nrun+1]
  unsigned char* st = reinterpret_cast<unsigned char*>(s_run_off + TILE_MAXRUN + 1);  // [hcap] types (!UC)

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int tx = t % g.ntx, ty = (t / g.ntx) % g.nty, tz = t / (g.ntx * g.nty);
  const int bx0 = tx * TBX - g.ox, by0 = ty * TBY - g.oy, bz0 = tz * TBZ - g.oz;
  const int xlo = max(0, bx0 - g.sx), xhi = min(g.mbx, bx0 + TBX + g.sx);
  const T org_x = sizeof(T) == 8 ? (T)((bx0 - g.sx + B.mbinlo[0]) * B.binsize[0]) : (T)0;
  const T org_y = sizeof(T) == 8 ? (T)((by0 - g.sy + B.mbinlo[1]) * B.binsize[1]) : (T)0;
  const T org_z = sizeof(T) == 8 ? (T)((bz0 - g.sz + B.mbinlo[2]) * B.binsize[2]) : (T)0;

  const int2* tr = tile_runs + (size_t)t * g.nrun;
  for (int p = threadIdx.x; p < g.nrun; p += blockDim.x) {
    const int2 r = tr[p];
    s_run_start[p] = r.x;
    s_run_off[p] = r.y;
  }
  if (threadIdx.x == 0) s_run_off[g.nrun] = H;
  __syncthreads();
  for (int e = threadIdx.x; e < g.nrun * WX1; e += blockDim.x) {
    const int p = e / WX1, wx = e - p * WX1;
    const int y = by0 - g.sy + p % g.nry, z = bz0 - g.sz + p / g.nry;
    int v = s_run_off[p + 1];
    if (y >= 0 && y < g.mby && z >= 0 && z < g.mbz && xlo + wx < xhi)
      v = s_run_off[p] + (bin_start[min(tile_bin_id(g, xlo + wx, y, z), mbins)] - s_run_start[p]);
    s_binoff[e] = v;
  }
  for (int p = w; p < g.nrun; p += nw) {
    const int start = s_run_start[p], off = s_run_off[p], len = s_run_off[p + 1] - off;
    for (int k = lane; k < len; k += 32) {
      const int id = __ldg(slots + start + k);
      const Vec4<T> v = ldg4(x + id);
      sx[off + k] = (float)(v.x - org_x);
      sy[off + k] = (float)(v.y - org_y);
      sz[off + k] = (float)(v.z - org_z);
      if (!UC) st[off + k] = (unsigned char)lane_to_type(v.w);
    }
  }
  // the last slot of the window (hcap >= atoms + 8) is a far-away atom: lanes past the end of a sweep test it
  const int far_slot = g.hcap - 1;
  if (threadIdx.x == 0) {
    sx[far_slot] = 3.0e18f; sy[far_slot] = 3.0e18f; sz[far_slot] = 3.0e18f;
    if (!UC) st[far_slot] = 0;
  }
  __syncthreads();

  const unsigned lt_mask = (1u << lane) - 1u;
  // dense intervals of this warp's current pair: {first index - prefix, class | slot0*4} and, apart, the prefix of lengths
  int2* dense = reinterpret_cast<int2*>(s_dense) + w * 32;
  int* dpref = reinterpret_cast<int*>(reinterpret_cast<int2*>(s_dense) + TB2_WARPS * 32) + w * 32;
  const int tile_q0 = tile_center[(size_t)t * TILE_NCENTER].z;   // first row of the tile (tile_table_kernel)
  unsigned short* rows_t = rows + (size_t)tile_q0 * tcap;
  asm volatile("" : "+l"(rows_t));   // keep the sum in registers (ptxas otherwise re-derives it in front of every store)
  const float bsy = (float)B.binsize[1], bsz = (float)B.binsize[2];
  const float ey0 = (float)((by0 - g.sy + B.mbinlo[1]) * B.binsize[1] - (double)org_y);
  const float ez0 = (float)((bz0 - g.sz + B.mbinlo[2]) * B.binsize[2] - (double)org_z);
  const float rc = B.rcull, rc2 = rc * rc, band = B.band;
  float fcut0 = (float)B.cut0;
  asm volatile("" : "+f"(fcut0));   // (keeps the FP64 -> FP32 conversion out of the sweep loop)
  int warp_max_h = 0, warp_max_f = 0;
  unsigned long long warp_total = 0ull;

  for (int cr = w; cr < TILE_NCENTER; cr += nw) {
    const int cy = cr % TBY, cz = cr / TBY;
    const int by = by0 + cy, bz = bz0 + cz;
    if (by < 0 || by >= g.mby || bz < 0 || bz >= g.mbz) continue;
    const int4 ce = tile_center[(size_t)t * TILE_NCENTER + cr];
    const int pc = (cy + g.sy) + (cz + g.sz) * g.nry;
    const int slot0_c = s_run_start[pc] - s_run_off[pc];
    const float ey_own = ey0 + (float)(cy + g.sy) * bsy, ez_own = ez0 + (float)(cz + g.sz) * bsz;
    for (int cx = 0; cx < TBX; cx++) {
      const int bx = bx0 + cx;
      if (bx < 0 || bx >= g.mbx) continue;
      const int xw = bx - xlo;
      const int own_lo = s_binoff[pc * WX1 + xw], own_hi = s_binoff[pc * WX1 + xw + 1];
      if (own_hi == own_lo) continue;
      // ---- lane r: stencil run r of this bin (which window bins it covers, its pencil) ----
      int r_row = 0, r_wlo = 0, r_whi = -1, r_info = 0, r_slot0 = 0;
      float r_ylo = 0.f, r_zlo = 0.f;
      if (lane < nsr) {
        const StencilRun run = sruns[lane];
        const int dxlo = run.off - (run.dz * g.mby + run.dy) * g.mbx;
        const int yy = by + run.dy, zz = bz + run.dz;
        const int wlo = xw + dxlo, whi = wlo + run.len;
        if (yy < 0 || yy >= g.mby || zz < 0 || zz >= g.mbz || wlo < 0 || xlo + whi > xhi) {
          // only matters if the bin owns local atoms (checked per atom below)
          r_whi = -2;
        } else {
          const int p = (cy + g.sy + run.dy) + (cz + g.sz + run.dz) * g.nry;
          r_row = p * WX1; r_wlo = wlo; r_whi = whi;
          r_ylo = ey_own + (float)run.dy * bsy;
          r_zlo = ez_own + (float)run.dz * bsz;
          r_slot0 = s_run_start[p] - s_run_off[p];
          const bool upper = run.dz > 0 || (run.dz == 0 && run.dy > 0);
          const bool ownp = run.dz == 0 && run.dy == 0;
          r_info = (upper ? 2 : 0) | (ownp ? 1 : 0);
        }
      }
      const bool bad_geom = __any_sync(0xffffffffu, r_whi == -2);

      // ---- the local atoms of the bin, TWO at a time: consecutive atoms of an x-sorted bin have almost the same
      //      candidates, so the pair shares one interval search per run (the union of the two intervals) and every
      //      sweep of 32 candidates is fetched once and tested against both ----
      for (int a = own_lo; a < own_hi;) {
        const int idA = __ldg(slots + slot0_c + a);
        if (idA >= nlocal) { a++; continue; }  // warp-uniform
        int idB = 0x7fffffff;
        if (pair && a + 1 < own_hi) idB = __ldg(slots + slot0_c + a + 1);
        const bool two = idB < nlocal;
        const int aA = a, aB = two ? a + 1 : a;
        a += two ? 2 : 1;
        if (bad_geom) { if (lane == 0) atomicOr(status, 2); continue; }
        // (a single atom: B sits far away and never finds a neighbor)
        const float xA = sx[aA], yA = sy[aA], zA = sz[aA];
        const float xB = two ? sx[aB] : -3.0e18f, yB = two ? sy[aB] : -3.0e18f, zB = two ? sz[aB] : -3.0e18f;
        const int tA = UC ? 0 : (int)st[aA], tB = UC ? 0 : (int)st[aB];
        const int qA = ce.z + (aA - ce.x), qB = ce.z + (aB - ce.x);
        // rows of a tile are contiguous: one 64-bit base per CTA, a 32-bit offset per atom
        unsigned oA = (unsigned)(qA - tile_q0) * (unsigned)tcap, oB = (unsigned)(qB - tile_q0) * (unsigned)tcap;
        asm volatile("" : "+r"(oA), "+r"(oB));

        // ---- candidate interval of the pair in every run ----
        int L = 0, len = 0;
        if (lane < nsr && r_whi >= 0) {
          float xa = 3.0e38f, xb = -3.0e38f;
          {
            const float gy = fmaxf(0.0f, fmaxf(r_ylo - yA, yA - (r_ylo + bsy)));
            const float gz = fmaxf(0.0f, fmaxf(r_zlo - zA, zA - (r_zlo + bsz)));
            const float rem = rc2 - gy * gy - gz * gz;
            if (rem > 0.0f) { const float xr = sqrtf(rem); xa = xA - xr; xb = xA + xr; }
          }
          if (two) {
            const float gy = fmaxf(0.0f, fmaxf(r_ylo - yB, yB - (r_ylo + bsy)));
            const float gz = fmaxf(0.0f, fmaxf(r_zlo - zB, zB - (r_zlo + bsz)));
            const float rem = rc2 - gy * gy - gz * gz;
            if (rem > 0.0f) { const float xr = sqrtf(rem); xa = fminf(xa, xB - xr); xb = fmaxf(xb, xB + xr); }
          }
          if (xa <= xb) {
            const int lo_i = s_binoff[r_row + r_wlo], hi_i = s_binoff[r_row + r_whi];
            int l0 = lo_i, l1 = hi_i;  // first index with sx >= xa
            while (l0 < l1) {
              const int mid = (l0 + l1) >> 1;
              if (sx[mid] < xa) l0 = mid + 1; else l1 = mid;
            }
            L = l0;
            l1 = hi_i;                  // first index with sx > xb
            while (l0 < l1) {
              const int mid = (l0 + l1) >> 1;
              if (sx[mid] <= xb) l0 = mid + 1; else l1 = mid;
            }
            len = l0 - L;
          }
        }
        const unsigned nz = __ballot_sync(0xffffffffu, len > 0);
        const int nd = __popc(nz);
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const int M = __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
        if (len > 0) {  // flags in the two low bits of .y, CSR slot of tile-local index 0 above them
          const int k = __popc(nz & lt_mask);
          dense[k] = make_int2(L - (incl - len), r_slot0 * 4 + r_info);
          dpref[k] = incl - len;
        }
        __syncwarp();
        const int my_dpref = lane < nd ? dpref[lane] : 0x7fffffff;

        int nA = 0, hA = 0, nB = 0, hB = 0;
        // one pass over the pair's candidates, 32 per sweep; EXACT adds the FP64 re-test of candidates inside the guard band
        auto sweep_all = [&](auto exact_tag) -> bool {
          constexpr bool EXACT = decltype(exact_tag)::value;
          nA = 0; hA = 0; nB = 0; hB = 0;
          bool closeany = false;
          const unsigned own_n = (unsigned)(own_hi - own_lo);
          for (int s0 = 0; s0 < M; s0 += 32) {
            const int n = s0 + lane;
            const bool valid = n < M;
            const bool starts_here = my_dpref >= s0 && my_dpref < s0 + 32;
            const unsigned starts = __reduce_or_sync(0xffffffffu, starts_here ? (1u << (my_dpref - s0)) : 0u);
            const int before = __popc(__ballot_sync(0xffffffffu, my_dpref < s0));
            const int rr = valid ? before + __popc(starts & ((2u << lane) - 1u)) - 1 : 0;
            const int2 dv = dense[rr];
            const int lc = valid ? dv.x + n : far_slot;
            const int info = dv.y;
            const float cxj = sx[lc], cyj = sy[lc], czj = sz[lc];
            T cutA = B.cut0, cutB = B.cut0;
            float fA = fcut0, fB = fcut0;
            if (!UC) {
              const int tj = (int)st[lc];
              cutA = __ldg(&cutneighsq[tA * ntypes + tj]); fA = (float)cutA;
              cutB = __ldg(&cutneighsq[tB * ntypes + tj]); fB = (float)cutB;
            }
            bool okA, okB;
            {
              const float dx = xA - cxj, dy = yA - cyj, dz = zA - czj;
              if (sizeof(T) == 4) {
                okA = rsq_unfused(dx, dy, dz) <= fA;
              } else {
                const float d = (dx * dx + dy * dy + dz * dz) - fA;
                okA = d < -band;
                const bool close = fabsf(d) <= band && lc != aA;
                if (!EXACT) closeany = closeany || close;
                if (EXACT) {
                  if (__any_sync(0xffffffffu, close)) {
                    if (close) okA = build2_exact_within<T>(x, idA, __ldg(slots + (info >> 2) + lc), cutA);
                  }
                }
              }
            }
            {
              const float dx = xB - cxj, dy = yB - cyj, dz = zB - czj;
              if (sizeof(T) == 4) {
                okB = rsq_unfused(dx, dy, dz) <= fB;
              } else {
                const float d = (dx * dx + dy * dy + dz * dz) - fB;
                okB = d < -band;
                const bool close = fabsf(d) <= band && lc != aB;
                if (!EXACT) closeany = closeany || close;
                if (EXACT) {
                  if (__any_sync(0xffffffffu, close)) {
                    if (close) okB = build2_exact_within<T>(x, idB, __ldg(slots + (info >> 2) + lc), cutB);
                  }
                }
              }
            }
            okA = okA && lc != aA;
            okB = okB && lc != aB;
            bool halfA = true, halfB = true;
            if (MODE == 1) {
              const bool own_bin = (info & 1) && (unsigned)(lc - own_lo) < own_n;
              halfA = halfB = (info & 2) != 0 || ((info & 1) && lc >= own_hi);
              if (__any_sync(0xffffffffu, own_bin && (okA || okB))) {  // within the bin the reference orders by atom id
                if (own_bin && (okA || okB)) {
                  const int id_j = __ldg(slots + slot0_c + lc);
                  halfA = id_j > idA;
                  if (halfA && okA && id_j >= nlocal) halfA = !build2_ghost_below<T>(x, idA, id_j);
                  halfB = id_j > idB;
                  if (halfB && okB && id_j >= nlocal) halfB = !build2_ghost_below<T>(x, idB, id_j);
                }
              }
            }
            if (MODE == 2) {
              halfA = halfB = false;
              if (okA || okB) {
                const int id_j = __ldg(slots + (info >> 2) + lc);
                halfA = id_j > idA;
                halfB = id_j > idB;
              }
            }
            const unsigned mA = __ballot_sync(0xffffffffu, okA);
            const unsigned mB = __ballot_sync(0xffffffffu, okB);
            const unsigned mhA = MODE == 0 ? mA : __ballot_sync(0xffffffffu, okA && halfA);
            const unsigned mhB = MODE == 0 ? mB : __ballot_sync(0xffffffffu, okB && halfB);
            if (okA) {
              const int pos = nA + __popc(mA & lt_mask);
              if (pos < tcap) stg_u16(rows_t + (oA + (unsigned)pos), (unsigned short)(lc | ((MODE != 0 && halfA) ? TILE_HALF_BIT : 0)));
            }
            if (okB) {
              const int pos = nB + __popc(mB & lt_mask);
              if (pos < tcap) stg_u16(rows_t + (oB + (unsigned)pos), (unsigned short)(lc | ((MODE != 0 && halfB) ? TILE_HALF_BIT : 0)));
            }
            nA += __popc(mA); hA += __popc(mhA);
            nB += __popc(mB); hB += __popc(mhB);
          }
          return __any_sync(0xffffffffu, closeany);
        };
        if (sizeof(T) == 4) {
          sweep_all(TileTag<false>());
        } else if (sweep_all(TileTag<false>())) {
          sweep_all(TileTag<true>());  // a candidate sat inside the guard band: redo this pair with exact FP64 tests
        }
        if (lane == 0) {
          numneigh_half[idA] = hA;
          row_atom[qA] = make_int2(idA, nA);
          if (two) {
            numneigh_half[idB] = hB;
            row_atom[qB] = make_int2(idB, nB);
          }
        }
        warp_max_h = max(warp_max_h, max(hA, two ? hB : 0));
        warp_max_f = max(warp_max_f, max(nA, two ? nB : 0));
        if (lane == 0) warp_total += (unsigned long long)(hA + (two ? hB : 0));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    warp_max_h = max(warp_max_h, __shfl_xor_sync(0xffffffffu, warp_max_h, o));
    warp_max_f = max(warp_max_f, __shfl_xor_sync(0xffffffffu, warp_max_f, o));
    warp_total += __shfl_xor_sync(0xffffffffu, warp_total, o);
  }
  if (lane == 0) {
    atomicMax(max_half, warp_max_h);
    atomicMax(max_full, warp_max_f);
    atomicAdd(total_half, warp_total);
  }
}

// ---------------------------------------------------------------------------------------
// Shared-memory image of a tile: run tables + SoA positions of the halo window.
// ---------------------------------------------------------------------------------------
template <class T> struct alignas(2 * sizeof(T)) Vec2 { T x, y; };

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
// PACKXY: x and y of an atom share one 2-lane record in the [sx, sx + 2*hcap) region (one LDS.128 / LDS.64 fetches both)
template <class T, bool TYPES, bool PACKXY = false>
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
      if (PACKXY) {
        Vec2<T> xy;
        xy.x = v.x; xy.y = v.y;
        reinterpret_cast<Vec2<T>*>(S.sx)[off + k] = xy;
      } else {
        S.sx[off + k] = v.x;
        S.sy[off + k] = v.y;
      }
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

// INTEG = 1 fuses the velocity-Verlet updates that follow the force into the kernel's epilogue (F_i is in registers):
// finalIntegrate of this step (ref/integrate.cpp:59-68, with sum m v^2 on thermo steps, ref/thermo.cpp:149-157) and
// initialIntegrate of the next (:46-57), same operation order as final_initial_integrate_kernel.  Other CTAs still
// read the old positions of their halo atoms, so the new positions go to a second buffer (x_out) that becomes the
// atom array after the launch; f is not written.
template <class T> struct VerletParams {
  Vec4<T>* v;
  Vec4<T>* x_out;
  T dt, dtforce, mass;
};

template <class T, int EV, int UNIFORM, int INTEG>
__global__ void __launch_bounds__(TILE_THREADS, 2)
force_lj_tile_kernel(const Vec4<T>* __restrict__ x, Vec4<T>* __restrict__ f, TileGeo g, const int2* __restrict__ tile_runs,
                     const int4* __restrict__ tile_center, const int2* __restrict__ tile_info, const int* __restrict__ slots,
                     const unsigned short* __restrict__ rows, const int2* __restrict__ row_atom, int tcap, int nlocal,
                     LJTileParams<T> P, VerletParams<T> VP, double* __restrict__ ev_out,
                     unsigned long long* __restrict__ prof /* {staging clocks, CTA clocks, CTAs}: -DMMD_KERNEL_PROFILE builds only */) {
  extern __shared__ __align__(16) unsigned char tile_smem_raw[];
  const int t = blockIdx.x;
  const int2 inf = tile_info[t];
  if (inf.y == 0) return;
#ifdef MMD_KERNEL_PROFILE
  const long long clk0 = clock64();
#endif
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
  tile_stage<T, !UNIFORM, true>(S, g, t, inf.x, tile_runs, slots, x);
#ifdef MMD_KERNEL_PROFILE
  const long long clk1 = clock64();
#endif
  const Vec2<T>* __restrict__ sxy = reinterpret_cast<const Vec2<T>*>(S.sx);

  double eng = 0.0, vir = 0.0, ke = 0.0;
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
      // an atom owns its row whatever the row's length (an isolated atom has an empty one and is still stored / integrated)
      const int id = (a < ce.y && ta.x >= 0 && ta.x < nlocal) ? ta.x : -1;
      const int cnt = id >= 0 ? max(ta.y, 0) : 0;
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
      const Vec2<T> xyi = sxy[aa];
      const T xi = xyi.x, yi = xyi.y, zi = S.sz[aa];
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
            const Vec2<T> xyj = sxy[lj[e]];
            dx[e] = xi - xyj.x;
            dy[e] = yi - xyj.y;
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
        if (INTEG) {
          Vec4<T> vi = VP.v[id];
          vi.x += VP.dtforce * fx;
          vi.y += VP.dtforce * fy;
          vi.z += VP.dtforce * fz;
          if (EV) ke += (double)((vi.x * vi.x + vi.y * vi.y + vi.z * vi.z) * VP.mass);
          vi.x += VP.dtforce * fx;
          vi.y += VP.dtforce * fy;
          vi.z += VP.dtforce * fz;
          Vec4<T> xo;
          xo.x = xi + VP.dt * vi.x;
          xo.y = yi + VP.dt * vi.y;
          xo.z = zi + VP.dt * vi.z;
          xo.w = x[id].w;  // the type lane travels with the atom
          VP.v[id] = vi;
          VP.x_out[id] = xo;
        } else {
          Vec4<T> out;
          out.x = fx; out.y = fy; out.z = fz; out.w = (T)0;
          f[id] = out;
        }
      }
    }
  }
  if (EV) {
    if (INTEG) {
      const double v3[3] = {eng * P.e_scale, vir * P.v_scale, ke};
      block_accumulate<3>(v3, ev_out);
    } else {
      const double v2[2] = {eng * P.e_scale, vir * P.v_scale};
      block_accumulate<2>(v2, ev_out);
    }
  }
#ifdef MMD_KERNEL_PROFILE
  if (prof) {
    __syncthreads();
    if (threadIdx.x == 0) {
      atomicAdd(prof + 0, (unsigned long long)(clk1 - clk0));
      atomicAdd(prof + 1, (unsigned long long)(clock64() - clk0));
      atomicAdd(prof + 2, 1ull);
    }
  }
#endif
}

// ---------------------------------------------------------------------------------------
// Export to the reference's row format (Neighbor::neighbors / numneigh): global atom ids, the
// half-list subset when the list was built for half neighbor semantics.  One CTA per tile.
// ---------------------------------------------------------------------------------------
constexpr int TILE_EXPORT_MAX = 512;  // longest row the sorting export handles
__global__ void __launch_bounds__(TILE_THREADS)
tile_rows_export_kernel(TileGeo g, const int2* __restrict__ tile_runs, const int4* __restrict__ tile_center,
                        const int2* __restrict__ tile_info, const int* __restrict__ slots, const int* __restrict__ oslot,
                        const unsigned short* __restrict__ rows, const int2* __restrict__ row_atom, int tcap, int nlocal,
                        int half_only, int* __restrict__ out, int out_stride, int out_cap, int* __restrict__ status) {
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
      if (oslot == nullptr) {  // windows in CSR order: rows already are in the reference's order
        for (int k = 0; k < cnt; k++) {
          const unsigned short e = row[k];
          if (half_only && !(e & TILE_HALF_BIT)) continue;
          const int loc = e & 0x7fff;
          if (loc < s_off[p]) p = 0;
          while (p + 1 < g.nrun && loc >= s_off[p + 1]) p++;
          const int j = slots[s_start[p] + (loc - s_off[p])];
          if (n < out_cap) out[(size_t)id * out_stride + n] = j;
          n++;
        }
        continue;
      }
      // x-sorted windows: the reference's row order is the order of the original CSR positions
      int key[TILE_EXPORT_MAX], val[TILE_EXPORT_MAX];
      for (int k = 0; k < cnt; k++) {
        const unsigned short e = row[k];
        if (half_only && !(e & TILE_HALF_BIT)) continue;
        const int loc = e & 0x7fff;
        if (loc < s_off[p]) p = 0;
        while (p + 1 < g.nrun && loc >= s_off[p + 1]) p++;
        const int sl = s_start[p] + (loc - s_off[p]);
        if (n >= TILE_EXPORT_MAX) { atomicOr(status, 16); break; }
        const int kk = oslot[sl], vv = slots[sl];
        int j = n - 1;
        while (j >= 0 && key[j] > kk) { key[j + 1] = key[j]; val[j + 1] = val[j]; j--; }
        key[j + 1] = kk; val[j + 1] = vv;
        n++;
      }
      for (int k = 0; k < n && k < out_cap; k++) out[(size_t)id * out_stride + k] = val[k];
    }
  }
}

}  // namespace mmd

// Binning, counting sort by cell, and the bin-and-stencil neighbor-list build.
// Replaces Neighbor::binatoms / Neighbor::build / Atom::sort (ref/neighbor.cpp:79-300,
// ref/atom.cpp:355-421).
//
// Device bin layout is CSR (bin_start[mbins+1], bin_atoms[nall]) instead of the reference's
// fixed-width rows with a doubling/retry protocol: one counting pass, one scan, one fill, no
// retry, no padding traffic.  Inside a bin, atom ids are sorted ascending, which is exactly the
// order the reference's single-thread append produces (ref/neighbor.cpp:238-251) -- so sort
// permutations, ghost lists and neighbor rows come out in the reference's order and can be
// compared index by index.
#pragma once
#include "common.cuh"

namespace mmd {

template <class T> struct BinGeo {
  T prd[3];
  T bininv[3];
  int nbin[3];
  int mbin[3];
  int mbinlo[3];
  int mbins;
};

// Neighbor::coord2bin, one axis (ref/neighbor.cpp:278-297).  Three branches, reciprocal
// multiply and C truncation are part of the bit-exact contract.  (c - prd) * bininv cannot be
// contracted into an FMA (it is sub-then-mul), c * bininv is a single multiply.
template <class T> __device__ __forceinline__ int axis_bin(T c, T prd, T bininv, int nbin, int mbinlo) {
  if (c >= prd) return (int)((c - prd) * bininv) + nbin - mbinlo;
  if (c >= (T)0) return (int)(c * bininv) - mbinlo;
  return (int)(c * bininv) - mbinlo - 1;
}
template <class T> __device__ __forceinline__ int coord2bin(const BinGeo<T>& g, T x, T y, T z) {
  const int ix = axis_bin(x, g.prd[0], g.bininv[0], g.nbin[0], g.mbinlo[0]);
  const int iy = axis_bin(y, g.prd[1], g.bininv[1], g.nbin[1], g.mbinlo[1]);
  const int iz = axis_bin(z, g.prd[2], g.bininv[2], g.nbin[2], g.mbinlo[2]);
  return iz * g.mbin[1] * g.mbin[0] + iy * g.mbin[0] + ix + 1;  // the "+1" is the reference's (:299)
}

// squared distance in the reference's evaluation order, every operation rounded separately
// (the reference build has no FMA), so the <= cutneighsq decision is reproduced exactly.
__device__ __forceinline__ double rsq_unfused(double dx, double dy, double dz) {
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
__device__ __forceinline__ float rsq_unfused(float dx, float dy, float dz) {
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---- binning ---------------------------------------------------------------------------
// status[0] |= 1 if an atom falls outside the bin grid (lost atom); it is then left unbinned.
template <class T>
__global__ void bin_count_kernel(const Vec4<T>* __restrict__ x, int n, BinGeo<T> g, int* __restrict__ atom_bin,
                                 int* __restrict__ bincount, int* __restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Vec4<T> p = x[i];
  int b = coord2bin(g, p.x, p.y, p.z);
  if (b < 0 || b >= g.mbins) {
    atomicOr(status, 1);
    b = -1;
  } else {
    atomicAdd(&bincount[b], 1);
  }
  atom_bin[i] = b;
}

__global__ void bin_fill_kernel(const int* __restrict__ atom_bin, int n, const int* __restrict__ bin_start,
                                int* __restrict__ cursor, int* __restrict__ bin_atoms) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = atom_bin[i];
  if (b < 0) return;
  const int slot = atomicAdd(&cursor[b], 1);
  bin_atoms[bin_start[b] + slot] = i;
}

// one thread per bin: insertion sort of its (few) atom ids => deterministic reference order
__global__ void bin_sort_kernel(const int* __restrict__ bin_start, int mbins, int* __restrict__ bin_atoms) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= mbins) return;
  const int s = bin_start[b], e = bin_start[b + 1];
  for (int a = s + 1; a < e; a++) {
    const int key = bin_atoms[a];
    int k = a - 1;
    while (k >= s && bin_atoms[k] > key) {
      bin_atoms[k + 1] = bin_atoms[k];
      k--;
    }
    bin_atoms[k + 1] = key;
  }
}

// reference-layout export of the bins: rows of width apb (Neighbor::bins)
__global__ void bins_to_rows_kernel(const int* __restrict__ bin_start, const int* __restrict__ bin_atoms, int mbins,
                                    int apb, int* __restrict__ rows) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= mbins) return;
  const int s = bin_start[b], e = bin_start[b + 1];
  for (int k = 0; k < apb; k++) rows[(size_t)b * apb + k] = (s + k < e) ? bin_atoms[s + k] : -1;
}
__global__ void bin_counts_from_start_kernel(const int* __restrict__ bin_start, int mbins, int* __restrict__ cnt) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < mbins) cnt[b] = bin_start[b + 1] - bin_start[b];
}

// ---- Atom::sort permutation --------------------------------------------------------------
// new[i] = old[bin_atoms[i]] for x (with type lane) and v.  (ref/atom.cpp:391-406)
template <class T>
__global__ void permute_atoms_kernel(const int* __restrict__ order, int n, const Vec4<T>* __restrict__ x_old,
                                     const Vec4<T>* __restrict__ v_old, Vec4<T>* __restrict__ x_new,
                                     Vec4<T>* __restrict__ v_new) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int src = order[i];
  x_new[i] = x_old[src];
  v_new[i] = v_old[src];
}

// ---- neighbor build ----------------------------------------------------------------------
// One warp per bin.  All local atoms of a bin share one candidate set (the stencil of the bin),
// so the warp loads each candidate once (lane = candidate, coalesced through the CSR bins) and
// tests it against every local atom of the bin (broadcast from shared memory).  Accepted
// candidates are appended with ballot/popc compaction, which keeps the reference's row order:
// stencil order, then bin order (ref/neighbor.cpp:141-183).
//
// MODE 0: full list              (skip j==i in the own bin)
// MODE 1: half list, ghost_newton (own bin: skip j<=i and ghosts lexicographically below i)
// MODE 2: half list, no ghost_newton (every bin: skip j<i; own bin: skip j==i)
constexpr int NB_WARPS = 8;       // warps (= bins) per block
constexpr int NB_CHUNK = 32;      // local atoms of one bin handled per sweep
constexpr int NB_MAXC = 768;      // candidate ids staged per warp and pass (a typical stencil holds 290 / 570)

template <class T, int MODE>
__global__ void __launch_bounds__(NB_WARPS * 32)
neigh_build_kernel(const Vec4<T>* __restrict__ x, int nlocal, const int* __restrict__ bin_start,
                   const int* __restrict__ bin_atoms, int mbins, const int2* __restrict__ runs, int nruns,
                   const T* __restrict__ cutneighsq, int ntypes, int* __restrict__ neighbors,
                   int* __restrict__ numneigh, int stride, int* __restrict__ max_n,
                   unsigned long long* __restrict__ total, int maxneighs) {
  __shared__ Vec4<T> s_xi[NB_WARPS][NB_CHUNK];
  __shared__ int s_id[NB_WARPS][NB_CHUNK];
  // The stencil's candidates arrive as ~13 (half) / ~25 (full) short runs of consecutive bins; testing them run by
  // run leaves most 32-lane sweeps half empty.  Their ids are first packed, in stencil order, into this
  // per-warp buffer (bit 31 marks candidates of the warp's own bin), then swept 32 at a time.
  __shared__ unsigned s_cand[NB_WARPS][NB_MAXC];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.x * NB_WARPS + w;
  if (b >= mbins) return;
  const int s0 = bin_start[b], s1 = bin_start[b + 1];
  if (s1 == s0) return;
  const unsigned lt_mask = (1u << lane) - 1u;

  for (int chunk0 = s0; chunk0 < s1; chunk0 += NB_CHUNK) {
    // stage this sweep's i-atoms; ids ascend inside a bin, so locals (id < nlocal) come first
    const int ci = chunk0 + lane;
    int my_id = (ci < s1) ? bin_atoms[ci] : 0x7fffffff;
    const bool is_local = my_id < nlocal;
    const int nloc = __popc(__ballot_sync(0xffffffffu, is_local));
    if (nloc == 0) break;
    if (is_local) {
      s_xi[w][lane] = x[my_id];
      s_id[w][lane] = my_id;
    }
    __syncwarp();
    int my_n = 0;  // row length of i-atom `lane`

    int r = 0, off = 0;  // next run / offset inside it still to be staged
    while (r < nruns) {
      // ---- pack candidate ids of as many runs as fit ----
      int filled = 0;
      while (r < nruns && filled < NB_MAXC) {
        const int2 run = runs[r];
        int blo = b + run.x, bhi = blo + run.y;
        blo = max(blo, 0);
        bhi = min(bhi, mbins);
        if (bhi <= blo) { r++; off = 0; continue; }
        const int c_begin = bin_start[blo] + off, c_end = bin_start[bhi];
        const int take = min(c_end - c_begin, NB_MAXC - filled);
        for (int k = lane; k < take; k += 32) {
          const int c = c_begin + k;
          const unsigned own = (c >= s0 && c < s1) ? 0x80000000u : 0u;
          s_cand[w][filled + k] = (unsigned)bin_atoms[c] | own;
        }
        filled += take;
        if (c_begin + take == c_end) { r++; off = 0; }
        else off += take;  // buffer full in the middle of a run: continue it in the next pass
      }
      __syncwarp();
      // ---- sweep the packed candidates, 32 per iteration ----
      for (int c0 = 0; c0 < filled; c0 += 32) {
        const int c = c0 + lane;
        const bool valid = c < filled;
        int j = 0;
        bool own_bin = false;
        Vec4<T> xj;
        xj.x = xj.y = xj.z = xj.w = (T)0;
        if (valid) {
          const unsigned e = s_cand[w][c];
          own_bin = (e & 0x80000000u) != 0u;
          j = (int)(e & 0x7fffffffu);
          xj = x[j];
        }
        const int tj = lane_to_type(xj.w);
        for (int t = 0; t < nloc; t++) {
          const Vec4<T> xi = s_xi[w][t];
          const int i = s_id[w][t];
          bool ok = valid;
          if (own_bin) {
            if (MODE == 0) ok = ok && (j != i);
            if (MODE == 2) ok = ok && (j > i);
            if (MODE == 1) {
              ok = ok && (j > i);
              if (j >= nlocal) {
                const bool below = (xj.z < xi.z) || (xj.z == xi.z && xj.y < xi.y) ||
                                   (xj.z == xi.z && xj.y == xi.y && xj.x < xi.x);
                ok = ok && !below;
              }
            }
          } else if (MODE == 2) {
            ok = ok && (j >= i);
          }
          const T rsq = rsq_unfused(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z);
          const int ti = lane_to_type(xi.w);
          ok = ok && (rsq <= __ldg(&cutneighsq[ti * ntypes + tj]));
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          const int base = __shfl_sync(0xffffffffu, my_n, t);
          if (ok) {
            const int pos = base + __popc(m & lt_mask);
            if (pos < maxneighs) neighbors[(size_t)i * stride + pos] = j;
          }
          if (lane == t) my_n += __popc(m);
        }
      }
      __syncwarp();
    }
    if (is_local) numneigh[my_id] = my_n;
    int mx = is_local ? my_n : 0;
    unsigned long long sum = is_local ? (unsigned long long)my_n : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if (lane == 0) {
      atomicMax(max_n, mx);
      atomicAdd(total, sum);
    }
    __syncwarp();
  }
}

}  // namespace mmd

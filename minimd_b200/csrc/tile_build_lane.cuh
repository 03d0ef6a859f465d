// Neighbor build on x-sorted halo windows, ONE LANE PER ATOM.
//
// Why (profiles/r1_tile_kernels.md): neigh_build_tile3_kernel gives one atom to a whole warp (lane = stencil run for the
// interval search, lane = candidate in the sweeps) and pays ~900 warp instructions per atom for interval compaction,
// prefix sums and ballot/popc row assembly -- 1.83 G warp instructions per rebuild, issue-bound.  Row order is free (the
// export restores the reference's order from the original CSR positions, the force kernel reads bank-dealt rows), so
// nothing forces a warp-wide compaction: here every lane owns one atom of a centre pencil and walks the stencil runs
// itself.  Consecutive lanes are consecutive atoms of an x-sorted pencil, so in every run their candidate intervals
// [x_i - x_r, x_i + x_r] are nearly the same index range: the shared-memory reads of a warp fall on neighbouring banks.
//
// Per lane and stencil run: slab distance in y/z -> half-width x_r of the candidate interval, ONE binary search for its
// first index, then a linear walk that ends at the first candidate beyond x_i + x_r.  Accepted candidates go through a
// 16-byte ring in shared memory and leave with one 128-bit store per 8 entries.
// Distance test, guard band with exact FP64 re-test (every rsq <= cutneighsq decision is the reference's,
// ref/neighbor.cpp:165,179), half-list flag (ref/neighbor.cpp:154-171), counters, row format and status bits are those of
// neigh_build_tile3_kernel (tile_kernels.cuh); rows come out in the same order.
#pragma once
#include "tile_kernels.cuh"

namespace mmd {

constexpr int TBL_THREADS = 256, TBL_WARPS = TBL_THREADS / 32;

template <class T> __host__ __device__ inline size_t buildl_smem_bytes(const TileGeo& g, int hcap, bool with_types) {
  return (size_t)hcap * 3 * sizeof(float) + (with_types ? (size_t)hcap : 0) + (size_t)g.nrun * (TBX + 2 * g.sx + 1) * sizeof(int) +
         (size_t)TBL_WARPS * 32 * (sizeof(int4) + sizeof(float2)) + (size_t)TBL_THREADS * sizeof(uint4) +
         (2 * TILE_MAXRUN + 1) * sizeof(int) + 96;
}

template <class T, int MODE, int UC>
__global__ void __launch_bounds__(TBL_THREADS, 4)
neigh_build_lane_kernel(const Vec4<T>* __restrict__ x, int nlocal, const int* __restrict__ bin_start,
                        const int* __restrict__ slots, int mbins, const StencilRun* __restrict__ sruns, int nsr,
                        const T* __restrict__ cutneighsq, int ntypes, TileGeo g, Build2Params<T> B,
                        const int2* __restrict__ tile_runs, const int4* __restrict__ tile_center,
                        const int2* __restrict__ tile_info, unsigned short* __restrict__ rows, int tcap,
                        int* __restrict__ numneigh_half, int2* __restrict__ row_atom, int* __restrict__ status,
                        int* __restrict__ max_half, int* __restrict__ max_full, unsigned long long* __restrict__ total_half) {
  extern __shared__ __align__(16) unsigned char bl_smem[];
  const int t = blockIdx.x;
  const int2 inf = tile_info[t];
  if (inf.y == 0) return;
  const int H = inf.x;
  const int WX1 = TBX + 2 * g.sx + 1;
  float* sx = reinterpret_cast<float*>(bl_smem);
  float* sy = sx + g.hcap;
  float* sz = sy + g.hcap;
  // per warp and stencil run of the current centre pencil: {row of the run's pencil in s_binoff, CSR slot of tile-local
  // index 0 of that pencil, flags (1 own pencil, 2 upper half of the stencil, 4 pencil inside the grid), dxlo | len << 16}
  int4* s_rtab = reinterpret_cast<int4*>(sz + g.hcap);                 // hcap is a multiple of 64: 16-byte aligned
  uint4* s_ring = reinterpret_cast<uint4*>(s_rtab + TBL_WARPS * 32);    // [threads] 8 pending row entries per lane
  float2* s_ryz = reinterpret_cast<float2*>(s_ring + TBL_THREADS);      // [warps][32] low y/z face of the run's pencil
  int* s_binoff = reinterpret_cast<int*>(s_ryz + TBL_WARPS * 32);       // [nrun][WX1] tile-local start index of each window bin
  int* s_run_start = s_binoff + g.nrun * WX1;                           // [nrun]
  int* s_run_off = s_run_start + TILE_MAXRUN;                           // [nrun+1]
  unsigned char* st = reinterpret_cast<unsigned char*>(s_run_off + TILE_MAXRUN + 1);  // [hcap] types (!UC)

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int tx = t % g.ntx, ty = (t / g.ntx) % g.nty, tz = t / (g.ntx * g.nty);
  const int bx0 = tx * TBX - g.ox, by0 = ty * TBY - g.oy, bz0 = tz * TBZ - g.oz;
  const int xlo = max(0, bx0 - g.sx), xhi = min(g.mbx, bx0 + TBX + g.sx);
  // window origin (FP64 build: keeps the FP32 images small; any point near the tile would do)
  const T org_x = sizeof(T) == 8 ? (T)((bx0 - g.sx + B.mbinlo[0]) * B.binsize[0]) : (T)0;
  const T org_y = sizeof(T) == 8 ? (T)((by0 - g.sy + B.mbinlo[1]) * B.binsize[1]) : (T)0;
  const T org_z = sizeof(T) == 8 ? (T)((bz0 - g.sz + B.mbinlo[2]) * B.binsize[2]) : (T)0;

  // ---- stage the window: run tables, bin offsets, FP32 images (+ types) ----
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
  __syncthreads();

  int4* rtab = s_rtab + w * 32;
  float2* ryz = s_ryz + w * 32;
  unsigned short* ring = reinterpret_cast<unsigned short*>(s_ring + threadIdx.x);
  const float bsy = (float)B.binsize[1], bsz = (float)B.binsize[2];
  const float ey0 = (float)((by0 - g.sy + B.mbinlo[1]) * B.binsize[1] - (double)org_y);
  const float ez0 = (float)((bz0 - g.sz + B.mbinlo[2]) * B.binsize[2] - (double)org_z);
  const float rc = B.rcull, rc2 = rc * rc, fcut0 = (float)B.cut0, band = B.band;
  int warp_max_h = 0, warp_max_f = 0;
  unsigned long long lane_total = 0ull;

  for (int cr = w; cr < TILE_NCENTER; cr += nw) {
    const int cy = cr % TBY, cz = cr / TBY;
    const int by = by0 + cy, bz = bz0 + cz;
    if (by < 0 || by >= g.mby || bz < 0 || bz >= g.mbz) continue;
    const int4 ce = tile_center[(size_t)t * TILE_NCENTER + cr];
    if (ce.y <= ce.x) continue;
    const int pc = (cy + g.sy) + (cz + g.sz) * g.nry;
    const int slot0_c = s_run_start[pc] - s_run_off[pc];
    // ---- the stencil runs as seen from this centre pencil (lane r prepares run r) ----
    __syncwarp();
    if (lane < nsr) {
      const StencilRun run = sruns[lane];
      const int dxlo = run.off - (run.dz * g.mby + run.dy) * g.mbx;
      const int yy = by + run.dy, zz = bz + run.dz;
      int flags = 0, row = 0, slot0 = 0;
      float ylo = 0.f, zlo = 0.f;
      if (yy >= 0 && yy < g.mby && zz >= 0 && zz < g.mbz) {
        const int p = (cy + g.sy + run.dy) + (cz + g.sz + run.dz) * g.nry;
        row = p * WX1;
        slot0 = s_run_start[p] - s_run_off[p];
        ylo = ey0 + (float)(cy + g.sy + run.dy) * bsy;
        zlo = ez0 + (float)(cz + g.sz + run.dz) * bsz;
        const bool upper = run.dz > 0 || (run.dz == 0 && run.dy > 0);
        const bool ownp = run.dz == 0 && run.dy == 0;
        flags = 4 | (upper ? 2 : 0) | (ownp ? 1 : 0);
      }
      rtab[lane] = make_int4(row, slot0, flags, (dxlo & 0xffff) | (run.len << 16));
      ryz[lane] = make_float2(ylo, zlo);
    }
    __syncwarp();
    const int cx_lo = max(0, bx0), cx_hi = min(g.mbx, bx0 + TBX);  // own bins of the pencil: [cx_lo, cx_hi)

    for (int a0 = ce.x; a0 < ce.y; a0 += 32) {
      const int a = a0 + lane;
      const bool have = a < ce.y;
      const int id_i = have ? __ldg(slots + slot0_c + a) : 0x7fffffff;
      bool live = id_i < nlocal;  // ghosts sit in bins too; they get no row
      // own bin of the atom (its x index inside the window) and that bin's index range
      int xw = cx_lo - xlo, own_lo = 0, own_hi = 0;
      {
        const int* bo = s_binoff + pc * WX1;
        for (int bx = cx_lo; bx < cx_hi; bx++) {
          const int lo = bo[bx - xlo], hi = bo[bx - xlo + 1];
          if (have && a >= lo && a < hi) { xw = bx - xlo; own_lo = lo; own_hi = hi; }
        }
      }
      const int aa = have ? a : ce.x;
      const float xi = sx[aa], yi = sy[aa], zi = sz[aa];
      const int ti = UC ? 0 : (int)st[aa];
      const int q = ce.z + (aa - ce.x);
      unsigned short* __restrict__ rowp = rows + (size_t)q * tcap;
      int n_t = 0, h_t = 0;
      bool bad = false;

      for (int r = 0; r < nsr; r++) {
        const int4 rt = rtab[r];
        const float2 yz = ryz[r];
        if (!(rt.z & 4)) { bad = bad || live; continue; }   // the run's pencil lies outside the bin grid
        const int dxlo = (int)(short)(rt.w & 0xffff), len = rt.w >> 16;
        const int wlo = xw + dxlo, whi = wlo + len;
        const bool out = wlo < 0 || xlo + whi > xhi;       // the run's bins leave the window along x
        bad = bad || (live && out);
        if (!live || out) continue;
        const float gy = fmaxf(0.0f, fmaxf(yz.x - yi, yi - (yz.x + bsy)));
        const float gz = fmaxf(0.0f, fmaxf(yz.y - zi, zi - (yz.y + bsz)));
        const float rem = rc2 - gy * gy - gz * gz;
        if (rem <= 0.0f) continue;
        const float xr = sqrtf(rem);
        const float xa = xi - xr, xb = xi + xr;
        const int hi_i = s_binoff[rt.x + whi];
        int k = s_binoff[rt.x + wlo];
        {  // first index with sx >= xa
          int l1 = hi_i;
          while (k < l1) {
            const int mid = (k + l1) >> 1;
            if (sx[mid] < xa) k = mid + 1; else l1 = mid;
          }
        }
        const bool ownp = (rt.z & 1) != 0, upper = (rt.z & 2) != 0;
        for (; k < hi_i; k++) {
          const float cxk = sx[k];
          if (cxk > xb) break;
          const float dx = xi - cxk, dy = yi - sy[k], dz = zi - sz[k];
          T cut = B.cut0;
          float fc = fcut0;
          if (!UC) { cut = __ldg(&cutneighsq[ti * ntypes + (int)st[k]]); fc = (float)cut; }
          bool ok;
          if (sizeof(T) == 4) {
            ok = rsq_unfused(dx, dy, dz) <= fc;
          } else {
            const float d = (dx * dx + dy * dy + dz * dz) - fc;
            ok = d < -band;
            // inside the guard band the FP32 images cannot decide: the reference's own FP64 arithmetic does
            if (fabsf(d) <= band && !(ownp && k == a)) ok = build2_exact_within<T>(x, id_i, __ldg(slots + rt.y + k), cut);
          }
          ok = ok && !(ownp && k == a);
          if (!ok) continue;
          bool half = true;
          if (MODE == 1) {
            half = upper || (ownp && k >= own_hi);
            if (ownp && k >= own_lo && k < own_hi) {  // within the bin the reference orders by atom id
              const int id_j = __ldg(slots + slot0_c + k);
              half = id_j > id_i;
              if (half && id_j >= nlocal) half = !build2_ghost_below<T>(x, id_i, id_j);
            }
          }
          if (MODE == 2) half = __ldg(slots + rt.y + k) > id_i;
          if (n_t < tcap) ring[n_t & 7] = (unsigned short)(k | ((MODE != 0 && half) ? TILE_HALF_BIT : 0));
          n_t++;
          h_t += half ? 1 : 0;
          if ((n_t & 7) == 0 && n_t <= tcap) *reinterpret_cast<uint4*>(rowp + n_t - 8) = s_ring[threadIdx.x];
        }
      }
      if (live && bad) { atomicOr(status, 2); live = false; }
      if (live) {
        if ((n_t & 7) != 0 && n_t < tcap) *reinterpret_cast<uint4*>(rowp + (n_t & ~7)) = s_ring[threadIdx.x];
        numneigh_half[id_i] = h_t;
        row_atom[q] = make_int2(id_i, n_t);
        warp_max_h = max(warp_max_h, h_t);
        warp_max_f = max(warp_max_f, n_t);
        lane_total += (unsigned long long)h_t;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    warp_max_h = max(warp_max_h, __shfl_xor_sync(0xffffffffu, warp_max_h, o));
    warp_max_f = max(warp_max_f, __shfl_xor_sync(0xffffffffu, warp_max_f, o));
    lane_total += __shfl_xor_sync(0xffffffffu, lane_total, o);
  }
  if (lane == 0) {
    atomicMax(max_half, warp_max_h);
    atomicMax(max_full, warp_max_f);
    atomicAdd(total_half, lane_total);
  }
}

}  // namespace mmd

// Lennard-Jones pair force over per-atom neighbor rows.
// Replaces ForceLJ::compute_halfneigh<EV,GN> (ref/force_lj.cpp:185-263; threaded variant
// :271-357) and ForceLJ::compute_fullneigh<EV> (:366-449).
//
// Work decomposition: a group of TPA lanes (TPA = 1..32, a power of two) owns one atom i.
// Lane s of the group reads neighbor ids k = s, s+TPA, ... of the row (consecutive lanes read
// consecutive ints: coalesced), gathers x_j with ONE 256-bit load (FP64) / 128-bit load (FP32)
// that also carries type_j, and accumulates its partial force; partials are combined with
// shuffles inside the group.  Half lists scatter -F to f_j with fire-and-forget REDG
// (FP32: one vector REDG.F32x4; FP64: three REDG.F64) and add the reduced F_i with one more.
// Energy / virial (EV) are reduced warp -> block -> one REDG per block in FP64 regardless of T.
#pragma once
#include "common.cuh"

namespace mmd {

template <class T> struct LJParams {
  // uniform fast path (all type pairs equal, which is what ref/ljs.cpp:299-305 + ForceLJ::setup
  // always produce); per-type tables are used otherwise
  T cutforcesq, sigma6, epsilon;
  const T* cutforcesq_tab;
  const T* sigma6_tab;
  const T* epsilon_tab;
  int ntypes;
};

constexpr int LJ_BLOCK = 256;

// HALF: 1 = half list (Newton's 3rd law scatter), 0 = full list.
// GN (half only): ghost_newton -- scatter to ghosts too; otherwise ghost pairs count half in EV.
// EV: accumulate energy/virial into ev_out[0..1].
// UNIFORM: all type pairs share one parameter set.
template <class T, int TPA, int HALF, int GN, int EV, int UNIFORM>
__global__ void __launch_bounds__(LJ_BLOCK)
force_lj_kernel(const Vec4<T>* __restrict__ x, Vec4<T>* __restrict__ f, const int* __restrict__ neighbors,
                const int* __restrict__ numneigh, int stride, int nlocal, LJParams<T> P, double* __restrict__ ev_out) {
  const int gtid = blockIdx.x * LJ_BLOCK + threadIdx.x;
  const int i = gtid / TPA;
  const int sub = threadIdx.x % TPA;
  const int lane = threadIdx.x & 31;
  const bool active = i < nlocal;
  // staging area of the warp-cooperative scatter (half lists only; unused and not allocated otherwise)
  __shared__ T s_val[HALF ? LJ_BLOCK / 32 : 1][HALF ? 96 : 1];
  __shared__ int s_idx[HALF ? LJ_BLOCK / 32 : 1][HALF ? 32 : 1];
  T* sv = s_val[HALF ? (threadIdx.x >> 5) : 0];
  int* sj = s_idx[HALF ? (threadIdx.x >> 5) : 0];

  T fx = 0, fy = 0, fz = 0;
  double eng = 0.0, vir = 0.0;

  Vec4<T> xi;
  xi.x = xi.y = xi.z = xi.w = (T)0;
  int cnt = 0;
  const int* __restrict__ row = neighbors + (size_t)(active ? i : 0) * stride;
  if (active) {
    xi = x[i];
    cnt = numneigh[i];
  }
  const int ti = lane_to_type(xi.w);
  // half lists: every lane of the warp runs the same number of iterations (the scatter is warp-cooperative)
  const int kend = HALF ? __reduce_max_sync(0xffffffffu, cnt) : cnt;
  for (int k0 = 0; k0 < kend; k0 += TPA) {
    const int k = k0 + sub;
    bool scatter = false;
    int j = 0;
    T sx = 0, sy = 0, sz = 0;
    if (k < cnt) {
      j = __ldg(row + k);
      const Vec4<T> xj = ldg4(x + j);
      const T dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
      const T rsq = dx * dx + dy * dy + dz * dz;
      T cut, s6, eps;
      if (UNIFORM) {
        cut = P.cutforcesq; s6 = P.sigma6; eps = P.epsilon;
      } else {
        const int tij = ti * P.ntypes + lane_to_type(xj.w);
        cut = __ldg(P.cutforcesq_tab + tij); s6 = __ldg(P.sigma6_tab + tij); eps = __ldg(P.epsilon_tab + tij);
      }
      if (rsq < cut) {
        const T sr2 = (T)1 / rsq;
        const T sr6 = sr2 * sr2 * sr2 * s6;
        const T force = (T)48 * sr6 * (sr6 - (T)0.5) * sr2 * eps;
        fx += dx * force;
        fy += dy * force;
        fz += dz * force;
        if (HALF) {
          scatter = GN || (j < nlocal);
          sx = -dx * force; sy = -dy * force; sz = -dz * force;
          if (EV) {
            const double scale = scatter ? 1.0 : 0.5;
            eng += scale * (double)((T)4 * sr6 * (sr6 - (T)1) * eps);
            vir += scale * (double)(rsq * force);
          }
        } else if (EV) {
          eng += (double)(sr6 * (sr6 - (T)1) * eps);
          vir += (double)(rsq * force);
        }
      }
    }
    if (HALF) warp_scatter3(f, sv, sj, lane, scatter, j, sx, sy, sz);
  }
  if (TPA > 1) {
    fx = group_sum<TPA>(fx);
    fy = group_sum<TPA>(fy);
    fz = group_sum<TPA>(fz);
  }
  if (HALF) {
    warp_scatter3(f, sv, sj, lane, active && sub == 0, i, fx, fy, fz);
  } else if (active && sub == 0) {
    Vec4<T> out;
    out.x = fx; out.y = fy; out.z = fz; out.w = (T)0;
    f[i] = out;
  }
  if (EV) {
    if (!HALF) { eng *= 4.0; vir *= 0.5; }  // ref/force_lj.cpp:441-442
    const double v2[2] = {eng, vir};
    block_accumulate<2>(v2, ev_out);
  }
}

}  // namespace mmd

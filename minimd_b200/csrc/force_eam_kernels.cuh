// Embedded-atom (Cu funcfl) force: density pass, embedding derivative, fp halo, pair pass.
// Replaces ForceEAM::compute_fullneigh (ref/force_eam.cpp:274-449) and
// ForceEAM::compute_halfneigh (ref/force_eam.cpp:94-270).
//
// Spline tables are re-packed at upload from the reference's 7 coefficients per knot
// (ref/force_eam.cpp:765-793) into aligned 4-lane records, one vector load each:
//   *_val[knot] = (c3,c4,c5,c6)  value cubic      ((c3*p + c4)*p + c5)*p + c6
//   *_der[knot] = (c0,c1,c2, 0)  derivative quad.  (c0*p + c1)*p + c2
// Same TPA-lanes-per-atom decomposition as the LJ kernel.
#pragma once
#include "common.cuh"

namespace mmd {

template <class T> struct EAMTables {
  const Vec4<T>* rho_val;   // [ntypes^2][nr+1]
  const Vec4<T>* rho_der;
  const Vec4<T>* z2_val;
  const Vec4<T>* z2_der;
  const Vec4<T>* frho_val;  // [ntypes^2][nrho+1]
  const Vec4<T>* frho_der;
  const T* cutforcesq_tab;  // [ntypes^2]
  T cutforcesq;             // uniform value
  T rdr, rdrho;
  int nr, nrho, ntypes;
};

constexpr int EAM_BLOCK = 256;

template <class T> __device__ __forceinline__ T cubic(const Vec4<T>& c, T p) { return ((c.x * p + c.y) * p + c.z) * p + c.w; }
template <class T> __device__ __forceinline__ T quad(const Vec4<T>& c, T p) { return (c.x * p + c.y) * p + c.z; }
__device__ __forceinline__ double real_sqrt(double v) { return sqrt(v); }
__device__ __forceinline__ float real_sqrt(float v) { return sqrtf(v); }

// knot lookup of the pair/density tables: p = r*rdr + 1; m = min(int(p), nr-1); p = min(p-m, 1)
// (ref/force_eam.cpp:152-156)
template <class T> __device__ __forceinline__ void knot_r(T r, T rdr, int nr, int& m, T& p) {
  p = r * rdr + (T)1;
  m = (int)p;
  m = m < nr - 1 ? m : nr - 1;
  p -= (T)m;
  p = p < (T)1 ? p : (T)1;
}

// embedding: fp = F'(rho), optional energy F(rho) (ref/force_eam.cpp:336-347; type_ii = type*type quirk)
template <class T, int UNIFORM>
__device__ __forceinline__ T embed(const EAMTables<T>& E, int type_i, T rho, bool want_e, double& e_out) {
  T p = rho * E.rdrho + (T)1;
  int m = (int)p;
  m = max(1, min(m, E.nrho - 1));
  p -= (T)m;
  p = p < (T)1 ? p : (T)1;
  const size_t row = (UNIFORM ? (size_t)0 : (size_t)(type_i * type_i) * (E.nrho + 1)) + m;
  if (want_e) e_out += (double)cubic(ldg4(E.frho_val + row), p);
  return quad(ldg4(E.frho_der + row), p);
}

// ---- pass 1 -------------------------------------------------------------------------------
// FULL=1: rho_i over the full row, then fp[i] (+ embedding energy) in the same kernel.
// FULL=0 (half list): rho[i] += partial, rho[j] += same for local j (REDG); embed runs later.
template <class T, int TPA, int FULL, int EV, int UNIFORM>
__global__ void __launch_bounds__(EAM_BLOCK)
eam_rho_kernel(const Vec4<T>* __restrict__ x, const int* __restrict__ neighbors, const int* __restrict__ numneigh,
               int stride, int nlocal, EAMTables<T> E, T* __restrict__ rho, T* __restrict__ fp,
               double* __restrict__ ev_out) {
  const int i = (blockIdx.x * EAM_BLOCK + threadIdx.x) / TPA;
  const int sub = threadIdx.x % TPA;
  const bool active = i < nlocal;
  T rhoi = 0;
  int ti = 0;
  if (active) {
    const Vec4<T> xi = x[i];
    ti = lane_to_type(xi.w);
    const int* __restrict__ row = neighbors + (size_t)i * stride;
    const int cnt = numneigh[i];
    for (int k = sub; k < cnt; k += TPA) {
      const int j = __ldg(row + k);
      const Vec4<T> xj = ldg4(x + j);
      const T dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
      const T rsq = dx * dx + dy * dy + dz * dz;
      const int tij = UNIFORM ? 0 : ti * E.ntypes + lane_to_type(xj.w);
      const T cut = UNIFORM ? E.cutforcesq : __ldg(E.cutforcesq_tab + tij);
      if (rsq < cut) {
        int m; T p;
        knot_r(real_sqrt(rsq), E.rdr, E.nr, m, p);
        const T val = cubic(ldg4(E.rho_val + (size_t)tij * (E.nr + 1) + m), p);
        rhoi += val;
        if (!FULL && j < nlocal) red_add1(rho + j, val);
      }
    }
  }
  if (TPA > 1) rhoi = group_sum<TPA>(rhoi);
  double e = 0.0;
  if (active && sub == 0) {
    if (FULL) fp[i] = embed<T, UNIFORM>(E, ti, rhoi, EV != 0, e);
    else red_add1(rho + i, rhoi);
  }
  if (FULL && EV) {
    const double a[1] = {e};
    block_accumulate<1>(a, ev_out);
  }
}

// half list: fp[i] = F'(rho[i]) once rho is complete (ref/force_eam.cpp:172-185)
template <class T, int EV, int UNIFORM>
__global__ void eam_embed_kernel(const Vec4<T>* __restrict__ x, int nlocal, EAMTables<T> E, const T* __restrict__ rho,
                                 T* __restrict__ fp, double* __restrict__ ev_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (i < nlocal) fp[i] = embed<T, UNIFORM>(E, lane_to_type(x[i].w), rho[i], EV != 0, e);
  if (EV) {
    const double a[1] = {e};
    block_accumulate<1>(a, ev_out);
  }
}

// ---- pass 2 -------------------------------------------------------------------------------
// pair forces (ref/force_eam.cpp:368-441 full, :194-267 half).  ev_out[0] += energy term,
// ev_out[1] += virial.  Energy convention (the caller finishes it):
//   full: ev_out[0] accumulates sum 0.5*phi  (then eng_vdwl = 2*(embed + that), :446)
//   half: ev_out[0] accumulates phi (local j) or 0.5*phi (ghost j) (then eng_vdwl = embed + that, :269)
template <class T, int TPA, int HALF, int EV, int UNIFORM>
__global__ void __launch_bounds__(EAM_BLOCK)
eam_pair_kernel(const Vec4<T>* __restrict__ x, Vec4<T>* __restrict__ f, const int* __restrict__ neighbors,
                const int* __restrict__ numneigh, int stride, int nlocal, EAMTables<T> E, const T* __restrict__ fp,
                double* __restrict__ ev_out) {
  const int i = (blockIdx.x * EAM_BLOCK + threadIdx.x) / TPA;
  const int sub = threadIdx.x % TPA;
  const int lane = threadIdx.x & 31;
  const bool active = i < nlocal;
  __shared__ T s_val[HALF ? EAM_BLOCK / 32 : 1][HALF ? 96 : 1];  // warp-cooperative scatter staging (common.cuh)
  __shared__ int s_idx[HALF ? EAM_BLOCK / 32 : 1][HALF ? 32 : 1];
  T* sv = s_val[HALF ? (threadIdx.x >> 5) : 0];
  int* sj = s_idx[HALF ? (threadIdx.x >> 5) : 0];
  T fx = 0, fy = 0, fz = 0;
  double eng = 0.0, vir = 0.0;
  Vec4<T> xi;
  xi.x = xi.y = xi.z = xi.w = (T)0;
  int cnt = 0;
  T fpi = 0;
  const int* __restrict__ row = neighbors + (size_t)(active ? i : 0) * stride;
  if (active) {
    xi = x[i];
    fpi = fp[i];
    cnt = numneigh[i];
  }
  const int ti = lane_to_type(xi.w);
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
      const int tij = UNIFORM ? 0 : ti * E.ntypes + lane_to_type(xj.w);
      const T cut = UNIFORM ? E.cutforcesq : __ldg(E.cutforcesq_tab + tij);
      if (rsq < cut) {
        const T r = real_sqrt(rsq);
        int m; T p;
        knot_r(r, E.rdr, E.nr, m, p);
        const size_t row_t = (size_t)tij * (E.nr + 1) + m;
        const T rhoip = quad(ldg4(E.rho_der + row_t), p);
        const T z2p = quad(ldg4(E.z2_der + row_t), p);
        const T z2 = cubic(ldg4(E.z2_val + row_t), p);
        const T recip = (T)1 / r;
        const T phi = z2 * recip;
        const T phip = z2p * recip - phi * recip;
        const T psip = fpi * rhoip + __ldg(fp + j) * rhoip + phip;
        T fpair = -psip * recip;
        fx += dx * fpair;
        fy += dy * fpair;
        fz += dz * fpair;
        if (HALF) {
          const bool local_j = j < nlocal;
          if (local_j) {
            scatter = true;
            sx = -dx * fpair; sy = -dy * fpair; sz = -dz * fpair;
          } else {
            fpair *= (T)0.5;
          }
          if (EV) {
            vir += (double)(rsq * fpair);
            eng += local_j ? (double)phi : 0.5 * (double)phi;
          }
        } else if (EV) {
          vir += (double)(rsq * ((T)0.5 * fpair));
          eng += 0.5 * (double)phi;
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
    const double v2[2] = {eng, vir};
    block_accumulate<2>(v2, ev_out);
  }
}

}  // namespace mmd

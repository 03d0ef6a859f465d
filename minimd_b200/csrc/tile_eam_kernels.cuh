// Embedded-atom force over tile-resident neighbor rows (tile_kernels.cuh): the two EAM passes as owner-computes
// shared-memory kernels.  Replaces ForceEAM::compute_fullneigh (ref/force_eam.cpp:274-449) and
// ForceEAM::compute_halfneigh (:94-270) for either list style: every atom visits its complete neighborhood, so
//   pass 1  rho_i = sum_j rho(r_ij) is complete in one lane group -> fp_i = F'(rho_i) and the embedding energy F(rho_i)
//           are produced by the same kernel (no rho array, no scatter to rho_j, no separate embed kernel);
//   halo    fp of the ghosts (ForceEAM::communicate, :851-914) -- unchanged, between the two launches;
//   pass 2  pair forces from positions AND fp staged in shared memory; F_i written once.
// Energy / virial: a pair met from both ends contributes 0.5*phi and 0.5*r^2*fpair each time, which is the
// full-list convention (:431-446) and sums to the half-list one (:246-260: phi for a local j, 0.5*phi for a ghost j
// seen from both owners); the caller applies eng_vdwl = embed + S (half) or 2*(embed + S) (full).
// Spline tables stay in global memory (L1-resident, 16 KB each; records as in force_eam_kernels.cuh).
#pragma once
#include "force_eam_kernels.cuh"
#include "tile_kernels.cuh"

namespace mmd {

constexpr int EAMT_TPA = 2;

// PASS 1: rho + embed      PASS 2: pair forces
template <class T, int PASS, int EV, int UNIFORM>
__global__ void __launch_bounds__(TILE_THREADS, PASS == 1 ? 2 : 1)
eam_tile_kernel(const Vec4<T>* __restrict__ x, Vec4<T>* __restrict__ f, TileGeo g, const int2* __restrict__ tile_runs,
                const int4* __restrict__ tile_center, const int2* __restrict__ tile_info, const int* __restrict__ slots,
                const unsigned short* __restrict__ rows, const int2* __restrict__ row_atom, int tcap, int nlocal,
                EAMTables<T> E, T* __restrict__ fp, double* __restrict__ ev_out /* [0] 0.5 phi, [1] virial, [3] embed */) {
  extern __shared__ __align__(16) unsigned char tile_smem_raw[];
  const int t = blockIdx.x;
  const int2 inf = tile_info[t];
  if (inf.y == 0) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int sub = lane & (EAMT_TPA - 1);
  constexpr int APP = 32 / EAMT_TPA;

  int cr = w;
  int4 ce = make_int4(0, 0, 0, 0);
  if (cr < TILE_NCENTER) ce = tile_center[(size_t)t * TILE_NCENTER + cr];
  int2 ta_n = make_int2(-1, 0);
  uint4 c_n = make_uint4(0, 0, 0, 0);
  {
    const int a = ce.x + lane / EAMT_TPA;
    if (a < ce.y) {
      const size_t q = (size_t)(ce.z + (a - ce.x));
      ta_n = __ldg(row_atom + q);
      c_n = ldg_row8(rows + q * tcap + sub * 8);
    }
  }

  TileSmem<T> S;
  S.carve(tile_smem_raw, g.hcap, true);
  T* sfp = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(S.st + g.hcap) + 15) & ~(uintptr_t)15);  // pass 2: fp of the halo window
  tile_stage<T, true>(S, g, t, inf.x, tile_runs, slots, x);
  if (PASS == 2) {
    for (int p = w; p < g.nrun; p += nw) {
      const int start = S.run_start[p], off = S.run_off[p], len = S.run_off[p + 1] - off;
      for (int k = lane; k < len; k += 32) sfp[off + k] = __ldg(fp + __ldg(slots + start + k));
    }
    __syncthreads();
  }

  double eng = 0.0, vir = 0.0, emb = 0.0;
  for (; cr < TILE_NCENTER; cr += nw) {
    if (cr != w) {
      ce = tile_center[(size_t)t * TILE_NCENTER + cr];
      ta_n = make_int2(-1, 0);
      const int a = ce.x + lane / EAMT_TPA;
      if (a < ce.y) {
        const size_t q = (size_t)(ce.z + (a - ce.x));
        ta_n = __ldg(row_atom + q);
        c_n = ldg_row8(rows + q * tcap + sub * 8);
      }
    }
    for (int a0 = ce.x; a0 < ce.y; a0 += APP) {
      const int a = a0 + lane / EAMT_TPA;
      const int2 ta = ta_n;
      uint4 nxt = c_n;
      // an atom owns its row whatever the row's length (an isolated atom has an empty one and is still stored / integrated)
      const int id = (a < ce.y && ta.x >= 0 && ta.x < nlocal) ? ta.x : -1;
      const int cnt = id >= 0 ? max(ta.y, 0) : 0;
      const int aa = id >= 0 ? a : ce.x;
      const unsigned short* __restrict__ row = rows + (size_t)(ce.z + (aa - ce.x)) * tcap;
      {
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
      const T fpi = PASS == 2 ? sfp[aa] : (T)0;
      T ax = 0, ay = 0, az = 0;  // pass 1: ax = rho_i; pass 2: force
      const int nch = (cnt + 7) >> 3;
      const int mine = nch > sub ? (nch - sub + EAMT_TPA - 1) / EAMT_TPA : 0;
      const int iters = __reduce_max_sync(0xffffffffu, mine);
      int k0 = sub * 8;
      for (int it = 0; it < iters; it++, k0 += 8 * EAMT_TPA) {
        const uint4 pk = nxt;
        if (k0 + 8 * EAMT_TPA < cnt) nxt = ldg_row8(row + k0 + 8 * EAMT_TPA);
        const unsigned wds[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
        for (int e = 0; e < 8; e++) {
          if (k0 + e < cnt) {
            const int lj = (int)((wds[e >> 1] >> ((e & 1) * 16)) & 0x7fff);
            const T dx = xi - S.sx[lj], dy = yi - S.sy[lj], dz = zi - S.sz[lj];
            const T rsq = dx * dx + dy * dy + dz * dz;
            const int tij = UNIFORM ? 0 : ti * E.ntypes + (int)S.st[lj];
            const T cut = UNIFORM ? E.cutforcesq : __ldg(E.cutforcesq_tab + tij);
            if (rsq < cut) {
              const T r = real_sqrt(rsq);
              int m; T p;
              knot_r(r, E.rdr, E.nr, m, p);
              const size_t row_t = (size_t)tij * (E.nr + 1) + m;
              if (PASS == 1) {
                ax += cubic(ldg4(E.rho_val + row_t), p);
              } else {
                const T rhoip = quad(ldg4(E.rho_der + row_t), p);
                const T z2p = quad(ldg4(E.z2_der + row_t), p);
                const T z2 = cubic(ldg4(E.z2_val + row_t), p);
                const T recip = (T)1 / r;
                const T phi = z2 * recip;
                const T phip = z2p * recip - phi * recip;
                const T psip = fpi * rhoip + sfp[lj] * rhoip + phip;
                const T fpair = -psip * recip;
                ax += dx * fpair;
                ay += dy * fpair;
                az += dz * fpair;
                if (EV) {
                  vir += (double)(rsq * ((T)0.5 * fpair));
                  eng += 0.5 * (double)phi;
                }
              }
            }
          }
        }
      }
      if (EAMT_TPA > 1) {
        ax = group_sum<EAMT_TPA>(ax);
        if (PASS == 2) { ay = group_sum<EAMT_TPA>(ay); az = group_sum<EAMT_TPA>(az); }
      }
      if (id >= 0 && sub == 0) {
        if (PASS == 1) {
          fp[id] = embed<T, UNIFORM>(E, ti, ax, EV != 0, emb);
        } else {
          Vec4<T> out;
          out.x = ax; out.y = ay; out.z = az; out.w = (T)0;
          f[id] = out;
        }
      }
    }
  }
  if (EV) {
    if (PASS == 1) {
      const double a1[1] = {emb};
      block_accumulate<1>(a1, ev_out + 3);
    } else {
      const double v2[2] = {eng, vir};
      block_accumulate<2>(v2, ev_out);
    }
  }
}

template <class T> __host__ __device__ inline size_t eam_tile_smem_bytes(int hcap, int pass) {
  return tile_smem_bytes<T>(hcap, true) + 32 + (pass == 2 ? (size_t)hcap * sizeof(T) : 0);
}

}  // namespace mmd

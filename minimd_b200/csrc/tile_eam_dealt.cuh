// Embedded-atom force over bank-dealt tile rows (tile_dealt_kernels.cuh): the two EAM passes of the full-neighbor
// formulation, ForceEAM::compute_fullneigh (ref/force_eam.cpp:274-449) -- and, evaluated from both ends, the pair set of
// compute_halfneigh (:94-270) -- as owner-computes shared-memory kernels.
//   pass 1  rho_i = sum_j rho(r_ij)  ->  fp_i = F'(rho_i) and the embedding energy F(rho_i) (:336-347), written in atom
//           order (for the fp halo, ForceEAM::communicate :851-914) and into the slot-ordered fp mirror
//   halo    fp of the ghosts -- between the two launches; the ghosts' values are then copied into the mirror
//   pass 2  pair forces (:368-441) from positions, fp and spline tables in shared memory; F_i written once, or consumed
//           by the fused velocity-Verlet epilogue (INTEG, as force_lj_dealt_kernel)
// Everything a pair needs sits in shared memory: the window's positions (bulk async copies from the slot-ordered
// mirror), its fp values (cp.async), and the spline tables of the pass.  The tables are repacked once per setup into
// arrays of coefficient PAIRS (16-byte records for FP64) so that a knot gather is an LDS.128 whose bank group is the
// knot index mod 8 -- the reference's 7-coefficient knots (ref/force_eam.cpp:765-793) would put eight lanes on four
// bank groups.  Pass 2 derives z2' from the value cubic, (3 c3 p + 2 c4) p + c5 times 1/dr, instead of gathering the
// derivative knots (the same polynomial, ref :789-792, up to rounding): one table gather less per pair.
// Uniform tables only (what ForceEAM::init_style produces: all type pairs share the Cu table, :753-760); per-type tables
// use the classic kernels.
// Energy / virial: a pair met from both ends contributes 0.5*phi and 0.5*r^2*fpair each time (:431-446); the caller
// applies eng_vdwl = embed + S (half-list semantics) or 2*(embed + S) (full).
#pragma once
#include "force_eam_kernels.cuh"
#include "tile_dealt_kernels.cuh"

namespace mmd {

// device blobs of pair-split spline coefficients; knots padded to a multiple of 4 so that every array is a multiple of
// 16 bytes in either precision
template <class T> struct EAMDealtTabs {
  const unsigned char* blob1;  // pass 1: rhoA[nkp] = (c3,c4) | rhoB[nkp] = (c5,c6)
  const unsigned char* blob2;  // pass 2: rdA[nkp] = (c0,c1) | z2A[nkp] = (c3,c4) | z2B[nkp] = (c5,c6) | rdB[nkp] = c2
  int nkp;
  __host__ __device__ size_t bytes1() const { return (size_t)nkp * 2 * sizeof(Vec2<T>); }
  __host__ __device__ size_t bytes2() const { return (size_t)nkp * (3 * sizeof(Vec2<T>) + sizeof(T)); }
};

template <class T> __host__ __device__ inline size_t eam_dealt_smem_bytes(int hcap, int scap, int pass, const EAMDealtTabs<T>& D) {
  return qwin_smem_bytes<T>(hcap, false, scap) + (pass == 2 ? (size_t)hcap * sizeof(T) + D.bytes2() : D.bytes1()) + 32;
}

__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
template <class T> __device__ __forceinline__ void cp_async_real(T* dst, const T* src) {
  if constexpr (sizeof(T) == 8) cp_async8(dst, src); else cp_async4(dst, src);
}

template <class T, int PASS, int EV, int INTEG, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 2 : 1)
eam_dealt_kernel(const Vec4<T>* __restrict__ x, Vec4<T>* __restrict__ f, TileGeo g, const int2* __restrict__ tile_runs,
                 const int4* __restrict__ tile_center, const int2* __restrict__ tile_info, XsMirror<T> xs_in,
                 const unsigned long long* __restrict__ rowsq, const int2* __restrict__ row_atom, int tcapq, int nlocal,
                 int scap, EAMTables<T> E, EAMDealtTabs<T> D, T* __restrict__ fp, T* __restrict__ fp_s,
                 VerletParams<T> VP, XsMirror<T> xs_out, double* __restrict__ ev_out /* [0] 0.5 phi, [1] virial, [2] ke, [3] embed */) {
  extern __shared__ __align__(16) unsigned char tile_smem_raw[];
  const int t = blockIdx.x;
  const int2 inf = tile_info[t];
  if (inf.y == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  const int p = lane & (QL - 1), qg = lane >> 3;
  const int wpr = tcapq / QB;
  const int H = inf.x;
  const unsigned long long sent4 = 0x0001000100010001ull * (unsigned long long)(g.hcap - 8 + p);

  QWin<T> S;
  S.carve(tile_smem_raw, g.hcap, scap);
  unsigned char* extra = tile_smem_raw + ((qwin_smem_bytes<T>(g.hcap, false, scap) + 15) & ~(size_t)15);
  T* sfp = reinterpret_cast<T*>(extra);                                   // pass 2: fp of the halo window
  unsigned char* tabs = extra + (PASS == 2 ? (size_t)g.hcap * sizeof(T) : 0);
  const Vec2<T>* tA = reinterpret_cast<const Vec2<T>*>(tabs);             // pass 1: rhoA         pass 2: rdA
  const Vec2<T>* tB = tA + D.nkp;                                          // pass 1: rhoB         pass 2: z2A
  const Vec2<T>* tC = tB + D.nkp;                                          //                      pass 2: z2B
  const T* tD = reinterpret_cast<const T*>(tC + D.nkp);                    //                      pass 2: rdB

  // ---- phase 1: tables, then the asynchronous copies (window positions, spline tables, fp) ----
  const int2* tr = tile_runs + (size_t)t * g.nrun;
  int2 my_run = make_int2(0, 0);
  int my_len = 0;
  if (tid < g.nrun) {
    my_run = __ldg(tr + tid);
    my_len = (tid + 1 < g.nrun ? __ldg(tr + tid + 1).y : H) - my_run.y;
    S.run_start[tid] = my_run.x;
    S.run_off[tid] = my_run.y;
  }
  if (tid == 0) {
    S.run_off[g.nrun] = H;
    mbar_init(S.bar, 1);
  }
  if (tid >= 32 && tid < 40) {  // the eight sentinel atoms, one per bank class
    Vec4<T> far;
    far.x = far.y = far.z = sentinel_coord<T>();
    far.w = type_to_lane<T>(0);
    S.put(g.hcap - 40 + tid, far, false);
    if (PASS == 2) sfp[g.hcap - 40 + tid] = (T)0;
  }
  if (w == 0) {
    int4 ce = make_int4(0, 0, 0, 0);
    int slot0 = 0;
    if (lane < TILE_NCENTER) {
      ce = __ldg(tile_center + (size_t)t * TILE_NCENTER + lane);
      const int2 r = __ldg(tr + (lane % TBY + g.sy) + (lane / TBY + g.sz) * g.nry);
      slot0 = r.x - r.y;
    }
    if (lane < TILE_NCENTER) S.pc[lane] = make_int4(ce.x, ce.y - ce.x, ce.z, slot0);
  }
  __syncthreads();
  const unsigned tab_bytes = (unsigned)(PASS == 1 ? D.bytes1() : D.bytes2());
  if (tid == 0) {
    mbar_expect_tx(S.bar, (unsigned)H * 16u + tab_bytes);
    bulk_g2s(tabs, PASS == 1 ? D.blob1 : D.blob2, tab_bytes, S.bar);
  }
  if (my_len > 0) bulk_g2s(S.rec + my_run.y, xs_in.rec + my_run.x, (unsigned)my_len * 16u, S.bar);
  if (sizeof(T) == 8 || PASS == 2) {
    for (int r = w; r < g.nrun; r += nw) {
      const int start = S.run_start[r], off = S.run_off[r], len = S.run_off[r + 1] - off;
      for (int k = lane; k < len; k += 32) {
        if constexpr (sizeof(T) == 8) cp_async8(S.z + off + k, xs_in.z + start + k);
        if (PASS == 2) cp_async_real(sfp + off + k, fp_s + start + k);
      }
    }
  }

  // row j of the tile -> {tile-local index of its atom, centre pencil}: the rows of a tile are numbered pencil by
  // pencil (tile_table_kernel), so a pass is simply four consecutive rows
  const int qtile0 = S.pc[0].z;
  const int nrows = S.pc[TILE_NCENTER - 1].z + S.pc[TILE_NCENTER - 1].y - qtile0;
  for (int c = w; c < TILE_NCENTER; c += nw) {
    const int4 pc = S.pc[c];
    for (int k = lane; k < pc.y; k += 32) S.stash_a[pc.z - qtile0 + k] = (pc.x + k) | (c << 16);
  }
  const int NP = (nrows + 3) >> 2;
  int2 ta_n = make_int2(-1, 0);
  unsigned long long w_n = sent4;
  int j_n = 0;
  bool have_n = false;
  auto locate = [&](int i) {
    const int j = i * 4 + qg;
    have_n = j < nrows;
    j_n = have_n ? j : 0;
    ta_n = make_int2(-1, 0);
    w_n = sent4;
    if (have_n) {
      ta_n = __ldg(row_atom + (qtile0 + j));
      w_n = ldg_rowq(rowsq + (size_t)(qtile0 + j) * wpr + p);
    }
  };
  locate(w);

  cp_async_wait_all();
  mbar_wait(S.bar, 0);
  __syncthreads();

  // ---- phase 2: the passes ----
  double eng = 0.0, vir = 0.0, ke = 0.0, emb = 0.0;
  const T cutsq = E.cutforcesq, rdr = E.rdr;
  const int nr = E.nr;
  for (int i = w; i < NP; i += nw) {
    const bool have = have_n;
    const int j = j_n, q = qtile0 + j;
    const int a = S.stash_a[j] & 0xffff;
    const int2 ta = ta_n;
    const bool own = have && ta.x >= 0 && ta.x < nlocal;
    const int n = own ? max(ta.y, 0) : 0;
    const int G = (n + QL - 1) / QL;
    const int nwords = (G + QB - 1) / QB;
    const unsigned long long* __restrict__ rowq = rowsq + (size_t)q * wpr + p;
    unsigned long long w0 = nwords > 0 ? w_n : sent4;
    unsigned long long w1 = sent4;
    if (nwords > 1) w1 = ldg_rowq(rowq + QL);
    if (PASS == 2 && INTEG && own && p == 1) prefetch_l2(VP.v + ta.x);
    locate(i + nw);
    const int gmax = __reduce_max_sync(0xffffffffu, G);
    T xi, yi, zi;
    S.get(a, xi, yi, zi);
    const T fpi = PASS == 2 ? sfp[a] : (T)0;
    T ax = 0, ay = 0, az = 0;  // pass 1: ax = rho_i; pass 2: the force
    auto pair = [&](int lj) {
      T xj, yj, zj;
      S.get(lj, xj, yj, zj);
      const T dx = xi - xj, dy = yi - yj, dz = zi - zj;
      const T rsq = dx * dx + dy * dy + dz * dz;
      if (rsq < cutsq) {
        const T r = real_sqrt(rsq);
        int m; T pp;
        knot_r(r, rdr, nr, m, pp);
        if (PASS == 1) {
          const Vec2<T> c34 = tA[m], c56 = tB[m];
          ax += ((c34.x * pp + c34.y) * pp + c56.x) * pp + c56.y;
        } else {
          const Vec2<T> d01 = tA[m], z34 = tB[m], z56 = tC[m];
          const T d2 = tD[m];
          const T rhoip = (d01.x * pp + d01.y) * pp + d2;
          const T z2 = ((z34.x * pp + z34.y) * pp + z56.x) * pp + z56.y;
          const T z2p = (((T)3 * z34.x * pp + (T)2 * z34.y) * pp + z56.x) * rdr;
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
    };
    const int nfull = gmax >> 2;
    for (int b = 0; b < nfull; b++) {
      const unsigned long long cur = w0;
      w0 = w1;
      w1 = sent4;
      if (b + 2 < nwords) w1 = ldg_rowq(rowq + (size_t)(b + 2) * QL);
      if (PASS == 1) {
#pragma unroll
        for (int e = 0; e < QB; e++) pair((int)((cur >> (16 * e)) & 0x7fffull));
      } else {  // the pair pass keeps two evaluations in flight (register budget of 64)
#pragma unroll 1
        for (int h = 0; h < QB; h += 2) {
          const unsigned two = (unsigned)(cur >> (16 * h));
          pair((int)(two & 0x7fffu));
          pair((int)((two >> 16) & 0x7fffu));
        }
      }
    }
    if (gmax & 2) {
      pair((int)(w0 & 0x7fffull));
      pair((int)((w0 >> 16) & 0x7fffull));
    }
    if (gmax & 1) pair((int)((w0 >> ((gmax & 2) * 16)) & 0x7fffull));
    if (PASS == 1) {
      ax = group_sum<QL>(ax);
      if (p == 0 && have) S.stash_f[j * 3] = ax;
    } else {
      // the three sums over the atom's eight lanes, as one butterfly that halves the number of values a lane carries
      // at every step: lanes 0,2,4 of the quarter warp end up with F_x, F_y, F_z
      const bool hi4 = (p & 4) != 0, hi2 = (p & 2) != 0;
      const T s1 = __shfl_xor_sync(0xffffffffu, hi4 ? ax : az, 4, 32);
      const T s2 = __shfl_xor_sync(0xffffffffu, ay, 4, 32);
      const T A = (hi4 ? az : ax) + s1;
      const T B = hi4 ? (T)0 : ay + s2;
      const T s3 = __shfl_xor_sync(0xffffffffu, hi2 ? A : B, 2, 32);
      T C = (hi2 ? B : A) + s3;
      C += __shfl_xor_sync(0xffffffffu, C, 1, 32);
      if (have && !(p & 1) && p < 6) S.stash_f[j * 3 + (p >> 1)] = C;
    }
  }
  __syncthreads();

  // ---- phase 3: one thread per row of the tile ----
  for (int j = tid; j < nrows; j += blockDim.x) {
    const int2 ta = __ldg(row_atom + (size_t)(qtile0 + j));
    if (ta.x < 0 || ta.x >= nlocal) continue;
    const int id = ta.x;
    const int ac = S.stash_a[j];
    const int a = ac & 0xffff;
    if (PASS == 1) {
      // fp = F'(rho), embedding energy F(rho) on thermo steps (ref/force_eam.cpp:336-347)
      const T fpv = embed<T, 1>(E, 0, S.stash_f[j * 3], EV != 0, emb);
      fp[id] = fpv;
      fp_s[S.pc[ac >> 16].w + a] = fpv;
    } else {
      const T fxa = S.stash_f[j * 3 + 0], fya = S.stash_f[j * 3 + 1], fza = S.stash_f[j * 3 + 2];
      if (INTEG) {
        T xa, ya, za;
        S.get(a, xa, ya, za);
        Vec4<T> vi = VP.v[id];
        vi.x += VP.dtforce * fxa;
        vi.y += VP.dtforce * fya;
        vi.z += VP.dtforce * fza;
        if (EV) ke += (double)((vi.x * vi.x + vi.y * vi.y + vi.z * vi.z) * VP.mass);
        vi.x += VP.dtforce * fxa;
        vi.y += VP.dtforce * fya;
        vi.z += VP.dtforce * fza;
        Vec4<T> xo;
        xo.x = xa + VP.dt * vi.x;
        xo.y = ya + VP.dt * vi.y;
        xo.z = za + VP.dt * vi.z;
        if constexpr (sizeof(T) == 8) xo.w = x[id].w; else xo.w = S.rec[a].w;
        VP.v[id] = vi;
        VP.x_out[id] = xo;
        xs_out.put_slot(S.pc[ac >> 16].w + a, xo);
      } else {
        Vec4<T> out;
        out.x = fxa; out.y = fya; out.z = fza; out.w = (T)0;
        f[id] = out;
      }
    }
  }
  if (EV) {
    if (PASS == 1) {
      const double a1[1] = {emb};
      block_accumulate<1>(a1, ev_out + 3);
    } else if (INTEG) {
      const double v3[3] = {eng, vir, ke};
      block_accumulate<3>(v3, ev_out);
    } else {
      const double v2[2] = {eng, vir};
      block_accumulate<2>(v2, ev_out);
    }
  }
}

// ghosts' fp (written in atom order by the fp halo) -> slot-ordered mirror
template <class T>
__global__ void fp_mirror_ghosts_kernel(const T* __restrict__ fp, int nlocal, int nghost, const int* __restrict__ slot_of,
                                        T* __restrict__ fp_s) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < nghost) fp_s[slot_of[nlocal + g]] = fp[nlocal + g];
}

}  // namespace mmd

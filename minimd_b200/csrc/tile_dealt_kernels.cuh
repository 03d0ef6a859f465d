// Bank-dealt neighbor rows and the pair-force kernel that walks them.
//
// Why (profiles/r1_tile_kernels.md): the first tile force kernel gave every lane its own row, so the 32 gathers of a warp
// instruction hit unrelated shared-memory banks: 97 M LSU wavefronts per launch at -s 80 of which 51 M were bank-conflict
// replays (LSU pipe 79 %, the binding resource).  The order of the entries inside a row is free (the export restores the
// reference's order from the original CSR position), so a row can be laid out for the banks:
//
//   * a QUARTER WARP (8 lanes) owns one atom and reads 8 entries of its row per step.  The shared-memory image keeps
//     (x,y) of an atom in one 16-byte record (FP32: x,y,z,type in one record), so the gather is an LDS.128, which the
//     hardware serves one quarter warp per phase -- conflicts can only arise among the 8 entries of ONE row;
//   * the row is sorted by bank class (tile-local index mod 8 = the 16-byte bank group of the record) and dealt round
//     robin into G = ceil(n/8) groups: group g holds the sorted entries g, g+G, g+2G, ... -- one per lane.  A bank class with
//     c entries puts ceil(c/G) of them into a group, i.e. one (no conflict) unless the class is over-represented.
//
// Storage: entry (group g, lane p) of a row is the 16-bit word (g>>2)*32 + p*4 + (g&3): lane p fetches the entries of
// four consecutive groups with one 64-bit load, a quarter warp reads 64 contiguous bytes.  Rows are tcapq entries long
// (a multiple of 32).  An entry is the tile-local index with the build's half-list flag in bit 15 (the force kernels
// mask it off, the export and the counters use it); every slot that holds no neighbor -- (g,p) with g >= G or
// p*G + g >= n, up to the end of the row's last 64-byte block -- holds a SENTINEL index (hcap-8+p), shared-memory slots
// the force kernel fills with a far-away position: the pair loop carries no validity logic at all, a sentinel pair
// simply fails the cutoff test.
//
// The kernel evaluates every atom's complete neighborhood ("owner computes", as tile_kernels.cuh): the pair set is the
// reference's (ref/force_lj.cpp:185-263 half, :366-449 full), forces/energy/virial agree up to summation order.
#pragma once
#include "tile_kernels.cuh"
#include "xs_mirror.cuh"

namespace mmd {

constexpr int QL = 8;            // lanes per atom
constexpr int QB = 4;            // groups per 64-bit row word
constexpr int QBLK = QL * QB;    // entries per row block (64 bytes)

__host__ __device__ inline int dealt_capacity(int longest_row) { return ((longest_row > 1 ? longest_row : 1) + QBLK - 1) / QBLK * QBLK; }

// ---- asynchronous staging primitives (sm_90+): mbarrier + bulk copy (TMA engine, SASS UBLKCP) + cp.async (LDGSTS) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// global -> shared bulk copy, completion counted in bytes on the mbarrier; dst, src and bytes are multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---------------------------------------------------------------------------------------
// rows in candidate order (neigh_build_*_kernel) -> bank-dealt rows.  A CTA takes 128 consecutive rows: ONE bulk copy
// (cp.async.bulk, completion on an mbarrier) brings them into shared memory while the threads prefill the sentinel
// pattern of the output; then ONE THREAD PER ROW counts the 8 classes and places every entry at its dealt position
// inside a second shared-memory copy, and the CTA writes the dealt rows out with 16-byte stores.  The staged rows keep
// their global stride (tcap/2 words, a multiple of 4), so thread r starts its walk at word (r - r*stride) mod 32 of
// its row and wraps: the threads of a warp then sit on 32 different banks.  Empty slots of a lane p hold the sentinel
// index sentinel0 + p: hcap-8 .. hcap-1 are eight far-away atoms, one per bank class, so that a sentinel read by lane
// p (which mostly holds entries of the classes around p) rarely collides with a real entry of its group.
// ---------------------------------------------------------------------------------------
constexpr int DEAL_THREADS = 128;
constexpr int DEAL_MAXROW = 255;  // longest row the 8-bit class counters handle
__host__ __device__ inline size_t deal_smem_bytes(int tcap, int tcapq) {
  return (size_t)DEAL_THREADS * ((tcap / 2) + (tcapq / 2 + 1)) * 4 + DEAL_THREADS * 4 + 16;
}

__global__ void __launch_bounds__(DEAL_THREADS)
tile_rows_deal_kernel(const unsigned short* __restrict__ rows, const int2* __restrict__ row_atom, int nrows, int tcap,
                      int nlocal, unsigned short* __restrict__ rowsq, int tcapq, int sentinel0) {
  extern __shared__ __align__(16) unsigned char deal_smem[];
  const int sstride = tcap / 2, dstride = tcapq / 2 + 1;  // words per row: staged rows as in global memory, dealt rows odd
  unsigned* s_src = reinterpret_cast<unsigned*>(deal_smem);
  unsigned* s_dst = s_src + DEAL_THREADS * sstride;
  int* s_n = reinterpret_cast<int*>(s_dst + DEAL_THREADS * dstride);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>((reinterpret_cast<size_t>(s_n + DEAL_THREADS) + 7) & ~(size_t)7);
  const int q0 = blockIdx.x * DEAL_THREADS;
  const int q = q0 + threadIdx.x;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    const unsigned bytes = (unsigned)min(DEAL_THREADS, nrows - q0) * (unsigned)tcap * 2u;  // tcap is a multiple of 8
    mbar_expect_tx(bar, bytes);
    bulk_g2s(s_src, rows + (size_t)q0 * tcap, bytes, bar);
  }
  int n = 0;
  if (q < nrows) {
    const int2 ta = row_atom[q];
    if (ta.x >= 0 && ta.x < nlocal) n = min(min(max(ta.y, 0), tcap), tcapq);
  }
  s_n[threadIdx.x] = n;
  // sentinel pattern of a dealt row (every thread fills its own row: odd word stride, conflict-free):
  // word k holds slots 2k, 2k+1, both of lane (k >> 1) & 7
  {
    unsigned* d = s_dst + threadIdx.x * dstride;
    for (int k = 0; k < tcapq / 2; k++) {
      const unsigned sv = (unsigned)(sentinel0 + ((k >> 1) & 7));
      d[k] = sv | (sv << 16);
    }
  }
  __syncthreads();      // the barrier is initialised, s_n is complete
  mbar_wait(bar, 0);    // the staged rows have landed
  if (n > 0) {
    // class counts, 8 bits each (n <= 255): classes 0-3 in c0, 4-7 in c1; two entries per shared-memory word
    const unsigned* __restrict__ srcw = s_src + threadIdx.x * sstride;
    unsigned c0 = 0u, c1 = 0u;
    const int nw2 = (n + 1) >> 1;
    // first word of this thread's walk.  The dealt positions must not depend on it (row numbers are handed out to the
    // tiles in launch order, so a row's thread index differs from run to run): entries are ranked within their class
    // in ROW order -- the walk starts with the rank of its first word and falls back to zero when it wraps
    const int rot = ((threadIdx.x - threadIdx.x * sstride) & 31) % nw2;
    unsigned snap0 = 0u, snap1 = 0u;   // class counts of the words [rot, nw2)
    for (int k = 0, kk = rot; k < nw2; k++) {
      const unsigned wd = srcw[kk];
      const unsigned i0 = 1u << ((wd & 3u) << 3), i1 = 1u << (((wd >> 16) & 3u) << 3);
      const bool two = 2 * kk + 1 < n;
      c0 += (wd & 4u) ? 0u : i0;
      c1 += (wd & 4u) ? i0 : 0u;
      c0 += (two && !(wd & 0x40000u)) ? i1 : 0u;
      c1 += (two && (wd & 0x40000u)) ? i1 : 0u;
      if (kk + 1 == nw2) { kk = 0; snap0 = c0; snap1 = c1; } else kk++;
    }
    // exclusive prefix over the classes (byte k of the product = sum of the bytes below k; totals < 256)
    const unsigned base0 = c0 * 0x01010100u;
    const unsigned base1 = c1 * 0x01010100u + ((c0 * 0x01010101u) >> 24) * 0x01010101u;
    unsigned p0 = base0 + (c0 - snap0), p1 = base1 + (c1 - snap1);   // + the class members in the words [0, rot): no byte carries
    const int G = (n + QL - 1) / QL;
    const unsigned invG = 65536u / (unsigned)G + 1u;   // t / G == (t * invG) >> 16 for t < 256, G <= 32 (checked exhaustively)
    unsigned short* dst = reinterpret_cast<unsigned short*>(s_dst + threadIdx.x * dstride);
    auto place = [&](unsigned ent) {
      const unsigned sh = (ent & 3u) << 3;
      const bool hi = (ent & 4u) != 0u;
      const unsigned pre = hi ? p1 : p0;
      const int t = (int)((pre >> sh) & 0xffu);
      const unsigned inc = 1u << sh;
      p0 += hi ? 0u : inc;
      p1 += hi ? inc : 0u;
      const int p = (int)(((unsigned)t * invG) >> 16);
      const int g = t - p * G;
      dst[((g >> 2) << 5) + (p << 2) + (g & 3)] = (unsigned short)ent;   // the half-list flag (bit 15) travels along
    };
    for (int k = 0, kk = rot; k < nw2; k++) {
      const unsigned wd = srcw[kk];
      place(wd & 0xffffu);
      if (2 * kk + 1 < n) place(wd >> 16);
      if (kk + 1 == nw2) { kk = 0; p0 = base0; p1 = base1; } else kk++;
    }
  }
  __syncthreads();
  // write out: ceil(n/32) blocks of 64 bytes per row; 8 rows per step, 16 lanes x 16 bytes per row
  {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll 4
    for (int r = ty; r < DEAL_THREADS; r += DEAL_THREADS / 16) {
      const int nb = ((s_n[r] + QBLK - 1) / QBLK) * QBLK;
      for (int ch = tx; ch * 8 < nb; ch += 16) {
        const unsigned* sp = s_dst + r * dstride + ch * 4;
        uint4 o;
        o.x = sp[0]; o.y = sp[1]; o.z = sp[2]; o.w = sp[3];
        *reinterpret_cast<uint4*>(rowsq + (size_t)(q0 + r) * tcapq + ch * 8) = o;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Shared-memory image of a halo window for the dealt-row kernels (record formats: xs_mirror.cuh).
//   FP64: rec[hcap] = (x,y) 16-byte records, z[hcap]; types (per-type parameter tables only) in st[hcap]
//   FP32: rec[hcap] = (x,y,z,type bits) 16-byte records
// followed by the CTA's force stash (one finished force per row of the tile: the store / Verlet epilogue runs one thread
// per atom after all rows are done), the centre-pencil tables and the mbarrier of the bulk copies.
// ---------------------------------------------------------------------------------------
template <class T> struct QWin {
  QRec<T>* rec;
  T* z;                // FP64 only
  T* stash_f;          // [scap][3]
  int* stash_a;        // [scap] row of the tile -> tile-local index of its atom | centre pencil << 16 (the pass table)
  int4* pc;            // [16] centre pencils: {first tile-local index, atoms, first row, CSR slot of tile-local index 0}
  int* pp;             // [17] (unused since the passes are four consecutive rows; keeps the layout of the tables)
  int* run_start;      // [TILE_MAXRUN]
  int* run_off;        // [TILE_MAXRUN + 1]
  unsigned long long* bar;
  unsigned char* st;   // FP64 + per-type tables only
  __device__ __forceinline__ void carve(unsigned char* base, int hcap, int scap) {
    rec = reinterpret_cast<QRec<T>*>(base);
    unsigned char* q = base + (size_t)hcap * 16;
    z = reinterpret_cast<T*>(q);
    if (sizeof(T) == 8) q += (size_t)hcap * 8;
    pc = reinterpret_cast<int4*>(q);
    q += TILE_NCENTER * sizeof(int4);
    bar = reinterpret_cast<unsigned long long*>(q);
    q += 16;
    stash_f = reinterpret_cast<T*>(q);
    q += (size_t)scap * 3 * sizeof(T);
    stash_a = reinterpret_cast<int*>(q);
    q += (size_t)scap * sizeof(int);
    pp = reinterpret_cast<int*>(q);
    run_start = pp + 20;
    run_off = run_start + TILE_MAXRUN;
    st = reinterpret_cast<unsigned char*>(run_off + TILE_MAXRUN + 1);
  }
  __device__ __forceinline__ void put(int k, const Vec4<T>& v, bool types) {
    if constexpr (sizeof(T) == 8) {
      QRec<T> r; r.x = v.x; r.y = v.y;
      rec[k] = r;
      z[k] = v.z;
      if (types) st[k] = (unsigned char)lane_to_type(v.w);
    } else {
      QRec<T> r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
      rec[k] = r;
    }
  }
  __device__ __forceinline__ void get(int k, T& x, T& y, T& zz) const {
    const QRec<T> r = rec[k];
    x = r.x; y = r.y;
    if constexpr (sizeof(T) == 8) zz = z[k]; else zz = r.z;
  }
  __device__ __forceinline__ void get(int k, T& x, T& y, T& zz, int& type) const {
    const QRec<T> r = rec[k];
    x = r.x; y = r.y;
    if constexpr (sizeof(T) == 8) { zz = z[k]; type = (int)st[k]; }
    else { zz = r.z; type = lane_to_type(r.w); }
  }
};
template <class T> __host__ __device__ inline size_t qwin_smem_bytes(int hcap, bool with_types, int scap) {
  return (size_t)hcap * (sizeof(T) == 8 ? 24 : 16) + TILE_NCENTER * sizeof(int4) + 16 + (size_t)scap * (3 * sizeof(T) + sizeof(int)) +
         (20 + 2 * TILE_MAXRUN + 1) * sizeof(int) + ((with_types && sizeof(T) == 8) ? (size_t)hcap : 0) + 16;
}

// one 64-bit row word = the entries of four consecutive groups; the L2::128B hint pulls the atom's next block too
__device__ __forceinline__ unsigned long long ldg_rowq(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.global.nc.L2::128B.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// reciprocal for the pair loop.  FP64: MUFU.RCP64H seed + one third-order (Halley) step, y(1 + e + e^2) with e = 1 - x*y:
// the seed is good to ~2^-22, so the result is good to ~2^-60 -- three DFMA instead of the four of two Newton steps,
// and no IEEE-division slow path.  FP32: the hardware reciprocal.
__device__ __forceinline__ double pair_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  const double e2 = fma(e, e, e);
  return fma(y, e2, y);
}
__device__ __forceinline__ float pair_rcp(float x) { return __frcp_rn(x); }

template <class T> struct LJDealtParams {
  T cutforcesq, sigma6, epsilon;
  T k48;   // 48 * epsilon * sigma6
  T kA, kB;  // 48 eps sigma6^2 and -24 eps sigma6: F/r = r^-8 (kA r^-6 + kB), the uniform-parameter form of the line below
  const T* cutforcesq_tab;
  const T* sigma6_tab;
  const T* epsilon_tab;
  int ntypes;
  double e_scale, v_scale;
};

// far-away position of the sentinel slot: its distance to any atom is finite (no inf/NaN) and beyond every cutoff
template <class T> __device__ __forceinline__ T sentinel_coord() { return sizeof(T) == 8 ? (T)1e150 : (T)1e18f; }

// one pair: neighbor at tile-local index lj
template <class T, int EV, int UNIFORM>
__device__ __forceinline__ void lj_pair(const QWin<T>& S, const LJDealtParams<T>& P, int lj, T xi, T yi, T zi, int ti,
                                        T& fx, T& fy, T& fz, double& eng, double& vir) {
  T xj, yj, zj;
  int tj = 0;
  if (UNIFORM) S.get(lj, xj, yj, zj); else S.get(lj, xj, yj, zj, tj);
  const T dx = xi - xj, dy = yi - yj, dz = zi - zj;
  const T rsq = dx * dx + dy * dy + dz * dz;
  T cut, s6, k48, eps;
  if (UNIFORM) {
    cut = P.cutforcesq; s6 = P.sigma6; k48 = P.k48; eps = P.epsilon;
  } else {
    const int tij = ti * P.ntypes + tj;
    cut = __ldg(P.cutforcesq_tab + tij); s6 = __ldg(P.sigma6_tab + tij); eps = __ldg(P.epsilon_tab + tij);
    k48 = (T)48 * eps * s6;
  }
  // evaluated unconditionally (r^2 > 0 for distinct atoms; the sentinel gives a finite 3e300), accumulated under the
  // cutoff predicate: no select, no branch
  const bool hit = rsq < cut;
  const T a1 = pair_rcp(rsq);
  const T a2 = a1 * a1;
  const T a3 = a2 * a1;
  // F/r = 48 eps sr6 (sr6 - 0.5) / r^2 with sr6 = sigma6 / r^6 (ref/force_lj.cpp:232-235)
  const T force = UNIFORM ? (a2 * a2) * (a3 * P.kA + P.kB) : (a2 * a2) * (a3 * s6 - (T)0.5) * k48;
  if (hit) {
    fx += dx * force;
    fy += dy * force;
    fz += dz * force;
  }
  if (EV) {
    const T sr6 = a3 * s6;
    if (hit) {
      eng += (double)((T)4 * sr6 * (sr6 - (T)1) * eps);
      vir += (double)(rsq * force);
    }
  }
}

// FP32, uniform parameters: TWO neighbors at a time on Blackwell's packed FP32 pipe (FFMA2 / FMUL2: one instruction,
// two IEEE FP32 results).  The FP32 launch is issue-bound, and the pair body is 17 FP32 instructions: packing two
// neighbors into every arithmetic instruction removes about a quarter of the kernel's instructions.  f2 = (fxA, fxB)
// etc. are pairs of partial sums, added up once per atom.  EV = 0 only (thermo steps take the scalar path).
__device__ __forceinline__ void lj_pair2_f32(const QWin<float>& S, const LJDealtParams<float>& P, int ljA, int ljB, float2 xi2,
                                             float2 yi2, float2 zi2, float2& fx2, float2& fy2, float2& fz2) {
  const QRec<float> a = S.rec[ljA], b = S.rec[ljB];
  const float2 m1 = make_float2(-1.0f, -1.0f);
  const float2 dx = __ffma2_rn(make_float2(a.x, b.x), m1, xi2);
  const float2 dy = __ffma2_rn(make_float2(a.y, b.y), m1, yi2);
  const float2 dz = __ffma2_rn(make_float2(a.z, b.z), m1, zi2);
  const float2 rsq = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
  const float2 a1 = make_float2(__frcp_rn(rsq.x), __frcp_rn(rsq.y));
  const float2 a2 = __fmul2_rn(a1, a1);
  const float2 a3 = __fmul2_rn(a2, a1);
  const float2 t = __ffma2_rn(a3, make_float2(P.kA, P.kA), make_float2(P.kB, P.kB));
  float2 force = __fmul2_rn(__fmul2_rn(a2, a2), t);
  force.x = rsq.x < P.cutforcesq ? force.x : 0.0f;
  force.y = rsq.y < P.cutforcesq ? force.y : 0.0f;
  fx2 = __ffma2_rn(dx, force, fx2);
  fy2 = __ffma2_rn(dy, force, fy2);
  fz2 = __ffma2_rn(dz, force, fz2);
}

// INTEG = 1: the velocity-Verlet halves that follow the force ride in the epilogue (see VerletParams, tile_kernels.cuh):
// new positions go to x_out[] AND to the other buffer of the slot-ordered mirror (xs_out), which the next launch stages.
//
// Phases of a CTA (one tile):
//   1. thread p < nrun issues ONE bulk copy for pencil run p of the halo window (mirror -> shared memory, 16-byte
//      records); FP64 z values follow as 8-byte cp.async.  Row words of the first passes are requested meanwhile.
//   2. passes of 4 atoms (a quarter warp each): four consecutive rows of the tile, dealt round robin over the 16 warps;
//      the three force sums of an atom go through one split butterfly into the CTA's stash.
//   3. epilogue with one thread per atom: store f, or finalIntegrate(n) + initialIntegrate(n+1) (ref/integrate.cpp:46-68).
template <class T, int EV, int UNIFORM, int INTEG>
__global__ void __launch_bounds__(TILE_THREADS, 2)
force_lj_dealt_kernel(const Vec4<T>* __restrict__ x, Vec4<T>* __restrict__ f, TileGeo g, const int2* __restrict__ tile_runs,
                      const int4* __restrict__ tile_center, const int2* __restrict__ tile_info, XsMirror<T> xs_in,
                      const unsigned char* __restrict__ types_s, const unsigned long long* __restrict__ rowsq,
                      const int2* __restrict__ row_atom, int tcapq, int nlocal, int scap, LJDealtParams<T> P,
                      VerletParams<T> VP, XsMirror<T> xs_out, double* __restrict__ ev_out,
                      const int* __restrict__ tile_list /* tiles that own local atoms */, GhostImages<T> GI,
                      unsigned long long* __restrict__ prof /* {staging clocks, CTA clocks, CTAs}: -DMMD_KERNEL_PROFILE builds only */) {
  extern __shared__ __align__(16) unsigned char tile_smem_raw[];
  // tile_list holds tiles that own local atoms only (tile_classify_kernel): no CTA exits early, and nothing has to
  // be read from global memory before the copies of the window can be described
  const int t = __ldg(tile_list + blockIdx.x);
#ifdef MMD_KERNEL_PROFILE
  const long long clk0 = clock64();
#endif
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  const int p = lane & (QL - 1), qg = lane >> 3;
  const int wpr = tcapq / QB;  // 64-bit words per row
  const unsigned long long sent4 = 0x0001000100010001ull * (unsigned long long)(g.hcap - 8 + p);  // lane p's sentinel

  QWin<T> S;
  S.carve(tile_smem_raw, g.hcap, scap);

  // ---- phase 1: the asynchronous copies of the halo window.  One barrier up front (mbarrier visible to everybody, no
  //      load in flight yet); after that every thread describes and issues its copies as soon as ITS table entries
  //      arrive -- one round trip for the tables, one for the data, no barrier in between ----
  if (tid == 0) mbar_init(S.bar, 1);
  if (tid >= 32 && tid < 40) {  // the eight sentinel atoms, one per bank class
    Vec4<T> far;
    far.x = far.y = far.z = sentinel_coord<T>();
    far.w = type_to_lane<T>(0);
    S.put(g.hcap - 40 + tid, far, !UNIFORM);
  }
  __syncthreads();
  const int2* tr = tile_runs + (size_t)t * g.nrun;
  const int4* tc = tile_center + (size_t)t * TILE_NCENTER;
  if (tid < g.nrun) {  // thread r: pencil run r of the window, ONE bulk copy of 16-byte records
    const int2 my_run = __ldg(tr + tid);
    const int my_len = (tid + 1 < g.nrun ? __ldg(tr + tid + 1).y : __ldg(tc).w) - my_run.y;
    if (tid == 0) mbar_expect_tx(S.bar, (unsigned)__ldg(tc).w * 16u);  // tile_center[0].w = atoms of the window
    if (my_len > 0) bulk_g2s(S.rec + my_run.y, xs_in.rec + my_run.x, (unsigned)my_len * 16u, S.bar);
  }
  if constexpr (sizeof(T) == 8) {  // FP64: z values (8-byte cp.async), run tables read straight from global
    for (int r = w; r < g.nrun; r += nw) {
      const int2 r0 = __ldg(tr + r);
      const int len = (r + 1 < g.nrun ? __ldg(tr + r + 1).y : __ldg(tc).w) - r0.y;
      for (int k = lane; k < len; k += 32) {
        cp_async8(S.z + r0.y + k, xs_in.z + r0.x + k);
        if (!UNIFORM) S.st[r0.y + k] = __ldg(types_s + r0.x + k);
      }
    }
  }

  // centre pencils, held by EVERY warp in the registers of lanes 0..15: {first tile-local index, atoms, first row,
  // CSR slot of tile-local index 0}
  int4 pcl = make_int4(0, 0, 0, 0);
  if (lane < TILE_NCENTER) {
    const int4 ce = __ldg(tc + lane);
    const int2 r = __ldg(tr + (lane % TBY + g.sy) + (lane / TBY + g.sz) * g.nry);
    pcl = make_int4(ce.x, ce.y - ce.x, ce.z, r.x - r.y);
  }
  const int qtile0 = __shfl_sync(0xffffffffu, pcl.z, 0);
  const int nrows = __shfl_sync(0xffffffffu, pcl.z + pcl.y, TILE_NCENTER - 1) - qtile0;
  if (w == 0 && lane < TILE_NCENTER) S.pc[lane] = pcl;  // for the epilogue (read after the barriers below)
  // row j of the tile -> {tile-local index of its atom, centre pencil}: the rows of a tile are numbered pencil by
  // pencil (tile_table_kernel), so a pass is simply four consecutive rows
  for (int c = w; c < TILE_NCENTER; c += nw) {
    const int lo = __shfl_sync(0xffffffffu, pcl.x, c), cnt = __shfl_sync(0xffffffffu, pcl.y, c);
    const int r0 = __shfl_sync(0xffffffffu, pcl.z, c) - qtile0;
    for (int k = lane; k < cnt; k += 32) S.stash_a[r0 + k] = (lo + k) | (c << 16);
  }
  const int NP = (nrows + 3) >> 2;

  // pass i of the tile: rows 4i .. 4i+3, one per quarter warp
  int2 ta_n = make_int2(-1, 0);
  unsigned long long w_n = sent4;
  int j_n = 0;
  bool have_n = false;
  auto locate = [&](int i) {  // fills the *_n state for pass i (warp-uniform i)
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

  if constexpr (sizeof(T) == 8) cp_async_wait_all();
  mbar_wait(S.bar, 0);
  __syncthreads();
#ifdef MMD_KERNEL_PROFILE
  const long long clk1 = clock64();
#endif

  // ---- phase 2: the passes ----
  double eng = 0.0, vir = 0.0, ke = 0.0;
  for (int i = w; i < NP; i += nw) {
    const bool have = have_n;
    const int j = j_n, q = qtile0 + j;
    const int a = S.stash_a[j] & 0xffff;
    const int2 ta = ta_n;
    // an atom owns its row whatever the row's length (an isolated atom has an empty one and still integrates)
    const bool own = have && ta.x >= 0 && ta.x < nlocal;
    const int n = own ? max(ta.y, 0) : 0;
    const int G = (n + QL - 1) / QL;          // groups of this atom's row
    const int nwords = (G + QB - 1) / QB;     // row words that hold them; beyond: all-sentinel words
    const unsigned long long* __restrict__ rowq = rowsq + (size_t)q * wpr + p;
    unsigned long long w0 = nwords > 0 ? w_n : sent4;
    unsigned long long w1 = sent4;
    if (nwords > 1) w1 = ldg_rowq(rowq + QL);
    if (INTEG && own && p == 1) prefetch_l2(VP.v + ta.x);
    locate(i + nw);  // next pass of this warp
    const int gmax = __reduce_max_sync(0xffffffffu, G);   // the quarter warps of a pass run in lock step
    T xi, yi, zi;
    int ti = 0;
    if (UNIFORM) S.get(a, xi, yi, zi); else S.get(a, xi, yi, zi, ti);
    T fx = 0, fy = 0, fz = 0;
    const int nfull = gmax >> 2;
    constexpr bool PACKED = sizeof(T) == 4 && UNIFORM && !EV;  // FP32: two neighbors per packed instruction
    float2 fx2 = make_float2(0.f, 0.f), fy2 = fx2, fz2 = fx2;
    float2 xi2, yi2, zi2;
    if constexpr (PACKED) { xi2 = make_float2(xi, xi); yi2 = make_float2(yi, yi); zi2 = make_float2(zi, zi); }
    for (int b = 0; b < nfull; b++) {
      const unsigned long long cur = w0;
      w0 = w1;
      w1 = sent4;
      if (b + 2 < nwords) w1 = ldg_rowq(rowq + (size_t)(b + 2) * QL);
      if constexpr (PACKED) {
        lj_pair2_f32(S, P, (int)(cur & 0x7fffull), (int)((cur >> 16) & 0x7fffull), xi2, yi2, zi2, fx2, fy2, fz2);
        lj_pair2_f32(S, P, (int)((cur >> 32) & 0x7fffull), (int)((cur >> 48) & 0x7fffull), xi2, yi2, zi2, fx2, fy2, fz2);
      } else {
        // four independent pair evaluations in flight: the FP64 dependency chains overlap
#pragma unroll
        for (int e = 0; e < QB; e++)
          lj_pair<T, EV, UNIFORM>(S, P, (int)((cur >> (16 * e)) & 0x7fffull), xi, yi, zi, ti, fx, fy, fz, eng, vir);
      }
    }
    if (gmax & 2) {
      if constexpr (PACKED) {
        lj_pair2_f32(S, P, (int)(w0 & 0x7fffull), (int)((w0 >> 16) & 0x7fffull), xi2, yi2, zi2, fx2, fy2, fz2);
      } else {
        lj_pair<T, EV, UNIFORM>(S, P, (int)(w0 & 0x7fffull), xi, yi, zi, ti, fx, fy, fz, eng, vir);
        lj_pair<T, EV, UNIFORM>(S, P, (int)((w0 >> 16) & 0x7fffull), xi, yi, zi, ti, fx, fy, fz, eng, vir);
      }
    }
    if (gmax & 1)
      lj_pair<T, EV, UNIFORM>(S, P, (int)((w0 >> ((gmax & 2) * 16)) & 0x7fffull), xi, yi, zi, ti, fx, fy, fz, eng, vir);
    if constexpr (PACKED) {
      fx += fx2.x + fx2.y;
      fy += fy2.x + fy2.y;
      fz += fz2.x + fz2.y;
    }
    // the three sums over the atom's eight lanes, as one butterfly that halves the number of values a lane carries at
    // every step: lanes 0,2,4 of the quarter warp end up with F_x, F_y, F_z
    {
      const bool hi4 = (p & 4) != 0, hi2 = (p & 2) != 0;
      const T s1 = __shfl_xor_sync(0xffffffffu, hi4 ? fx : fz, 4, 32);
      const T s2 = __shfl_xor_sync(0xffffffffu, fy, 4, 32);
      const T A = (hi4 ? fz : fx) + s1;        // lanes 0-3: x, lanes 4-7: z
      const T B = hi4 ? (T)0 : fy + s2;        // lanes 0-3: y
      const T s3 = __shfl_xor_sync(0xffffffffu, hi2 ? A : B, 2, 32);
      T C = (hi2 ? B : A) + s3;                // lanes 0,1: x   2,3: y   4,5: z   6,7: 0
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
    const T fxa = S.stash_f[j * 3 + 0], fya = S.stash_f[j * 3 + 1], fza = S.stash_f[j * 3 + 2];
    if (INTEG) {
      const int ac = S.stash_a[j];
      const int a = ac & 0xffff;
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
      if constexpr (sizeof(T) == 8) xo.w = x[id].w;  // the type lane travels with the atom
      else xo.w = S.rec[a].w;
      VP.v[id] = vi;
      VP.x_out[id] = xo;
      xs_out.put_slot(S.pc[ac >> 16].w + a, xo);
      if (GI.start) {  // the atom's periodic images: the forward halo of the next step, done here
        const int ib = __ldg(GI.start + id), ie = __ldg(GI.start + id + 1);
        for (int k = ib; k < ie; k++) {
          const int2 im = __ldg(GI.list + k);
          const int sx = (im.y & 3) - 1, sy = ((im.y >> 2) & 3) - 1, sz = ((im.y >> 4) & 3) - 1;
          Vec4<T> pg = xo;
          if (sx) pg.x = pg.x + sx * GI.prd[0];
          if (sy) pg.y = pg.y + sy * GI.prd[1];
          if (sz) pg.z = pg.z + sz * GI.prd[2];
          const int ga = GI.nlocal + im.x;
          VP.x_out[ga] = pg;
          xs_out.put_atom(ga, pg);
        }
      }
    } else {
      Vec4<T> out;
      out.x = fxa; out.y = fya; out.z = fza; out.w = (T)0;
      f[id] = out;
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
    if (tid == 0) {
      atomicAdd(prof + 0, (unsigned long long)(clk1 - clk0));
      atomicAdd(prof + 1, (unsigned long long)(clock64() - clk0));
      atomicAdd(prof + 2, 1ull);
    }
  }
#endif
}

// ---------------------------------------------------------------------------------------
// Interior / boundary split for several ranks: a tile whose halo window holds no ghost atom needs nothing from the
// forward halo, so its force can run while the halo of the same step is still in flight.  One warp per tile scans the
// window's slots; lists[0 .. counts[0]) receives the interior tiles, lists[ntiles .. ntiles + counts[1]) the others,
// lists[2*ntiles .. 2*ntiles + counts[2]) every tile that owns local atoms (what a single launch covers).
// ---------------------------------------------------------------------------------------
// send_flag (may be null): 1 for every local atom that appears in a send list.  A tile that OWNS such an atom counts as
// boundary too, so the forward halo of step n+1 only reads positions written by the boundary kernel of step n.
__global__ void __launch_bounds__(128)
tile_classify_kernel(TileGeo g, const int2* __restrict__ tile_runs, const int4* __restrict__ tile_center,
                     const int2* __restrict__ tile_info, const int* __restrict__ slots, int nlocal,
                     const unsigned char* __restrict__ send_flag, int* __restrict__ lists, int* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (t >= g.ntiles) return;
  const int2 inf = tile_info[t];
  if (inf.y == 0) return;
  const int2* tr = tile_runs + (size_t)t * g.nrun;
  // the window's runs are contiguous slot ranges; run p ends where run p+1 starts in tile-local numbering
  bool ghost = false;
  for (int p = 0; p < g.nrun && !ghost; p++) {
    const int2 r = tr[p];
    const int len = (p + 1 < g.nrun ? tr[p + 1].y : inf.x) - r.y;
    bool gl = false;
    for (int k = lane; k < len; k += 32) gl = gl || (__ldg(slots + r.x + k) >= nlocal);
    ghost = __any_sync(0xffffffffu, gl);
  }
  if (!ghost && send_flag) {  // own atoms: the centre pencils' index ranges
    for (int c = 0; c < TILE_NCENTER && !ghost; c++) {
      const int4 ce = tile_center[(size_t)t * TILE_NCENTER + c];
      const int2 r = tr[(c % TBY + g.sy) + (c / TBY + g.sz) * g.nry];
      bool gl = false;
      for (int a = ce.x + lane; a < ce.y; a += 32) {
        const int id = __ldg(slots + r.x + (a - r.y));
        gl = gl || (id < nlocal && send_flag[id] != 0);
      }
      ghost = __any_sync(0xffffffffu, gl);
    }
  }
  if (lane == 0) {
    const int k = atomicAdd(counts + (ghost ? 1 : 0), 1);
    lists[(ghost ? g.ntiles : 0) + k] = t;
    lists[2 * g.ntiles + atomicAdd(counts + 2, 1)] = t;
  }
}

// send_flag[list[k]] = 1 for the local atoms of one send list
__global__ void mark_send_atoms_kernel(const int* __restrict__ list, int count, int nlocal, unsigned char* __restrict__ send_flag) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < count) {
    const int i = list[k];
    if (i < nlocal) send_flag[i] = 1;
  }
}

}  // namespace mmd

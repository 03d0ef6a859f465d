// Velocity-Verlet halves, the thermo kinetic-energy reduction, PBC wrap, and the halo
// (ghost atom) kernels.  Replaces Integrate::initialIntegrate/finalIntegrate
// (ref/integrate.cpp:46-68), Thermo::temperature's loop (ref/thermo.cpp:149-157), Atom::pbc
// (ref/atom.cpp:106-122) and Atom::pack_comm/unpack_comm/pack_reverse/unpack_reverse/
// pack_border/unpack_border (ref/atom.cpp:135-226) as used by Comm (ref/comm.cpp:276-883).
#pragma once
#include "common.cuh"
#include "scan.cuh"
#include "xs_mirror.cuh"

namespace mmd {

// ---- integrate --------------------------------------------------------------------------
// One atom per thread, one aligned vector access per array.  ZERO_F additionally clears f
// after it has been consumed (the force prologue of the half-list kernels, fused here so the
// step needs no separate memset over the local atoms).
template <class T, int ZERO_F>
__global__ void initial_integrate_kernel(Vec4<T>* __restrict__ x, Vec4<T>* __restrict__ v, Vec4<T>* __restrict__ f,
                                         int nlocal, T dt, T dtforce) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  Vec4<T> xi = x[i], vi = v[i];
  const Vec4<T> fi = f[i];
  vi.x += dtforce * fi.x;
  vi.y += dtforce * fi.y;
  vi.z += dtforce * fi.z;
  xi.x += dt * vi.x;
  xi.y += dt * vi.y;
  xi.z += dt * vi.z;
  x[i] = xi;
  v[i] = vi;
  if (ZERO_F) {
    Vec4<T> z; z.x = z.y = z.z = z.w = (T)0;
    f[i] = z;
  }
}

// KE: also accumulate sum_i (v.v)*mass of the UPDATED velocities into *ke (fused thermo).
template <class T, int KE>
__global__ void final_integrate_kernel(Vec4<T>* __restrict__ v, const Vec4<T>* __restrict__ f, int nlocal, T dtforce,
                                       T mass, double* __restrict__ ke) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (i < nlocal) {
    Vec4<T> vi = v[i];
    const Vec4<T> fi = f[i];
    vi.x += dtforce * fi.x;
    vi.y += dtforce * fi.y;
    vi.z += dtforce * fi.z;
    v[i] = vi;
    if (KE) e = (double)((vi.x * vi.x + vi.y * vi.y + vi.z * vi.z) * mass);
  }
  if (KE) {
    const double a[1] = {e};
    block_accumulate<1>(a, ke);
  }
}

// finalIntegrate of step n fused with initialIntegrate of step n+1 (they are adjacent in the time loop and
// use the same force): one pass over x, v, f instead of two.  The two velocity half-kicks stay two
// separate roundings, so the result is bit-identical to running the two kernels back to back.
template <class T, int KE>
__global__ void final_initial_integrate_kernel(Vec4<T>* __restrict__ x, Vec4<T>* __restrict__ v,
                                               const Vec4<T>* __restrict__ f, int nlocal, T dt, T dtforce, T mass,
                                               double* __restrict__ ke) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (i < nlocal) {
    Vec4<T> xi = x[i], vi = v[i];
    const Vec4<T> fi = f[i];
    vi.x += dtforce * fi.x;
    vi.y += dtforce * fi.y;
    vi.z += dtforce * fi.z;
    if (KE) e = (double)((vi.x * vi.x + vi.y * vi.y + vi.z * vi.z) * mass);
    vi.x += dtforce * fi.x;
    vi.y += dtforce * fi.y;
    vi.z += dtforce * fi.z;
    xi.x += dt * vi.x;
    xi.y += dt * vi.y;
    xi.z += dt * vi.z;
    x[i] = xi;
    v[i] = vi;
  }
  if (KE) {
    const double a[1] = {e};
    block_accumulate<1>(a, ke);
  }
}

template <class T>
__global__ void sum_mv2_kernel(const Vec4<T>* __restrict__ v, int nlocal, T mass, double* __restrict__ out) {
  double e = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlocal; i += gridDim.x * blockDim.x) {
    const Vec4<T> vi = v[i];
    e += (double)((vi.x * vi.x + vi.y * vi.y + vi.z * vi.z) * mass);
  }
  const double a[1] = {e};
  block_accumulate<1>(a, out);
}

// ---- PBC ----------------------------------------------------------------------------------
// Two sequential tests per axis, in the reference's order (ref/atom.cpp:106-122).
template <class T> __global__ void pbc_kernel(Vec4<T>* __restrict__ x, int nlocal, T xprd, T yprd, T zprd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlocal) return;
  Vec4<T> p = x[i];
  if (p.x < (T)0) p.x += xprd;
  if (p.x >= xprd) p.x -= xprd;
  if (p.y < (T)0) p.y += yprd;
  if (p.y >= yprd) p.y -= yprd;
  if (p.z < (T)0) p.z += zprd;
  if (p.z >= zprd) p.z -= zprd;
  x[i] = p;
}

// ---- halo: per-step forward / reverse ------------------------------------------------------
// A "swap pair" is the two swaps of one dimension layer (send down / send up); they read the
// same source range and write disjoint ghost ranges, so one launch covers both.
struct SwapPairDev {
  const int* list[2];  // Comm::sendlist
  int count[2];        // Comm::sendnum (self swap: == recvnum)
  int first[2];        // Comm::firstrecv
  int any[2];          // Comm::pbc_any
  int flag[2][3];      // Comm::pbc_flagx/y/z
};

// self-swap forward: x[first+k] = x[list[k]] + flag*prd, type carried along in the 4th lane
// (pack_comm + unpack_comm fused, no staging buffer).  ZERO_F also clears the ghost's force
// (half-list prologue, ref/force_lj.cpp:195-199).
template <class T, int ZERO_F>
__global__ void halo_forward_self_kernel(Vec4<T>* __restrict__ x, Vec4<T>* __restrict__ f, SwapPairDev sp, T xprd,
                                         T yprd, T zprd, XsMirror<T> M) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  int s = 0;
  if (k >= sp.count[0]) { k -= sp.count[0]; s = 1; }
  if (k >= sp.count[s]) return;
  Vec4<T> p = x[sp.list[s][k]];
  if (sp.any[s]) {
    p.x = p.x + sp.flag[s][0] * xprd;
    p.y = p.y + sp.flag[s][1] * yprd;
    p.z = p.z + sp.flag[s][2] * zprd;
  }
  x[sp.first[s] + k] = p;
  M.put_atom(sp.first[s] + k, p);
  if (ZERO_F) {
    Vec4<T> z; z.x = z.y = z.z = z.w = (T)0;
    f[sp.first[s] + k] = z;
  }
}

// ---------------------------------------------------------------------------------------
// One-launch forward halo when every swap is a self swap (single rank): each ghost is resolved, at ghost-rebuild
// time, to the LOCAL atom it ultimately copies and to the accumulated periodic shift (a ghost of a ghost made in the
// y swap from an x-swap ghost carries both shifts).  Every swap shifts one coordinate only, so the composed copy
// x[src] + shift*prd is bit-identical to the reference's six ordered swaps (ref/comm.cpp:276-317).
// shift is packed as three 2-bit fields holding flag+1.
// ---------------------------------------------------------------------------------------
__global__ void ghost_resolve_kernel(const int* __restrict__ list, int count, int first, int nlocal, int fx, int fy, int fz,
                                     int* __restrict__ src, int* __restrict__ shift) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const int s = list[k];
  const int g = first + k - nlocal;
  int sx = fx, sy = fy, sz = fz, from = s;
  if (s >= nlocal) {
    const int pk = shift[s - nlocal];
    sx += (pk & 3) - 1; sy += ((pk >> 2) & 3) - 1; sz += ((pk >> 4) & 3) - 1;
    from = src[s - nlocal];
  }
  src[g] = from;
  shift[g] = (sx + 1) | ((sy + 1) << 2) | ((sz + 1) << 4);
}
template <class T>
__global__ void halo_forward_resolved_kernel(Vec4<T>* __restrict__ x, int nlocal, int nghost, const int* __restrict__ src,
                                             const int* __restrict__ shift, T xprd, T yprd, T zprd, XsMirror<T> M) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  Vec4<T> p = x[src[g]];
  const int pk = shift[g];
  const int sx = (pk & 3) - 1, sy = ((pk >> 2) & 3) - 1, sz = ((pk >> 4) & 3) - 1;
  if (sx) p.x = p.x + sx * xprd;
  if (sy) p.y = p.y + sy * yprd;
  if (sz) p.z = p.z + sz * zprd;
  x[nlocal + g] = p;
  M.put_atom(nlocal + g, p);
}

// self-swap reverse: f[list[k]] += f[first+k]  (pack_reverse + unpack_reverse fused).
// REDG because the two swaps of a pair (and, in tiny boxes, one list) may name an atom twice.
template <class T>
__global__ void halo_reverse_self_kernel(Vec4<T>* __restrict__ f, SwapPairDev sp) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  int s = 0;
  if (k >= sp.count[0]) { k -= sp.count[0]; s = 1; }
  if (k >= sp.count[s]) return;
  const Vec4<T> g = f[sp.first[s] + k];
  red_add3(f + sp.list[s][k], g.x, g.y, g.z);
}

// remote swaps: pack to / unpack from a contiguous buffer that NCCL moves (3 reals per atom,
// the reference's wire format, ref/atom.cpp:135-195)
template <class T>
__global__ void halo_pack_x_kernel(const Vec4<T>* __restrict__ x, const int* __restrict__ list, int n, int any, int fx,
                                   int fy, int fz, T xprd, T yprd, T zprd, T* __restrict__ buf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  Vec4<T> p = x[list[k]];
  if (any) {
    p.x = p.x + fx * xprd;
    p.y = p.y + fy * yprd;
    p.z = p.z + fz * zprd;
  }
  buf[3 * k + 0] = p.x;
  buf[3 * k + 1] = p.y;
  buf[3 * k + 2] = p.z;
}
template <class T, int ZERO_F>
__global__ void halo_unpack_x_kernel(Vec4<T>* __restrict__ x, Vec4<T>* __restrict__ f, int first, int n,
                                     const T* __restrict__ buf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  Vec4<T> p = x[first + k];  // keeps the type lane set by borders
  p.x = buf[3 * k + 0];
  p.y = buf[3 * k + 1];
  p.z = buf[3 * k + 2];
  x[first + k] = p;
  if (ZERO_F) {
    Vec4<T> z; z.x = z.y = z.z = z.w = (T)0;
    f[first + k] = z;
  }
}
template <class T>
__global__ void halo_pack_f_kernel(const Vec4<T>* __restrict__ f, int first, int n, T* __restrict__ buf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const Vec4<T> g = f[first + k];
  buf[3 * k + 0] = g.x;
  buf[3 * k + 1] = g.y;
  buf[3 * k + 2] = g.z;
}
template <class T>
__global__ void halo_unpack_f_kernel(Vec4<T>* __restrict__ f, const int* __restrict__ list, int n,
                                     const T* __restrict__ buf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  red_add3(f + list[k], buf[3 * k + 0], buf[3 * k + 1], buf[3 * k + 2]);
}

// both swaps of one dimension layer in one launch (they are independent: same source range, disjoint
// ghost ranges), so a remote dimension costs pack -> ONE NCCL group (2 sends + 2 recvs) -> unpack.
// Buffer layout: [3*count0 reals of swap 0 | 3*count1 reals of swap 1].
template <class T>
__global__ void halo_pack_x_pair_kernel(const Vec4<T>* __restrict__ x, SwapPairDev sp, T xprd, T yprd, T zprd,
                                        T* __restrict__ buf) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  int k = g, s = 0;
  if (k >= sp.count[0]) { k -= sp.count[0]; s = 1; }
  if (k >= sp.count[s]) return;
  Vec4<T> p = x[sp.list[s][k]];
  if (sp.any[s]) {
    p.x = p.x + sp.flag[s][0] * xprd;
    p.y = p.y + sp.flag[s][1] * yprd;
    p.z = p.z + sp.flag[s][2] * zprd;
  }
  buf[(size_t)3 * g + 0] = p.x;
  buf[(size_t)3 * g + 1] = p.y;
  buf[(size_t)3 * g + 2] = p.z;
}
// ---------------------------------------------------------------------------------------
// Forward halo over peer memory (NVLink / NVSwitch), one rank per GPU.
// Every rank owns a receive window that its neighbors map through CUDA IPC.  The sender's kernel packs the
// positions of both swaps of a dimension (Atom::pack_comm, ref/atom.cpp:135-151) and stores them straight into the
// two neighbors' windows; the last block to finish publishes the call's epoch in the neighbors' flag words
// (system-scope release).  The receiver's unpack kernel spins on its own flag words (system-scope acquire) and
// then scatters the window into the ghost slots (Atom::unpack_comm, :153-170).  No host round trip, no staging
// copy, no NCCL launch: two short kernels per dimension.  Windows are double-buffered by epoch parity; a sender
// can reuse a buffer only two calls later, after it has itself consumed the receiver's next message, which the
// receiver sent after finishing the unpack of the buffer in question (stream order on both sides).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

template <class T>
__global__ void halo_p2p_send_kernel(const Vec4<T>* __restrict__ x, SwapPairDev sp, T xprd, T yprd, T zprd,
                                     T* __restrict__ dst0, T* __restrict__ dst1, unsigned long long* flag0,
                                     unsigned long long* flag1, unsigned long long epoch, unsigned* __restrict__ done) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  int k = g, s = 0;
  if (k >= sp.count[0]) { k -= sp.count[0]; s = 1; }
  if (k < sp.count[s]) {
    Vec4<T> p = x[sp.list[s][k]];
    if (sp.any[s]) {
      p.x = p.x + sp.flag[s][0] * xprd;
      p.y = p.y + sp.flag[s][1] * yprd;
      p.z = p.z + sp.flag[s][2] * zprd;
    }
    T* d = (s ? dst1 : dst0) + (size_t)3 * k;
    d[0] = p.x;
    d[1] = p.y;
    d[2] = p.z;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(done, 1u);
    if (t == gridDim.x - 1) {
      *done = 0u;
      __threadfence_system();
      st_release_sys(flag0, epoch);
      st_release_sys(flag1, epoch);
    }
  }
}

// status |= 4 when a neighbor's message did not arrive within ~2 s (a rank died or the ranks diverged)
template <class T, int ZERO_F>
__global__ void halo_p2p_unpack_kernel(Vec4<T>* __restrict__ x, Vec4<T>* __restrict__ f, int first0, int n0, int first1,
                                       int n1, const T* __restrict__ src0, const T* __restrict__ src1,
                                       const unsigned long long* flag0, const unsigned long long* flag1,
                                       unsigned long long epoch, int* __restrict__ status, XsMirror<T> M) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag0) < epoch || ld_acquire_sys(flag1) < epoch) {
      if (clock64() - t0 > 4000000000ll) {
        atomicOr(status, 4);
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n0 + n1) return;
  const int dst = g < n0 ? first0 + g : first1 + (g - n0);
  const T* b = g < n0 ? src0 + (size_t)3 * g : src1 + (size_t)3 * (g - n0);
  Vec4<T> p = x[dst];  // keeps the type lane set by borders
  p.x = __ldcg(b + 0);  // the window is written by another GPU: read it from L2, never from a stale L1 line
  p.y = __ldcg(b + 1);
  p.z = __ldcg(b + 2);
  x[dst] = p;
  M.put_atom(dst, p);
  if (ZERO_F) {
    Vec4<T> z; z.x = z.y = z.z = z.w = (T)0;
    f[dst] = z;
  }
}

template <class T, int ZERO_F>
__global__ void halo_unpack_x_pair_kernel(Vec4<T>* __restrict__ x, Vec4<T>* __restrict__ f, int first0, int n0, int first1,
                                          int n1, const T* __restrict__ buf, XsMirror<T> M) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n0 + n1) return;
  const int dst = g < n0 ? first0 + g : first1 + (g - n0);
  Vec4<T> p = x[dst];  // keeps the type lane set by borders
  p.x = buf[(size_t)3 * g + 0];
  p.y = buf[(size_t)3 * g + 1];
  p.z = buf[(size_t)3 * g + 2];
  x[dst] = p;
  M.put_atom(dst, p);
  if (ZERO_F) {
    Vec4<T> z; z.x = z.y = z.z = z.w = (T)0;
    f[dst] = z;
  }
}
template <class T>
__global__ void halo_pack_f_pair_kernel(const Vec4<T>* __restrict__ f, int first0, int n0, int first1, int n1,
                                        T* __restrict__ buf) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n0 + n1) return;
  const Vec4<T> v = f[g < n0 ? first0 + g : first1 + (g - n0)];
  buf[(size_t)3 * g + 0] = v.x;
  buf[(size_t)3 * g + 1] = v.y;
  buf[(size_t)3 * g + 2] = v.z;
}
// buf: [3*count0 | 3*count1] force contributions for the atoms of send list 0 / 1
template <class T>
__global__ void halo_unpack_f_pair_kernel(Vec4<T>* __restrict__ f, SwapPairDev sp, const T* __restrict__ buf) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  int k = g, s = 0;
  if (k >= sp.count[0]) { k -= sp.count[0]; s = 1; }
  if (k >= sp.count[s]) return;
  red_add3(f + sp.list[s][k], buf[(size_t)3 * g + 0], buf[(size_t)3 * g + 1], buf[(size_t)3 * g + 2]);
}
template <class T>
__global__ void gather_scalar_pair_kernel(const T* __restrict__ a, SwapPairDev sp, T* __restrict__ buf) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  int k = g, s = 0;
  if (k >= sp.count[0]) { k -= sp.count[0]; s = 1; }
  if (k >= sp.count[s]) return;
  buf[g] = a[sp.list[s][k]];
}

// scalar per-atom forward halo (EAM fp, ForceEAM::communicate ref/force_eam.cpp:851-914)
template <class T>
__global__ void halo_forward_scalar_self_kernel(T* __restrict__ a, SwapPairDev sp) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  int s = 0;
  if (k >= sp.count[0]) { k -= sp.count[0]; s = 1; }
  if (k >= sp.count[s]) return;
  a[sp.first[s] + k] = a[sp.list[s][k]];
}
template <class T>
__global__ void gather_scalar_kernel(const T* __restrict__ a, const int* __restrict__ list, int n, T* __restrict__ buf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) buf[k] = a[list[k]];
}

// ---- borders: ordered slab selection --------------------------------------------------------
// Atoms i in [nfirst,nlast) with lo <= x[dim] <= hi (both inclusive, ref/comm.cpp:776) are
// appended to the send list IN INDEX ORDER (the reference's order), for both swaps of a pair at
// once.  Three phases share the generic scan spine: per-tile counts, spine, ordered scatter.
constexpr int BORDER_THREADS = 256;

template <class T>
__device__ __forceinline__ T coord_of(const Vec4<T>& p, int dim) { return dim == 0 ? p.x : (dim == 1 ? p.y : p.z); }

template <class T>
__global__ void border_count_kernel(const Vec4<T>* __restrict__ x, int nfirst, int nlast, int dim, T lo0, T hi0, T lo1,
                                    T hi1, int* __restrict__ tile_counts0, int* __restrict__ tile_counts1) {
  const int i = nfirst + blockIdx.x * BORDER_THREADS + threadIdx.x;
  int m0 = 0, m1 = 0;
  if (i < nlast) {
    const T c = coord_of(x[i], dim);
    m0 = (c >= lo0 && c <= hi0);
    m1 = (c >= lo1 && c <= hi1);
  }
  int t0, t1;
  block_exclusive_scan(m0, &t0);
  block_exclusive_scan(m1, &t1);
  if (threadIdx.x == 0) {
    tile_counts0[blockIdx.x] = t0;
    tile_counts1[blockIdx.x] = t1;
  }
}

// writes list entries; for self swaps also materialises the ghost (pack_border+unpack_border).
// first1 (start of the second swap's ghosts) = first0 + total0 is passed by the host after it
// has read the totals (it needs them anyway to grow the arrays).
template <class T>
__global__ void border_scatter_kernel(Vec4<T>* __restrict__ x, int nfirst, int nlast, int dim, T lo0, T hi0, T lo1,
                                      T hi1, const int* __restrict__ tile_off0, const int* __restrict__ tile_off1,
                                      int* __restrict__ list0, int* __restrict__ list1, int self0, int self1, int first0,
                                      int first1, SwapPairDev sp, T xprd, T yprd, T zprd, T* __restrict__ buf0,
                                      T* __restrict__ buf1) {
  const int i = nfirst + blockIdx.x * BORDER_THREADS + threadIdx.x;
  int m0 = 0, m1 = 0;
  Vec4<T> p;
  p.x = p.y = p.z = p.w = (T)0;
  if (i < nlast) {
    p = x[i];
    const T c = coord_of(p, dim);
    m0 = (c >= lo0 && c <= hi0);
    m1 = (c >= lo1 && c <= hi1);
  }
  int t0, t1;
  const int e0 = block_exclusive_scan(m0, &t0) + tile_off0[blockIdx.x];
  const int e1 = block_exclusive_scan(m1, &t1) + tile_off1[blockIdx.x];
  const T prd[3] = {xprd, yprd, zprd};
  if (m0) {
    list0[e0] = i;
    Vec4<T> q = p;
    if (sp.any[0]) { q.x = q.x + sp.flag[0][0] * prd[0]; q.y = q.y + sp.flag[0][1] * prd[1]; q.z = q.z + sp.flag[0][2] * prd[2]; }
    if (self0) x[first0 + e0] = q;
    else { buf0[4 * e0 + 0] = q.x; buf0[4 * e0 + 1] = q.y; buf0[4 * e0 + 2] = q.z; buf0[4 * e0 + 3] = (T)lane_to_type(q.w); }
  }
  if (m1) {
    list1[e1] = i;
    Vec4<T> q = p;
    if (sp.any[1]) { q.x = q.x + sp.flag[1][0] * prd[0]; q.y = q.y + sp.flag[1][1] * prd[1]; q.z = q.z + sp.flag[1][2] * prd[2]; }
    if (self1) x[first1 + e1] = q;
    else { buf1[4 * e1 + 0] = q.x; buf1[4 * e1 + 1] = q.y; buf1[4 * e1 + 2] = q.z; buf1[4 * e1 + 3] = (T)lane_to_type(q.w); }
  }
}

// unpack_border for remote swaps: 4 reals per atom (x,y,z,type), ref/atom.cpp:215-226
template <class T>
__global__ void border_unpack_kernel(Vec4<T>* __restrict__ x, int first, int n, const T* __restrict__ buf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  Vec4<T> p;
  p.x = buf[4 * k + 0];
  p.y = buf[4 * k + 1];
  p.z = buf[4 * k + 2];
  p.w = type_to_lane<T>((int)buf[4 * k + 3]);
  x[first + k] = p;
}

// ---- exchange: atom migration between sub-boxes (Comm::exchange, ref/comm.cpp:364-597) -----------
// Leavers (x[dim] outside [lo,hi)) are packed in index order, 7 reals per atom (x,y,z,vx,vy,vz,type:
// Atom::pack_exchange, ref/atom.cpp:228-240); the holes they leave below the new nlocal are filled, in
// ascending order, with the surviving atoms of the tail -- the reference's hole-filling rule
// (ref/comm.cpp:470-488).  Arrivals that fall inside my sub-box are appended in arrival order.
template <class T>
__global__ void exch_flag_kernel(const Vec4<T>* __restrict__ x, int n, int dim, T lo, T hi, int* __restrict__ leaves) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T c = coord_of(x[i], dim);
  leaves[i] = (c < lo || c >= hi) ? 1 : 0;
}

// pos = exclusive scan of leaves.  nkeep = n - (number of leavers).
template <class T>
__global__ void exch_pack_kernel(const Vec4<T>* __restrict__ x, const Vec4<T>* __restrict__ v, int n,
                                 const int* __restrict__ leaves, const int* __restrict__ pos, int nkeep,
                                 T* __restrict__ buf, int* __restrict__ hole_idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !leaves[i]) return;
  const int r = pos[i];
  const Vec4<T> xi = x[i], vi = v[i];
  T* b = buf + (size_t)7 * r;
  b[0] = xi.x; b[1] = xi.y; b[2] = xi.z;
  b[3] = vi.x; b[4] = vi.y; b[5] = vi.z;
  b[6] = (T)lane_to_type(xi.w);
  if (i < nkeep) hole_idx[r] = i;  // every leaver before a hole is itself below nkeep, so r is the hole's rank
}

// survivors in [nkeep, n) move down into the holes; pos[nkeep] = number of holes
template <class T>
__global__ void exch_fill_kernel(Vec4<T>* __restrict__ x, Vec4<T>* __restrict__ v, int n, const int* __restrict__ leaves,
                                 const int* __restrict__ pos, int nkeep, const int* __restrict__ hole_idx) {
  const int i = nkeep + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || leaves[i]) return;
  const int nholes = pos[nkeep];
  const int t = (i - nkeep) - (pos[i] - nholes);  // rank among the surviving tail atoms
  const int dst = hole_idx[t];
  x[dst] = x[i];
  v[dst] = v[i];
}

template <class T>
__global__ void exch_recv_flag_kernel(const T* __restrict__ buf, int n, int dim, T lo, T hi, int* __restrict__ mine) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T c = buf[(size_t)7 * i + dim];
  mine[i] = (c >= lo && c < hi) ? 1 : 0;
}

// Atom::unpack_exchange (ref/atom.cpp:242-254)
template <class T>
__global__ void exch_unpack_kernel(const T* __restrict__ buf, int n, const int* __restrict__ mine,
                                   const int* __restrict__ pos, int first, Vec4<T>* __restrict__ x,
                                   Vec4<T>* __restrict__ v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !mine[i]) return;
  const T* b = buf + (size_t)7 * i;
  Vec4<T> xi, vi;
  xi.x = b[0]; xi.y = b[1]; xi.z = b[2]; xi.w = type_to_lane<T>((int)b[6]);
  vi.x = b[3]; vi.y = b[4]; vi.z = b[5]; vi.w = (T)0;
  x[first + pos[i]] = xi;
  v[first + pos[i]] = vi;
}

}  // namespace mmd

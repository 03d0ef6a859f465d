// Slot-ordered mirror of the positions: the same coordinates as the atom array x[], laid out in the order of the tile
// kernels' slot map (tile_kernels.cuh: CSR bins, each bin sorted by x), in exactly the record format the force kernel
// keeps in shared memory.  Every (y,z) pencil run of a halo window is then ONE contiguous range of records, so a CTA
// stages its window with bulk asynchronous copies (cp.async.bulk / cp.async) instead of an indexed gather
// (slots[] -> x[id] -> st.shared), see force_lj_dealt_kernel.
//   FP64: rec[slot] = (x, y) 16-byte records, z[slot] 8-byte;   FP32: rec[slot] = (x, y, z, type bits)
// The mirror is rebuilt with the neighbor list (xs_fill_kernel) and kept current by every kernel that moves atoms between
// two rebuilds: the force kernel's Verlet epilogue (local atoms), the forward-halo kernels (ghosts) and, for the unfused
// integrators, xs_refresh_kernel.  slot_of[atom] is the inverse of the slot map.
#pragma once
#include "common.cuh"

namespace mmd {

template <class T> struct QRec;
template <> struct alignas(16) QRec<double> { double x, y; };
template <> struct alignas(16) QRec<float> { float x, y, z, w; };

template <class T> struct XsMirror {
  QRec<T>* rec;        // nullptr: no mirror to maintain
  T* z;                // FP64 only
  const int* slot_of;  // atom -> slot
  __device__ __forceinline__ void put_slot(int s, const Vec4<T>& p) const {
    if constexpr (sizeof(T) == 8) {
      QRec<T> r; r.x = p.x; r.y = p.y;
      rec[s] = r;
      z[s] = p.z;
    } else {
      QRec<T> r; r.x = p.x; r.y = p.y; r.z = p.z; r.w = p.w;
      rec[s] = r;
    }
  }
  __device__ __forceinline__ void put_atom(int atom, const Vec4<T>& p) const {
    if (rec) put_slot(__ldg(slot_of + atom), p);
  }
};

// mirror of all binned atoms + inverse slot map + slot-ordered types (per-type parameter tables, FP64)
template <class T>
__global__ void xs_fill_kernel(const Vec4<T>* __restrict__ x, const int* __restrict__ slots, int n, XsMirror<T> M,
                               int* __restrict__ slot_of, unsigned char* __restrict__ types) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int id = slots[s];
  const Vec4<T> p = ldg4(x + id);
  M.put_slot(s, p);
  if (slot_of) slot_of[id] = s;
  if (types) types[s] = (unsigned char)lane_to_type(p.w);
}

// atoms [first, first + count) moved outside the fused kernels: copy them into the mirror
template <class T>
__global__ void xs_refresh_kernel(const Vec4<T>* __restrict__ x, int first, int count, XsMirror<T> M) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  M.put_atom(first + k, x[first + k]);
}

// ---------------------------------------------------------------------------------------
// Ghost images (single rank, every swap a self swap): ghost g is a copy of local atom src[g] shifted by whole box
// lengths (ghost_resolve_kernel).  Inverted into a CSR list per local atom, the Verlet epilogue of the force kernel
// writes an atom's ghost copies together with its new position -- the per-step forward halo
// (Comm::communicate, ref/comm.cpp:276-317) needs no launch of its own.  Same arithmetic as
// halo_forward_resolved_kernel (x + s * prd per shifted coordinate): bit-identical ghosts.
// ---------------------------------------------------------------------------------------
template <class T> struct GhostImages {
  const int* start;   // [nlocal + 1]; nullptr: no images to write
  const int2* list;   // {ghost index (0-based behind the local atoms), packed shift: 2 bits per axis holding s + 1}
  T prd[3];
  int nlocal;
};
__global__ void ghost_image_count_kernel(const int* __restrict__ src, int nghost, int* __restrict__ count) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < nghost) atomicAdd(count + src[g], 1);
}
__global__ void ghost_image_fill_kernel(const int* __restrict__ src, const int* __restrict__ shift, int nghost,
                                        const int* __restrict__ start, int* __restrict__ cursor, int2* __restrict__ list) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nghost) return;
  const int s = src[g];
  list[start[s] + atomicAdd(cursor + s, 1)] = make_int2(g, shift[g]);
}

}  // namespace mmd

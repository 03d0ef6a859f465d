// Exclusive prefix sum over int arrays (bin counts, border flags), hand-written:
// tile sums -> single-block spine scan -> tile apply.  No library (cub/thrust) calls.
#pragma once
#include "common.cuh"

namespace mmd {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048 ints per block

__device__ __forceinline__ int warp_inclusive_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive scan of one int per thread across a block of up to 1024 threads; returns the
// exclusive prefix, *block_total gets the sum (valid in every thread).
__device__ __forceinline__ int block_exclusive_scan(int v, int* block_total) {
  __shared__ int warp_sums[32];
  __shared__ int total_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  int inc = warp_inclusive_scan(v, lane);
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarp ? warp_sums[lane] : 0;
    int winc = warp_inclusive_scan(w, lane);
    if (lane < nwarp) warp_sums[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) total_s = winc;
  }
  __syncthreads();
  int r = inc - v + warp_sums[warp];
  *block_total = total_s;
  __syncthreads();  // smem reusable by the next call
  return r;
}

// phase 1: per-tile sum (+ optional running maximum of the elements)
__global__ void scan_tile_sums_kernel(const int* __restrict__ in, int n, int* __restrict__ tile_sums,
                                      int* __restrict__ max_elem) {
  const int base = blockIdx.x * SCAN_TILE;
  int s = 0, m = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    int idx = base + k * SCAN_THREADS + threadIdx.x;
    int v = idx < n ? in[idx] : 0;
    s += v;
    m = max(m, v);
  }
  int total;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
  if (max_elem) {
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 16));
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 8));
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 4));
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(max_elem, m);
  }
}

// phase 2: one block turns tile sums into exclusive tile offsets (any tile count) and
// writes the grand total to *total_out.
// blockIdx.x selects one of several independent arrays laid out `stride` ints apart.
__global__ void scan_spine_kernel(int* __restrict__ tile_sums_all, int ntiles, int* __restrict__ total_out_all,
                                  int stride) {
  int* __restrict__ tile_sums = tile_sums_all + (size_t)blockIdx.x * stride;
  int* __restrict__ total_out = total_out_all + blockIdx.x;
  int carry = 0;
  for (int base = 0; base < ntiles; base += blockDim.x) {
    int idx = base + threadIdx.x;
    int v = idx < ntiles ? tile_sums[idx] : 0;
    int total;
    int ex = block_exclusive_scan(v, &total);
    if (idx < ntiles) tile_sums[idx] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

// phase 3: exclusive scan inside each tile + tile offset. out may alias in.
__global__ void scan_apply_kernel(const int* __restrict__ in, int n, const int* __restrict__ tile_offsets,
                                  int* __restrict__ out) {
  // blocked arrangement: thread t owns SCAN_ITEMS consecutive elements
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    int idx = base + k;
    v[k] = idx < n ? in[idx] : 0;
    s += v[k];
  }
  int total;
  int ex = block_exclusive_scan(s, &total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    int idx = base + k;
    if (idx < n) out[idx] = ex;
    ex += v[k];
  }
}

}  // namespace mmd

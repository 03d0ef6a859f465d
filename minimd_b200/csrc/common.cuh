// Shared device/host helpers for the miniMD B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mmd {

// ---------------------------------------------------------------------------------------
// Device atom record: position (or velocity / force) padded to 4 lanes so that one atom is
// exactly one aligned vector access: FP64 -> 32 B = one DRAM/L2 sector, moved by a single
// 256-bit LDG.E.ENL2.256 / STG.E.ENL2.256 on sm_100a; FP32 -> 16 B (LDG.E.128).
// For positions the 4th lane carries the atom type (integer bits), so the pair loop needs
// one gather per neighbor instead of two (ref gathers x[j*PAD..] and type[j] separately,
// ref/force_lj.cpp:223-227).  This is the reference's own -DPAD4 idea (ref/types.h:77-81).
// ---------------------------------------------------------------------------------------
template <class T> struct Vec4;
template <> struct alignas(32) Vec4<double> { double x, y, z, w; };
template <> struct alignas(16) Vec4<float> { float x, y, z, w; };

template <class T> __host__ __device__ __forceinline__ T type_to_lane(int t);
template <> __host__ __device__ __forceinline__ double type_to_lane<double>(int t) {
  union { long long i; double d; } u; u.i = (long long)t; return u.d;
}
template <> __host__ __device__ __forceinline__ float type_to_lane<float>(int t) {
  union { int i; float f; } u; u.i = t; return u.f;
}
__host__ __device__ __forceinline__ int lane_to_type(double w) {
  union { long long i; double d; } u; u.d = w; return (int)u.i;
}
__host__ __device__ __forceinline__ int lane_to_type(float w) {
  union { int i; float f; } u; u.f = w; return u.i;
}

// read-only gather through the non-coherent path
template <class T> __device__ __forceinline__ Vec4<T> ldg4(const Vec4<T>* p) { return __ldg(p); }
template <> __device__ __forceinline__ Vec4<double> ldg4<double>(const Vec4<double>* p) {
  Vec4<double> r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
template <> __device__ __forceinline__ Vec4<float> ldg4<float>(const Vec4<float>* p) {
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  Vec4<float> r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
}

// fire-and-forget force scatter: f[j].xyz -= / += (no return value => REDG, not ATOMG)
__device__ __forceinline__ void red_add3(Vec4<double>* p, double a, double b, double c) {
  double* q = reinterpret_cast<double*>(p);
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(q), "d"(a) : "memory");
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(q + 1), "d"(b) : "memory");
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(q + 2), "d"(c) : "memory");
}
__device__ __forceinline__ void red_add3(Vec4<float>* p, float a, float b, float c) {
  // one 128-bit vector reduction (REDG.E.ADD.F32x4) instead of three scalar ones
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(0.0f) : "memory");
}
__device__ __forceinline__ void red_add1(double* p, double a) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(a) : "memory");
}
__device__ __forceinline__ void red_add1(float* p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

// ---------------------------------------------------------------------------------------
// Warp-cooperative force scatter for FP64 (half neighbor lists).
// Measured on B200 (tools/microbench/red_bench.cu): RED.E.ADD.F64 throughput is ~222 G
// instruction-lanes/s when every lane of an instruction targets its own 32-byte sector, but lanes of
// ONE instruction that fall into the same sector are merged: emitting x,y,z of an atom from three
// adjacent lanes of a single RED moves 2.3x more pair updates per second than three separate
// REDs per lane.  So the 32 (fx,fy,fz,j) records of a warp are transposed through shared memory
// into 96 (address, value) elements and emitted as 3 RED instructions of 32 lanes, 3 lanes per atom.
// Must be called by all 32 lanes (warp-uniform control flow); `on` selects the lanes that contribute.
// sv: 96 doubles, sj: 32 ints of shared memory private to the calling warp.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_scatter3(Vec4<double>* __restrict__ f, double* sv, int* sj, int lane, bool on,
                                              int j, double a, double b, double c) {
  const unsigned mask = __ballot_sync(0xffffffffu, on);
  if (mask == 0u) return;
  if (on) {
    sv[3 * lane + 0] = a;
    sv[3 * lane + 1] = b;
    sv[3 * lane + 2] = c;
    sj[lane] = j;
  }
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const int e = 32 * r + lane;
    const int p = e / 3;
    const int comp = e - 3 * p;
    if ((mask >> p) & 1u) red_add1(reinterpret_cast<double*>(f + sj[p]) + comp, sv[e]);
  }
  __syncwarp();
}
// FP32: one 128-bit vector reduction per pair already covers the whole atom record
__device__ __forceinline__ void warp_scatter3(Vec4<float>* __restrict__ f, float*, int*, int, bool on, int j, float a,
                                              float b, float c) {
  if (on) red_add3(f + j, a, b, c);
}

// ---------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------
template <int WIDTH, class T> __device__ __forceinline__ T group_sum(T v) {
#pragma unroll
  for (int o = WIDTH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 32);
  return v;
}

// Block-wide sum of NV doubles per thread; result added to out[0..NV) with one REDG per
// value per block.  Must be called by all threads of the block.
template <int NV> __device__ __forceinline__ void block_accumulate(const double (&val)[NV], double* out) {
  __shared__ double red_smem[NV][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  double v[NV];
#pragma unroll
  for (int k = 0; k < NV; k++) v[k] = group_sum<32>(val[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) red_smem[k][warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      double s = lane < nwarp ? red_smem[k][lane] : 0.0;
      s = group_sum<32>(s);
      if (lane == 0) red_add1(out + k, s);
    }
  }
}

__host__ __device__ __forceinline__ int div_up(long long a, int b) { return (int)((a + b - 1) / b); }

}  // namespace mmd

// Microbenchmark: throughput of FP64 global reductions (RED.E.ADD.F64) on sm_100a for the access
// shapes the half-list force kernel can produce.  Answers: is the cost per LANE or per 32-byte SECTOR?
//   A  scatter3      : each lane owns one pair, issues 3 REDs (x,y,z of atom j)          [current kernel]
//   B  triple-lanes  : 3 adjacent lanes carry x,y,z of ONE atom in a single RED instruction
//   C  one-comp      : each lane 1 RED (x only): lane-op rate reference
//   D  f32x4         : one red.v4.f32 per lane (FP32 kernel shape)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bench red_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

struct alignas(32) D4 { double x, y, z, w; };

__device__ __forceinline__ unsigned hash(unsigned a) {
  a ^= a >> 16; a *= 0x7feb352du; a ^= a >> 15; a *= 0x846ca68bu; a ^= a >> 16; return a;
}
__device__ __forceinline__ void red(double* p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

// j index: near i (window) to mimic bin-sorted locality, or fully random
__device__ __forceinline__ int pick(int i, int k, int n, int window) {
  unsigned h = hash(i * 977u + k * 7919u);
  if (window <= 0) return h % n;
  int j = i + (int)(h % (2 * window)) - window;
  return j < 0 ? j + n : (j >= n ? j - n : j);
}

__global__ void kA(D4* f, int n, int per, int window) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < per; k++) {
    int j = pick(i, k, n, window);
    double* q = &f[j].x;
    red(q, 1.0); red(q + 1, 2.0); red(q + 2, 3.0);
  }
}
__global__ void kB(D4* f, int n, int per, int window) {
  // thread t: pair = t/3 ... uses 3x the threads, each doing one component; lanes 30,31 of a warp idle
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int warp = t >> 5, lane = t & 31;
  if (lane >= 30) return;
  int i = warp * 10 + lane / 3, c = lane % 3;
  if (i >= n) return;
  for (int k = 0; k < per; k++) {
    int j = pick(i, k, n, window);
    red(&f[j].x + c, 1.0 + c);
  }
}
__global__ void kC(D4* f, int n, int per, int window) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < per; k++) {
    int j = pick(i, k, n, window);
    red(&f[j].x, 1.0);
  }
}
__global__ void kD(float4* f, int n, int per, int window) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < per; k++) {
    int j = pick(i, k, n, window);
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(f + j), "f"(1.f), "f"(2.f), "f"(3.f), "f"(0.f) : "memory");
  }
}

template <class F> float timeit(F fn) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  fn(); cudaDeviceSynchronize();
  cudaEventRecord(a); fn(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  const int n = 2048000, per = 28;  // 28 in-cutoff pairs per atom, as in the LJ half list
  D4* f; cudaMalloc(&f, sizeof(D4) * n); cudaMemset(f, 0, sizeof(D4) * n);
  float4* g; cudaMalloc(&g, sizeof(float4) * n); cudaMemset(g, 0, sizeof(float4) * n);
  for (int window : {2000, 60000, 0}) {
    float a = timeit([&] { kA<<<(n + 255) / 256, 256>>>(f, n, per, window); });
    int tB = ((n + 9) / 10) * 32;
    float b = timeit([&] { kB<<<(tB + 255) / 256, 256>>>(f, n, per, window); });
    float c = timeit([&] { kC<<<(n + 255) / 256, 256>>>(f, n, per, window); });
    float d = timeit([&] { kD<<<(n + 255) / 256, 256>>>(g, n, per, window); });
    double pairs = (double)n * per;
    printf("window %6d: A scatter3 %.3f ms (%.1f Gpair/s, %.1f Glane-op/s) | B triple-lanes %.3f ms (%.1f Gpair/s) | "
           "C one-comp %.3f ms (%.1f Glane-op/s) | D f32x4 %.3f ms (%.1f Gpair/s)\n",
           window, a, pairs / a / 1e6, 3 * pairs / a / 1e6, b, pairs / b / 1e6, c, pairs / c / 1e6, d, pairs / d / 1e6);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}

// Microbenchmark: the second roof of the pair-force kernels -- vector FP64 throughput of one B200 (SURVEY.md 8d asks for
// a measured FP64 FMA peak before any FP64 fraction is quoted; MEASURED_PEAKS.json has only HBM and bf16 tensor numbers).
//   fma    : 8 independent DFMA chains per thread (issue-limited peak of the FP64 pipe)
//   rcp    : the pair loop's reciprocal (MUFU.RCP64H + 3 DFMA, tile_dealt_kernels.cuh pair_rcp) -- chains per thread as above
//   pair   : the LJ pair evaluation exactly as lj_pair<double,0,1> does it, operands in registers (no memory):
//            pair evaluations per second the FP64 pipe can sustain, the compute floor of force_lj_dealt_kernel
// Also checks pair_rcp against IEEE division over the r^2 range of the force loop (max relative error printed).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_fma_bench fp64_fma_bench.cu
// Output: one JSON line (kept in profiles/r2_fp64_peak.json).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ double pair_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  const double e2 = fma(e, e, e);
  return fma(y, e2, y);
}

__global__ void k_fma(double* out, int iters, double a, double b) {
  double c[8];
#pragma unroll
  for (int k = 0; k < 8; k++) c[k] = threadIdx.x * 1e-3 + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = fma(c[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += c[k];
  if (s == 1.2345e300) out[0] = s;
}

__global__ void k_rcp(double* out, int iters, double a) {
  double c[8];
#pragma unroll
  for (int k = 0; k < 8; k++) c[k] = 1.0 + threadIdx.x * 1e-3 + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) c[k] = pair_rcp(c[k]) + a;
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += c[k];
  if (s == 1.2345e300) out[0] = s;
}

// 4 pairs in flight per thread, as in the kernel's unrolled row word
__global__ void k_pair(double* out, int iters, double cut, double s6, double k48) {
  double xi = 0.1 * threadIdx.x, yi = 0.2, zi = 0.3;
  double xj[4], yj[4], zj[4];
#pragma unroll
  for (int e = 0; e < 4; e++) { xj[e] = xi + 0.9 + 0.1 * e; yj[e] = yi + 0.5 * e; zj[e] = zi - 0.4 * e; }
  double fx = 0, fy = 0, fz = 0;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const double dx = xi - xj[e], dy = yi - yj[e], dz = zi - zj[e];
      const double rsq = dx * dx + dy * dy + dz * dz;
      const bool hit = rsq < cut;
      const double a1 = pair_rcp(rsq);
      const double a2 = a1 * a1;
      const double a3 = a2 * a1;
      const double force = (a2 * a2) * (a3 * s6 - 0.5) * k48;
      if (hit) { fx += dx * force; fy += dy * force; fz += dz * force; }
      xj[e] += 1e-9 * force;  // keeps the loop body live without adding FP64 work of note
    }
  }
  if (fx + fy + fz == 1.2345e300) out[0] = fx;
}

__global__ void k_rcp_error(double lo, double hi, int n, double* maxerr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = lo * pow(hi / lo, (double)i / (double)(n - 1));
  const double ref = 1.0 / x;
  const double err = fabs(pair_rcp(x) - ref) / ref;
  unsigned long long* p = reinterpret_cast<unsigned long long*>(maxerr);
  atomicMax(p, (unsigned long long)__double_as_longlong(err));  // non-negative doubles order like their bit patterns
}

template <class F> static double time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();  // warm-up
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double* out;
  cudaMalloc(&out, 64);
  cudaMemset(out, 0, 64);
  const int threads = 512, blocks = sms * 4, iters = 4096;
  const double lanes = (double)threads * blocks;
  const double t_fma = time_ms([&] { k_fma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
  const double t_rcp = time_ms([&] { k_rcp<<<blocks, threads>>>(out, iters, 1.0); });
  const double t_pair = time_ms([&] { k_pair<<<blocks, threads>>>(out, iters, 6.25, 1.0, 48.0); });
  const double fma_per_s = lanes * iters * 8 / (t_fma * 1e-3);
  const double rcp_per_s = lanes * iters * 8 / (t_rcp * 1e-3);
  const double pair_per_s = lanes * iters * 4 / (t_pair * 1e-3);
  double* d_err = out + 1;
  k_rcp_error<<<4096, 256>>>(0.5, 8.0, 4096 * 256, d_err);
  double h_err = 0;
  cudaMemcpy(&h_err, d_err, sizeof(double), cudaMemcpyDeviceToHost);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_khz\": %d, \"fp64_fma_per_s\": %.4e, \"fp64_tflops\": %.2f, "
         "\"fma_per_clk_per_sm\": %.1f, \"pair_rcp_per_s\": %.4e, \"lj_pairs_per_s\": %.4e, \"fp64_ops_per_pair\": 19, "
         "\"pair_rcp_max_rel_err\": %.3e, \"how\": \"tools/microbench/fp64_fma_bench.cu: 8 independent DFMA chains per "
         "thread, %d blocks x %d threads, best of 5 (CUDA events)\"}\n",
         prop.name, sms, clk, fma_per_s, 2 * fma_per_s / 1e12, fma_per_s / (sms * (clk * 1e3)), rcp_per_s, pair_per_s, h_err,
         blocks, threads);
  return 0;
}

// Microbenchmark: cost of the neighbor-position gather of a full-list FP64 LJ force evaluation on sm_100a
// when positions come (G) from global memory through L1TEX (one 32-byte sector per neighbor, what
// force_lj_kernel does today) versus (S) from a per-CTA shared-memory copy of the tile's halo window,
// addressed by 16-bit tile-local indices.  Index pattern mimics -s 80: tile of 8x4x4 bins (~7.1 atoms
// per bin), halo 12x8x8 bins stored as 64 pencil runs, ~75 neighbors per atom spread over 25 runs.
//   S0: smem AoS 32 B records (x,y,z,type)      S1: smem SoA (three double arrays)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tile_gather_bench tile_gather_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

struct alignas(32) D4 { double x, y, z, w; };

#ifndef BX
#define BX 8
#endif
constexpr int RUN_TILE = (BX * 71 + 9) / 10;                            // atoms of one centre run inside the tile
constexpr int RUNS = 64, RUN_ATOMS = RUN_TILE + 28, HALO = RUNS * RUN_ATOMS;
constexpr int TILE_ATOMS = 16 * RUN_TILE;
constexpr int NEIGH = 76, STRIDE = 80;                               // 16-bit entries per row (16-B multiple)

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

#ifndef MATH
#define MATH 1
#endif
__device__ __forceinline__ double frcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0); y = fma(y, e, y);
  e = fma(-x, y, 1.0); y = fma(y, e, y);
  return y;
}
__device__ __forceinline__ void lj(double dx, double dy, double dz, double& fx, double& fy, double& fz) {
  if (MATH == 0) { fx += dx; fy += dy; fz += dz; return; }
  const double rsq = dx * dx + dy * dy + dz * dz;
  if (rsq < 6.25 && rsq > 0.0) {
    const double sr2 = MATH == 2 ? frcp(rsq) : 1.0 / rsq;
    const double sr6 = sr2 * sr2 * sr2;
    const double force = 48.0 * sr6 * (sr6 - 0.5) * sr2;
    fx += dx * force; fy += dy * force; fz += dz * force;
  }
}

template <int TPA> __device__ __forceinline__ double gsum(double v) {
#pragma unroll
  for (int o = TPA / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 32);
  return v;
}

// ---- S: shared-memory gathers -------------------------------------------------------------------
template <int TPA, int SOA>
__global__ void __launch_bounds__(512) kS(const D4* __restrict__ xh /* [tiles][HALO] */, const unsigned short* __restrict__ rows,
                                          const int* __restrict__ own /* [TILE_ATOMS] local index of each tile atom */,
                                          D4* __restrict__ f) {
  extern __shared__ __align__(32) unsigned char smem[];
  D4* sA = reinterpret_cast<D4*>(smem);
  double* sx = reinterpret_cast<double*>(smem);
  double* sy = sx + HALO;
  double* sz = sy + HALO;
  const int tile = blockIdx.x;
  const D4* src = xh + (size_t)tile * HALO;
  for (int s = threadIdx.x; s < HALO; s += blockDim.x) {
    D4 v;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(src + s));
    if (SOA) { sx[s] = v.x; sy[s] = v.y; sz[s] = v.z; } else sA[s] = v;
  }
  __syncthreads();
  const int sub = threadIdx.x % TPA;
  for (int a = threadIdx.x / TPA; a < TILE_ATOMS; a += blockDim.x / TPA) {
    const int li = own[a];
    double xi, yi, zi;
    if (SOA) { xi = sx[li]; yi = sy[li]; zi = sz[li]; } else { xi = sA[li].x; yi = sA[li].y; zi = sA[li].z; }
    const unsigned short* row = rows + ((size_t)tile * TILE_ATOMS + a) * STRIDE;
    double fx = 0, fy = 0, fz = 0;
    // each lane pulls 8 entries (16 B) at a time
    for (int k0 = sub * 8; k0 < NEIGH; k0 += TPA * 8) {
      const uint4 pk = __ldg(reinterpret_cast<const uint4*>(row + k0));
      const unsigned w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int lj_ = (w[e >> 1] >> ((e & 1) * 16)) & 0xffff;
        if (k0 + e < NEIGH) {
          double xj, yj, zj;
          if (SOA) { xj = sx[lj_]; yj = sy[lj_]; zj = sz[lj_]; } else { const D4 v = sA[lj_]; xj = v.x; yj = v.y; zj = v.z; }
          lj(xi - xj, yi - yj, zi - zj, fx, fy, fz);
        }
      }
    }
    fx = gsum<TPA>(fx); fy = gsum<TPA>(fy); fz = gsum<TPA>(fz);
    if (sub == 0) { D4 o; o.x = fx; o.y = fy; o.z = fz; o.w = 0; f[(size_t)tile * TILE_ATOMS + a] = o; }
  }
}

// ---- G: global gathers (today's shape): rows hold 32-bit global ids into the per-tile halo copy --
template <int TPA>
__global__ void __launch_bounds__(256) kG(const D4* __restrict__ xh, const int* __restrict__ rows32, const int* __restrict__ own,
                                          D4* __restrict__ f, int natoms) {
  const int g = (blockIdx.x * 256 + threadIdx.x) / TPA;
  const int sub = threadIdx.x % TPA;
  if (g >= natoms) return;
  const int tile = g / TILE_ATOMS, a = g % TILE_ATOMS;
  const D4* base = xh + (size_t)tile * HALO;
  const D4 vi = base[own[a]];
  const int* row = rows32 + (size_t)g * STRIDE;
  double fx = 0, fy = 0, fz = 0;
  for (int k = sub; k < NEIGH; k += TPA) {
    const int j = __ldg(row + k);
    D4 v;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(base + j));
    lj(vi.x - v.x, vi.y - v.y, vi.z - v.z, fx, fy, fz);
  }
  fx = gsum<TPA>(fx); fy = gsum<TPA>(fy); fz = gsum<TPA>(fz);
  if (sub == 0) { D4 o; o.x = fx; o.y = fy; o.z = fz; o.w = 0; f[g] = o; }
}

static unsigned rng_state = 12345u;
static unsigned rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

int main(int argc, char** argv) {
  const int tiles = argc > 1 ? atoi(argv[1]) : 2246;
  const size_t natoms = (size_t)tiles * TILE_ATOMS;
  printf("BX %d MATH %d tiles %d, atoms %zu, halo %d atoms/tile, %d neighbors/atom (full list)\n", BX, MATH, tiles, natoms, HALO, NEIGH);
  // tile-local geometry: run r = (ry, rz) in 8x8; centre runs ry,rz in 2..5; atoms of a run ordered along x
  std::vector<int> own(TILE_ATOMS);
  std::vector<unsigned short> row16((size_t)TILE_ATOMS * STRIDE, 0);
  {
    int a = 0;
    for (int rz = 2; rz < 6; rz++)
      for (int ry = 2; ry < 6; ry++)
        for (int k = 0; k < RUN_TILE; k++, a++) {
          const int xs = 14 + k;                     // slot along the run (2 halo bins = 14 atoms before the tile)
          own[a] = (rz * 8 + ry) * RUN_ATOMS + xs;
          int n = 0;
          for (int dz = -2; dz <= 2; dz++)
            for (int dy = -2; dy <= 2; dy++) {
              const int r = (rz + dz) * 8 + (ry + dy);
              int cnt = 3 + ((dz == 0 && dy == 0) ? 1 : 0);
              int picks[4];
              for (int c = 0; c < cnt; c++) picks[c] = xs - 17 + (int)(rnd() % 35);
              for (int c = 0; c < cnt; c++) for (int d = c + 1; d < cnt; d++) if (picks[d] < picks[c]) { int t = picks[c]; picks[c] = picks[d]; picks[d] = t; }
              for (int c = 0; c < cnt && n < NEIGH; c++) {
                int s = picks[c]; if (s < 0) s = 0; if (s >= RUN_ATOMS) s = RUN_ATOMS - 1;
                int li = r * RUN_ATOMS + s; if (li == own[a]) li++;
                row16[(size_t)a * STRIDE + n++] = (unsigned short)li;
              }
            }
        }
  }
  std::vector<unsigned short> rows16(natoms * STRIDE);
  std::vector<int> rows32(natoms * STRIDE);
  for (int t = 0; t < tiles; t++)
    for (size_t e = 0; e < (size_t)TILE_ATOMS * STRIDE; e++) {
      rows16[(size_t)t * TILE_ATOMS * STRIDE + e] = row16[e];
      rows32[(size_t)t * TILE_ATOMS * STRIDE + e] = row16[e];
    }
  std::vector<D4> xh((size_t)tiles * HALO);
  for (size_t i = 0; i < xh.size(); i++) {
    const int li = i % HALO, r = li / RUN_ATOMS, s = li % RUN_ATOMS;
    xh[i].x = s * 0.29 + (rnd() % 1000) * 1e-4; xh[i].y = (r % 8) * 2.04 + (rnd() % 1000) * 2e-3; xh[i].z = (r / 8) * 2.04 + (rnd() % 1000) * 2e-3; xh[i].w = 0;
  }
  D4 *dx, *df; unsigned short* dr16; int *dr32, *down;
  CK(cudaMalloc(&dx, xh.size() * sizeof(D4))); CK(cudaMalloc(&df, natoms * sizeof(D4)));
  CK(cudaMalloc(&dr16, rows16.size() * 2)); CK(cudaMalloc(&dr32, rows32.size() * 4)); CK(cudaMalloc(&down, TILE_ATOMS * 4));
  CK(cudaMemcpy(dx, xh.data(), xh.size() * sizeof(D4), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dr16, rows16.data(), rows16.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dr32, rows32.data(), rows32.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(down, own.data(), TILE_ATOMS * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto launch) {
    for (int w = 0; w < 2; w++) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e9;
    for (int r = 0; r < 5; r++) {
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    printf("%-28s %.3f ms  (%.1f G pair-slots/s)\n", name, best, natoms * (double)NEIGH / best * 1e-6);
  };
  const int smA = HALO * 32, smS = HALO * 24;
  CK(cudaFuncSetAttribute(kS<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smA));
  CK(cudaFuncSetAttribute(kS<8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smA));
  CK(cudaFuncSetAttribute(kS<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smA));
  CK(cudaFuncSetAttribute(kS<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smS));
  CK(cudaFuncSetAttribute(kS<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smS));
  CK(cudaFuncSetAttribute(kS<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smS));
  CK(cudaFuncSetAttribute(kS<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smS));
  for (int th : {512}) {
    printf("-- block %d\n", th);
    timeit("S AoS32 TPA2", [&] { kS<2, 0><<<tiles, th, smA>>>(dx, dr16, down, df); });
    timeit("S AoS32 TPA4", [&] { kS<4, 0><<<tiles, th, smA>>>(dx, dr16, down, df); });
    timeit("S AoS32 TPA8", [&] { kS<8, 0><<<tiles, th, smA>>>(dx, dr16, down, df); });
    timeit("S SoA   TPA1", [&] { kS<1, 1><<<tiles, th, smS>>>(dx, dr16, down, df); });
    timeit("S SoA   TPA2", [&] { kS<2, 1><<<tiles, th, smS>>>(dx, dr16, down, df); });
    timeit("S SoA   TPA4", [&] { kS<4, 1><<<tiles, th, smS>>>(dx, dr16, down, df); });
    timeit("S SoA   TPA8", [&] { kS<8, 1><<<tiles, th, smS>>>(dx, dr16, down, df); });
  }
  const int nb2 = (int)((natoms * 2 + 255) / 256), nb4 = (int)((natoms * 4 + 255) / 256);
  timeit("G global TPA2", [&] { kG<2><<<nb2, 256>>>(dx, dr32, down, df, (int)natoms); });
  timeit("G global TPA4", [&] { kG<4><<<nb4, 256>>>(dx, dr32, down, df, (int)natoms); });
  return 0;
}

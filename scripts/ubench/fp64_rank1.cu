// Microbenchmark 2: (a) symmetric rank-1 accumulation acc[tri(i,j)] += (w a_i) a_j with 10-vector a
// (the K2a hot-loop pattern: 55 accumulators, 3 distinct 64-bit operands per DFMA), per warps/SM;
// (b) DFMA and DMMA issued concurrently from different warps of the SAME scheduler, work balanced,
// to tell whether m8n8k4 DMMA runs on the FP64 vector pipe or beside it.
#include <cstdio>
#include <cuda_runtime.h>

template <int NV>
__global__ void rank1(double* out, int iters, double seed) {
  constexpr int NA = NV * (NV + 1) / 2;
  double acc[NA], a[NV];
#pragma unroll
  for (int i = 0; i < NA; ++i) acc[i] = 0.0;
#pragma unroll
  for (int i = 0; i < NV; ++i) a[i] = seed + 0.01 * i + 1e-3 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NV; ++i) a[i] = a[i] * 0.999999;   // NV DMUL: keep a changing
    int k = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const double s = seed * a[i];                          // NV DMUL
#pragma unroll
      for (int j = i; j < NV; ++j) acc[k++] = fma(s, a[j], acc[k]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NA; ++i) s += acc[i];
  if (s == 12345.678) *out = s;
}

__global__ void mixed(double* out, int iters, int mode, double seed) {
  // warps 0-3: one per scheduler, warps 4-7: one per scheduler.  mode 0: all DFMA; 1: all DMMA;
  // 2: warps 0-3 DFMA (8x iterations... same FMA count as a DMMA warp), warps 4-7 DMMA
  const int warp = threadIdx.x >> 5;
  const bool mma = mode == 1 || (mode == 2 && warp >= 4);
  if (mma) {
    double c[8][2];
    double a = seed, b = 1.0 - 1e-9;
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = seed + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) *out = s;
  } else {
    double a[8];
    const double x = seed, y = 1.0 - 1e-9;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + i;
    for (int it = 0; it < iters * 8; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], y, x);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 12345.678) *out = s;
  }
}

template <typename F>
float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  double* d; cudaMalloc(&d, 8);
  const int iters = 4000;
  for (int warps : {4, 8, 12, 16}) {
    float ms = time_it([&] { rank1<10><<<nsm, warps * 32>>>(d, iters, 0.5); });
    double ops = (double)nsm * warps * 32 * iters * (55 + 20);
    if (cudaGetLastError() != cudaSuccess) { printf("rank1<10> warps %d: launch failed\n", warps); } else printf("rank1<10> warps/SM %2d: %.3f ms  %.1f fp64 ops/clk/SM (peak 64)\n", warps, ms, ops / (ms * 1e-3) / nsm / 1.965e9);
    ms = time_it([&] { rank1<12><<<nsm, warps * 32>>>(d, iters, 0.5); });
    ops = (double)nsm * warps * 32 * iters * (78 + 24);
    if (cudaGetLastError() != cudaSuccess) { printf("rank1<12> warps %d: launch failed\n", warps); } else printf("rank1<12> warps/SM %2d: %.3f ms  %.1f fp64 ops/clk/SM (peak 64)\n", warps, ms, ops / (ms * 1e-3) / nsm / 1.965e9);
  }
  for (int mode = 0; mode < 3; ++mode) {
    float ms = time_it([&] { mixed<<<nsm, 256>>>(d, 5000, mode, 0.5); });
    double fma_total = (double)nsm * 8 * 256.0 * 8 * 5000;   // every warp performs the same FMA count in all modes
    printf("mixed mode %d: %.3f ms  %.1f FMA/clk/SM\n", mode, ms, fma_total / (ms * 1e-3) / nsm / 1.965e9);
  }
  return 0;
}

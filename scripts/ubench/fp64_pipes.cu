// Microbenchmark: FP64 vector (DFMA) and FP64 tensor (DMMA m8n8k4) throughput on one B200,
// alone and concurrently, as a function of resident warps per SM and ILP.  Establishes the
// FP64 compute ceiling used in DESIGN.md next to the HBM roofline.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipes fp64_pipes.cu && ./fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__device__ __forceinline__ void dfma_body(double* out, int iters, double seed) {
  double a[ILP];
  const double x = seed, y = 1.0 - 1e-9;
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = seed + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], y, x);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  if (s == 12345.678) *out = s;
}

template <int ILP>
__device__ __forceinline__ void dmma_body(double* out, int iters, double seed) {
  double c[ILP][2];
  double a = seed, b = 1.0 - 1e-9;
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = seed + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) *out = s;
}

// mode 0: all warps DFMA; 1: all warps DMMA; 2: even warps DFMA, odd warps DMMA
template <int ILP>
__global__ void k(double* out, int iters, int mode, double seed) {
  const int warp = threadIdx.x >> 5;
  const bool mma = mode == 1 || (mode == 2 && (warp & 1));
  if (mma) dmma_body<ILP>(out, iters, seed); else dfma_body<ILP>(out, iters, seed);
}

template <int ILP>
void run(int warps, int mode, int nsm, double clk_ghz) {
  double* d;
  cudaMalloc(&d, 8);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<ILP><<<nsm, warps * 32>>>(d, 100, mode, 0.5);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<ILP><<<nsm, warps * 32>>>(d, iters, mode, 0.5);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int wf = mode == 0 ? warps : mode == 1 ? 0 : (warps + 1) / 2;
  int wm = warps - wf;
  double fma_f = (double)nsm * wf * 32 * ILP * (double)iters;         // scalar FMAs
  double fma_m = (double)nsm * wm * 256.0 * ILP * (double)iters;      // m8n8k4 = 256 FMAs
  double tf = 2.0 * (fma_f + fma_m) / (ms * 1e-3) / 1e12;
  double per_clk_sm = (fma_f + fma_m) / (ms * 1e-3) / nsm / (clk_ghz * 1e9);
  printf("mode %d warps/SM %2d ILP %2d : %8.3f ms  %7.2f TFLOP/s  (DFMA %.2f + DMMA %.2f TF)  %.1f FMA/clk/SM @%.3f GHz\n", mode, warps,
         ILP, ms, tf, 2 * fma_f / (ms * 1e-3) / 1e12, 2 * fma_m / (ms * 1e-3) / 1e12, per_clk_sm, clk_ghz);
  cudaFree(d);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double ghz = clk_khz * 1e-6;
  printf("%s  SMs %d  clock %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
  const int nsm = p.multiProcessorCount;
  for (int mode = 0; mode < 3; ++mode) {
    for (int warps : {4, 8, 16, 32}) {
      run<1>(warps, mode, nsm, ghz);
      run<4>(warps, mode, nsm, ghz);
      run<8>(warps, mode, nsm, ghz);
      run<16>(warps, mode, nsm, ghz);
    }
  }
  return 0;
}

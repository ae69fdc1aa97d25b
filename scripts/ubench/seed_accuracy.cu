// Accuracy of the MUFU-seeded reciprocal / reciprocal square root used in K2p (k2_producer.cuh, mcba_math.cuh)
// with and without their last Newton step: max and mean relative error against correctly rounded results.
#include <cmath>
#include <cstdio>
#include <vector>
__device__ __forceinline__ double rsqrt_seed(double t) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t)); return y; }
__device__ __forceinline__ double rcp_seed(double t) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t)); return y; }
__global__ void k(const double* t, int n, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = t[i];
  double y = rsqrt_seed(v);
  out[0 * n + i] = y;
  double e = fma(-v, y * y, 1.0);
  y = fma(y * e, fma(0.375, e, 0.5), y);
  out[1 * n + i] = y;                       // cubic step only
  e = fma(-v, y * y, 1.0);
  out[2 * n + i] = fma(0.5 * y, e, y);      // + quadratic step (library)
  double r = rcp_seed(v);
  out[3 * n + i] = r;
  double d = fma(-v, r, 1.0);
  d = fma(d, d, d);
  r = fma(r, d, r);
  out[4 * n + i] = r;                       // cubic step only
  d = fma(-v, r, 1.0);
  out[5 * n + i] = fma(r, d, r);            // + quadratic step (library)
}
int main() {
  const int n = 1 << 22;
  std::vector<double> t(n);
  for (int i = 0; i < n; ++i) {
    t[i] = 1.0 + (double)i / n * (i % 3 == 0 ? 1.0 : (i % 3 == 1 ? 30.0 : 4000.0)) + 1e-9 * i;
    if (i % 5 == 4) t[i] = ldexp(t[i], (i / 5) % 400 - 100);   // wide exponent range: 2^-100 .. 2^300
  }
  double *dt, *dout;
  cudaMalloc(&dt, n * 8); cudaMalloc(&dout, 6ull * n * 8);
  cudaMemcpy(dt, t.data(), n * 8, cudaMemcpyHostToDevice);
  k<<<(n + 255) / 256, 256>>>(dt, n, dout);
  std::vector<double> o(6ull * n);
  cudaMemcpy(o.data(), dout, 6ull * n * 8, cudaMemcpyDeviceToHost);
  const char* names[6] = {"rsqrt seed", "rsqrt seed+cubic", "rsqrt seed+cubic+quadratic", "rcp seed", "rcp seed+cubic", "rcp seed+cubic+quadratic"};
  for (int v = 0; v < 6; ++v) {
    long double mx = 0, sum = 0;
    for (int i = 0; i < n; ++i) {
      const long double ref = v < 3 ? 1.0L / sqrtl((long double)t[i]) : 1.0L / (long double)t[i];
      const long double err = fabsl(((long double)o[(size_t)v * n + i] - ref) / ref);
      mx = err > mx ? err : mx; sum += err;
    }
    printf("%-28s max rel err %.3Le (%.2Lf ulp of 2^-53)  mean %.3Le\n", names[v], mx, mx / 1.1102230246251565e-16L, sum / n);
  }
  return 0;
}

// Device helpers shared by the K2a kernels (see k2_frames.cu for the math).
#pragma once
#include "mcba_internal.h"
#include "mcba_obs.cuh"

namespace mcba {

__host__ __device__ constexpr int tri12(int i, int j) { return i * 12 - (i * (i - 1)) / 2 + (j - i); }
__host__ __device__ constexpr int sym12(int i, int j) { return i <= j ? tri12(i, j) : tri12(j, i); }
__host__ __device__ constexpr int tri6(int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); }
__host__ __device__ constexpr int sym6(int i, int j) { return i <= j ? tri6(i, j) : tri6(j, i); }
constexpr int kQ = 78;  // offset of q inside the 96-value accumulator

template <bool IsU>
__device__ __forceinline__ void accumulate_row(double (&acc)[kUPad], const double (&a)[10], double wh,
                                               double gf) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const int I = IsU ? kIdxU[i] : kIdxV[i];
    const double s = wh * a[i];
    acc[kQ + I] = fma(gf, a[i], acc[kQ + I]);
#pragma unroll
    for (int j = i; j < 10; ++j) {
      const int J = IsU ? kIdxU[j] : kIdxV[j];
      acc[tri12(I, J)] = fma(s, a[j], acc[tri12(I, J)]);
    }
  }
}

// Sum over the 32 lanes of v[Base + l] delivered to lane l (recursive halving:
// 31 shuffles instead of 32 x 5).
template <int Base>
__device__ __forceinline__ double lane_transpose_sum32(const double (&v)[kUPad], int lane) {
  double w[16];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const double keep = up ? v[Base + 16 + i] : v[Base + i];
      const double send = up ? v[Base + i] : v[Base + 16 + i];
      w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1) {
    const bool up = lane & half;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const double keep = up ? w[half + i] : w[i];
      const double send = up ? w[i] : w[half + i];
      w[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return w[0];
}

// Walk the N corners of one (camera, frame) pair: A = sum w a a^T, q = -sum rho' f a.
__device__ __forceinline__ void accumulate_pair(const K2Params& p, const Intr& in, const double (&Rcf)[9],
                                                const double (&tcf)[3], const double2* __restrict__ ob,
                                                const double* __restrict__ s_obj, double (&acc)[kUPad],
                                                double& cost_acc, double& sumsq_acc, double& cnt_acc) {
  const int N = p.N;
  double2 o = ob[0];
  for (int n = 0; n < N; ++n) {
    const double2 cur = o;
    if (n + 1 < N) o = ob[(size_t)(n + 1) * kTile];
    const bool hu = cur.x == cur.x, hv = cur.y == cur.y;
    if (hu | hv) {
      double pu, pv, au[10], av[10];
      project_jac(in, Rcf, tcf, s_obj[3 * n], s_obj[3 * n + 1], s_obj[3 * n + 2], pu, pv, au, av);
      const double fu = hu ? cur.x - pu : 0.0, fv = hv ? cur.y - pv : 0.0;
      double rho, wg, wh;
      robust_weights(p.loss, fu, p.inv_c, p.c2, rho, wg, wh);
      if (!hu) wh = 0.0;
      cost_acc += hu ? rho : 0.0;
      accumulate_row<true>(acc, au, wh, -wg * fu);
      robust_weights(p.loss, fv, p.inv_c, p.c2, rho, wg, wh);
      if (!hv) wh = 0.0;
      cost_acc += hv ? rho : 0.0;
      accumulate_row<false>(acc, av, wh, -wg * fv);
      sumsq_acc += fma(fu, fu, fv * fv);
      cnt_acc += (hu ? 1.0 : 0.0) + (hv ? 1.0 : 0.0);
    }
  }
}

// V = P'^T V'' P', g = P'^T g'' with P' = blkdiag(J_l(rho), I); damping lambda * D_f^2 with the
// running Marquardt scaling; 6x6 Cholesky; returns L^-1 (packed lower) and y = L^-1 g.
__device__ __forceinline__ void pose_block_factor(const double (&Vpp)[21], const double (&gpp)[6],
                                                  const double (&Jl)[9], double lambda, double* d2p,
                                                  bool write_d2, double (&Linv)[21], double (&yv)[6],
                                                  double (&gp)[6], double& gmax) {
  double V[36];
  double T1[9];  // V''_ee J
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      T1[3 * i + j] = Vpp[sym6(i, 0)] * Jl[j] + Vpp[sym6(i, 1)] * Jl[3 + j] + Vpp[sym6(i, 2)] * Jl[6 + j];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      V[6 * i + j] = Jl[i] * T1[j] + Jl[3 + i] * T1[3 + j] + Jl[6 + i] * T1[6 + j];
      const double vt = Jl[i] * Vpp[tri6(0, 3 + j)] + Jl[3 + i] * Vpp[tri6(1, 3 + j)] + Jl[6 + i] * Vpp[tri6(2, 3 + j)];
      V[6 * i + 3 + j] = vt;
      V[6 * (3 + j) + i] = vt;
      V[6 * (3 + i) + 3 + j] = Vpp[sym6(3 + i, 3 + j)];
    }
    gp[i] = Jl[i] * gpp[0] + Jl[3 + i] * gpp[1] + Jl[6 + i] * gpp[2];
    gp[3 + i] = gpp[3 + i];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double d2 = fmax(d2p[i * 32], V[7 * i]);
    if (write_d2) d2p[i * 32] = d2;
    if (d2 == 0.0) d2 = 1.0;
    V[7 * i] = fma(lambda, d2, V[7 * i]);
    gmax = fmax(gmax, fabs(gp[i]));
  }
  double Lm[21];  // Cholesky factor, 1/L_jj stored on the diagonal; empty frames at lambda = 0 give zeros
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = V[7 * j];
#pragma unroll
    for (int k = 0; k < j; ++k) s -= Lm[j * (j + 1) / 2 + k] * Lm[j * (j + 1) / 2 + k];
    const double inv = s > 0.0 ? rsqrt(s) : 0.0;
    Lm[j * (j + 1) / 2 + j] = inv;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double t = V[6 * i + j];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= Lm[i * (i + 1) / 2 + k] * Lm[j * (j + 1) / 2 + k];
      Lm[i * (i + 1) / 2 + j] = t * inv;
    }
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    Linv[j * (j + 1) / 2 + j] = Lm[j * (j + 1) / 2 + j];
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double t = 0.0;
#pragma unroll
      for (int k = j; k < i; ++k) t += Lm[i * (i + 1) / 2 + k] * Linv[k * (k + 1) / 2 + j];
      Linv[i * (i + 1) / 2 + j] = -t * Lm[i * (i + 1) / 2 + i];
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double t = 0.0;
#pragma unroll
    for (int j = 0; j <= i; ++j) t += Linv[i * (i + 1) / 2 + j] * gp[j];
    yv[i] = t;
  }
}

// One row of Z_cf = (W' P') L^-T from one row of W' = A[:,ext] E'.
__device__ __forceinline__ void z_row(const double (&b)[6], const double (&Jl)[9], const double (&Linv)[21],
                                      double (&z)[6]) {
  double w[6];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    w[k] = b[0] * Jl[k] + b[1] * Jl[3 + k] + b[2] * Jl[6 + k];
    w[3 + k] = b[3 + k];
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double t = 0.0;
#pragma unroll
    for (int j = 0; j <= k; ++j) t += w[j] * Linv[k * (k + 1) / 2 + j];
    z[k] = t;
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

}  // namespace mcba

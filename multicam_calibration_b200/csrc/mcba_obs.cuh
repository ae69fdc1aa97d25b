// Per-observation forward model + raw Jacobian rows (see mcba_math.cuh header).
#pragma once
#include "mcba_math.cuh"

namespace mcba {

// Sparse raw Jacobian rows of the PREDICTION.  Row u has structural zeros at
// (fy, cy), row v at (fx, cx), so each is stored as 10
// values mapped to the 12 raw columns by kIdxU / kIdxV:
//   [ d/df, d/dc(=1), d/dk1, d/dk2, m(3), G(3) ]
__device__ constexpr int kIdxU[10] = {0, 2, 4, 5, 6, 7, 8, 9, 10, 11};
__device__ constexpr int kIdxV[10] = {1, 3, 4, 5, 6, 7, 8, 9, 10, 11};

struct Intr {
  double fx, fy, cx, cy, k1, k2;
};

// Camera-frame point from the composed transform X_c = Rcf Q + tcf.
__device__ __forceinline__ void project(const Intr& in, const double Rcf[9], const double tcf[3],
                                        double qx, double qy, double qz, double& pu, double& pv) {
  const double X = fma(Rcf[0], qx, fma(Rcf[1], qy, fma(Rcf[2], qz, tcf[0])));
  const double Y = fma(Rcf[3], qx, fma(Rcf[4], qy, fma(Rcf[5], qz, tcf[1])));
  const double Z = fma(Rcf[6], qx, fma(Rcf[7], qy, fma(Rcf[8], qz, tcf[2])));
  const double iz = fast_rcp(Z);   // branch-free (K1 is bound by its instruction count, not by HBM)
  const double x = X * iz, y = Y * iz;
  const double r2 = fma(x, x, y * y);
  const double d = fma(r2, fma(in.k2, r2, in.k1), 1.0);
  pu = fma(in.fx, x * d, in.cx);
  pv = fma(in.fy, y * d, in.cy);
}

// Prediction and raw Jacobian rows for one observation.
__device__ __forceinline__ void project_jac(const Intr& in, const double Rcf[9], const double tcf[3],
                                            double qx, double qy, double qz, double& pu, double& pv,
                                            double (&au)[10], double (&av)[10]) {
  const double X = fma(Rcf[0], qx, fma(Rcf[1], qy, fma(Rcf[2], qz, tcf[0])));
  const double Y = fma(Rcf[3], qx, fma(Rcf[4], qy, fma(Rcf[5], qz, tcf[1])));
  const double Z = fma(Rcf[6], qx, fma(Rcf[7], qy, fma(Rcf[8], qz, tcf[2])));
  const double iz = 1.0 / Z;
  const double x = X * iz, y = Y * iz;
  const double r2 = fma(x, x, y * y);
  const double d = fma(r2, fma(in.k2, r2, in.k1), 1.0);
  const double dp = fma(2.0 * in.k2, r2, in.k1);
  const double xd = x * d, yd = y * d;
  pu = fma(in.fx, xd, in.cx);
  pv = fma(in.fy, yd, in.cy);
  // intrinsics
  au[0] = xd;
  au[1] = 1.0;
  au[2] = in.fx * x * r2;
  au[3] = au[2] * r2;
  av[0] = yd;
  av[1] = 1.0;
  av[2] = in.fy * y * r2;
  av[3] = av[2] * r2;
  // A = d(u,v)/d(x,y)
  const double xy2 = 2.0 * x * y * dp;
  const double A00 = in.fx * fma(2.0 * x * x, dp, d), A01 = in.fx * xy2;
  const double A10 = in.fy * xy2, A11 = in.fy * fma(2.0 * y * y, dp, d);
  const double su = fma(A00, x, A01 * y), sv = fma(A10, x, A11 * y);
  // G = A B with B = [[1,0,-x],[0,1,-y]] / Z ;  m = X_c x G = (x,y,1) x (A.0, A.1, -s)
  au[4] = -fma(y, su, A01);
  au[5] = fma(x, su, A00);
  au[6] = fma(x, A01, -y * A00);
  au[7] = A00 * iz;
  au[8] = A01 * iz;
  au[9] = -su * iz;
  av[4] = -fma(y, sv, A11);
  av[5] = fma(x, sv, A10);
  av[6] = fma(x, A11, -y * A10);
  av[7] = A10 * iz;
  av[8] = A11 * iz;
  av[9] = -sv * iz;
}

}  // namespace mcba

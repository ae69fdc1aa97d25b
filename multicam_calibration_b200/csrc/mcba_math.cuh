// Device-side fp64 geometry shared by every kernel of the bundle-adjustment path.
//
// Forward model (reference geometry.py:8-35, 277-325; bundle_adjustment.py:27-29):
//   X_w = R(rho_f) X_o + tau_f ;  X_c = R(r_c) X_w + t_c ;  x = X/Z, y = Y/Z
//   d = 1 + k1 r^2 + k2 r^4 ;  u = fx x d + cx ;  v = fy y d + cy
// Analytic derivatives follow SURVEY.md Appendix A, re-expressed with LEFT
// perturbations so that the per-observation work is a "raw" 12-vector
//   a = [ d(u|v)/d(fx,fy,cx,cy,k1,k2) | m = X_c x G | G = d(u|v)/dX_c ]
// and the maps to the true parameters are per-camera / per-frame constants:
//   d/dr_c  = (m - t_c x G)^T J_l(r_c)          d/dt_c  = G^T
//   d/drho  = (m - t_cf x G)^T R_c J_l(rho)     d/dtau  = G^T R_c     (t_cf = R_c tau + t_c)
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace mcba {

constexpr int kCamBlock = 12;   // fx fy cx cy k1 k2 rx ry rz tx ty tz  (bundle_adjustment.py:155)
constexpr int kPoseBlock = 6;   // rho(3) tau(3)                          (bundle_adjustment.py:156)
constexpr int kTile = 32;       // frames per tile = lanes per warp

// 1/z for finite z of either sign in the normal range: MUFU.RCP64H seed (relative error <= 1e-6) and ONE
// cubic Newton step, y (1 + e + e^2): error e^3 ~ 1e-18, i.e. already rounding-limited.  Measured on
// B200 over 4 M arguments and 400 binades (scripts/ubench/seed_accuracy.cu): max 1.00 ulp, the same as
// with the further quadratic step CUDA's own 1.0/z appends (which this function carried until round 2).
__device__ __forceinline__ double fast_rcp(double z) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(z));
  double e = fma(-z, y, 1.0);
  e = fma(e, e, e);
  return fma(y, e, y);
}

// Per-camera constants derived from the 12 camera parameters once per evaluation.
struct CamConst {
  double fx, fy, cx, cy, k1, k2;
  double R[9];    // R(r_c), row major
  double t[3];    // t_c
  double Jl[9];   // left Jacobian J_l(r_c):  exp(r + dr) ~ exp(J_l dr) exp(r)
  double tJ[9];   // [t_c]x J_l(r_c)
};

__device__ __forceinline__ void rodrigues(const double r[3], double R[9]) {
  // geometry.py:22-34: K = [r]x / theta (theta == 0 -> divide by 1), R = I + sin K + (1 - cos) K^2
  const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  const double inv = (th == 0.0) ? 1.0 : 1.0 / th;
  const double kx = r[0] * inv, ky = r[1] * inv, kz = r[2] * inv;
  double s, c;
  sincos(th, &s, &c);
  const double oc = 1.0 - c;
  // K^2 = k k^T - |k|^2 I  (|k| = 1 unless theta == 0, where K = 0)
  const double n2 = kx * kx + ky * ky + kz * kz;
  R[0] = 1.0 + oc * (kx * kx - n2);
  R[1] = -s * kz + oc * kx * ky;
  R[2] = s * ky + oc * kx * kz;
  R[3] = s * kz + oc * kx * ky;
  R[4] = 1.0 + oc * (ky * ky - n2);
  R[5] = -s * kx + oc * ky * kz;
  R[6] = -s * ky + oc * kx * kz;
  R[7] = s * kx + oc * ky * kz;
  R[8] = 1.0 + oc * (kz * kz - n2);
}

// J_l(r) = I + a [r]x + b [r]x^2,  a = (1 - cos th)/th^2,  b = (th - sin th)/th^3
__device__ __forceinline__ void so3_left_jacobian(const double r[3], double J[9]) {
  const double t2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  const double th = sqrt(t2);
  double a, b;
  if (th < 0.1) {
    a = 0.5 + t2 * (-1.0 / 24 + t2 * (1.0 / 720 + t2 * (-1.0 / 40320 + t2 * (1.0 / 3628800))));
    b = 1.0 / 6 + t2 * (-1.0 / 120 + t2 * (1.0 / 5040 + t2 * (-1.0 / 362880 + t2 * (1.0 / 39916800))));
  } else {
    const double sh = sin(0.5 * th);
    a = 2.0 * sh * sh / t2;
    b = (th - sin(th)) / (t2 * th);
  }
  const double x = r[0], y = r[1], z = r[2];
  // [r]x^2 = r r^T - t2 I
  J[0] = 1.0 + b * (x * x - t2);
  J[1] = -a * z + b * x * y;
  J[2] = a * y + b * x * z;
  J[3] = a * z + b * x * y;
  J[4] = 1.0 + b * (y * y - t2);
  J[5] = -a * x + b * y * z;
  J[6] = -a * y + b * x * z;
  J[7] = a * x + b * y * z;
  J[8] = 1.0 + b * (z * z - t2);
}

// C = A B (3x3 row major)
__device__ __forceinline__ void mat3_mul(const double A[9], const double B[9], double C[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

// C = [t]x B
__device__ __forceinline__ void cross_mat3(const double t[3], const double B[9], double C[9]) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    C[j] = t[1] * B[6 + j] - t[2] * B[3 + j];
    C[3 + j] = t[2] * B[j] - t[0] * B[6 + j];
    C[6 + j] = t[0] * B[3 + j] - t[1] * B[j];
  }
}

__device__ __forceinline__ void mat3_vec(const double A[9], const double v[3], double o[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}

// Robust loss on one scalar residual f (scipy optimize/_lsq/least_squares.py:195-201,
// common.py:720-731).  z = (f/C)^2.  Returns rho(z) C^2 and the two weights the
// normal equations need:  wg = rho'  (gradient  J^T (rho' f)),
// wh = max(rho' + 2 rho'' z, EPS)  (Gauss-Newton  J^T diag(wh) J).
// soft_l1: rho' = (1+z)^-1/2 and rho' + 2 rho'' z = (1+z)^-3/2 exactly.
//
// kLossIrls (bit 8 of the loss code) selects the majorising Gauss-Newton weight
// wh = rho' instead (iteratively re-weighted least squares): same gradient, cost
// and stationary points, but a Hessian model that does not collapse when the
// residuals are still large (rho' + 2 rho'' z -> z^-3/2), which is what makes
// scipy's Triggs-scaled model crawl far from the minimum.
enum Loss : int { kLossLinear = 0, kLossSoftL1 = 1, kLossIrls = 0x100 };

__device__ __forceinline__ void robust_weights(int loss, double f, double inv_c, double c2,
                                               double& rho, double& wg, double& wh) {
  if ((loss & 0xff) == kLossLinear) {
    rho = f * f;
    wg = 1.0;
    wh = 1.0;
    return;
  }
  const double fs = f * inv_c;
  const double t = 1.0 + fs * fs;
  const double b = rsqrt(t);
  rho = 2.0 * (t * b - 1.0) * c2;
  wg = b;
  wh = (loss & kLossIrls) ? b : fmax(b * b * b, 2.220446049250313e-16);
}

}  // namespace mcba

// Reprojection-error QC: numeric core of viz.plot_residuals (viz.py:155-177), the step after
// bundle_adjust.  Per (camera, frame) whose board corners were all detected:
//   undistort the detections (OpenCV's 5-step inversion, geometry.py:328-358)  ->
//   homography detections -> board plane, as cv2.findHomography(src, dst) with the default method
//   computes it: both point sets rounded to float32, Hartley-normalised DLT (9x9 normal matrix,
//   eigenvector of the smallest eigenvalue by cyclic Jacobi), then at most 10 iterations of
//   OpenCV's LMSolver on the reprojection error of the 8 free entries  ->
//   cv2.perspectiveTransform of the distortion-free projections of the board corners (made by
//   project_points_multi with dist = None) into the board's own coordinates  ->
//   distance of every transferred corner to the true one.
// One thread per (camera, frame); the N <= 128 rounded detections of the frame live in local memory
// (1 KB at 5 x 7 corners), the small dense algebra (9x9 Jacobi, 8x8 Cholesky) in registers / local
// memory.  OpenCV is a binary dependency of the reference: the algorithm is restated from its
// documentation (oracle/np_oracle.py::find_homography) and pinned numerically against cv2 4.13 through
// tests/golden/qc.npz.
#include "mcba_internal.h"

namespace mcba {

constexpr int kQcMaxPoints = 128;
constexpr int kQcMaxCams = 32;

struct QcCam {
  double fx, fy, cx, cy, skew, k1, k2, p1, p2, k3;
};
struct QcParams {
  const double* uvs;      // (C,F,N,2) detections, NaN = missing
  const double* reproj;   // (C,F,N,2) distortion-free projections of the board corners
  const double* obj;      // (N,3)
  double* transformed;    // (C,F,N,2)
  double* err;            // (C,F,N)
  int C, N;
  long long F;
  QcCam cam[kQcMaxCams];
};

__device__ __forceinline__ void qc_undistort(const QcCam& k, double u, double v, double& uo, double& vo) {
  const double x0 = (u - k.cx) / k.fx, y0 = (v - k.cy) / k.fy;
  double x = x0, y = y0;
#pragma unroll
  for (int it = 0; it < 5; ++it) {
    const double r2 = x * x + y * y;
    const double icd = 1.0 / (1.0 + ((k.k3 * r2 + k.k2) * r2 + k.k1) * r2);
    const double dx = 2.0 * k.p1 * x * y + k.p2 * (r2 + 2.0 * x * x);
    const double dy = k.p1 * (r2 + 2.0 * y * y) + 2.0 * k.p2 * x * y;
    x = (x0 - dx) * icd;
    y = (y0 - dy) * icd;
  }
  uo = x * k.fx + y * k.skew + k.cx;
  vo = y * k.fy + k.cy;
}

// Eigenvector of the smallest eigenvalue of a symmetric 9x9 matrix (cyclic Jacobi).
__device__ void smallest_eigenvector9(double (&A)[9][9], double (&out)[9]) {
  double V[9][9];
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0.0, dia = 0.0;
    for (int i = 0; i < 9; ++i) {
      dia += A[i][i] * A[i][i];
      for (int j = i + 1; j < 9; ++j) off += A[i][j] * A[i][j];
    }
    if (off <= 1e-34 * dia) break;
    for (int p = 0; p < 8; ++p) {
      for (int q = p + 1; q < 9; ++q) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = rsqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 9; ++k) {   // columns p, q
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 9; ++k) {   // rows p, q
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 9; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
    }
  }
  int best = 0;
  for (int i = 1; i < 9; ++i)
    if (A[i][i] < A[best][best]) best = i;
  for (int k = 0; k < 9; ++k) out[k] = V[k][best];
}

// Lower Cholesky factor of the symmetric 8x8 matrix in A (full storage); false when not positive definite.
__device__ bool cholesky8(double (&A)[8][8]) {
  for (int j = 0; j < 8; ++j) {
    double d = A[j][j];
    for (int k = 0; k < j; ++k) d -= A[j][k] * A[j][k];
    if (!(d > 0.0)) return false;
    const double l = sqrt(d), inv = 1.0 / l;
    A[j][j] = l;
    for (int i = j + 1; i < 8; ++i) {
      double t = A[i][j];
      for (int k = 0; k < j; ++k) t -= A[i][k] * A[j][k];
      A[i][j] = t * inv;
    }
  }
  return true;
}
__device__ void cholesky8_solve(const double (&L)[8][8], const double (&b)[8], double (&x)[8]) {
  double y[8];
  for (int i = 0; i < 8; ++i) {
    double t = b[i];
    for (int k = 0; k < i; ++k) t -= L[i][k] * y[k];
    y[i] = t / L[i][i];
  }
  for (int i = 7; i >= 0; --i) {
    double t = y[i];
    for (int k = i + 1; k < 8; ++k) t -= L[k][i] * x[k];
    x[i] = t / L[i][i];
  }
}

// Residual sum of squares of H (8 free entries) and, when want_jac, J^T J (full 8x8) and J^T r:
// OpenCV's HomographyRefineCallback::compute.
__device__ double qc_lm_compute(const double (&h)[8], const double* Mx, const double* My, const double* mx, const double* my,
                                int N, bool want_jac, double (&A)[8][8], double (&v)[8], double& rmax) {
  double S = 0.0;
  rmax = 0.0;
  if (want_jac) {
    for (int i = 0; i < 8; ++i) {
      v[i] = 0.0;
      for (int j = 0; j < 8; ++j) A[i][j] = 0.0;
    }
  }
  for (int n = 0; n < N; ++n) {
    const double X = Mx[n], Y = My[n];
    double ww = h[6] * X + h[7] * Y + 1.0;
    ww = fabs(ww) > 2.220446049250313e-16 ? 1.0 / ww : 0.0;
    const double xi = (h[0] * X + h[1] * Y + h[2]) * ww, yi = (h[3] * X + h[4] * Y + h[5]) * ww;
    const double rx = xi - mx[n], ry = yi - my[n];
    S += rx * rx + ry * ry;
    rmax = fmax(rmax, fmax(fabs(rx), fabs(ry)));
    if (want_jac) {
      const double jx[8] = {X * ww, Y * ww, ww, 0.0, 0.0, 0.0, -X * ww * xi, -Y * ww * xi};
      const double jy[8] = {0.0, 0.0, 0.0, X * ww, Y * ww, ww, -X * ww * yi, -Y * ww * yi};
      for (int i = 0; i < 8; ++i) {
        v[i] += jx[i] * rx + jy[i] * ry;
        for (int j = i; j < 8; ++j) A[i][j] += jx[i] * jx[j] + jy[i] * jy[j];
      }
    }
  }
  if (want_jac)
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < i; ++j) A[i][j] = A[j][i];
  return S;
}

__global__ void __launch_bounds__(64) homography_transfer_kernel(const __grid_constant__ QcParams p) {
  extern __shared__ double s_obj[];   // [4][N]: exact board x, y | float32-rounded x, y
  const int N = p.N;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const double ox = p.obj[3 * n], oy = p.obj[3 * n + 1];
    s_obj[n] = ox;
    s_obj[N + n] = oy;
    s_obj[2 * N + n] = (double)(float)ox;
    s_obj[3 * N + n] = (double)(float)oy;
  }
  __syncthreads();
  const long long pair = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (pair >= p.C * p.F) return;
  const int cam = (int)(pair / p.F);
  const QcCam& k = p.cam[cam];
  const double2* uv = reinterpret_cast<const double2*>(p.uvs) + pair * N;
  const double2* rep = reinterpret_cast<const double2*>(p.reproj) + pair * N;
  double2* out = reinterpret_cast<double2*>(p.transformed) + pair * N;
  double* err = p.err + pair * N;
  const double* mx = s_obj + 2 * N;
  const double* my = s_obj + 3 * N;
  const double qnan = nan("");

  // ---- undistorted detections, rounded to float32 (findHomography converts its inputs)
  double Mx[kQcMaxPoints], My[kQcMaxPoints];
  bool valid = true;
  for (int n = 0; n < N; ++n) {
    const double2 o = uv[n];
    if (!(o.x == o.x && o.y == o.y)) { valid = false; break; }
    double u, v;
    qc_undistort(k, o.x, o.y, u, v);
    Mx[n] = (double)(float)u;
    My[n] = (double)(float)v;
  }
  double H[9];
  if (valid) {
    // ---- normalised DLT (HomographyEstimatorCallback::runKernel)
    double cMx = 0, cMy = 0, cmx = 0, cmy = 0;
    for (int n = 0; n < N; ++n) { cMx += Mx[n]; cMy += My[n]; cmx += mx[n]; cmy += my[n]; }
    cMx /= N; cMy /= N; cmx /= N; cmy /= N;
    double sMx = 0, sMy = 0, smx = 0, smy = 0;
    for (int n = 0; n < N; ++n) {
      sMx += fabs(Mx[n] - cMx); sMy += fabs(My[n] - cMy);
      smx += fabs(mx[n] - cmx); smy += fabs(my[n] - cmy);
    }
    const double tiny = 2.220446049250313e-16;
    if (fabs(sMx) < tiny || fabs(sMy) < tiny || fabs(smx) < tiny || fabs(smy) < tiny) valid = false;
    if (valid) {
      sMx = N / sMx; sMy = N / sMy; smx = N / smx; smy = N / smy;
      double LtL[9][9];
      for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j) LtL[i][j] = 0.0;
      for (int n = 0; n < N; ++n) {
        const double x = (mx[n] - cmx) * smx, y = (my[n] - cmy) * smy;
        const double X = (Mx[n] - cMx) * sMx, Y = (My[n] - cMy) * sMy;
        const double Lx[9] = {X, Y, 1.0, 0.0, 0.0, 0.0, -x * X, -x * Y, -x};
        const double Ly[9] = {0.0, 0.0, 0.0, X, Y, 1.0, -y * X, -y * Y, -y};
        for (int i = 0; i < 9; ++i)
          for (int j = i; j < 9; ++j) LtL[i][j] += Lx[i] * Lx[j] + Ly[i] * Ly[j];
      }
      for (int i = 0; i < 9; ++i)
        for (int j = 0; j < i; ++j) LtL[i][j] = LtL[j][i];
      double h0[9];
      smallest_eigenvector9(LtL, h0);
      // H = invHnorm * H0 * Hnorm2,  invHnorm = [[1/smx,0,cmx],[0,1/smy,cmy],[0,0,1]],  Hnorm2 = [[sMx,0,-cMx sMx],[0,sMy,-cMy sMy],[0,0,1]]
      double T[9];
      for (int j = 0; j < 3; ++j) {
        T[j] = h0[j] / smx + cmx * h0[6 + j];
        T[3 + j] = h0[3 + j] / smy + cmy * h0[6 + j];
        T[6 + j] = h0[6 + j];
      }
      for (int i = 0; i < 3; ++i) {
        H[3 * i] = T[3 * i] * sMx;
        H[3 * i + 1] = T[3 * i + 1] * sMy;
        H[3 * i + 2] = -T[3 * i] * cMx * sMx - T[3 * i + 1] * cMy * sMy + T[3 * i + 2];
      }
      const double inv = 1.0 / H[8];
      for (int i = 0; i < 9; ++i) H[i] *= inv;

      // ---- refinement: OpenCV LMSolver (Nash's Marquardt variant), <= 10 iterations, tolerance FLT_EPSILON
      double x[8], A[8][8], v[8], D[8], rmax;
      for (int i = 0; i < 8; ++i) x[i] = H[i];
      double S = qc_lm_compute(x, Mx, My, mx, my, N, true, A, v, rmax);
      for (int i = 0; i < 8; ++i) D[i] = A[i][i];
      double lam = 1.0, lc = 0.75;
      const double feps = 1.1920928955078125e-07;
      for (int it = 0; it < 10;) {
        double Ap[8][8], d[8], xd[8];
        for (int i = 0; i < 8; ++i)
          for (int j = 0; j < 8; ++j) Ap[i][j] = A[i][j] + (i == j ? lam * D[i] : 0.0);
        if (!cholesky8(Ap)) break;
        cholesky8_solve(Ap, v, d);
        for (int i = 0; i < 8; ++i) xd[i] = x[i] - d[i];
        double Atmp[8][8], vtmp[8], rmax_d;
        const double Sd = qc_lm_compute(xd, Mx, My, mx, my, N, false, Atmp, vtmp, rmax_d);
        double dS = 0.0, dv = 0.0, dmax = 0.0;
        for (int i = 0; i < 8; ++i) {
          double Ad = 0.0;
          for (int j = 0; j < 8; ++j) Ad += A[i][j] * d[j];
          dS += d[i] * (2.0 * v[i] - Ad);
          dv += d[i] * v[i];
          dmax = fmax(dmax, fabs(d[i]));
        }
        const double R = (S - Sd) / (fabs(dS) > tiny ? dS : 1.0);
        if (R > 0.75) {
          lam *= 0.5;
          if (lam < lc) lam = 0.0;
        } else if (R < 0.25) {
          double nu = (Sd - S) / (fabs(dv) > tiny ? dv : 1.0) + 2.0;
          nu = fmin(fmax(nu, 2.0), 10.0);
          if (lam == 0.0) {
            double L[8][8];
            for (int i = 0; i < 8; ++i)
              for (int j = 0; j < 8; ++j) L[i][j] = A[i][j];
            if (!cholesky8(L)) break;
            double maxval = tiny;   // max diag(A^-1) = max_i |L^-T e_i ... |: column i of A^-1 by two triangular solves
            for (int i = 0; i < 8; ++i) {
              double e[8], col[8];
              for (int j = 0; j < 8; ++j) e[j] = j == i ? 1.0 : 0.0;
              cholesky8_solve(L, e, col);
              maxval = fmax(maxval, fabs(col[i]));
            }
            lam = lc = 1.0 / maxval;
            nu *= 0.5;
          }
          lam *= nu;
        }
        if (Sd < S) {
          S = Sd;
          for (int i = 0; i < 8; ++i) x[i] = xd[i];
          S = qc_lm_compute(x, Mx, My, mx, my, N, true, A, v, rmax);
        }
        ++it;
        if (!(dmax >= feps && rmax >= feps)) break;
      }
      for (int i = 0; i < 8; ++i) H[i] = x[i];
      H[8] = 1.0;
    }
  }
  // ---- transfer (cv2.perspectiveTransform) and distance to the true corner
  for (int n = 0; n < N; ++n) {
    double2 t = make_double2(qnan, qnan);
    double e = qnan;
    if (valid) {
      const double2 r = rep[n];
      double w = H[6] * r.x + H[7] * r.y + H[8];
      w = fabs(w) > 2.220446049250313e-16 ? 1.0 / w : 0.0;
      t.x = (H[0] * r.x + H[1] * r.y + H[2]) * w;
      t.y = (H[3] * r.x + H[4] * r.y + H[5]) * w;
      const double dx = t.x - s_obj[n], dy = t.y - s_obj[N + n];
      e = sqrt(dx * dx + dy * dy);
    }
    out[n] = t;
    err[n] = e;
  }
}

}  // namespace mcba

using namespace mcba;

extern "C" int mcba_homography_transfer(int device, void* stream, const double* d_uvs, const double* d_reproj,
                                        const double* d_obj, int C, int64_t F, int N, const double* h_K,
                                        const double* h_dist, double* d_transformed, double* d_err) {
  if (!d_uvs || !d_reproj || !d_obj || !h_K || !h_dist || !d_transformed || !d_err || C < 1 || F < 0 || N < 4) {
    set_error("mcba_homography_transfer: bad arguments (at least 4 points per frame)");
    return MCBA_ERR_ARG;
  }
  if (C > kQcMaxCams || N > kQcMaxPoints) {
    set_error("mcba_homography_transfer: at most 32 cameras and 128 points per frame");
    return MCBA_ERR_ARG;
  }
  MCBA_CUDA(cudaSetDevice(device));
  if (F == 0) return MCBA_OK;
  QcParams p;
  p.uvs = d_uvs; p.reproj = d_reproj; p.obj = d_obj; p.transformed = d_transformed; p.err = d_err;
  p.C = C; p.N = N; p.F = F;
  for (int c = 0; c < C; ++c) {
    const double* K = h_K + 9 * c;
    const double* d = h_dist + 5 * c;
    p.cam[c] = QcCam{K[0], K[4], K[2], K[5], K[1], d[0], d[1], d[2], d[3], d[4]};
  }
  const long long pairs = (long long)C * F;
  const int threads = 64;
  homography_transfer_kernel<<<(unsigned)((pairs + threads - 1) / threads), threads, sizeof(double) * 4 * N, (cudaStream_t)stream>>>(p);
  MCBA_CUDA(cudaGetLastError());
  return MCBA_OK;
}

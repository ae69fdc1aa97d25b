// K3a: the damped reduced camera system  (S + lambda D_c^2) delta_cam = -b  solved by ONE CTA
// (replaces the cuSOLVER potrf + potrs + damping launches of round 1: five library / helper kernels,
// ~90 us of launch and library latency per LM iteration for a 72 x 72 system, by one ~10 us kernel).
// Math: SURVEY.md Appendix A ("Schur form"), scipy's x_scale='jac' analogue (common.py:598-610).
//
// The lower triangle lives packed in shared memory (12C <= 192: 161 KB), augmented by one more row
// that holds the right-hand side -b: a blocked right-looking Cholesky of the augmented matrix leaves
// y = L^-1 (-b) in that row, so the forward substitution costs nothing extra.  Per panel of 8 columns:
//   1. one thread factors the 8 x 8 diagonal block in registers (the sequential rsqrt chain is the
//      critical path of any Cholesky; everything else hangs off it),
//   2. one thread per row below solves its 8 entries against the block (broadcast reads of L11),
//   3. the trailing update C_ij -= L_i L_j^T runs on the FP64 tensor path, one warp per 8 x 8 tile
//      (two DMMA m8n8k4 per tile, same fragment mapping as the SYRK in k2_schur.cu).
// The backward substitution L^T delta = y walks the blocks from the last to the first.
// A non-positive (or NaN) pivot is reported in info[0] like potrf's (the LM loop then rejects the step).
#include <cstdlib>

#include "k2_common.cuh"

namespace mcba {

constexpr int kSolveThreads = 512;
constexpr int kSolveMaxN = 192;   // 16 cameras; larger systems take the library path

__host__ __device__ __forceinline__ int pk(int i, int j) { return i * (i + 1) / 2 + j; }   // j <= i

__device__ __forceinline__ void solve_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct SolveParams {
  const double* red;    // packed reduced system [S | b | g | diag | ...]
  long long offS, offB, offDiag;
  int nc;
  double lambda;
  double* D2cam;        // running max of diag(U) (Marquardt scaling), updated here
  double* dcam;         // out: delta_cam (true basis), nc
  int* info;            // out: 0, or 1 + index of the first non-positive pivot
};

__global__ void __launch_bounds__(kSolveThreads, 1) solve_reduced_kernel(const SolveParams p) {
  extern __shared__ double A[];                 // packed lower triangle of the augmented matrix, n1p rows
  __shared__ double s_inv[kSolveMaxN + 8];      // 1 / L_jj
  __shared__ double s_y[kSolveMaxN + 8];
  __shared__ int s_info;
  const int nc = p.nc, n1p = (nc + 1 + 7) & ~7, nb = n1p / 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  if (tid == 0) s_info = 0;

  // ---- load: S + lambda D^2 (lower triangle), row nc = -b, unit diagonal on the padding
  for (int e = tid; e < n1p * n1p; e += blockDim.x) {
    const int i = e / n1p, j = e - i * n1p;
    if (j > i) continue;
    double v;
    if (i < nc) {
      v = p.red[p.offS + (size_t)i * nc + j];
      if (i == j) {
        double d2 = fmax(p.D2cam[i], p.red[p.offDiag + i]);
        p.D2cam[i] = d2;
        if (d2 == 0.0) d2 = 1.0;
        v = fma(p.lambda, d2, v);
      }
    } else if (i == nc) {
      v = j < nc ? -p.red[p.offB + j] : 1.0;
    } else {
      v = i == j ? 1.0 : 0.0;
    }
    A[pk(i, j)] = v;
  }
  __syncthreads();

  // ---- blocked Cholesky of the leading nc columns
  for (int k0 = 0; k0 < nc; k0 += 8) {
    const int w = nc - k0 < 8 ? nc - k0 : 8;   // pivot columns of this panel
    if (tid == 0) {
      double a[36];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c <= r) a[pk(r, c)] = A[pk(k0 + r, k0 + c)];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < w) {
          const double d = a[pk(j, j)];
          if (!(d > 0.0) && s_info == 0) s_info = k0 + j + 1;
          const double inv = rsqrt(d);
          s_inv[k0 + j] = inv;
          a[pk(j, j)] = d * inv;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i > j) a[pk(i, j)] *= inv;
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (i > j && c > j && c <= i) a[pk(i, c)] = fma(-a[pk(i, j)], a[pk(c, j)], a[pk(i, c)]);
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c <= r) A[pk(k0 + r, k0 + c)] = a[pk(r, c)];
    }
    __syncthreads();
    // rows below the block: x L11^T = a  (one thread per row; L11 is read as broadcasts)
    for (int i = k0 + 8 + tid; i < n1p; i += blockDim.x) {
      double x[8];
      double* row = A + pk(i, k0);
#pragma unroll
      for (int c = 0; c < 8; ++c) x[c] = row[c];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < w) {
          double t = x[j];
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < j) t = fma(-x[c], A[pk(k0 + j, k0 + c)], t);
          x[j] = t * s_inv[k0 + j];
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) row[c] = x[c];
    }
    __syncthreads();
    // trailing update on the FP64 tensor path: tile (ib, jb), kb < jb <= ib < nb
    const int kb = k0 >> 3, m = nb - kb - 1, nt = m * (m + 1) / 2;
    for (int t = warp; t < nt; t += nwarps) {
      int r = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
      while (r * (r + 1) / 2 > t) --r;
      while ((r + 1) * (r + 2) / 2 <= t) ++r;
      const int c = t - r * (r + 1) / 2;
      const int i0 = (kb + 1 + r) * 8, j0 = (kb + 1 + c) * 8;
      const int fr = lane >> 2, fk = k0 + (lane & 3);
      const double a0 = -A[pk(i0 + fr, fk)], a1 = -A[pk(i0 + fr, fk + 4)];
      const double b0 = A[pk(j0 + fr, fk)], b1 = A[pk(j0 + fr, fk + 4)];
      const int cr = i0 + fr, cc = j0 + 2 * (lane & 3);
      const bool v0 = cc <= cr, v1 = cc + 1 <= cr;
      double c0 = v0 ? A[pk(cr, cc)] : 0.0, c1 = v1 ? A[pk(cr, cc + 1)] : 0.0;
      solve_dmma(c0, c1, a0, b0);
      solve_dmma(c0, c1, a1, b1);
      if (v0) A[pk(cr, cc)] = c0;
      if (v1) A[pk(cr, cc + 1)] = c1;
    }
    __syncthreads();
  }

  // ---- backward substitution L^T delta = y, y = row nc of the factor
  for (int i = tid; i < nc; i += blockDim.x) s_y[i] = A[pk(nc, i)];
  __syncthreads();
  for (int k0 = ((nc - 1) >> 3) << 3; k0 >= 0; k0 -= 8) {
    const int w = nc - k0 < 8 ? nc - k0 : 8;
    if (tid == 0) {
      double d[8];
#pragma unroll
      for (int j = 7; j >= 0; --j) {
        if (j < w) {
          double t = s_y[k0 + j];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i > j && i < w) t = fma(-A[pk(k0 + i, k0 + j)], d[i], t);
          d[j] = t * s_inv[k0 + j];
          s_y[k0 + j] = d[j];
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < k0; i += blockDim.x) {
      double t = s_y[i];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < w) t = fma(-A[pk(k0 + j, i)], s_y[k0 + j], t);
      s_y[i] = t;
    }
    __syncthreads();
  }
  for (int i = tid; i < nc; i += blockDim.x) p.dcam[i] = s_y[i];
  if (tid == 0) p.info[0] = s_info;
}

// ------------------------------------------------------------------ library path (12C > 192, or MCBA_CUSOLVER=1 for A/B runs)
__global__ void damp_kernel(const double* __restrict__ red, long long offS, long long offB, long long offDiag,
                            int nc, double lambda, double* __restrict__ D2cam, double* __restrict__ Sd,
                            double* __restrict__ rhs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc * nc) return;
  const int r = i / nc, q = i % nc;
  double v = red[offS + i];
  if (r == q) {
    double d2 = fmax(D2cam[r], red[offDiag + r]);
    D2cam[r] = d2;
    if (d2 == 0.0) d2 = 1.0;
    v = fma(lambda, d2, v);
    rhs[r] = -red[offB + r];
  }
  Sd[i] = v;
}

bool solve_uses_library(int nc) {
  static const bool force = getenv("MCBA_CUSOLVER") != nullptr;
  return force || nc > kSolveMaxN;
}

int solve_reduced(mcba_handle* h, double lambda) {
  const Layout& L = h->L;
  const int nc = L.nc;
  if (!solve_uses_library(nc)) {
    SolveParams p;
    p.red = h->d_red; p.offS = L.offS; p.offB = L.offB; p.offDiag = L.offDiag; p.nc = nc; p.lambda = lambda;
    p.D2cam = h->d_D2cam; p.dcam = h->d_dcam; p.info = h->d_info;
    const int n1p = (nc + 1 + 7) & ~7;
    const size_t smem = sizeof(double) * (size_t)n1p * (n1p + 1) / 2;
    MCBA_CUDA(cudaFuncSetAttribute(solve_reduced_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    solve_reduced_kernel<<<1, kSolveThreads, smem, h->stream>>>(p);
    h->launches++;
    MCBA_CUDA(cudaGetLastError());
    return MCBA_OK;
  }
  if (!h->solver) {   // created on first use: the common sizes never touch the library
    if (cusolverDnCreate(&h->solver) != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnCreate failed"); return MCBA_ERR_SOLVER; }
    if (cusolverDnDpotrf_bufferSize(h->solver, CUBLAS_FILL_MODE_LOWER, nc, h->d_Sd, nc, &h->lwork) != CUSOLVER_STATUS_SUCCESS) {
      set_error("cusolverDnDpotrf_bufferSize failed");
      return MCBA_ERR_SOLVER;
    }
    MCBA_CUDA(cudaMalloc((void**)&h->d_work, sizeof(double) * (size_t)(h->lwork > 0 ? h->lwork : 1)));
  }
  cusolverDnSetStream(h->solver, h->stream);
  damp_kernel<<<(nc * nc + 255) / 256, 256, 0, h->stream>>>(h->d_red, L.offS, L.offB, L.offDiag, nc, lambda,
                                                            h->d_D2cam, h->d_Sd, h->d_dcam);
  h->launches++;
  MCBA_CUDA(cudaGetLastError());
  cusolverStatus_t st = cusolverDnDpotrf(h->solver, CUBLAS_FILL_MODE_LOWER, nc, h->d_Sd, nc, h->d_work, h->lwork, h->d_info);
  if (st != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnDpotrf failed"); return MCBA_ERR_SOLVER; }
  st = cusolverDnDpotrs(h->solver, CUBLAS_FILL_MODE_LOWER, nc, 1, h->d_Sd, nc, h->d_dcam, nc, h->d_info + 1);
  if (st != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnDpotrs failed"); return MCBA_ERR_SOLVER; }
  h->launches += 4;  // potrf + potrs kernels (library; approximate count)
  return MCBA_OK;
}

}  // namespace mcba

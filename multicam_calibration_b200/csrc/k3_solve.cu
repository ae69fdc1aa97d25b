// K3a: the damped reduced camera system  (S + lambda D_c^2) delta_cam = -b  solved by ONE CTA
// (replaces the cuSOLVER potrf + potrs + damping launches of round 1: five library / helper kernels,
// ~90 us of launch and library latency per LM iteration for a 72 x 72 system).
// Math: SURVEY.md Appendix A ("Schur form"), scipy's x_scale='jac' analogue (common.py:598-610).
//
// The lower triangle lives packed in shared memory (12C <= 192: 161 KB), augmented by one more row
// that holds the right-hand side -b: a right-looking blocked Cholesky of the augmented matrix leaves
// y = L^-1 (-b) in that row, so the forward substitution costs nothing extra.  The sequential chain
// of a Cholesky (pivot -> rsqrt -> scale -> update -> next pivot, 12C links) is what bounds a system
// this small, so everything else is taken off it:
//   * panel of 8 columns: ONE thread per row of the block column; every thread factors the 8 x 8
//     diagonal block redundantly in registers (broadcast reads) and solves its own row against it in
//     the same unrolled loop -- no hand-over between a factoring thread and the row threads, and the
//     reciprocal square root is a MUFU seed with two corrections instead of the library call;
//   * trailing matrix: the 8 x 8 tiles are dealt out to the 16 warps ONCE and their accumulators stay
//     in registers for the whole factorisation (FP64 tensor path, DMMA m8n8k4, same fragment mapping
//     as the SYRK in k2_schur.cu); a tile goes back to shared memory only when its block column is
//     the next panel.  Two CTA barriers per panel.
//   * L^T delta = y: one warp, y in registers, 12C steps of (broadcast delta_j, row j of L is
//     contiguous in the packed layout): no CTA barrier at all.
// A non-positive (or NaN) pivot is reported in info[0] like potrf's (the LM loop then rejects the step).
#include <cstdio>
#include <cstdlib>

#include "k2_common.cuh"

namespace mcba {

#ifdef MCBA_SOLVE_TIMING
__device__ long long g_solve_clk[8];
#define MCBA_STAMP(i) do { if (threadIdx.x == 0) g_solve_clk[i] = clock64(); } while (0)
#else
#define MCBA_STAMP(i) do {} while (0)
#endif

constexpr int kSolveThreads = 512;
constexpr int kSolveWarps = kSolveThreads / 32;
constexpr int kSolveMaxN = 192;   // 16 cameras; larger systems take the library path

__host__ __device__ __forceinline__ int pk(int i, int j) { return i * (i + 1) / 2 + j; }   // j <= i

__device__ __forceinline__ void solve_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 1/sqrt(t), t > 0 in the normal range: MUFU.RSQ64H seed, one cubic and one quadratic correction
// (non-positive or NaN t gives NaN / inf, which the pivot test reports)
__device__ __forceinline__ double solve_rsqrt(double t) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t));
  double e = fma(-t, y * y, 1.0);
  y = fma(y * e, fma(0.375, e, 0.5), y);
  e = fma(-t, y * y, 1.0);
  return fma(0.5 * y, e, y);
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

// kT: trailing tiles per warp (3: up to 6 cameras, 8: up to 10, 19: up to 16 -- 24 * 25 / 2 = 300 tiles of the
// 200-row augmented system over 16 warps; that variant spills part of its accumulators to local memory)
template <int kT>
__global__ void __launch_bounds__(kSolveThreads, 1) solve_reduced_kernel(const SolveParams p) {
  extern __shared__ double A[];                 // packed lower triangle of the augmented matrix, n1p rows
  __shared__ double s_inv[kSolveMaxN + 8];      // 1 / L_jj
  __shared__ int s_info;
  const int nc = p.nc, n1p = (nc + 1 + 7) & ~7, nb = n1p / 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_info = 0;
  MCBA_STAMP(0);

  // ---- load: S (lower triangle), row nc = -b, unit diagonal on the padding; 8 independent loads per
  //      thread and pass.  The damping of the diagonal is added after the barrier by one thread per row.
  double damp = 0.0;
  if (tid < nc) {
    double d2 = fmax(p.D2cam[tid], p.red[p.offDiag + tid]);
    p.D2cam[tid] = d2;
    if (d2 == 0.0) d2 = 1.0;
    damp = p.lambda * d2;
  }
  const int n_pk = n1p * (n1p + 1) / 2;
  for (int base = 0; base < n_pk; base += 8 * kSolveThreads) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = base + u * kSolveThreads + tid;
      v[u] = 0.0;
      if (e < n_pk) {
        int i = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
        while (i * (i + 1) / 2 > e) --i;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        const int j = e - i * (i + 1) / 2;
        if (i < nc) v[u] = p.red[p.offS + (size_t)i * nc + j];
        else if (i == nc) v[u] = j < nc ? -p.red[p.offB + j] : 1.0;
        else v[u] = i == j ? 1.0 : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = base + u * kSolveThreads + tid;
      if (e < n_pk) A[e] = v[u];
    }
  }
  __syncthreads();
  MCBA_STAMP(1);
  if (tid < nc) A[pk(tid, tid)] += damp;

  // ---- this warp's tiles of the trailing matrix: tile (ib, jb), 1 <= jb <= ib < nb, numbered block
  //      column by block column and dealt round-robin (the tiles still active at any panel are then
  //      spread evenly over the warps); accumulators in registers until the tile's column is the panel
  int tij[kT];
  double acc[kT][2];
  const int n_tiles = (nb - 1) * nb / 2;
  const int fr = lane >> 2, fc = 2 * (lane & 3);
  __syncthreads();
#pragma unroll
  for (int s = 0; s < kT; ++s) {
    const int t = warp + s * kSolveWarps;
    tij[s] = -1;
    acc[s][0] = acc[s][1] = 0.0;
    if (t < n_tiles) {
      // column jb holds nb - jb tiles; find it
      int jb = 1, rem = t;
      while (rem >= nb - jb) { rem -= nb - jb; ++jb; }
      const int ib = jb + rem;
      tij[s] = (ib << 8) | jb;
      const int cr = ib * 8 + fr, cc = jb * 8 + fc;
      if (cc <= cr) acc[s][0] = A[pk(cr, cc)];
      if (cc + 1 <= cr) acc[s][1] = A[pk(cr, cc + 1)];
    }
  }

  MCBA_STAMP(2);
  // ---- blocked Cholesky of the leading nc columns
  for (int k0 = 0; k0 < nc; k0 += 8) {
    const int w = nc - k0 < 8 ? nc - k0 : 8;   // pivot columns of this panel
    const int i = k0 + tid;                    // one thread per row of the block column
    const int rr = tid;                        // < 8: a row of the diagonal block itself
    double a[36];
    int bad = 0;
    if (i < n1p) {
      double x[8];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c <= r) a[pk(r, c)] = A[pk(k0 + r, k0 + c)];
      double* row = A + pk(i, k0);
      if (rr >= 8) {
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = row[c];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < w) {
          const double d = a[pk(j, j)];
          if (!(d > 0.0) && bad == 0) bad = k0 + j + 1;
          const double inv = solve_rsqrt(d);
          a[pk(j, j)] = d * inv;
#pragma unroll
          for (int r = 0; r < 8; ++r)
            if (r > j) a[pk(r, j)] *= inv;
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (r > j && c > j && c <= r) a[pk(r, c)] = fma(-a[pk(r, j)], a[pk(c, j)], a[pk(r, c)]);
          if (rr >= 8) {   // own row against column j of the factor
            double t = x[j];
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c < j) t = fma(-x[c], a[pk(j, c)], t);
            x[j] = t * inv;
          }
          if (rr == 0) s_inv[k0 + j] = inv;
        }
      }
      if (rr >= 8) {
#pragma unroll
        for (int c = 0; c < 8; ++c) row[c] = x[c];
      }
    }
    __syncthreads();
    // the factored diagonal block goes back only now: until the barrier the other warps were still
    // reading the unfactored one (nobody reads it during the trailing update)
    if (rr < 8) {
      if (rr == 0 && bad && s_info == 0) s_info = bad;
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r == rr) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c <= r) A[pk(k0 + r, k0 + c)] = a[pk(r, c)];
        }
    }
    // trailing update on the FP64 tensor path: C_ij -= L_i L_j^T for this warp's tiles right of the panel
    const int kb = k0 >> 3;
    const int fk = k0 + (lane & 3);
#pragma unroll
    for (int s = 0; s < kT; ++s) {
      const int jb = tij[s] & 0xff, ib = tij[s] >> 8;
      if (tij[s] >= 0 && jb > kb) {     // warp-uniform
        const double a0 = -A[pk(ib * 8 + fr, fk)], a1 = -A[pk(ib * 8 + fr, fk + 4)];
        const double b0 = A[pk(jb * 8 + fr, fk)], b1 = A[pk(jb * 8 + fr, fk + 4)];
        solve_dmma(acc[s][0], acc[s][1], a0, b0);
        solve_dmma(acc[s][0], acc[s][1], a1, b1);
        if (jb == kb + 1) {             // the next panel: back to shared memory
          const int cr = ib * 8 + fr, cc = jb * 8 + fc;
          if (cc <= cr) A[pk(cr, cc)] = acc[s][0];
          if (cc + 1 <= cr) A[pk(cr, cc + 1)] = acc[s][1];
        }
      }
    }
    __syncthreads();
  }

  MCBA_STAMP(3);
  // ---- backward substitution L^T delta = y (y = row nc of the factor): warp 0, y in registers
  if (warp == 0) {
    constexpr int kSlots = (kSolveMaxN + 31) / 32;
    double y[kSlots];
#pragma unroll
    for (int s = 0; s < kSlots; ++s) y[s] = (lane + 32 * s) < nc ? A[pk(nc, lane + 32 * s)] : 0.0;
    for (int j = nc - 1; j >= 0; --j) {
      const int sj = j >> 5, lj = j & 31;
      double yj = 0.0;
#pragma unroll
      for (int s = 0; s < kSlots; ++s)
        if (s == sj) yj = y[s];
      const double dj = __shfl_sync(0xffffffffu, yj, lj) * s_inv[j];
      const double* Lj = A + pk(j, 0);
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        const int i = lane + 32 * s;
        if (s <= sj) {
          if (i < j) y[s] = fma(-Lj[i], dj, y[s]);
          else if (i == j) y[s] = dj;
        }
      }
    }
#pragma unroll
    for (int s = 0; s < kSlots; ++s)
      if (lane + 32 * s < nc) p.dcam[lane + 32 * s] = y[s];
    if (lane == 0) p.info[0] = s_info;
    MCBA_STAMP(4);
  }
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
    const int nb = n1p / 8, per_warp = ((nb - 1) * nb / 2 + kSolveWarps - 1) / kSolveWarps;
#define MCBA_SOLVE(T)                                                                                            \
  do {                                                                                                           \
    MCBA_CUDA(cudaFuncSetAttribute(solve_reduced_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    solve_reduced_kernel<T><<<1, kSolveThreads, smem, h->stream>>>(p);                                           \
  } while (0)
    if (per_warp <= 3) MCBA_SOLVE(3);
    else if (per_warp <= 8) MCBA_SOLVE(8);
    else MCBA_SOLVE(19);
#undef MCBA_SOLVE
#ifdef MCBA_SOLVE_TIMING
    {
      long long c[8];
      cudaStreamSynchronize(h->stream);
      cudaMemcpyFromSymbol(c, g_solve_clk, sizeof(c));
      fprintf(stderr, "solve phases (cycles): load %lld  setup %lld  panels %lld  backward %lld\n", c[1] - c[0], c[2] - c[1], c[3] - c[2], c[4] - c[3]);
    }
#endif
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

// K2a: fused residual + analytic Jacobian + robust scaling + per-frame Schur
// elimination (replaces scipy's finite-difference Jacobian + LSMR for the path
// bundle_adjustment.py:299-313; math in SURVEY.md Appendix A).
//
// Mapping: one frame tile = 32 consecutive frames; one warp per camera; lane =
// frame.  A lane owns one (camera, frame) pair and walks its N board corners,
// reading the tiled observation SoA with fully coalesced 16-byte loads
// ([tile][camera][corner][lane] -> 512 contiguous bytes per warp load).
// Per pair it accumulates the raw 12x12 Gauss-Newton block A = sum w a a^T and
// q = -sum w' f a in registers (no cross-lane traffic in the hot loop), then
//   * adds (A, q) to the camera's U / gradient through a register transpose-
//     reduction (3 values per lane),
//   * maps the extrinsic part to the pose basis (W' = A[:,ext] E', V'' = E'^T A_ee E'),
//   * after a CTA-wide exchange of V'' factors V_f + lambda D_f^2 = L L^T (6x6, per lane),
//   * writes Z_cf = W_cf L^-T, y_f = L^-1 g_f, L^-1 for the SYRK and back-substitution.
// Nothing per-observation is written: algorithmic HBM traffic is 16 B/observation.
#include <cstdlib>

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

template <bool kMulti>
__global__ void __launch_bounds__(256, 1) k2_frames_kernel(const K2Params p) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = p.C, N = p.N, NW = p.nwarps, nc = 12 * C;
  double* s_obj = smem;                               // N*3
  double* s_V = s_obj + ((3 * N + 1) & ~1);           // [NW][27][32]
  double* s_U = s_V + (size_t)NW * 27 * 32;           // [C][96]
  for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) s_obj[i] = p.obj[i];
  for (int i = threadIdx.x; i < C * kUPad; i += blockDim.x) s_U[i] = 0.0;
  double cost_acc = 0.0, sumsq_acc = 0.0, cnt_acc = 0.0, gmax = 0.0;
  __syncthreads();

  for (long long tile = blockIdx.x; tile < p.nTiles; tile += gridDim.x) {
    const long long f = tile * kTile + lane;
    const bool fvalid = f < p.F;
    double pose[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) pose[i] = fvalid ? p.x[(size_t)nc + f * 6 + i] : 0.0;
    double Rp[9];
    rodrigues(pose, Rp);
    double B[72];  // W' = A[:,ext] E'  (raw camera rows x pose-perturbation columns)

    const int ngroups = kMulti ? p.ngroups : 1;
    for (int g = 0; g < ngroups; ++g) {
      const int c = g * NW + warp;
      if (c < C) {
        const CamConst& cam = p.cams[c];
        Intr in{cam.fx, cam.fy, cam.cx, cam.cy, cam.k1, cam.k2};
        double Rc[9], Rcf[9], tcf[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) Rc[i] = cam.R[i];
        mat3_mul(Rc, Rp, Rcf);
        mat3_vec(Rc, pose + 3, tcf);
        tcf[0] += cam.t[0];
        tcf[1] += cam.t[1];
        tcf[2] += cam.t[2];

        double acc[kUPad];
#pragma unroll
        for (int i = 0; i < kUPad; ++i) acc[i] = 0.0;

        // ---------------- phase 1: walk the board corners ----------------
        const double2* ob = p.obs + ((size_t)(tile * C + c) * N) * kTile + lane;
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

        // ---------------- phase 2a: camera block U, gradient ----------------
        {
          const double r0 = lane_transpose_sum32<0>(acc, lane);
          const double r1 = lane_transpose_sum32<32>(acc, lane);
          const double r2 = lane_transpose_sum32<64>(acc, lane);
          double* u = s_U + c * kUPad;   // this warp is the only writer of camera c in this CTA
          u[lane] += r0;
          u[32 + lane] += r1;
          u[64 + lane] += r2;
        }

        // ---------------- phase 2b: extrinsic columns -> pose basis ----------------
        // E' = [[Rc, 0], [K, Rc]],  K = [t_cf]x Rc   (rows: raw [m | G], cols: [eps | tau])
        double K[9];
        cross_mat3(tcf, Rc, K);
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          double am[3], ag[3];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            am[j] = acc[sym12(i, 6 + j)];
            ag[j] = acc[sym12(i, 9 + j)];
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            B[i * 6 + k] = am[0] * Rc[k] + am[1] * Rc[3 + k] + am[2] * Rc[6 + k] + ag[0] * K[k] +
                           ag[1] * K[3 + k] + ag[2] * K[6 + k];
            B[i * 6 + 3 + k] = ag[0] * Rc[k] + ag[1] * Rc[3 + k] + ag[2] * Rc[6 + k];
          }
        }
        // V'' = E'^T B[ext rows]  (upper triangle) and g'' = E'^T q_ext -> shared
        double* sv = s_V + (size_t)warp * 27 * 32 + lane;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
#pragma unroll
          for (int n = a; n < 6; ++n) {
            double v;
            if (a < 3) {
              v = Rc[a] * B[36 + n] + Rc[3 + a] * B[42 + n] + Rc[6 + a] * B[48 + n] +
                  K[a] * B[54 + n] + K[3 + a] * B[60 + n] + K[6 + a] * B[66 + n];
            } else {
              v = Rc[a - 3] * B[54 + n] + Rc[3 + a - 3] * B[60 + n] + Rc[6 + a - 3] * B[66 + n];
            }
            if (kMulti && g > 0) sv[tri6(a, n) * 32] += v; else sv[tri6(a, n) * 32] = v;
          }
          double gv;
          if (a < 3) {
            gv = Rc[a] * acc[kQ + 6] + Rc[3 + a] * acc[kQ + 7] + Rc[6 + a] * acc[kQ + 8] +
                 K[a] * acc[kQ + 9] + K[3 + a] * acc[kQ + 10] + K[6 + a] * acc[kQ + 11];
          } else {
            gv = Rc[a - 3] * acc[kQ + 9] + Rc[3 + a - 3] * acc[kQ + 10] + Rc[6 + a - 3] * acc[kQ + 11];
          }
          if (kMulti && g > 0) sv[(21 + a) * 32] += gv; else sv[(21 + a) * 32] = gv;
        }
        if (kMulti && fvalid) {  // stage W' through the Z buffer (camera count exceeds the warps)
          double* z = p.Z + (size_t)f * 6 * nc + c * 12;
#pragma unroll
          for (int k = 0; k < 6; ++k)
#pragma unroll
            for (int i = 0; i < 12; ++i) z[(size_t)k * nc + i] = B[i * 6 + k];
        }
      } else if (!kMulti || g == 0) {
        // idle warp slot (C < warps): contribute zeros
        double* sv = s_V + (size_t)warp * 27 * 32 + lane;
#pragma unroll
        for (int i = 0; i < 27; ++i) sv[i * 32] = 0.0;
      }
    }
    __syncthreads();

    // ---------------- per-frame step (every warp, redundantly; lane = frame) ----------------
    double Jl[9];
    so3_left_jacobian(pose, Jl);
    double Linv[21];  // L^-1, lower triangle, row-major packed: idx(i,j) = i(i+1)/2 + j
    double yv[6];
    {
      double Vpp[21], gpp[6];
#pragma unroll
      for (int i = 0; i < 21; ++i) Vpp[i] = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) gpp[i] = 0.0;
      for (int w = 0; w < NW; ++w) {
        const double* sv = s_V + (size_t)w * 27 * 32 + lane;
#pragma unroll
        for (int i = 0; i < 21; ++i) Vpp[i] += sv[i * 32];
#pragma unroll
        for (int i = 0; i < 6; ++i) gpp[i] += sv[(21 + i) * 32];
      }
      // true pose basis: P' = blkdiag(J_l(rho), I):  V = P'^T V'' P',  g = P'^T g''
      double V[36], gp[6];
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
          const double vt = Jl[i] * Vpp[tri6(0, 3 + j)] + Jl[3 + i] * Vpp[tri6(1, 3 + j)] +
                            Jl[6 + i] * Vpp[tri6(2, 3 + j)];
          V[6 * i + 3 + j] = vt;
          V[6 * (3 + j) + i] = vt;
          V[6 * (3 + i) + 3 + j] = Vpp[sym6(3 + i, 3 + j)];
        }
        gp[i] = Jl[i] * gpp[0] + Jl[3 + i] * gpp[1] + Jl[6 + i] * gpp[2];
        gp[3 + i] = gpp[3 + i];
      }
      // damping with the running Marquardt scaling D_f^2 = max over evaluations of diag(V_f)
      double* d2p = p.D2pose + (size_t)tile * 6 * 32 + lane;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        double d2 = fmax(d2p[i * 32], V[7 * i]);
        if (warp == 0) d2p[i * 32] = d2;
        if (d2 == 0.0) d2 = 1.0;
        V[7 * i] = fma(p.lambda, d2, V[7 * i]);
        gmax = fmax(gmax, fabs(gp[i]));
      }
      if (warp == 0 && fvalid) {
#pragma unroll
        for (int i = 0; i < 6; ++i) p.gpose[(size_t)f * 6 + i] = gp[i];
      }
      // Cholesky V = L L^T (in place, lower), guarded for empty frames at lambda = 0
      double Lm[21];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        double s = V[7 * j];
#pragma unroll
        for (int k = 0; k < j; ++k) s -= Lm[j * (j + 1) / 2 + k] * Lm[j * (j + 1) / 2 + k];
        const double inv = s > 0.0 ? rsqrt(s) : 0.0;
        Lm[j * (j + 1) / 2 + j] = inv;  // store 1/L_jj on the diagonal
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
          double t = V[6 * i + j];
#pragma unroll
          for (int k = 0; k < j; ++k) t -= Lm[i * (i + 1) / 2 + k] * Lm[j * (j + 1) / 2 + k];
          Lm[i * (i + 1) / 2 + j] = t * inv;
        }
      }
      // L^-1 (lower): Linv_ii = 1/L_ii ; Linv_ij = -(sum_{k=j}^{i-1} L_ik Linv_kj) / L_ii
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
      if (warp == 0) {
        double* lo = p.Linv + (size_t)tile * 21 * 32 + lane;
#pragma unroll
        for (int i = 0; i < 21; ++i) lo[i * 32] = Linv[i];
        if (fvalid) {
          double* yo = p.y + (size_t)f * 6;
#pragma unroll
          for (int i = 0; i < 6; i += 2) *reinterpret_cast<double2*>(yo + i) = make_double2(yv[i], yv[i + 1]);
        }
      }
    }

    // ---------------- phase 3: Z_cf = (W' P') L^-T ----------------
    for (int g = 0; g < ngroups; ++g) {
      const int c = g * NW + warp;
      if (c < C && fvalid) {
        double* z = p.Z + (size_t)f * 6 * nc + c * 12;
        if (kMulti) {
#pragma unroll
          for (int k = 0; k < 6; ++k)
#pragma unroll
            for (int i = 0; i < 12; ++i) B[i * 6 + k] = z[(size_t)k * nc + i];
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          double w[6];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            w[k] = B[i * 6] * Jl[k] + B[i * 6 + 1] * Jl[3 + k] + B[i * 6 + 2] * Jl[6 + k];
            w[3 + k] = B[i * 6 + 3 + k];
          }
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            double t = 0.0;
#pragma unroll
            for (int j = 0; j <= k; ++j) t += w[j] * Linv[k * (k + 1) / 2 + j];
            B[i * 6 + k] = t;
          }
        }
#pragma unroll
        for (int k = 0; k < 6; ++k)
#pragma unroll
          for (int i = 0; i < 12; i += 2)
            *reinterpret_cast<double2*>(z + (size_t)k * nc + i) = make_double2(B[i * 6 + k], B[(i + 1) * 6 + k]);
      }
    }
    __syncthreads();  // s_V is rewritten by the next tile
  }

  // ---------------- CTA epilogue: partial sums ----------------
  __syncthreads();
  double* pu = p.partU + (size_t)blockIdx.x * C * kUPad;
  for (int i = threadIdx.x; i < C * kUPad; i += blockDim.x) pu[i] = s_U[i];
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    cost_acc += __shfl_xor_sync(0xffffffffu, cost_acc, off);
    sumsq_acc += __shfl_xor_sync(0xffffffffu, sumsq_acc, off);
    cnt_acc += __shfl_xor_sync(0xffffffffu, cnt_acc, off);
    gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, off));
  }
  double* s_red = s_V;  // reuse
  if (lane == 0) {
    s_red[warp * 4 + 0] = cost_acc;
    s_red[warp * 4 + 1] = sumsq_acc;
    s_red[warp * 4 + 2] = cnt_acc;
    s_red[warp * 4 + 3] = gmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, cn = 0, gm = 0;
    for (int w = 0; w < NW; ++w) {
      a += s_red[w * 4];
      b += s_red[w * 4 + 1];
      cn += s_red[w * 4 + 2];
      gm = fmax(gm, s_red[w * 4 + 3]);
    }
    double* ps = p.partS + (size_t)blockIdx.x * kRsNum;
    ps[kRsCost] = 0.5 * a;
    ps[kRsSumSq] = b;
    ps[kRsCount] = cn;
    ps[kRsGmaxPose] = gm;  // every warp computed the same per-frame gradients
  }
}

size_t k2_frames_smem(int C, int N, int nwarps) {
  return sizeof(double) * (((3 * N + 1) & ~1) + (size_t)nwarps * 27 * 32 + (size_t)C * kUPad);
}

int launch_k2_frames(mcba_handle* h, const double* x, double lambda, int loss, double f_scale) {
  const Layout& L = h->L;
  K2Params p;
  p.C = L.C;
  p.N = L.N;
  p.nwarps = L.C < 8 ? L.C : 8;
  p.ngroups = (L.C + p.nwarps - 1) / p.nwarps;
  p.F = L.F;
  p.nTiles = L.nTiles;
  p.obs = h->d_obs_tiled;
  p.obj = h->d_obj;
  p.x = x;
  p.cams = h->d_cams;
  p.lambda = lambda;
  p.loss = loss;
  p.inv_c = 1.0 / f_scale;
  p.c2 = f_scale * f_scale;
  p.Z = h->d_Z;
  p.Linv = h->d_Linv;
  p.y = h->d_y;
  p.gpose = h->d_gpose;
  p.D2pose = h->d_D2pose;
  p.partU = h->d_partU;
  p.partS = h->d_partS;
  // C <= 6: warp-specialised producer/consumer kernel (k2_frames_ws.cu); MCBA_K2_LEGACY=1 forces this one
  static const bool legacy = getenv("MCBA_K2_LEGACY") != nullptr;
  if (L.C <= 6 && !legacy) return launch_k2_frames_ws(h, p);
  const size_t smem = k2_frames_smem(L.C, L.N, p.nwarps);
  const int grid = h->grid_frames;
  if (p.ngroups == 1) {
    MCBA_CUDA(cudaFuncSetAttribute(k2_frames_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k2_frames_kernel<false><<<grid, p.nwarps * 32, smem, h->stream>>>(p);
  } else {
    MCBA_CUDA(cudaFuncSetAttribute(k2_frames_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k2_frames_kernel<true><<<grid, p.nwarps * 32, smem, h->stream>>>(p);
  }
  h->launches++;
  MCBA_CUDA(cudaGetLastError());
  return MCBA_OK;
}

}  // namespace mcba

// K2: fused residual + analytic Jacobian + robust scaling + per-frame Schur elimination
// (replaces scipy's finite-difference Jacobian + LSMR for the path
// bundle_adjustment.py:299-313; math in SURVEY.md Appendix A), as two kernels:
//
//   K2p (k2_producer.cuh)  per (camera, frame) pair: walk the board corners, accumulate the raw
//       12x12 block A_cf / q_cf in registers, add it to the camera partial sums, hand the
//       extrinsic part (63 doubles per pair) to K2c.            FP64-pipe bound.
//   K2c (below)            per frame: map the hand-off to the pose basis (W' = A[:,ext] E',
//       V'' = E'^T A_ee E'), sum over cameras, factor V_f + lambda D_f^2 = L L^T (6x6, per lane),
//       write Z_cf = W_cf L^-T, y_f = L^-1 g_f, L^-1 for the SYRK and the back-substitution,
//       and the per-tile sums Z y.  Every global access is a coalesced 256-byte row
//       (lane = frame is the fastest index of H, Z, y, L^-1).  HBM bound: reads 504 B,
//       writes 576 B per pair.
//
// A rejected LM step only changes lambda: K2c is re-run on the same hand-off, K2p is not.
#include <cstdlib>

#include "k2_producer.cuh"

namespace mcba {

// ------------------------------------------------------------------ K2c
struct K2CParams {
  int C;
  long long F, nTiles;
  const double* H;       // [tile][c][63][32]
  const double* x;       // 12C + 6F
  const CamConst* cams;
  double lambda;
  double* Z;             // [tile][row 12C][k 6][lane 32]
  double* Linv;          // [tile][21][32]
  double* y;             // [tile][6][32]
  double* partZy;        // [tile][12C]  sum over the tile's frames of Z_f y_f
  double* gpose;         // [f][6]
  double* D2pose;        // [tile][6][32] running max of diag(V_f)
  double* partG;         // [tile] max |g_pose|
};

// E' = [[Rc, 0], [K, Rc]],  K = [t_cf]x Rc   (rows: raw [m | G], cols: pose perturbation [eps | tau])
__device__ __forceinline__ void pose_map(const CamConst& cam, const double (&pose)[6], double (&Rc)[9], double (&K)[9]) {
  double tcf[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) Rc[i] = cam.R[i];
  mat3_vec(Rc, pose + 3, tcf);
  tcf[0] += cam.t[0];
  tcf[1] += cam.t[1];
  tcf[2] += cam.t[2];
  cross_mat3(tcf, Rc, K);
}

template <int kW>
__global__ void __launch_bounds__(kW * 32, 2) k2c_kernel(const K2CParams p) {
  __shared__ double s_V[kW][27][32];
  __shared__ double s_g[kW];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = p.C, nc = 12 * C;
  const long long tile = blockIdx.x;
  const long long f = tile * kTile + lane;
  const bool fvalid = f < p.F;
  double pose[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) pose[i] = fvalid ? p.x[(size_t)nc + f * 6 + i] : 0.0;
  double Jl[9];
  so3_left_jacobian(pose, Jl);

  // ---- V'' = sum_c E'^T A_ee E', g'' = sum_c E'^T q_e over this warp's cameras
  double Vpp[21], gpp[6];
#pragma unroll
  for (int i = 0; i < 21; ++i) Vpp[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) gpp[i] = 0.0;
  for (int c = warp; c < C; c += kW) {
    double Rc[9], K[9];
    pose_map(p.cams[c], pose, Rc, K);
    const double* h = p.H + ((size_t)(tile * C + c) * kHandoff) * kTile + lane;
    double Aee[21], qe[6];
#pragma unroll
    for (int i = 0; i < 21; ++i) Aee[i] = h[(size_t)(36 + i) * kTile];
#pragma unroll
    for (int i = 0; i < 6; ++i) qe[i] = h[(size_t)(57 + i) * kTile];
    double Be[36];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        Be[r * 6 + k] = Aee[sym6(r, 0)] * Rc[k] + Aee[sym6(r, 1)] * Rc[3 + k] + Aee[sym6(r, 2)] * Rc[6 + k] +
                        Aee[sym6(r, 3)] * K[k] + Aee[sym6(r, 4)] * K[3 + k] + Aee[sym6(r, 5)] * K[6 + k];
        Be[r * 6 + 3 + k] = Aee[sym6(r, 3)] * Rc[k] + Aee[sym6(r, 4)] * Rc[3 + k] + Aee[sym6(r, 5)] * Rc[6 + k];
      }
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) {
#pragma unroll
      for (int n = a; n < 6; ++n) {
        double v;
        if (a < 3) {
          v = Rc[a] * Be[n] + Rc[3 + a] * Be[6 + n] + Rc[6 + a] * Be[12 + n] + K[a] * Be[18 + n] +
              K[3 + a] * Be[24 + n] + K[6 + a] * Be[30 + n];
        } else {
          v = Rc[a - 3] * Be[18 + n] + Rc[3 + a - 3] * Be[24 + n] + Rc[6 + a - 3] * Be[30 + n];
        }
        Vpp[tri6(a, n)] += v;
      }
      if (a < 3) {
        gpp[a] += Rc[a] * qe[0] + Rc[3 + a] * qe[1] + Rc[6 + a] * qe[2] + K[a] * qe[3] + K[3 + a] * qe[4] +
                  K[6 + a] * qe[5];
      } else {
        gpp[a] += Rc[a - 3] * qe[3] + Rc[3 + a - 3] * qe[4] + Rc[6 + a - 3] * qe[5];
      }
    }
  }
  // ---- exchange the partial sums (every warp then holds the full pose block; fixed order)
  if (kW > 1) {
#pragma unroll
    for (int i = 0; i < 21; ++i) s_V[warp][i][lane] = Vpp[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) s_V[warp][21 + i][lane] = gpp[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 21; ++i) Vpp[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) gpp[i] = 0.0;
#pragma unroll
    for (int w = 0; w < kW; ++w) {
#pragma unroll
      for (int i = 0; i < 21; ++i) Vpp[i] += s_V[w][i][lane];
#pragma unroll
      for (int i = 0; i < 6; ++i) gpp[i] += s_V[w][21 + i][lane];
    }
  }
  double Linv[21], yv[6], gp[6], gmax = 0.0;
  pose_block_factor(Vpp, gpp, Jl, p.lambda, p.D2pose + (size_t)tile * 6 * 32 + lane, warp == 0, Linv, yv, gp, gmax);
  if (warp == 0) {
    double* lo = p.Linv + (size_t)tile * 21 * 32 + lane;
#pragma unroll
    for (int i = 0; i < 21; ++i) lo[i * 32] = Linv[i];
    double* yo = p.y + (size_t)tile * 6 * 32 + lane;
#pragma unroll
    for (int i = 0; i < 6; ++i) yo[i * 32] = yv[i];
    if (fvalid) {
#pragma unroll
      for (int i = 0; i < 6; i += 2)
        *reinterpret_cast<double2*>(p.gpose + (size_t)f * 6 + i) = make_double2(gp[i], gp[i + 1]);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, off));
    if (lane == 0) p.partG[tile] = gmax;
  }
  // ---- Z_cf = (A[:,ext] E' P') L^-T for this warp's cameras, one raw camera row at a time
  for (int c = warp; c < C; c += kW) {
    double Rc[9], K[9];
    pose_map(p.cams[c], pose, Rc, K);
    const double* h = p.H + ((size_t)(tile * C + c) * kHandoff) * kTile + lane;
    double* z = p.Z + ((size_t)(tile * nc + c * 12) * 6) * kTile + lane;
    double zy[12];
#pragma unroll
    for (int row = 0; row < 12; ++row) {
      double am[3], ag[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        am[j] = row < 6 ? h[(size_t)(row * 6 + j) * kTile] : h[(size_t)(36 + sym6(row - 6, j)) * kTile];
        ag[j] = row < 6 ? h[(size_t)(row * 6 + 3 + j) * kTile] : h[(size_t)(36 + sym6(row - 6, 3 + j)) * kTile];
      }
      double b[6], zr[6];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        b[k] = am[0] * Rc[k] + am[1] * Rc[3 + k] + am[2] * Rc[6 + k] + ag[0] * K[k] + ag[1] * K[3 + k] +
               ag[2] * K[6 + k];
        b[3 + k] = ag[0] * Rc[k] + ag[1] * Rc[3 + k] + ag[2] * Rc[6 + k];
      }
      z_row(b, Jl, Linv, zr);
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        z[(size_t)(row * 6 + k) * kTile] = zr[k];   // padded lanes carry zeros (their hand-off is zero)
        t = fma(zr[k], yv[k], t);
      }
      zy[row] = t;
    }
#pragma unroll
    for (int row = 0; row < 12; ++row) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) zy[row] += __shfl_xor_sync(0xffffffffu, zy[row], off);
    }
    if (lane < 12) {
      double v = zy[0];
#pragma unroll
      for (int row = 1; row < 12; ++row) v = lane == row ? zy[row] : v;
      p.partZy[(size_t)tile * nc + c * 12 + lane] = v;
    }
  }
}

// ------------------------------------------------------------------ launchers
int k2_producer_grid(const Layout& L, int n_sm, int* warps) {
  // 8 warps per CTA (2 per scheduler at 255 registers) once that still fills the GPU,
  // otherwise narrower CTAs so that small problems spread over more SMs.
  int w = 8;
  while (w > 1 && (long long)L.C * ((L.nTiles + w - 1) / w) < n_sm) w >>= 1;
  *warps = w;
  const long long groups = (long long)L.C * ((L.nTiles + w - 1) / w);
  return (int)(groups < n_sm ? groups : n_sm);
}

template <int kLoss, int kWarps>
static int launch_k2p_t(mcba_handle* h, const K2PParams& p) {
  const size_t smem = k2p_smem(p.C, p.N, kWarps);
  MCBA_CUDA(cudaFuncSetAttribute(k2p_kernel<kLoss, kWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k2p_kernel<kLoss, kWarps><<<h->grid_frames, kWarps * 32, smem, h->stream>>>(p);
  h->launches++;
  MCBA_CUDA(cudaGetLastError());
  return MCBA_OK;
}

template <int kLoss>
static int launch_k2p_w(mcba_handle* h, const K2PParams& p) {
  switch (h->prod_warps) {
    case 8: return launch_k2p_t<kLoss, 8>(h, p);
    case 4: return launch_k2p_t<kLoss, 4>(h, p);
    case 2: return launch_k2p_t<kLoss, 2>(h, p);
    default: return launch_k2p_t<kLoss, 1>(h, p);
  }
}

int launch_k2_producer(mcba_handle* h, const double* x, int loss, double f_scale) {
  const Layout& L = h->L;
  K2PParams p;
  p.C = L.C;
  p.N = L.N;
  p.F = L.F;
  p.nTiles = L.nTiles;
  p.nGroups = (long long)L.C * ((L.nTiles + h->prod_warps - 1) / h->prod_warps);
  p.obs = h->d_obs_tiled;
  p.obj = h->d_obj;
  p.x = x;
  p.cams = h->d_cams;
  p.inv_c = 1.0 / f_scale;
  p.c2 = f_scale * f_scale;
  p.H = h->d_H;
  p.partU = h->d_partU;
  p.partS = h->d_partS;
  if ((loss & 0xff) == kLossLinear) return launch_k2p_w<kLossLinear>(h, p);
  if (loss & kLossIrls) return launch_k2p_w<kLossSoftL1 | kLossIrls>(h, p);
  return launch_k2p_w<kLossSoftL1>(h, p);
}

int launch_k2_consumer(mcba_handle* h, const double* x, double lambda) {
  const Layout& L = h->L;
  K2CParams p;
  p.C = L.C;
  p.F = L.F;
  p.nTiles = L.nTiles;
  p.H = h->d_H;
  p.x = x;
  p.cams = h->d_cams;
  p.lambda = lambda;
  p.Z = h->d_Z;
  p.Linv = h->d_Linv;
  p.y = h->d_y;
  p.gpose = h->d_gpose;
  p.D2pose = h->d_D2pose;
  p.partG = h->d_partG;
  p.partZy = h->d_partZy;
  const int grid = (int)L.nTiles;
  if (L.C >= 6) k2c_kernel<6><<<grid, 192, 0, h->stream>>>(p);
  else if (L.C >= 3) k2c_kernel<3><<<grid, 96, 0, h->stream>>>(p);
  else k2c_kernel<1><<<grid, 32, 0, h->stream>>>(p);
  h->launches++;
  MCBA_CUDA(cudaGetLastError());
  return MCBA_OK;
}

}  // namespace mcba

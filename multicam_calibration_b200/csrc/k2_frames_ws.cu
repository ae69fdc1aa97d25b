// K2a, warp-specialised variant for C <= 6 cameras (the BASELINE configs 1-3, 5).
//
// Same mathematics and outputs as k2_frames_kernel (k2_frames.cu).  The CTA has
// C producer warps (warp = camera, lane = frame) that only walk the board
// corners and accumulate the raw 12x12 block A_cf, and 2 consumer warps that
// turn the extrinsic part of A_cf into the pose blocks (V_f, W_cf), factor
// V_f + lambda D_f^2 and emit Z_cf, y_f, L^-1.  Producers hand A[:,ext] and q_ext
// (63 doubles per pair) to the consumers through a double-buffered shared-memory
// tile guarded by named barriers, so the per-tile Schur work overlaps the
// corner loop of the next tile instead of serialising with it, and the two
// schedulers that host a single camera warp get a consumer warp as well
// (warps 0..5 -> SMSP 0,1,2,3,0,1; consumers 6,7 -> SMSP 2,3).
#include "k2_common.cuh"

namespace mcba {

constexpr int kHand = 63;        // 36 (A_int,ext) + 21 (A_ext,ext upper) + 6 (q_ext)
constexpr int kBarFull = 1;      // +buffer
constexpr int kBarEmpty = 3;     // +buffer
constexpr int kBarCons = 5;

__global__ void __launch_bounds__(256, 1) k2_frames_ws_kernel(const K2Params p) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = p.C, N = p.N, nc = 12 * C;
  const int nthreads = (C + 2) * 32;
  double* s_obj = smem;                                   // N*3
  double* s_U = s_obj + ((3 * N + 1) & ~1);               // [C][96]
  double* s_V = s_U + (size_t)C * kUPad;                  // [2][27][32]
  double* s_A = s_V + 2 * 27 * 32;                        // [2][C][63][32]
  for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) s_obj[i] = p.obj[i];
  for (int i = threadIdx.x; i < C * kUPad; i += blockDim.x) s_U[i] = 0.0;
  double cost_acc = 0.0, sumsq_acc = 0.0, cnt_acc = 0.0, gmax = 0.0;
  __syncthreads();
  const long long first = blockIdx.x;
  const int n_it = first < p.nTiles ? (int)((p.nTiles - first + gridDim.x - 1) / gridDim.x) : 0;

  if (warp < C) {
    // ============================ producer: camera = warp ============================
    const int c = warp;
    const CamConst& cam = p.cams[c];
    const Intr in{cam.fx, cam.fy, cam.cx, cam.cy, cam.k1, cam.k2};
    for (int it = 0; it < n_it; ++it) {
      const long long tile = first + (long long)it * gridDim.x;
      const int buf = it & 1;
      const long long f = tile * kTile + lane;
      const bool fvalid = f < p.F;
      double pose[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) pose[i] = fvalid ? p.x[(size_t)nc + f * 6 + i] : 0.0;
      double Rp[9], Rcf[9], tcf[3];
      rodrigues(pose, Rp);
      mat3_mul(cam.R, Rp, Rcf);
      mat3_vec(cam.R, pose + 3, tcf);
      tcf[0] += cam.t[0];
      tcf[1] += cam.t[1];
      tcf[2] += cam.t[2];
      double acc[kUPad];
#pragma unroll
      for (int i = 0; i < kUPad; ++i) acc[i] = 0.0;
      accumulate_pair(p, in, Rcf, tcf, p.obs + ((size_t)(tile * C + c) * N) * kTile + lane, s_obj, acc, cost_acc,
                      sumsq_acc, cnt_acc);
      {
        const double r0 = lane_transpose_sum32<0>(acc, lane);
        const double r1 = lane_transpose_sum32<32>(acc, lane);
        const double r2 = lane_transpose_sum32<64>(acc, lane);
        double* u = s_U + c * kUPad;  // this warp is the only writer of camera c in this CTA
        u[lane] += r0;
        u[32 + lane] += r1;
        u[64 + lane] += r2;
      }
      if (it >= 2) named_bar_sync(kBarEmpty + buf, nthreads);   // consumers are done with this buffer
      double* sa = s_A + ((size_t)(buf * C + c) * kHand) * 32 + lane;
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) sa[(i * 6 + j) * 32] = acc[tri12(i, 6 + j)];
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int s = r; s < 6; ++s) sa[(36 + tri6(r, s)) * 32] = acc[tri12(6 + r, 6 + s)];
#pragma unroll
      for (int r = 0; r < 6; ++r) sa[(57 + r) * 32] = acc[kQ + 6 + r];
      __threadfence_block();
      named_bar_arrive(kBarFull + buf, nthreads);
    }
  } else {
    // ============================ consumers: pose blocks ============================
    const int cw = warp - C;   // 0 or 1
    for (int it = 0; it < n_it; ++it) {
      const long long tile = first + (long long)it * gridDim.x;
      const int buf = it & 1;
      const long long f = tile * kTile + lane;
      const bool fvalid = f < p.F;
      double pose[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) pose[i] = fvalid ? p.x[(size_t)nc + f * 6 + i] : 0.0;
      double Jl[9];
      so3_left_jacobian(pose, Jl);
      named_bar_sync(kBarFull + buf, nthreads);
      // ---- V'' = sum_c E'^T A_ee E', g'' = sum_c E'^T q_e over this consumer's cameras
      double Vpp[21], gpp[6];
#pragma unroll
      for (int i = 0; i < 21; ++i) Vpp[i] = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) gpp[i] = 0.0;
      for (int c = cw; c < C; c += 2) {
        const CamConst& cam = p.cams[c];
        double Rc[9], tcf[3], K[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Rc[i] = cam.R[i];
        mat3_vec(Rc, pose + 3, tcf);
        tcf[0] += cam.t[0];
        tcf[1] += cam.t[1];
        tcf[2] += cam.t[2];
        cross_mat3(tcf, Rc, K);   // E' = [[Rc, 0], [K, Rc]]
        const double* sa = s_A + ((size_t)(buf * C + c) * kHand) * 32 + lane;
        double Aee[21], qe[6];
#pragma unroll
        for (int i = 0; i < 21; ++i) Aee[i] = sa[(36 + i) * 32];
#pragma unroll
        for (int i = 0; i < 6; ++i) qe[i] = sa[(57 + i) * 32];
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
      // ---- exchange the two partial sums
      {
        double* mine = s_V + (size_t)cw * 27 * 32 + lane;
#pragma unroll
        for (int i = 0; i < 21; ++i) mine[i * 32] = Vpp[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) mine[(21 + i) * 32] = gpp[i];
        named_bar_sync(kBarCons, 64);
        const double* other = s_V + (size_t)(1 - cw) * 27 * 32 + lane;
#pragma unroll
        for (int i = 0; i < 21; ++i) Vpp[i] += other[i * 32];
#pragma unroll
        for (int i = 0; i < 6; ++i) gpp[i] += other[(21 + i) * 32];
        named_bar_sync(kBarCons, 64);   // both have read before either overwrites
      }
      double Linv[21], yv[6], gp[6];
      pose_block_factor(Vpp, gpp, Jl, p.lambda, p.D2pose + (size_t)tile * 6 * 32 + lane, cw == 0, Linv, yv, gp, gmax);
      if (cw == 0) {
        double* lo = p.Linv + (size_t)tile * 21 * 32 + lane;
#pragma unroll
        for (int i = 0; i < 21; ++i) lo[i * 32] = Linv[i];
        if (fvalid) {
#pragma unroll
          for (int i = 0; i < 6; i += 2) {
            *reinterpret_cast<double2*>(p.y + (size_t)f * 6 + i) = make_double2(yv[i], yv[i + 1]);
            *reinterpret_cast<double2*>(p.gpose + (size_t)f * 6 + i) = make_double2(gp[i], gp[i + 1]);
          }
        }
      }
      // ---- Z_cf = (A[:,ext] E' P') L^-T for this consumer's cameras
      for (int c = cw; c < C; c += 2) {
        const CamConst& cam = p.cams[c];
        double Rc[9], tcf[3], K[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Rc[i] = cam.R[i];
        mat3_vec(Rc, pose + 3, tcf);
        tcf[0] += cam.t[0];
        tcf[1] += cam.t[1];
        tcf[2] += cam.t[2];
        cross_mat3(tcf, Rc, K);
        const double* sa = s_A + ((size_t)(buf * C + c) * kHand) * 32 + lane;
        double* z = p.Z + (size_t)f * 6 * nc + c * 12;
#pragma unroll
        for (int i = 0; i < 12; i += 2) {
          double zr[2][6];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int row = i + h;
            double am[3], ag[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              am[j] = row < 6 ? sa[(row * 6 + j) * 32] : sa[(36 + sym6(row - 6, j)) * 32];
              ag[j] = row < 6 ? sa[(row * 6 + 3 + j) * 32] : sa[(36 + sym6(row - 6, 3 + j)) * 32];
            }
            double b[6];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              b[k] = am[0] * Rc[k] + am[1] * Rc[3 + k] + am[2] * Rc[6 + k] + ag[0] * K[k] + ag[1] * K[3 + k] +
                     ag[2] * K[6 + k];
              b[3 + k] = ag[0] * Rc[k] + ag[1] * Rc[3 + k] + ag[2] * Rc[6 + k];
            }
            z_row(b, Jl, Linv, zr[h]);
          }
          if (fvalid) {
#pragma unroll
            for (int k = 0; k < 6; ++k)
              *reinterpret_cast<double2*>(z + (size_t)k * nc + i) = make_double2(zr[0][k], zr[1][k]);
          }
        }
      }
      if (it + 2 < n_it) named_bar_arrive(kBarEmpty + buf, nthreads);
    }
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
  double* s_red = s_V;
  if (lane == 0) {
    s_red[warp * 4 + 0] = cost_acc;
    s_red[warp * 4 + 1] = sumsq_acc;
    s_red[warp * 4 + 2] = cnt_acc;
    s_red[warp * 4 + 3] = gmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, cn = 0, gm = 0;
    for (int w = 0; w < C + 2; ++w) {
      a += s_red[w * 4];
      b += s_red[w * 4 + 1];
      cn += s_red[w * 4 + 2];
      gm = fmax(gm, s_red[w * 4 + 3]);
    }
    double* ps = p.partS + (size_t)blockIdx.x * kRsNum;
    ps[kRsCost] = 0.5 * a;
    ps[kRsSumSq] = b;
    ps[kRsCount] = cn;
    ps[kRsGmaxPose] = gm;
  }
}

int launch_k2_frames_ws(mcba_handle* h, const K2Params& p) {
  const size_t smem = sizeof(double) * (((3 * p.N + 1) & ~1) + (size_t)p.C * kUPad + 2 * 27 * 32 +
                                        (size_t)2 * p.C * kHand * 32);
  MCBA_CUDA(cudaFuncSetAttribute(k2_frames_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k2_frames_ws_kernel<<<h->grid_frames, (p.C + 2) * 32, smem, h->stream>>>(p);
  h->launches++;
  MCBA_CUDA(cudaGetLastError());
  return MCBA_OK;
}

}  // namespace mcba

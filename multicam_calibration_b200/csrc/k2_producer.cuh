// K2p: the per-observation half of the fused residual + analytic Jacobian + Schur pass
// (replaces scipy's finite-difference Jacobian, bundle_adjustment.py:299-313 /
// scipy _numdiff.py:288; math in SURVEY.md Appendix A and mcba_math.cuh).
//
// Work unit = (camera c, tile of 32 consecutive frames); lane = frame.  A lane owns one
// (camera, frame) pair, walks its N board corners with coalesced 16-byte loads of the tiled
// observation SoA ([tile][camera][corner][lane]) and accumulates the raw 12x12 Gauss-Newton
// block  A_cf = sum_n w a a^T,  q_cf = -sum_n rho' f a  in registers (86 live accumulators,
// layout in k2_common.cuh).  At the end of the unit
//   * the full block is added to the camera's U / gradient partial sums (register
//     transpose-reduction: 4 values per lane, then a fixed-order sum over the CTA's warps),
//   * the extrinsic part (A[:,ext], A_ext,ext, q_ext: 63 doubles per pair) is handed to the
//     per-frame Schur kernel K2c through the coalesced hand-off buffer H[tile][c][63][lane].
// Nothing per-observation is written: algorithmic HBM traffic is 16 B/observation.
//
// Scheduling: only LIVE units (camera has observations in the tile; frames are sorted by
// visibility mask, k1_residual.cu) are walked.  They are ordered camera-major in groups of kWarps
// units; every CTA owns a contiguous, equally sized range of groups, so the camera index is
// CTA-uniform and changes at most a few times per CTA.  All kWarps warps of the CTA walk corners (no consumer warps, no named barriers: the
// per-frame Schur step is the separate kernel K2c).  The corner loop is ONE branch-free basic
// block: missing observations are handled with selects, the reciprocal and rsqrt are inlined
// Newton iterations on MUFU seeds (no slow-path calls), the loss is a template parameter.
// Measured on B200 (scripts/ubench/k2p_variants.cu, 6 x 50k x 35, 20 % missing views): 0.232 ms,
// ~61 % of the FP64 pipe; walking u and v rows separately with half the accumulators live (to
// free registers for software pipelining) was 20 % slower and is not kept.
#pragma once
#include "k2_common.cuh"

namespace mcba {

struct K2PParams {
  int C, N;
  long long F, nTiles;
  const double2* obs;         // tiled [tile][c][n][lane]
  const int* perm;            // tile slot -> frame index in x (-1 = padding)
  const int* units;           // [C][nTiles] live tiles per camera, compacted (build_units_kernel)
  const int* unit_count;      // [C] live units per camera
  const int* gprefix;         // [C + 1] groups of kWarps units before camera c; [C] = total
  const double* obj;          // (N,3)
  const double* x;            // 12C + 6F
  CamConst* cams;             // per-camera constants: built by every CTA for itself, published by CTA 0 for K2c / finalize / K3
  double inv_c, c2;           // 1/f_scale, f_scale^2
  double* H;                  // [tile][c][63][32]
  double* partU;              // [grid][C][kAcc]
  double* partS;              // [grid][kRsNum]
  const unsigned long long* n_scalars;   // finite observation scalars of this rank's frames
};

template <bool IsU>
__device__ __forceinline__ void accumulate_row(double (&acc)[kAcc], const double (&a)[10], double wh, double gf) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const int I = IsU ? kIdxU[i] : kIdxV[i];
    const double s = wh * a[i];
    acc[acc_slot_q(I)] = fma(gf, a[i], acc[acc_slot_q(I)]);
#pragma unroll
    for (int j = i; j < 10; ++j) {
      const int J = IsU ? kIdxU[j] : kIdxV[j];
      acc[acc_slot(I, J)] = fma(s, a[j], acc[acc_slot(I, J)]);
    }
  }
}

// Sum over the 32 lanes of v[Base + l] delivered to lane l (recursive halving: 31 shuffles
// instead of 32 x 5).
template <int Base>
__device__ __forceinline__ double lane_transpose_sum32(const double (&v)[kAcc], int lane) {
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

// 1/sqrt(t) for t >= 1: MUFU.RSQ64H seed (relative error <= 1e-6) and ONE cubic correction
// y (1 + e/2 + 3 e^2/8): max 1.01 ulp on B200 (scripts/ubench/seed_accuracy.cu); the quadratic step that
// followed it until round 2 cost four dependent FP64 instructions per weight and measured 1.48 ulp.
__device__ __forceinline__ double fast_rsqrt(double t) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t));
  const double e = fma(-t, y * y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
}

// Robust weights of one scalar residual (see robust_weights in mcba_math.cuh), branch-free on
// the data: f == 0 for a missing scalar gives a zero gradient term by itself; only the Gauss-Newton
// weight needs the validity select.  The cost is accumulated in the cheapest form the loss allows:
//   soft_l1:  rho = 2 c^2 (sqrt(1 + z) - 1) with sqrt(1 + z) = t * rsqrt(t): ONE fma per scalar adds
//             t b to the running sum; the "- 1" (every walked slot, missing ones included: they add
//             t b = 1) and the factor 2 c^2 are applied once per thread (k2p_cost_of_thread);
//   linear:   rho = f^2, one fma.
template <int kLoss>
__device__ __forceinline__ void robust_weights_t(double f, bool valid, double inv_c, double& cost_acc, double& wg,
                                                 double& wh) {
  if ((kLoss & 0xff) == kLossLinear) {
    cost_acc = fma(f, f, cost_acc);
    wg = 1.0;
    wh = valid ? 1.0 : 0.0;
  } else {
    const double fs = f * inv_c;
    const double t = fma(fs, fs, 1.0);
    const double b = fast_rsqrt(t);
    cost_acc = fma(t, b, cost_acc);
    wg = b;
    const double w3 = (kLoss & kLossIrls) ? b : fmax(b * b * b, 2.220446049250313e-16);
    wh = valid ? w3 : 0.0;
  }
}
// sum of rho over the slots a thread walked, from its running sum and the number of scalar slots
template <int kLoss>
__device__ __forceinline__ double k2p_cost_of_thread(double cost_acc, double slots, double c2) {
  return (kLoss & 0xff) == kLossLinear ? cost_acc : 2.0 * c2 * (cost_acc - slots);
}

#ifndef MCBA_K2P_LAUNDER
#define MCBA_K2P_LAUNDER 1
#endif
#ifndef MCBA_K2P_VOLATILE_SR
#define MCBA_K2P_VOLATILE_SR 1
#endif
#if MCBA_K2P_VOLATILE_SR
typedef const volatile double* SrPtr;
#else
typedef const double* __restrict__ SrPtr;
#endif

// The camera intrinsics reach the corner loop through an opaque FP64 add: uses inside the loop
// then depend on a fixed-latency ALU result, not on the scoreboard of the global load that
// fetched them before the loop.  (ptxas otherwise makes the first FP64 instruction of every
// iteration wait on that scoreboard, which it shares with the in-loop observation prefetch:
// 26 % of all stall samples in profiles/r01_b.)
struct IntrReg {
  double fx, fy, cx, cy, k1, k2;
};
__device__ __forceinline__ double opaque(double v) {
#if MCBA_K2P_LAUNDER
  asm volatile("add.f64 %0, %0, 0d0000000000000000;" : "+d"(v));
#endif
  return v;
}

// Shared part of the projection of one corner: camera-frame point, distortion, d(u,v)/d(x,y).
struct Proj {
  double x, y, iz, r2, d, A00, A01, A10, A11, su, sv, pu, pv;
};

// 1 / Z_c of a corner: issued ONE corner ahead of its use (it depends only on the pose, not on
// the previous corner), which takes the MUFU + 5 dependent FMAs of the reciprocal off the
// per-corner critical path at the price of one carried double.
__device__ __forceinline__ double inverse_depth(SrPtr sR, double qx, double qy, double qz) {
  return fast_rcp(fma(sR[6 * 32], qx, fma(sR[7 * 32], qy, fma(sR[8 * 32], qz, sR[11 * 32]))));
}

__device__ __forceinline__ void project_shared(const IntrReg& cam, SrPtr sR, double qx, double qy, double qz,
                                               double iz, Proj& o) {
  // sR: this lane's [Rcf (9) | tcf (3)], stride 32 doubles between entries
  const double X = fma(sR[0 * 32], qx, fma(sR[1 * 32], qy, fma(sR[2 * 32], qz, sR[9 * 32])));
  const double Y = fma(sR[3 * 32], qx, fma(sR[4 * 32], qy, fma(sR[5 * 32], qz, sR[10 * 32])));
  o.iz = iz;
  o.x = X * o.iz;
  o.y = Y * o.iz;
  o.r2 = fma(o.x, o.x, o.y * o.y);
  o.d = fma(o.r2, fma(cam.k2, o.r2, cam.k1), 1.0);
  const double dp = fma(2.0 * cam.k2, o.r2, cam.k1);
  o.pu = fma(cam.fx, o.x * o.d, cam.cx);
  o.pv = fma(cam.fy, o.y * o.d, cam.cy);
  const double xy2 = 2.0 * o.x * o.y * dp;
  o.A00 = cam.fx * fma(2.0 * o.x * o.x, dp, o.d);
  o.A01 = cam.fx * xy2;
  o.A10 = cam.fy * xy2;
  o.A11 = cam.fy * fma(2.0 * o.y * o.y, dp, o.d);
  o.su = fma(o.A00, o.x, o.A01 * o.y);
  o.sv = fma(o.A10, o.x, o.A11 * o.y);
}

// Raw Jacobian row of u (kU) or v: [d/df, d/dc = 1, d/dk1, d/dk2, m (3), G (3)]  (mcba_obs.cuh).
template <bool kU>
__device__ __forceinline__ void jac_row(const IntrReg& cam, const Proj& p, double (&a)[10]) {
  const double f = kU ? cam.fx : cam.fy, w = kU ? p.x : p.y;
  const double Aa = kU ? p.A00 : p.A10, Ab = kU ? p.A01 : p.A11, s = kU ? p.su : p.sv;
  a[0] = w * p.d;
  a[1] = 1.0;
  a[2] = f * w * p.r2;
  a[3] = a[2] * p.r2;
  a[4] = -fma(p.y, s, Ab);
  a[5] = fma(p.x, s, Aa);
  a[6] = fma(p.x, Ab, -p.y * Aa);
  a[7] = Aa * p.iz;
  a[8] = Ab * p.iz;
  a[9] = -s * p.iz;
}

// One unit (camera, 32 frames): walk the N corners, both rows of each observation.
// MCBA_K2P_VARIANT (A/B builds, scripts/ubench/Makefile): 0 = projection at the top of the iteration with only
// the inverse depth one corner ahead; 1 (default) = the normalised image point of the next corner is formed
// while the current one accumulates (-2.4 % on B200); 2 = the whole projection and both robust weights one
// corner ahead (slower: the carried state spills).  All three give bit-identical sums.
#ifndef MCBA_K2P_VARIANT
#define MCBA_K2P_VARIANT 1
#endif
#if MCBA_K2P_VARIANT == 0
template <int kLoss>
__device__ __forceinline__ void walk_corners(const K2PParams& p, const IntrReg& cam, SrPtr sR,
                                             const double2* __restrict__ ob, double2 o0, double2 o1,
                                             const double* __restrict__ s_obj,
                                             double (&acc)[kAcc], double& cost_acc, double& sumsq_acc,
                                             double& cnt_acc) {
  const int N = p.N;
  double iz_next = inverse_depth(sR, s_obj[0], s_obj[1], s_obj[2]);
#pragma unroll 1
  for (int n = 0; n < N; ++n) {
    const double2 cur = o0;
    o0 = o1;
    if (n + 2 < N) o1 = ob[(size_t)(n + 2) * kTile];
    const double iz = iz_next;
    const int nn = n + 1 < N ? n + 1 : n;
    iz_next = inverse_depth(sR, s_obj[3 * nn], s_obj[3 * nn + 1], s_obj[3 * nn + 2]);
    Proj pr;
    project_shared(cam, sR, s_obj[3 * n], s_obj[3 * n + 1], s_obj[3 * n + 2], iz, pr);
    {
      const bool hu = cur.x == cur.x;
      const double fu = hu ? cur.x - pr.pu : 0.0;
      double wg, wh, au[10];
      robust_weights_t<kLoss>(fu, hu, p.inv_c, cost_acc, wg, wh);
      jac_row<true>(cam, pr, au);
      sumsq_acc = fma(fu, fu, sumsq_acc);
      accumulate_row<true>(acc, au, wh, -wg * fu);
    }
    {
      const bool hv = cur.y == cur.y;
      const double fv = hv ? cur.y - pr.pv : 0.0;
      double wg, wh, av[10];
      robust_weights_t<kLoss>(fv, hv, p.inv_c, cost_acc, wg, wh);
      jac_row<false>(cam, pr, av);
      sumsq_acc = fma(fv, fv, sumsq_acc);
      accumulate_row<false>(acc, av, wh, -wg * fv);
    }
  }
}

#elif MCBA_K2P_VARIANT == 1
// Variant 1: normalised image point (x, y) and inverse depth of corner n+1 are formed while corner n
// accumulates (carried: 3 doubles).
__device__ __forceinline__ void project_tail(const IntrReg& cam, double x, double y, double iz, Proj& o) {
  o.iz = iz; o.x = x; o.y = y;
  o.r2 = fma(o.x, o.x, o.y * o.y);
  o.d = fma(o.r2, fma(cam.k2, o.r2, cam.k1), 1.0);
  const double dp = fma(2.0 * cam.k2, o.r2, cam.k1);
  o.pu = fma(cam.fx, o.x * o.d, cam.cx);
  o.pv = fma(cam.fy, o.y * o.d, cam.cy);
  const double xy2 = 2.0 * o.x * o.y * dp;
  o.A00 = cam.fx * fma(2.0 * o.x * o.x, dp, o.d);
  o.A01 = cam.fx * xy2;
  o.A10 = cam.fy * xy2;
  o.A11 = cam.fy * fma(2.0 * o.y * o.y, dp, o.d);
  o.su = fma(o.A00, o.x, o.A01 * o.y);
  o.sv = fma(o.A10, o.x, o.A11 * o.y);
}
__device__ __forceinline__ void point_xy(SrPtr sR, double qx, double qy, double qz, double& x, double& y, double& iz) {
  const double X = fma(sR[0 * 32], qx, fma(sR[1 * 32], qy, fma(sR[2 * 32], qz, sR[9 * 32])));
  const double Y = fma(sR[3 * 32], qx, fma(sR[4 * 32], qy, fma(sR[5 * 32], qz, sR[10 * 32])));
  iz = inverse_depth(sR, qx, qy, qz);
  x = X * iz;
  y = Y * iz;
}
template <int kLoss>
__device__ __forceinline__ void walk_corners(const K2PParams& p, const IntrReg& cam, SrPtr sR,
                                             const double2* __restrict__ ob, double2 o0, double2 o1,
                                             const double* __restrict__ s_obj,
                                             double (&acc)[kAcc], double& cost_acc, double& sumsq_acc,
                                             double& cnt_acc) {
  // o0, o1: the unit's first two observations, requested by the caller BEFORE its pose work
  const int N = p.N;
  double xn, yn, izn;
  point_xy(sR, s_obj[0], s_obj[1], s_obj[2], xn, yn, izn);
#pragma unroll 1
  for (int n = 0; n < N; ++n) {
    const double2 cur = o0;
    o0 = o1;
    if (n + 2 < N) o1 = ob[(size_t)(n + 2) * kTile];
    Proj pr;
    project_tail(cam, xn, yn, izn, pr);
    const int nn = n + 1 < N ? n + 1 : n;
    point_xy(sR, s_obj[3 * nn], s_obj[3 * nn + 1], s_obj[3 * nn + 2], xn, yn, izn);
    {
      const bool hu = cur.x == cur.x;
      const double fu = hu ? cur.x - pr.pu : 0.0;
      double wg, wh, au[10];
      robust_weights_t<kLoss>(fu, hu, p.inv_c, cost_acc, wg, wh);
      jac_row<true>(cam, pr, au);
      sumsq_acc = fma(fu, fu, sumsq_acc);
      accumulate_row<true>(acc, au, wh, -wg * fu);
    }
    {
      const bool hv = cur.y == cur.y;
      const double fv = hv ? cur.y - pr.pv : 0.0;
      double wg, wh, av[10];
      robust_weights_t<kLoss>(fv, hv, p.inv_c, cost_acc, wg, wh);
      jac_row<false>(cam, pr, av);
      sumsq_acc = fma(fv, fv, sumsq_acc);
      accumulate_row<false>(acc, av, wh, -wg * fv);
    }
  }
}
#elif MCBA_K2P_VARIANT == 2
// Variant 2: the whole projection AND both robust weights of corner n+1 are formed while corner n
// accumulates (carried: Proj + 4 weights).
struct RowW { double wh, gf; };
template <int kLoss>
__device__ __forceinline__ void corner_front(const K2PParams& p, const IntrReg& cam, SrPtr sR, const double* __restrict__ s_obj,
                                             int n, double2 cur, Proj& pr, RowW& wu, RowW& wv, double& cost_acc,
                                             double& sumsq_acc, double& cnt_acc) {
  const double iz = inverse_depth(sR, s_obj[3 * n], s_obj[3 * n + 1], s_obj[3 * n + 2]);
  project_shared(cam, sR, s_obj[3 * n], s_obj[3 * n + 1], s_obj[3 * n + 2], iz, pr);
  {
    const bool hu = cur.x == cur.x;
    const double fu = hu ? cur.x - pr.pu : 0.0;
    double wg;
    robust_weights_t<kLoss>(fu, hu, p.inv_c, cost_acc, wg, wu.wh);
    wu.gf = -wg * fu;
    sumsq_acc = fma(fu, fu, sumsq_acc);
  }
  {
    const bool hv = cur.y == cur.y;
    const double fv = hv ? cur.y - pr.pv : 0.0;
    double wg;
    robust_weights_t<kLoss>(fv, hv, p.inv_c, cost_acc, wg, wv.wh);
    wv.gf = -wg * fv;
    sumsq_acc = fma(fv, fv, sumsq_acc);
  }
}
template <int kLoss>
__device__ __forceinline__ void walk_corners(const K2PParams& p, const IntrReg& cam, SrPtr sR,
                                             const double2* __restrict__ ob, double2 o0, double2 o1,
                                             const double* __restrict__ s_obj,
                                             double (&acc)[kAcc], double& cost_acc, double& sumsq_acc,
                                             double& cnt_acc) {
  const int N = p.N;
  Proj pr;
  RowW wu, wv;
  corner_front<kLoss>(p, cam, sR, s_obj, 0, o0, pr, wu, wv, cost_acc, sumsq_acc, cnt_acc);
#pragma unroll 1
  for (int n = 0; n < N; ++n) {
    // past the last corner the look-ahead runs on a missing observation: it adds nothing to the sums
    const double2 nxt = n + 1 < N ? o1 : make_double2(nan(""), nan(""));
    if (n + 2 < N) o1 = ob[(size_t)(n + 2) * kTile];
    double au[10], av[10];
    jac_row<true>(cam, pr, au);
    jac_row<false>(cam, pr, av);
    const RowW cu = wu, cv = wv;
    corner_front<kLoss>(p, cam, sR, s_obj, n + 1 < N ? n + 1 : n, nxt, pr, wu, wv, cost_acc, sumsq_acc, cnt_acc);
    accumulate_row<true>(acc, au, cu.wh, cu.gf);
    accumulate_row<false>(acc, av, cv.wh, cv.gf);
  }
}
#endif

// The same walk for a HALF unit (16 frames): lanes 0-15 take the first half of the corners of
// their frame, lanes 16-31 the second half of the same frames; the two partial blocks are added
// with one shuffle pass afterwards.  Used for the units left over after the last full round of the
// static schedule, which would otherwise cost every CTA a whole extra round for a fraction of the
// warps (6.36 units per warp -> 7 rounds at BASELINE configs[2]).
__host__ __device__ constexpr bool acc_slot_used(int i) { return i < 21 || (i >= 32 && i < 53) || (i >= 64 && i < 108); }

template <int kLoss>
__device__ __forceinline__ void walk_corners_half(const K2PParams& p, const IntrReg& cam, SrPtr sR,
                                                  const double2* __restrict__ ob, const double* __restrict__ s_obj,
                                                  int lane, double (&acc)[kAcc], double& cost_acc, double& sumsq_acc,
                                                  double& cnt_acc) {
  const int N = p.N, Nh = (N + 1) >> 1;
  const int n0 = (lane & 16) ? Nh : 0;
  const double2 missing = make_double2(nan(""), nan(""));
  auto obs_at = [&](int n) { return n < N ? ob[(size_t)n * kTile] : missing; };
  auto corner = [&](int n) { return n < N ? n : N - 1; };
  // three observations in flight (the full walk keeps two): every warp of the GPU enters the tail round
  // at the same moment with cold loads, and a half unit has only ~18 corners to amortise them over
  double2 o0 = obs_at(n0);
  double2 o1 = obs_at(n0 + 1);
  double2 o2 = obs_at(n0 + 2);
  double iz_next = inverse_depth(sR, s_obj[3 * corner(n0)], s_obj[3 * corner(n0) + 1], s_obj[3 * corner(n0) + 2]);
#pragma unroll 1
  for (int i = 0; i < Nh; ++i) {
    const int n = corner(n0 + i), nn = corner(n0 + i + 1);
    const double2 cur = o0;
    o0 = o1;
    o1 = o2;
    o2 = obs_at(n0 + i + 3);
    const double iz = iz_next;
    iz_next = inverse_depth(sR, s_obj[3 * nn], s_obj[3 * nn + 1], s_obj[3 * nn + 2]);
    Proj pr;
    project_shared(cam, sR, s_obj[3 * n], s_obj[3 * n + 1], s_obj[3 * n + 2], iz, pr);
    {
      const bool hu = cur.x == cur.x;
      const double fu = hu ? cur.x - pr.pu : 0.0;
      double wg, wh, au[10];
      robust_weights_t<kLoss>(fu, hu, p.inv_c, cost_acc, wg, wh);
      jac_row<true>(cam, pr, au);
      sumsq_acc = fma(fu, fu, sumsq_acc);
      accumulate_row<true>(acc, au, wh, -wg * fu);
    }
    {
      const bool hv = cur.y == cur.y;
      const double fv = hv ? cur.y - pr.pv : 0.0;
      double wg, wh, av[10];
      robust_weights_t<kLoss>(fv, hv, p.inv_c, cost_acc, wg, wh);
      jac_row<false>(cam, pr, av);
      sumsq_acc = fma(fv, fv, sumsq_acc);
      accumulate_row<false>(acc, av, wh, -wg * fv);
    }
  }
  // add the two corner halves; lanes 16-31 then carry zeros so that the lane reductions count once
#pragma unroll
  for (int i = 0; i < kAcc; ++i) {
    if (acc_slot_used(i)) {
      const double other = __shfl_xor_sync(0xffffffffu, acc[i], 16);
      acc[i] = (lane & 16) ? 0.0 : acc[i] + other;
    }
  }
}

template <int kLoss, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) k2p_kernel(const K2PParams p) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = p.C, N = p.N, nc = 12 * C;
  double* s_obj = smem;                                  // 3N (+pad)
  double* s_U = s_obj + ((3 * N + 1) & ~1);              // [C][kAcc]   CTA partial sums of U, g
  double* s_Uw = s_U + (size_t)C * kAcc;                 // [kWarps][kAcc] per-group staging
  double* s_R = s_Uw + (size_t)kWarps * kAcc;            // [kWarps][12][32] per-lane Rcf | tcf
  CamConst* s_cam = reinterpret_cast<CamConst*>(s_R + (size_t)kWarps * 12 * 32);   // [C]
  if (threadIdx.x < C) {   // camera constants in the kernel itself: no separate launch in front of the walk
    const double* q = p.x + 12 * threadIdx.x;
    CamConst k;
    k.fx = q[0]; k.fy = q[1]; k.cx = q[2]; k.cy = q[3]; k.k1 = q[4]; k.k2 = q[5];
    const double r[3] = {q[6], q[7], q[8]};
    k.t[0] = q[9]; k.t[1] = q[10]; k.t[2] = q[11];
    rodrigues(r, k.R);
    so3_left_jacobian(r, k.Jl);
    cross_mat3(k.t, k.Jl, k.tJ);
    s_cam[threadIdx.x] = k;
    if (blockIdx.x == 0) p.cams[threadIdx.x] = k;
  }
  for (int i = threadIdx.x; i < 3 * N; i += blockDim.x) s_obj[i] = p.obj[i];
  for (int i = threadIdx.x; i < C * kAcc; i += blockDim.x) s_U[i] = 0.0;
  double cost_acc = 0.0, sumsq_acc = 0.0, cnt_acc = 0.0;   // cnt_acc: scalar slots this thread walked (k2p_cost_of_thread)
  double* sR = s_R + (size_t)warp * 12 * 32 + lane;
  __syncthreads();

  // contiguous range of camera-major groups of live units for this CTA (all CTAs get the same
  // number of fully populated groups: static, balanced, and the partial sums stay deterministic)
  const int nGroups = p.gprefix[C];
  // Leftover of the last full round: when at most half a round of groups remains and they belong
  // to one camera, every CTA runs `rounds` full groups and the leftover units are split into half
  // units (16 frames, corners divided over the two half-warps) spread over ALL CTAs.
  const int rounds = nGroups / (int)gridDim.x, left_groups = nGroups - rounds * (int)gridDim.x;
  int tail_c = 0;
  while (tail_c + 1 < C && p.gprefix[tail_c + 1] <= rounds * (int)gridDim.x) ++tail_c;
  const int tail_k0 = (rounds * (int)gridDim.x - p.gprefix[tail_c]) * kWarps;        // first leftover unit of camera tail_c
  const int tail_units = left_groups > 0 ? p.unit_count[tail_c] - tail_k0 : 0;
  const int tail_per_cta = (2 * tail_units + (int)gridDim.x - 1) / (int)gridDim.x;   // half units per CTA
  const bool split_tail = kWarps == 8 && rounds >= 1 && left_groups > 0 && tail_per_cta <= kWarps &&
                          p.gprefix[tail_c + 1] >= nGroups;
  const int g_begin = split_tail ? rounds * (int)blockIdx.x : (int)(((long long)nGroups * blockIdx.x) / gridDim.x);
  const int g_end = split_tail ? rounds * ((int)blockIdx.x + 1) : (int)(((long long)nGroups * (blockIdx.x + 1)) / gridDim.x);
  int c = 0;
  for (int g = g_begin; g < g_end; ++g) {
    while (c + 1 < C && p.gprefix[c + 1] <= g) ++c;     // CTA-uniform
    const int k = (g - p.gprefix[c]) * kWarps + warp;
    const bool live = k < p.unit_count[c];
    const long long tile = live ? p.units[(long long)c * p.nTiles + k] : 0;
    const CamConst& cam = s_cam[c];
    double* uw = s_Uw + warp * kAcc;
    if (live) {
      const long long f = p.perm[tile * kTile + lane];
      const double2* ob = p.obs + ((size_t)(tile * C + c) * N) * kTile + lane;
      const double2 o0 = ob[0];                       // in flight during the pose loads and the Rodrigues work
      const double2 o1 = N > 1 ? ob[kTile] : o0;
      const bool fvalid = f >= 0;
      {
        double pose[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) pose[i] = fvalid ? p.x[(size_t)nc + f * 6 + i] : 0.0;
        double Rp[9], Rcf[9], tcf[3];
        rodrigues(pose, Rp);
        mat3_mul(cam.R, Rp, Rcf);
        mat3_vec(cam.R, pose + 3, tcf);
#pragma unroll
        for (int i = 0; i < 9; ++i) sR[i * 32] = Rcf[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) sR[(9 + i) * 32] = tcf[i] + cam.t[i];
      }
      __syncwarp();
      double* h = p.H + ((size_t)(tile * C + c) * kHandoff) * kTile + lane;
      double acc[kAcc];
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;
      const IntrReg in{opaque(cam.fx), opaque(cam.fy), opaque(cam.cx), opaque(cam.cy), opaque(cam.k1), opaque(cam.k2)};
      walk_corners<kLoss>(p, in, sR, ob, o0, o1, s_obj, acc, cost_acc, sumsq_acc, cnt_acc);
      cnt_acc += 2.0 * (MCBA_K2P_VARIANT == 2 ? N + 1 : N);
      uw[lane] = lane_transpose_sum32<0>(acc, lane);
      uw[32 + lane] = lane_transpose_sum32<32>(acc, lane);
      uw[64 + lane] = lane_transpose_sum32<64>(acc, lane);
      uw[96 + lane] = lane_transpose_sum32<96>(acc, lane);
      // hand-off to K2c: A[int rows, ext], A_ext,ext (upper 21), q_ext
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) h[(size_t)(i * 6 + j) * kTile] = acc[acc_slot(i, 6 + j)];
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int q = r; q < 6; ++q) h[(size_t)(36 + tri6(r, q)) * kTile] = acc[acc_slot(6 + r, 6 + q)];
#pragma unroll
      for (int r = 0; r < 6; ++r) h[(size_t)(57 + r) * kTile] = acc[acc_slot_q(6 + r)];
    } else {
#pragma unroll
      for (int b = 0; b < 4; ++b) uw[32 * b + lane] = 0.0;
    }
    // camera block: sum over the CTA's warps in a fixed order
    __syncthreads();
    for (int i = threadIdx.x; i < kAcc; i += kWarps * 32) {
      double s = s_U[c * kAcc + i];
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += s_Uw[w * kAcc + i];
      s_U[c * kAcc + i] = s;
    }
    __syncthreads();
  }

  if (split_tail) {
    const int hu = (int)blockIdx.x * tail_per_cta + warp;
    const bool live = warp < tail_per_cta && hu < 2 * tail_units;
    const int cc = tail_c;
    const CamConst& cam = s_cam[cc];
    double* uw = s_Uw + warp * kAcc;
    if (live) {
      const long long tile = p.units[(long long)cc * p.nTiles + tail_k0 + (hu >> 1)];
      const int slot = (hu & 1) * 16 + (lane & 15);          // this lane's frame slot of the tile
      const long long f = p.perm[tile * kTile + slot];
      const bool fvalid = f >= 0;
      {
        double pose[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) pose[i] = fvalid ? p.x[(size_t)nc + f * 6 + i] : 0.0;
        double Rp[9], Rcf[9], tcf[3];
        rodrigues(pose, Rp);
        mat3_mul(cam.R, Rp, Rcf);
        mat3_vec(cam.R, pose + 3, tcf);
#pragma unroll
        for (int i = 0; i < 9; ++i) sR[i * 32] = Rcf[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) sR[(9 + i) * 32] = tcf[i] + cam.t[i];
      }
      __syncwarp();
      const double2* ob = p.obs + ((size_t)(tile * C + cc) * N) * kTile + slot;
      double* h = p.H + ((size_t)(tile * C + cc) * kHandoff) * kTile + slot;
      double acc[kAcc];
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;
      const IntrReg in{opaque(cam.fx), opaque(cam.fy), opaque(cam.cx), opaque(cam.cy), opaque(cam.k1), opaque(cam.k2)};
      walk_corners_half<kLoss>(p, in, sR, ob, s_obj, lane, acc, cost_acc, sumsq_acc, cnt_acc);
      cnt_acc += 2.0 * ((N + 1) >> 1);
      uw[lane] = lane_transpose_sum32<0>(acc, lane);
      uw[32 + lane] = lane_transpose_sum32<32>(acc, lane);
      uw[64 + lane] = lane_transpose_sum32<64>(acc, lane);
      uw[96 + lane] = lane_transpose_sum32<96>(acc, lane);
      if (lane < 16) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 6; ++j) h[(size_t)(i * 6 + j) * kTile] = acc[acc_slot(i, 6 + j)];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int q = r; q < 6; ++q) h[(size_t)(36 + tri6(r, q)) * kTile] = acc[acc_slot(6 + r, 6 + q)];
#pragma unroll
        for (int r = 0; r < 6; ++r) h[(size_t)(57 + r) * kTile] = acc[acc_slot_q(6 + r)];
      }
    } else {
#pragma unroll
      for (int b = 0; b < 4; ++b) uw[32 * b + lane] = 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kAcc; i += kWarps * 32) {
      double s = s_U[cc * kAcc + i];
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += s_Uw[w * kAcc + i];
      s_U[cc * kAcc + i] = s;
    }
    __syncthreads();
  }

  // ---------------- CTA epilogue: partial sums ----------------
  double* pu = p.partU + (size_t)blockIdx.x * C * kAcc;
  for (int i = threadIdx.x; i < C * kAcc; i += blockDim.x) pu[i] = s_U[i];
  cost_acc = k2p_cost_of_thread<kLoss>(cost_acc, cnt_acc, p.c2);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    cost_acc += __shfl_xor_sync(0xffffffffu, cost_acc, off);
    sumsq_acc += __shfl_xor_sync(0xffffffffu, sumsq_acc, off);
  }
  __syncthreads();
  double* s_red = s_Uw;
  if (lane == 0) {
    s_red[warp * 4 + 0] = cost_acc;
    s_red[warp * 4 + 1] = sumsq_acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < kWarps; ++w) {
      a += s_red[w * 4];
      b += s_red[w * 4 + 1];
    }
    double* ps = p.partS + (size_t)blockIdx.x * kRsNum;
    ps[kRsCost] = 0.5 * a;
    ps[kRsSumSq] = b;
    // the number of finite scalars does not change between evaluations: counted once when the observations
    // were tiled (tile_observations_kernel), reported by CTA 0
    ps[kRsCount] = blockIdx.x == 0 ? (double)*p.n_scalars : 0.0;
    ps[kRsGmaxPose] = 0.0;   // pose gradients are K2c's (partG)
  }
}

inline size_t k2p_smem(int C, int N, int warps) {
  return sizeof(double) * (((3 * N + 1) & ~1) + (size_t)C * kAcc + (size_t)warps * kAcc + (size_t)warps * 12 * 32) +
         sizeof(CamConst) * (size_t)C;
}

}  // namespace mcba

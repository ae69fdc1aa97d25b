// EXPERIMENT (round 2, not part of the library; results in scripts/experiments/README.md):
// K2p, warp-specialised build (same inputs, outputs and per-lane arithmetic as k2_producer.cuh).
//
// The one-role kernel keeps the 86 accumulators of A_cf / q_cf (172 registers) AND the projection chain
// of the corner in the same thread: at 255 registers only two warps fit per scheduler and little of the
// next corner's chain (reciprocal, rsqrt, ~100 dependent FP64 instructions) can be in flight while the
// current corner accumulates -- 61 % of the FP64 pipe (profiles/r02_ncu_full_summary.json).  Here every
// scheduler runs THREE warps of 168 registers, one per role, on the same (camera, 32-frame tile) units:
//   * warps 0-3, PROJECT: pose -> R_cf, t_cf once per unit (the next unit's tile, frame order and pose are
//     requested a unit ahead), then two corners per step: observation load, projection, distortion, robust
//     weights, both raw Jacobian rows, cost / RMS bookkeeping; the rows go into a shared-memory ring;
//   * warps 4-7, ACCUMULATE U: the u row of every corner -- accumulator blocks 0 (columns fx, cx) and
//     2-3 (columns k1, k2, m, G): 65 accumulators, 75 FMAs per corner, nothing else;
//   * warps 8-11, ACCUMULATE V: the v row -- blocks 1 (fy, cy) and 2-3.
// At the end of a unit the V warp hands its partial sums of the shared blocks 2-3 (44 doubles per lane) to
// the U warp through shared memory (two named barriers per pair), each warp adds its own blocks to the
// camera sums (register transpose-reduction, lane-private running sums in shared memory) and writes its
// part of the 63-double hand-off to K2c.  No value is computed twice; the rows are the only extra traffic
// (22 doubles per lane and corner written once, read once).
// Ring protocol, per scheduler triple: monotone counters.  prod = corners written, cons_u / cons_v =
// corners read by each accumulator; the projector writes a stage once seq - min(cons) < kWsStages, an
// accumulator reads it once prod > seq.  Counters and stage data go through volatile accesses (program order
// kept by ptxas); the projector fences (CTA scope) between its data and its counter; an accumulator
// publishes its counter after the FMAs that consumed the stage have issued, i.e. after its loads returned.
// Camera sums: every triple parks its running sums in `partW` when its camera changes and the CTA adds the
// four triples in a fixed order at the end: deterministic, same partU / partS layout as the one-role kernel.
#pragma once
#include "k2_producer.cuh"

namespace mcba {

// sin and cos of th >= 0 without the call into CUDA's Payne-Hanek slow path: three-constant Cody-Waite
// reduction to [-pi/4, pi/4] and the fdlibm kernel polynomials (<= 1.4 ulp for th <= 1e3 against 200-bit
// references).  The first design of this experiment needed it (ptxas 12.9 crashes on a call inside a
// setmaxnreg region); kept so that the projector warp has no call at all.
__device__ __forceinline__ void sincos_inline(double th, double* s_out, double* c_out) {
  th = th > 1.0e5 ? fma(-floor(th * 0.15915494309189535), 6.283185307179586, th) : th;
  const double k = rint(th * 6.36619772367581382433e-01);
  double r = fma(-k, 1.57079632673412561417e+00, th);
  r = fma(-k, 6.07710050630396597660e-11, r);
  r = fma(-k, 2.02226624871116645580e-21, r);
  const double z = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, z, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  const double s = fma(r * z, ps, r);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double c = fma(z * z, pc, fma(-0.5, z, 1.0));
  const int q = (int)k & 3;
  const double ss = (q & 1) ? c : s, cc = (q & 1) ? s : c;
  *s_out = (q & 2) ? -ss : ss;
  *c_out = (q == 1 || q == 2) ? -cc : cc;
}
__device__ __forceinline__ void rodrigues_inline(const double r[3], double R[9]) {
  const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  const double inv = (th == 0.0) ? 1.0 : 1.0 / th;
  const double kx = r[0] * inv, ky = r[1] * inv, kz = r[2] * inv;
  double s, c;
  sincos_inline(th, &s, &c);
  const double oc = 1.0 - c;
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

struct K2PWsParams : K2PParams {
  double* partW = nullptr;   // [grid][4][C][kAcc] scratch: the triples' camera sums before the CTA adds them
};

#ifndef MCBA_WS_STAGES
#define MCBA_WS_STAGES 6
#endif
#ifndef MCBA_WS_FENCE
#define MCBA_WS_FENCE 1
#endif
#ifndef MCBA_WS_DEBUG
#define MCBA_WS_DEBUG 0
#endif
constexpr int kWsStages = MCBA_WS_STAGES;   // corners in flight per triple
constexpr int kWsVec = 10;                  // double2 per (corner, lane): 5 for row u, 5 for row v
constexpr int kWsTriples = 4;
constexpr int kWsAccWarps = kWsTriples;     // partW rows per CTA
constexpr int kWsThreads = 384;
constexpr int kWsShared = 44;               // accumulators of blocks 2-3 (36 of A + 8 of q)
constexpr unsigned kWsStageBytes = kWsVec * kTile * 16u;

__device__ __forceinline__ unsigned ws_ld_flag(unsigned addr) {
  unsigned v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void ws_st_flag(unsigned addr, unsigned v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ double2 ws_ld2(unsigned addr) {
  double2 v;
  asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void ws_st2(unsigned addr, double a, double b) {
  asm volatile("st.volatile.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void ws_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void ws_bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// Camera and tile of the j-th live unit of the camera-major list (c only moves forward).
__device__ __forceinline__ long long ws_unit_tile(const K2PWsParams& p, const int* s_ucum, int j, int& c) {
  while (j >= s_ucum[c + 1]) ++c;
  return p.units[(long long)c * p.nTiles + (j - s_ucum[c])];
}

// What the projector hands over per corner: the shared projection values and the two rows' weights; the
// accumulator warps form their own row from them (12 FMAs each, jac_row).
struct WsRows {
  Proj pr;
  double whu, gfu, whv, gfv;
};

// One corner: the arithmetic of walk_corners (k2_producer.cuh, variant 1) up to the projection and weights.
template <int kLoss>
__device__ __forceinline__ void ws_corner(const K2PWsParams& p, const IntrReg& in, const double (&Rt)[12], double qx,
                                          double qy, double qz, double2 cur, WsRows& o, double& cost_acc,
                                          double& sumsq_acc, double& cnt_acc) {
  const double X = fma(Rt[0], qx, fma(Rt[1], qy, fma(Rt[2], qz, Rt[9])));
  const double Y = fma(Rt[3], qx, fma(Rt[4], qy, fma(Rt[5], qz, Rt[10])));
  const double iz = fast_rcp(fma(Rt[6], qx, fma(Rt[7], qy, fma(Rt[8], qz, Rt[11]))));
  project_tail(in, X * iz, Y * iz, iz, o.pr);
  {
    const bool hu = cur.x == cur.x;
    const double fu = hu ? cur.x - o.pr.pu : 0.0;
    double rho, wg;
    robust_weights_t<kLoss>(fu, hu, p.inv_c, p.c2, rho, wg, o.whu);
    cost_acc += rho;
    sumsq_acc = fma(fu, fu, sumsq_acc);
    cnt_acc += hu ? 1.0 : 0.0;
    o.gfu = -wg * fu;
  }
  {
    const bool hv = cur.y == cur.y;
    const double fv = hv ? cur.y - o.pr.pv : 0.0;
    double rho, wg;
    robust_weights_t<kLoss>(fv, hv, p.inv_c, p.c2, rho, wg, o.whv);
    cost_acc += rho;
    sumsq_acc = fma(fv, fv, sumsq_acc);
    cnt_acc += hv ? 1.0 : 0.0;
    o.gfv = -wg * fv;
  }
}

__device__ __forceinline__ void ws_store_rows(unsigned st, const WsRows& r) {
  ws_st2(st + 0 * 512, r.pr.x, r.pr.y);
  ws_st2(st + 1 * 512, r.pr.iz, r.pr.r2);
  ws_st2(st + 2 * 512, r.pr.d, r.pr.A00);
  ws_st2(st + 3 * 512, r.pr.A01, r.pr.su);
  ws_st2(st + 4 * 512, r.whu, r.gfu);
  ws_st2(st + 5 * 512, r.pr.x, r.pr.y);
  ws_st2(st + 6 * 512, r.pr.iz, r.pr.r2);
  ws_st2(st + 7 * 512, r.pr.d, r.pr.A10);
  ws_st2(st + 8 * 512, r.pr.A11, r.pr.sv);
  ws_st2(st + 9 * 512, r.whv, r.gfv);
}

// ------------------------------------------------------------------ projector warp
template <int kLoss>
__device__ __forceinline__ void ws_project_role(const K2PWsParams& p, const CamConst* s_cam, const int* s_ucum,
                                                const double* s_obj, unsigned ring, unsigned flags, int j0, int j1,
                                                int lane, double& cost_acc, double& sumsq_acc, double& cnt_acc) {
  const int C = p.C, N = p.N, nc = 12 * C;
  const int nu = j1 - j0;
  if (nu <= 0) return;
  // Units: `cur` is the one under work, `a` the one after it (its pose is requested when `cur` starts),
  // `b` the one after that (its tile and frame index are requested when `cur` starts).
  long long tile_a, tile_b = 0, f_b = -1;
  int c_cur = 0, c_a = 0, c_b = 0;
  double pose_pending[6];
  tile_a = ws_unit_tile(p, s_ucum, j0, c_a);
  {
    const long long f = p.perm[tile_a * kTile + lane];
#pragma unroll
    for (int i = 0; i < 6; ++i) pose_pending[i] = f >= 0 ? p.x[(size_t)nc + f * 6 + i] : 0.0;
  }
  if (nu > 1) {
    c_b = c_a;
    tile_b = ws_unit_tile(p, s_ucum, j0 + 1, c_b);
    f_b = p.perm[tile_b * kTile + lane];
  }
  const double2 missing = make_double2(nan(""), nan(""));
  const double2* ob_a = p.obs + ((size_t)(tile_a * C + c_a) * N) * kTile + lane;   // observations of unit `a`
  // first two corners of the first unit
  double2 n0 = ob_a[0], n1 = N > 1 ? ob_a[kTile] : missing;
  unsigned seq = 0, cons_seen = 0, stage = 0;
  const unsigned lane_off = (unsigned)lane * 16u;
  for (int u = 0; u < nu; ++u) {
    // ---- start of unit u: `a` becomes `cur`
    const double2* ob = ob_a;
    c_cur = c_a;
    double Rt[12];
    IntrReg in;
    {
      const CamConst& cam = s_cam[c_cur];
      double Rp[9], Rcf[9], tcf[3];
      rodrigues_inline(pose_pending, Rp);
      mat3_mul(cam.R, Rp, Rcf);
      mat3_vec(cam.R, pose_pending + 3, tcf);
#pragma unroll
      for (int i = 0; i < 9; ++i) Rt[i] = Rcf[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) Rt[9 + i] = tcf[i] + cam.t[i];
      in = IntrReg{opaque(cam.fx), opaque(cam.fy), opaque(cam.cx), opaque(cam.cy), opaque(cam.k1), opaque(cam.k2)};
    }
    tile_a = tile_b;
    c_a = c_b;
    ob_a = p.obs + ((size_t)(tile_a * C + c_a) * N) * kTile + lane;
#pragma unroll
    for (int i = 0; i < 6; ++i) pose_pending[i] = f_b >= 0 ? p.x[(size_t)nc + f_b * 6 + i] : 0.0;
    if (u + 2 < nu) {
      tile_b = ws_unit_tile(p, s_ucum, j0 + u + 2, c_b);
      f_b = p.perm[tile_b * kTile + lane];
    } else {
      f_b = -1;
    }
    const bool more_units = u + 1 < nu;
    // ---- corners, two per step
    const double2* ob_next = more_units ? ob_a : ob;   // where the look-ahead of the last pair points
    for (int n = 0; n < N; n += 2) {
      const double2 cur0 = n0, cur1 = n1;
      // the pair after this one: same unit, or the first pair of the next unit (branch-free; an index past
      // the unit is clamped and its value masked by `two` at its own step)
      {
        const bool last = n + 2 >= N;
        const double2* q0 = last ? ob_next : ob + (size_t)(n + 2) * kTile;
        const int i1 = last ? (N > 1 ? 1 : 0) : (n + 3 < N ? 1 : 0);
        n0 = q0[0];
        n1 = q0[(size_t)i1 * kTile];
      }
      const bool two = n + 1 < N;
      const int m = two ? n + 1 : n;
      WsRows r0, r1;
      ws_corner<kLoss>(p, in, Rt, s_obj[3 * n], s_obj[3 * n + 1], s_obj[3 * n + 2], cur0, r0, cost_acc, sumsq_acc,
                       cnt_acc);
      // a missing observation (NaN) adds nothing to the scalar sums
      ws_corner<kLoss>(p, in, Rt, s_obj[3 * m], s_obj[3 * m + 1], s_obj[3 * m + 2], two ? cur1 : missing, r1, cost_acc,
                       sumsq_acc, cnt_acc);
      // ring space for both corners
      const unsigned last = seq + (two ? 1u : 0u);
      while (last - cons_seen >= (unsigned)kWsStages) {
        cons_seen = min(ws_ld_flag(flags + 4), ws_ld_flag(flags + 8));
        if (last - cons_seen >= (unsigned)kWsStages) __nanosleep(64);
      }
      ws_store_rows(ring + stage * kWsStageBytes + lane_off, r0);
      stage = stage + 1 == (unsigned)kWsStages ? 0u : stage + 1;
      if (two) {
        ws_store_rows(ring + stage * kWsStageBytes + lane_off, r1);
        stage = stage + 1 == (unsigned)kWsStages ? 0u : stage + 1;
      }
      seq = last + 1;
      __syncwarp();
#if MCBA_WS_FENCE
      asm volatile("fence.acq_rel.cta;" ::: "memory");
#endif
      ws_st_flag(flags, seq);   // every lane stores the same value: no lane test
    }
  }
}

// ------------------------------------------------------------------ accumulator warps
// kV = false: row u (blocks 0, 2, 3; owner of the shared blocks), kV = true: row v (blocks 1, 2, 3).
template <bool kV>
__device__ __forceinline__ void ws_accumulate_role(const K2PWsParams& p, const CamConst* s_cam, const int* s_ucum, unsigned ring, unsigned flags,
                                                   double* uw, double* xchg, double* partW_triple, int bar_full,
                                                   int bar_free, int j0, int j1, int lane) {
  const int C = p.C, N = p.N;
  const int nu = j1 - j0;
  if (nu <= 0) return;
  constexpr int b0 = kV ? 1 : 0;   // own exclusive block
#pragma unroll
  for (int b = 0; b < 4; ++b)
    if (b == b0 || (!kV && b >= 2)) uw[32 * b + lane] = 0.0;
  unsigned ready = 0, seq = 0, stage = 0;
  int c = 0, c_prev = -1;
  const unsigned my_row = ring + (unsigned)lane * 16u + (kV ? 5u * 512u : 0u);
  const unsigned my_cons = flags + (kV ? 8u : 4u);
  for (int u = 0; u < nu; ++u) {
    const long long tile = ws_unit_tile(p, s_ucum, j0 + u, c);
    if (c != c_prev) {
      if (c_prev >= 0) {
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (b == b0 || (!kV && b >= 2)) {
            partW_triple[(size_t)c_prev * kAcc + 32 * b + lane] = uw[32 * b + lane];
            uw[32 * b + lane] = 0.0;
          }
      }
      c_prev = c;
    }
    const double fcam = opaque(kV ? s_cam[c].fy : s_cam[c].fx);
    double acc[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;
#pragma unroll 1
    for (int n = 0; n < N; ++n) {
      while (ready <= seq) {
        ready = ws_ld_flag(flags);
        if (ready <= seq) __nanosleep(32);
      }
      const unsigned st = my_row + stage * kWsStageBytes;
      const double2 r0 = ws_ld2(st + 0 * 512), r1 = ws_ld2(st + 1 * 512), r2 = ws_ld2(st + 2 * 512),
                    r3 = ws_ld2(st + 3 * 512), r4 = ws_ld2(st + 4 * 512);
      double a[10];
      {
        // jac_row (k2_producer.cuh) from the handed-over values: x y | iz r2 | d Aa | Ab s
        const double x = r0.x, y = r0.y, iz = r1.x, rr = r1.y, d = r2.x, Aa = r2.y, Ab = r3.x, sv = r3.y;
        const double w = kV ? y : x;
        a[0] = w * d;
        a[1] = 1.0;
        a[2] = fcam * w * rr;
        a[3] = a[2] * rr;
        a[4] = -fma(y, sv, Ab);
        a[5] = fma(x, sv, Aa);
        a[6] = fma(x, Ab, -y * Aa);
        a[7] = Aa * iz;
        a[8] = Ab * iz;
        a[9] = -sv * iz;
      }
#if MCBA_WS_DEBUG == 1
      acc[64] += a[0] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7] + a[8] + a[9] + r4.x + r4.y;
#else
      accumulate_row<!kV>(acc, a, r4.x, r4.y);
#endif
      ++seq;
      stage = stage + 1 == (unsigned)kWsStages ? 0u : stage + 1;
      // every load of this corner has returned (its values were just used): the stage is free.  All lanes
      // store the same value: no lane test.
      ws_st_flag(my_cons, seq);
    }
    // ---- end of the unit
    double* h = p.H + ((size_t)(tile * C + c) * kHandoff) * kTile + lane;
    if (kV) {
      if (u > 0) ws_bar_sync(bar_free);   // the U warp has read the previous unit's sums
#pragma unroll
      for (int i = 0; i < kWsShared; ++i) xchg[i * kTile + lane] = acc[64 + i];
      ws_bar_arrive(bar_full);
      uw[32 + lane] += lane_transpose_sum32<32>(acc, lane);
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        h[(size_t)(1 * 6 + j) * kTile] = acc[acc_slot(1, 6 + j)];
        h[(size_t)(3 * 6 + j) * kTile] = acc[acc_slot(3, 6 + j)];
      }
    } else {
      uw[lane] += lane_transpose_sum32<0>(acc, lane);
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        h[(size_t)(0 * 6 + j) * kTile] = acc[acc_slot(0, 6 + j)];
        h[(size_t)(2 * 6 + j) * kTile] = acc[acc_slot(2, 6 + j)];
      }
      ws_bar_sync(bar_full);
#pragma unroll
      for (int i = 0; i < kWsShared; ++i) acc[64 + i] += xchg[i * kTile + lane];
      if (u + 1 < nu) ws_bar_arrive(bar_free);
      uw[64 + lane] += lane_transpose_sum32<64>(acc, lane);
      uw[96 + lane] += lane_transpose_sum32<96>(acc, lane);
#pragma unroll
      for (int i = 4; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) h[(size_t)(i * 6 + j) * kTile] = acc[acc_slot(i, 6 + j)];
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int q = r; q < 6; ++q) h[(size_t)(36 + tri6(r, q)) * kTile] = acc[acc_slot(6 + r, 6 + q)];
#pragma unroll
      for (int r = 0; r < 6; ++r) h[(size_t)(57 + r) * kTile] = acc[acc_slot_q(6 + r)];
    }
  }
#pragma unroll
  for (int b = 0; b < 4; ++b)
    if (b == b0 || (!kV && b >= 2)) partW_triple[(size_t)c_prev * kAcc + 32 * b + lane] = uw[32 * b + lane];
}

struct K2PWsShared {
  size_t ring, xchg, obj, uw, red, cam, ucum, flags, crange, total;   // byte offsets into dynamic shared memory
};
__host__ __device__ inline K2PWsShared k2p_ws_layout(int C, int N) {
  K2PWsShared s;
  size_t o = 0;
  s.ring = o; o += (size_t)kWsTriples * kWsStages * kWsStageBytes;
  s.xchg = o; o += sizeof(double) * kWsTriples * kWsShared * kTile;
  s.obj = o; o += sizeof(double) * ((3 * N + 1) & ~1);
  s.uw = o; o += sizeof(double) * kWsTriples * kAcc;
  s.red = o; o += sizeof(double) * kWsTriples * 4;
  s.cam = o; o += sizeof(CamConst) * (size_t)C;
  s.ucum = o; o += sizeof(int) * (size_t)((C + 2) & ~1);
  s.flags = o; o += sizeof(unsigned) * kWsTriples * 4;   // per triple: prod, cons_u, cons_v, pad
  s.crange = o; o += sizeof(int) * kWsTriples * 2;
  s.total = o;
  return s;
}
inline size_t k2p_ws_smem(int C, int N) { return k2p_ws_layout(C, N).total; }

template <int kLoss>
__global__ void __launch_bounds__(kWsThreads, 1) k2p_ws_kernel(const K2PWsParams p) {
  extern __shared__ __align__(16) unsigned char smem_ws[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int role = warp >> 2, t = warp & 3;   // role 0 project, 1 accumulate u, 2 accumulate v; triple t
  const int C = p.C, N = p.N;
  const K2PWsShared L = k2p_ws_layout(C, N);
  double* s_obj = reinterpret_cast<double*>(smem_ws + L.obj);
  CamConst* s_cam = reinterpret_cast<CamConst*>(smem_ws + L.cam);
  int* s_ucum = reinterpret_cast<int*>(smem_ws + L.ucum);
  unsigned* s_flags = reinterpret_cast<unsigned*>(smem_ws + L.flags);
  int* s_crange = reinterpret_cast<int*>(smem_ws + L.crange);
  double* s_uw = reinterpret_cast<double*>(smem_ws + L.uw);
  double* s_red = reinterpret_cast<double*>(smem_ws + L.red);
  double* s_xchg = reinterpret_cast<double*>(smem_ws + L.xchg);
  if (threadIdx.x < C) {
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
  if (threadIdx.x == 32) {
    int a = 0;
    for (int c = 0; c < C; ++c) { s_ucum[c] = a; a += p.unit_count[c]; }
    s_ucum[C] = a;
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kWsTriples * 4) s_flags[threadIdx.x - 64] = 0u;
  __syncthreads();
  const int total_units = s_ucum[C];
  const long long gw = (long long)blockIdx.x * kWsTriples + t, nW = (long long)gridDim.x * kWsTriples;
  const int j0 = (int)((long long)total_units * gw / nW), j1 = (int)((long long)total_units * (gw + 1) / nW);
  const unsigned ring = (unsigned)__cvta_generic_to_shared(smem_ws + L.ring) + (unsigned)t * (kWsStages * kWsStageBytes);
  const unsigned flags = (unsigned)__cvta_generic_to_shared(s_flags + 4 * t);
  double* partW_triple = p.partW + ((size_t)blockIdx.x * kWsTriples + t) * C * kAcc;
  if (role == 0) {
    double cost_acc = 0.0, sumsq_acc = 0.0, cnt_acc = 0.0;
    if (lane == 0) {
      int c0 = 1, c1 = 0;
      if (j1 > j0) {
        c0 = 0;
        while (j0 >= s_ucum[c0 + 1]) ++c0;
        c1 = c0;
        while (j1 - 1 >= s_ucum[c1 + 1]) ++c1;
      }
      s_crange[2 * t] = c0;
      s_crange[2 * t + 1] = c1;
    }
    ws_project_role<kLoss>(p, s_cam, s_ucum, s_obj, ring, flags, j0, j1, lane, cost_acc, sumsq_acc, cnt_acc);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      cost_acc += __shfl_xor_sync(0xffffffffu, cost_acc, off);
      sumsq_acc += __shfl_xor_sync(0xffffffffu, sumsq_acc, off);
      cnt_acc += __shfl_xor_sync(0xffffffffu, cnt_acc, off);
    }
    if (lane == 0) {
      double* r = s_red + t * 4;
      r[0] = cost_acc; r[1] = sumsq_acc; r[2] = cnt_acc;
    }
  } else if (role == 1) {
    ws_accumulate_role<false>(p, s_cam, s_ucum, ring, flags, s_uw + t * kAcc, s_xchg + (size_t)t * kWsShared * kTile,
                              partW_triple, 1 + 2 * t, 2 + 2 * t, j0, j1, lane);
  } else {
    ws_accumulate_role<true>(p, s_cam, s_ucum, ring, flags, s_uw + t * kAcc, s_xchg + (size_t)t * kWsShared * kTile,
                             partW_triple, 1 + 2 * t, 2 + 2 * t, j0, j1, lane);
  }
  __syncthreads();
  // ---------------- CTA epilogue: the four triples' camera sums in a fixed order, scalars ----------------
  const double* pw = p.partW + (size_t)blockIdx.x * kWsTriples * C * kAcc;
  double* pu = p.partU + (size_t)blockIdx.x * C * kAcc;
  for (int i = threadIdx.x; i < C * kAcc; i += blockDim.x) {
    const int c = i / kAcc;
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kWsTriples; ++w)
      if (c >= s_crange[2 * w] && c <= s_crange[2 * w + 1]) s += pw[(size_t)w * C * kAcc + i];
    pu[i] = s;
  }
  if (threadIdx.x == 0) {
    double a = 0, b = 0, cn = 0;
    for (int w = 0; w < kWsTriples; ++w) {
      a += s_red[w * 4];
      b += s_red[w * 4 + 1];
      cn += s_red[w * 4 + 2];
    }
    double* ps = p.partS + (size_t)blockIdx.x * kRsNum;
    ps[kRsCost] = 0.5 * a;
    ps[kRsSumSq] = b;
    ps[kRsCount] = cn;
    ps[kRsGmaxPose] = 0.0;
  }
}

}  // namespace mcba

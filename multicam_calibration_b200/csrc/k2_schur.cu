// K2b: S_raw = sum_f Z_f Z_f^T and Z_f y_f (register-tiled SYRK over the frame
// axis), the finalisation of the packed reduced camera system in the true
// parameter basis, the damped Cholesky solve and K3, the pose back-substitution.
// Math: SURVEY.md Appendix A ("Schur form for this problem").
#include "k2_common.cuh"

namespace mcba {

__host__ __device__ inline int tile_index(int bi, int bj, int nb) { return bi * nb - (bi * (bi - 1)) / 2 + (bj - bi); }

// ------------------------------------------------------------------ SYRK
// S_raw = sum_f Z_f Z_f^T (upper block triangle) and sum_f Z_f y_f.
// Slot s < nT owns the 6x6 block (bi <= bj) of S_raw, slots nT .. nT+nb-1 own 6 entries
// of Z y.  The K axis (frame, pose column) is split over KG thread groups; each
// persistent CTA streams its contiguous range of Z through a 2-stage shared-memory
// ring filled by TMA bulk copies (cp.async.bulk + mbarrier) so loads overlap the FMAs.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int kSyrkStages = 2;

__global__ void __launch_bounds__(384, 1) k2_syrk_kernel(const double* __restrict__ Z, const double* __restrict__ y,
                                                         int nc, int nb, int nT, long long F, int FB, int slots_pad,
                                                         int KG, long long frames_per_cta, double* __restrict__ part) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const size_t slab_doubles = (size_t)FB * 6 * nc;
  const size_t stage_doubles = slab_doubles + (size_t)FB * 6;     // Z slab + y slab
  double* stages = reinterpret_cast<double*>(smem_raw);
  __shared__ unsigned long long full_bar[kSyrkStages];

  const int slot_local = threadIdx.x % slots_pad, kg = threadIdx.x / slots_pad;
  const int slot = blockIdx.y * slots_pad + slot_local;
  int kind = 2, bi = 0, bj = 0;
  if (slot_local < slots_pad && slot < nT) {
    kind = 0;
    int rem = slot;
    while (rem >= nb - bi) { rem -= nb - bi; ++bi; }
    bj = bi + rem;
  } else if (slot < nT + nb) {
    kind = 1;
    bi = slot - nT;
  }
  double acc[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) acc[i] = 0.0;

  const long long f_begin = blockIdx.x * frames_per_cta;
  long long f_end = f_begin + frames_per_cta;
  if (f_end > F) f_end = F;
  const int n_chunks = f_end > f_begin ? (int)((f_end - f_begin + FB - 1) / FB) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kSyrkStages; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int chunk) {
    const int s = chunk % kSyrkStages;
    const long long f0 = f_begin + (long long)chunk * FB;
    const int nfr = (int)((f_end - f0) < FB ? (f_end - f0) : FB);
    double* dst = stages + (size_t)s * stage_doubles;
    const unsigned zb = (unsigned)(nfr * 6 * nc * sizeof(double)), yb = (unsigned)(nfr * 6 * sizeof(double));
    mbar_expect_tx(&full_bar[s], zb + yb);
    tma_load_1d(dst, Z + (size_t)f0 * 6 * nc, zb, &full_bar[s]);
    tma_load_1d(dst + slab_doubles, y + (size_t)f0 * 6, yb, &full_bar[s]);
  };
  if (threadIdx.x == 0) {
    for (int c = 0; c < kSyrkStages && c < n_chunks; ++c) issue(c);
  }
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int s = chunk % kSyrkStages;
    mbar_wait(&full_bar[s], (unsigned)((chunk / kSyrkStages) & 1));
    const long long f0 = f_begin + (long long)chunk * FB;
    const int nk = (int)((f_end - f0) < FB ? (f_end - f0) : FB) * 6;
    const double* slab = stages + (size_t)s * stage_doubles;
    if (kind == 0) {
      const double* ra = slab + 6 * bi;
      const double* rb = slab + 6 * bj;
#pragma unroll 2
      for (int kk = kg; kk < nk; kk += KG) {
        const double2 a0 = *reinterpret_cast<const double2*>(ra + (size_t)kk * nc);
        const double2 a1 = *reinterpret_cast<const double2*>(ra + (size_t)kk * nc + 2);
        const double2 a2 = *reinterpret_cast<const double2*>(ra + (size_t)kk * nc + 4);
        const double2 b0 = *reinterpret_cast<const double2*>(rb + (size_t)kk * nc);
        const double2 b1 = *reinterpret_cast<const double2*>(rb + (size_t)kk * nc + 2);
        const double2 b2 = *reinterpret_cast<const double2*>(rb + (size_t)kk * nc + 4);
        const double a[6] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y};
        const double b[6] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y};
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 6; ++j) acc[i * 6 + j] = fma(a[i], b[j], acc[i * 6 + j]);
      }
    } else if (kind == 1) {
      const double* ra = slab + 6 * bi;
      const double* ysm = slab + slab_doubles;
      for (int kk = kg; kk < nk; kk += KG) {
        const double yk = ysm[kk];
#pragma unroll
        for (int i = 0; i < 6; ++i) acc[i] = fma(ra[(size_t)kk * nc + i], yk, acc[i]);
      }
    }
    __syncthreads();   // everyone is done with stage s: refill it
    if (threadIdx.x == 0 && chunk + kSyrkStages < n_chunks) issue(chunk + kSyrkStages);
  }

  // reduce the KG partial accumulators through shared memory, then one partial per CTA
  double* red = stages;   // [KG][slots_pad][36]
  if (kind != 2) {
#pragma unroll
    for (int i = 0; i < 36; ++i) red[((size_t)kg * slots_pad + slot_local) * 36 + i] = acc[i];
  }
  __syncthreads();
  double* out = part + (size_t)blockIdx.x * ((size_t)nT * 36 + (size_t)nb * 6);
  for (int e = threadIdx.x; e < slots_pad * 36; e += blockDim.x) {
    const int sl = e / 36, i = e % 36;
    const int gs = blockIdx.y * slots_pad + sl;
    if (gs >= nT + nb) continue;
    double v = 0.0;
    for (int g = 0; g < KG; ++g) v += red[((size_t)g * slots_pad + sl) * 36 + i];
    if (gs < nT) out[(size_t)gs * 36 + i] = v;
    else if (i < 6) out[(size_t)nT * 36 + (gs - nT) * 6 + i] = v;
  }
}

struct SyrkConfig {
  int threads, gy, FB, slots_pad, KG;
  size_t smem;
};

static SyrkConfig syrk_config(int nc) {
  SyrkConfig c;
  const int nb = nc / 6, nT = nb * (nb + 1) / 2, slots = nT + nb;
  c.gy = (slots + 287) / 288;
  c.slots_pad = (((slots + c.gy - 1) / c.gy) + 31) / 32 * 32;
  c.KG = 384 / c.slots_pad;
  if (c.KG < 1) c.KG = 1;
  c.threads = c.slots_pad * c.KG;
  c.FB = nc <= 96 ? 16 : (nc <= 192 ? 8 : 4);
  const size_t stage = sizeof(double) * ((size_t)c.FB * 6 * nc + (size_t)c.FB * 6);
  const size_t red = sizeof(double) * (size_t)c.KG * c.slots_pad * 36;
  c.smem = kSyrkStages * stage > red ? kSyrkStages * stage : red;
  return c;
}

int syrk_grid(int nc, long long F, int n_sm) {
  const SyrkConfig c = syrk_config(nc);
  long long g = (F + c.FB - 1) / c.FB;
  return (int)(g < n_sm ? g : n_sm);
}

int launch_k2_syrk(mcba_handle* h) {
  const Layout& L = h->L;
  const SyrkConfig c = syrk_config(L.nc);
  const int nb = L.nc / 6, nT = nb * (nb + 1) / 2;
  const int gx = h->grid_syrk;
  long long fpc = (L.F + gx - 1) / gx;
  fpc = (fpc + c.FB - 1) / c.FB * c.FB;   // whole chunks per CTA keep every TMA source 16-byte aligned
  MCBA_CUDA(cudaFuncSetAttribute(k2_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
  k2_syrk_kernel<<<dim3(gx, c.gy), c.threads, c.smem, h->stream>>>(h->d_Z, h->d_y, L.nc, nb, nT, L.F, c.FB,
                                                                  c.slots_pad, c.KG, fpc, h->d_partSyrk);
  h->launches++;
  MCBA_CUDA(cudaGetLastError());
  return MCBA_OK;
}

// ------------------------------------------------------------------ finalize
// Block (c, c'), c <= c':  S0_cc' = T_c^T ( [c==c'] U_raw,c - (Z Z^T)_cc' ) T_c'
// T_c = [[I6,0,0],[0,Jl,0],[0,[t]x Jl,I3]]  (rows raw [intr | m | G], cols true [intr | r | t]).
struct FinalizeParams {
  int C, nc, nb, nT, nPartU, nPartSyrk, nPartScal, rank;
  const CamConst* cams;
  const double* partU;     // [nPartU][C][kAcc]
  const double* partS;     // [nPartScal][kRsNum]  (K2p: cost, sum f^2, count)
  const double* partG;     // [nPartG] max |pose gradient| per tile (K2c)
  long long nPartG;
  const double* partSyrk;  // [nPartSyrk][nT*36 + nb*6]
  double* red;
  long long offS, offB, offG, offDiag, offScal, offRank;
};

__device__ __forceinline__ void build_T(const CamConst& cam, double* T /*[144]*/, int tid) {
  if (tid < 144) {
    const int r = tid / 12, q = tid % 12;
    double v = (r == q) ? 1.0 : 0.0;
    if (r >= 6 && r < 9 && q >= 6 && q < 9) v = cam.Jl[(r - 6) * 3 + (q - 6)];
    if (r >= 9 && q >= 6 && q < 9) v = cam.tJ[(r - 9) * 3 + (q - 6)];
    T[tid] = v;
  }
}

__global__ void __launch_bounds__(160) finalize_kernel(const FinalizeParams p) {
  __shared__ double M[144], X[144], Tc[144], Tp[144], vec[24];
  const int tid = threadIdx.x;
  const int blk = blockIdx.x;
  if (blk == p.C * p.C) {  // scalars
    double a = 0, b = 0, k = 0, g = 0;
    for (int i = tid; i < p.nPartScal; i += blockDim.x) {
      const double* s = p.partS + (size_t)i * kRsNum;
      a += s[kRsCost]; b += s[kRsSumSq]; k += s[kRsCount];
    }
    for (long long i = tid; i < p.nPartG; i += blockDim.x) g = fmax(g, p.partG[i]);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      b += __shfl_xor_sync(0xffffffffu, b, off);
      k += __shfl_xor_sync(0xffffffffu, k, off);
      g = fmax(g, __shfl_xor_sync(0xffffffffu, g, off));
    }
    if ((tid & 31) == 0) { M[(tid >> 5) * 4] = a; M[(tid >> 5) * 4 + 1] = b; M[(tid >> 5) * 4 + 2] = k; M[(tid >> 5) * 4 + 3] = g; }
    __syncthreads();
    if (tid == 0) {
      a = b = k = g = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += M[w * 4]; b += M[w * 4 + 1]; k += M[w * 4 + 2]; g = fmax(g, M[w * 4 + 3]); }
      double* sc = p.red + p.offScal;
      for (int i = 0; i < kRsNum; ++i) sc[i] = 0.0;
      sc[kRsCost] = a; sc[kRsSumSq] = b; sc[kRsCount] = k;
      for (int i = 0; i < kMaxRanks; ++i) p.red[p.offRank + i] = 0.0;
      p.red[p.offRank + p.rank] = g;
    }
    return;
  }
  const int c = blk / p.C, cp = blk % p.C;
  if (c > cp) return;
  const size_t strideSyrk = (size_t)p.nT * 36 + (size_t)p.nb * 6;
  build_T(p.cams[c], Tc, tid);
  build_T(p.cams[cp], Tp, tid);
  double uraw = 0.0;
  if (tid < 144) {
    const int i = tid / 12, j = tid % 12;
    const int r = 12 * c + i, q = 12 * cp + j;
    int bi = r / 6, bj = q / 6, e;
    if (bi <= bj) e = (r % 6) * 6 + (q % 6);
    else { const int tmp = bi; bi = bj; bj = tmp; e = (q % 6) * 6 + (r % 6); }
    const double* src = p.partSyrk + (size_t)tile_index(bi, bj, p.nb) * 36 + e;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int k = 0;
    for (; k + 4 <= p.nPartSyrk; k += 4) {
      s0 += src[(size_t)k * strideSyrk];
      s1 += src[(size_t)(k + 1) * strideSyrk];
      s2 += src[(size_t)(k + 2) * strideSyrk];
      s3 += src[(size_t)(k + 3) * strideSyrk];
    }
    for (; k < p.nPartSyrk; ++k) s0 += src[(size_t)k * strideSyrk];
    double s = -((s0 + s1) + (s2 + s3));
    if (c == cp) {
      const int slot = acc_slot(i, j);   // -1: structurally zero product (fx.fy, fx.cy, cx.fy, cx.cy)
      if (slot >= 0) {
        const double* us = p.partU + (size_t)c * kAcc + slot;
        for (int k2 = 0; k2 < p.nPartU; ++k2) uraw += us[(size_t)k2 * p.C * kAcc];
      }
      s += uraw;
    }
    M[tid] = s;
  }
  __syncthreads();
  if (tid < 144) {
    const int i = tid / 12, j = tid % 12;
    double s = 0;
#pragma unroll
    for (int k = 0; k < 12; ++k) s += M[i * 12 + k] * Tp[k * 12 + j];
    X[tid] = s;
  }
  __syncthreads();
  if (tid < 144) {
    const int i = tid / 12, j = tid % 12;
    double s = 0;
#pragma unroll
    for (int k = 0; k < 12; ++k) s += Tc[k * 12 + i] * X[k * 12 + j];
    const int r = 12 * c + i, q = 12 * cp + j;
    if (c != cp || i <= j) {   // diagonal blocks: one thread writes both mirror entries (exact symmetry)
      p.red[p.offS + (size_t)r * p.nc + q] = s;
      p.red[p.offS + (size_t)q * p.nc + r] = s;
    }
  }
  if (c != cp) return;
  // diagonal block extras: diag(T^T U_raw T), b, g_cam
  __syncthreads();
  if (tid < 144) M[tid] = uraw;
  if (tid >= 144 && tid < 156) {
    const int i = tid - 144;
    double graw = 0, zy = 0;
    const double* gs = p.partU + (size_t)c * kAcc + acc_slot_q(i);
    for (int k = 0; k < p.nPartU; ++k) graw += gs[(size_t)k * p.C * kAcc];
    const double* zs = p.partSyrk + (size_t)p.nT * 36 + (size_t)(2 * c + i / 6) * 6 + (i % 6);
    for (int k = 0; k < p.nPartSyrk; ++k) zy += zs[(size_t)k * strideSyrk];
    vec[i] = graw;
    vec[12 + i] = graw - zy;
  }
  __syncthreads();
  if (tid < 144) {
    const int i = tid / 12, j = tid % 12;
    double s = 0;
#pragma unroll
    for (int k = 0; k < 12; ++k) s += M[i * 12 + k] * Tc[k * 12 + j];
    X[tid] = s;
  }
  __syncthreads();
  if (tid < 12) {
    double d = 0, g = 0, b = 0;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      d += Tc[k * 12 + tid] * X[k * 12 + tid];
      g += Tc[k * 12 + tid] * vec[k];
      b += Tc[k * 12 + tid] * vec[12 + k];
    }
    p.red[p.offDiag + 12 * c + tid] = d;
    p.red[p.offG + 12 * c + tid] = g;
    p.red[p.offB + 12 * c + tid] = b;
  }
}

// Deterministic sum over per-CTA partial buffers: out[e] = sum_p part[p][e].
// 64 elements x 4 partial-groups per CTA, 8 independent loads in flight per thread.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ partA, int nPartA, int lenA,
                                                              const double* __restrict__ partB, int nPartB, int lenB,
                                                              double* __restrict__ out) {
  __shared__ double s[4][64];
  const int el = threadIdx.x & 63, pg = threadIdx.x >> 6;
  const int e = blockIdx.x * 64 + el;
  double v = 0.0;
  if (e < lenA + lenB) {
    const double* src = e < lenA ? partA + e : partB + (e - lenA);
    const int np = e < lenA ? nPartA : nPartB;
    const size_t stride = e < lenA ? (size_t)lenA : (size_t)lenB;
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int p = pg;
    for (; p + 28 < np; p += 32) {
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] += __ldcg(src + (size_t)(p + 4 * u) * stride);
    }
    for (; p < np; p += 4) a[0] += __ldcg(src + (size_t)p * stride);
    v = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  }
  s[pg][el] = v;
  __syncthreads();
  if (pg == 0 && e < lenA + lenB) out[e] = (s[0][el] + s[1][el]) + (s[2][el] + s[3][el]);
}

int launch_finalize(mcba_handle* h) {
  const Layout& L = h->L;
  FinalizeParams p;
  p.C = L.C; p.nc = L.nc; p.nb = L.nc / 6; p.nT = p.nb * (p.nb + 1) / 2;
  {
    const int lenA = p.nT * 36 + p.nb * 6, lenB = L.C * kAcc;
    reduce_partials_kernel<<<(lenA + lenB + 63) / 64, 256, 0, h->stream>>>(h->d_partSyrk, h->grid_syrk, lenA, h->d_partU,
                                                                      h->grid_frames, lenB, h->d_Sraw);
    h->launches++;
    MCBA_CUDA(cudaGetLastError());
    p.partSyrk = h->d_Sraw;
    p.partU = h->d_Sraw + lenA;
  }
  p.nPartU = 1; p.nPartSyrk = 1; p.nPartScal = h->grid_frames; p.rank = h->rank;
  p.cams = h->d_cams; p.partS = h->d_partS; p.partG = h->d_partG; p.nPartG = L.nTiles;
  p.red = h->d_red;
  p.offS = L.offS; p.offB = L.offB; p.offG = L.offG; p.offDiag = L.offDiag; p.offScal = L.offScal; p.offRank = L.offRank;
  finalize_kernel<<<L.C * L.C + 1, 160, 0, h->stream>>>(p);
  h->launches++;
  MCBA_CUDA(cudaGetLastError());
  return MCBA_OK;
}

// ------------------------------------------------------------------ damped system
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

int solve_reduced(mcba_handle* h, double lambda) {
  const Layout& L = h->L;
  const int nc = L.nc;
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

// ------------------------------------------------------------------ K3 back-substitution
// delta_f = -L^-T (y_f + Z_f^T delta_raw),  x_new = x + delta.  192 threads = 32 frames x 6.
struct BackParams {
  int C, nc, rank;
  long long F, nTiles;
  const CamConst* cams;
  const double* x;
  double* x_new;
  const double* dcam;   // true-basis camera step (12C)
  const double* Z; const double* Linv; const double* y; const double* gpose; const double* D2pose;
  const double* D2cam; const double* gcam;
  double* part; unsigned int* counter; double* out;  // out[0..3] = |dx|^2, |x|^2, g.dx, dx D2 dx
};

__global__ void __launch_bounds__(192) backsub_kernel(const BackParams p) {
  extern __shared__ double smem[];
  double* draw = smem;            // [nc]
  double* sv = draw + p.nc;       // [192]
  __shared__ double s_red[6 * 4];
  __shared__ bool s_last;
  const int tid = threadIdx.x, nc = p.nc;
  for (int r = tid; r < nc; r += blockDim.x) {
    const int c = r / 12, i = r % 12;
    const double* d = p.dcam + 12 * c;
    const CamConst& cam = p.cams[c];
    double v;
    if (i < 6) v = d[i];
    else if (i < 9) v = cam.Jl[(i - 6) * 3] * d[6] + cam.Jl[(i - 6) * 3 + 1] * d[7] + cam.Jl[(i - 6) * 3 + 2] * d[8];
    else v = cam.tJ[(i - 9) * 3] * d[6] + cam.tJ[(i - 9) * 3 + 1] * d[7] + cam.tJ[(i - 9) * 3 + 2] * d[8] + d[i];
    draw[r] = v;
  }
  __syncthreads();
  double dd = 0, xx = 0, gd = 0, dDd = 0;
  const int fl = tid / 6, k = tid % 6;
  for (long long tile = blockIdx.x; tile < p.nTiles; tile += gridDim.x) {
    const long long f = tile * kTile + fl;
    double v = 0.0;
    if (f < p.F) {
      const double2* zr = reinterpret_cast<const double2*>(p.Z + ((size_t)f * 6 + k) * nc);
      double s0 = 0, s1 = 0;
      for (int r = 0; r < nc / 2; ++r) {
        const double2 z = zr[r];
        s0 = fma(z.x, draw[2 * r], s0);
        s1 = fma(z.y, draw[2 * r + 1], s1);
      }
      v = p.y[(size_t)f * 6 + k] + (s0 + s1);
    }
    sv[tid] = v;
    __syncthreads();
    if (f < p.F) {
      const double* li = p.Linv + (size_t)tile * 21 * kTile + fl;
      double d = 0.0;
      for (int j = k; j < 6; ++j) d -= li[(j * (j + 1) / 2 + k) * kTile] * sv[fl * 6 + j];
      const size_t xi = (size_t)nc + (size_t)f * 6 + k;
      const double xo = p.x[xi];
      p.x_new[xi] = xo + d;
      double d2 = p.D2pose[(size_t)tile * 6 * kTile + k * kTile + fl];
      if (d2 == 0.0) d2 = 1.0;
      dd = fma(d, d, dd);
      xx = fma(xo, xo, xx);
      gd = fma(p.gpose[(size_t)f * 6 + k], d, gd);
      dDd = fma(d2 * d, d, dDd);
    }
    __syncthreads();
  }
  if (blockIdx.x == 0) {
    for (int r = tid; r < nc; r += blockDim.x) {
      const double d = p.dcam[r], xo = p.x[r];
      p.x_new[r] = xo + d;
      if (p.rank == 0) {  // camera terms are counted once across ranks
        double d2 = p.D2cam[r];
        if (d2 == 0.0) d2 = 1.0;
        dd = fma(d, d, dd);
        xx = fma(xo, xo, xx);
        gd = fma(p.gcam[r], d, gd);
        dDd = fma(d2 * d, d, dDd);
      }
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    dd += __shfl_xor_sync(0xffffffffu, dd, off);
    xx += __shfl_xor_sync(0xffffffffu, xx, off);
    gd += __shfl_xor_sync(0xffffffffu, gd, off);
    dDd += __shfl_xor_sync(0xffffffffu, dDd, off);
  }
  const int lane = tid & 31, warp = tid >> 5;
  if (lane == 0) { s_red[warp * 4] = dd; s_red[warp * 4 + 1] = xx; s_red[warp * 4 + 2] = gd; s_red[warp * 4 + 3] = dDd; }
  __syncthreads();
  if (tid == 0) {
    double a = 0, b = 0, c = 0, d = 0;
    for (int w = 0; w < 6; ++w) { a += s_red[w * 4]; b += s_red[w * 4 + 1]; c += s_red[w * 4 + 2]; d += s_red[w * 4 + 3]; }
    double* o = p.part + (size_t)blockIdx.x * 4;
    o[0] = a; o[1] = b; o[2] = c; o[3] = d;
    __threadfence();
    s_last = atomicAdd(p.counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    double a = 0, b = 0, c = 0, d = 0;
    for (unsigned i = tid; i < gridDim.x; i += blockDim.x) {
      a += __ldcg(p.part + i * 4); b += __ldcg(p.part + i * 4 + 1); c += __ldcg(p.part + i * 4 + 2); d += __ldcg(p.part + i * 4 + 3);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      b += __shfl_xor_sync(0xffffffffu, b, off);
      c += __shfl_xor_sync(0xffffffffu, c, off);
      d += __shfl_xor_sync(0xffffffffu, d, off);
    }
    __syncthreads();
    if (lane == 0) { s_red[warp * 4] = a; s_red[warp * 4 + 1] = b; s_red[warp * 4 + 2] = c; s_red[warp * 4 + 3] = d; }
    __syncthreads();
    if (tid == 0) {
      a = b = c = d = 0;
      for (int w = 0; w < 6; ++w) { a += s_red[w * 4]; b += s_red[w * 4 + 1]; c += s_red[w * 4 + 2]; d += s_red[w * 4 + 3]; }
      p.out[0] = a; p.out[1] = b; p.out[2] = c; p.out[3] = d;
      *p.counter = 0;
    }
  }
}

int launch_backsub(mcba_handle* h, const double* x, double* x_new, double lambda) {
  (void)lambda;
  const Layout& L = h->L;
  BackParams p;
  p.C = L.C; p.nc = L.nc; p.rank = h->rank; p.F = L.F; p.nTiles = L.nTiles;
  p.cams = h->d_cams; p.x = x; p.x_new = x_new; p.dcam = h->d_dcam;
  p.Z = h->d_Z; p.Linv = h->d_Linv; p.y = h->d_y; p.gpose = h->d_gpose; p.D2pose = h->d_D2pose;
  p.D2cam = h->d_D2cam; p.gcam = h->d_red + L.offG;
  p.part = h->d_scal + 64 + 3 * 4096;                                     // [grid_back][4]
  p.counter = reinterpret_cast<unsigned int*>(h->d_scal + 33);
  p.out = h->d_scal + 8;
  backsub_kernel<<<h->grid_back, 192, sizeof(double) * (L.nc + 192), h->stream>>>(p);
  h->launches++;
  MCBA_CUDA(cudaGetLastError());
  return MCBA_OK;
}

}  // namespace mcba

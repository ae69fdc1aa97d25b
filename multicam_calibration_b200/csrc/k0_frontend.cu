// K0: the front end of bundle_adjust on the device -- frame
// eligibility, reprojection-error outlier filter and the gather of the kept frames
// (bundle_adjustment.py:265-296, everything up to the RNG sub-sampling, which stays on the
// host because it consumes numpy's global generator).
//
//   eligible[f]   = more than one camera sees the COMPLETE board in frame f           (:266)
//   err[c,f,n]    = || observed - predicted ||_2 on eligible frames (NaN if either scalar is) (:269-276)
//   worst[f]      = nanmax_c nanmean_n err[c,f,n]                                      (:279)
//   threshold     = caller's, or 5 * nanmedian(err)                                    (:281-282)
//   use[f]        = eligible[f] and not (nan_to_num(worst[f]) > threshold)             (:284-285)
//
// The median is exact (radix sort of the error array, NaNs last, middle element or the mean of
// the two middle elements like numpy); at BASELINE configs[2] the numpy version of this block
// costs ~1.2 s of host time against a 4 ms solve.
#include <cstring>

#include <cub/device/device_radix_sort.cuh>

#include "mcba_internal.h"
#include "mcba_obs.cuh"

namespace mcba {

__global__ void prep_cameras_frontend_kernel(const double* __restrict__ x, int C, CamConst* __restrict__ cams) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double* p = x + 12 * c;
  CamConst k;
  k.fx = p[0]; k.fy = p[1]; k.cx = p[2]; k.cy = p[3]; k.k1 = p[4]; k.k2 = p[5];
  const double r[3] = {p[6], p[7], p[8]};
  k.t[0] = p[9]; k.t[1] = p[10]; k.t[2] = p[11];
  rodrigues(r, k.R);
  so3_left_jacobian(r, k.Jl);
  cross_mat3(k.t, k.Jl, k.tJ);
  cams[c] = k;
}

// one warp per (c,f) row: complete[c*F + f] = 1 when no scalar of the row is NaN
__global__ void complete_rows_kernel(const double* __restrict__ uvs, long long rows, int N,
                                     unsigned char* __restrict__ complete) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp; row < rows; row += nwarps) {
    const double* p = uvs + row * 2 * N;
    bool bad = false;
    for (int s = lane; s < 2 * N; s += 32) bad |= !(p[s] == p[s]);
    const bool any_bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) complete[row] = any_bad ? 0 : 1;
  }
}

__global__ void eligible_kernel(const unsigned char* __restrict__ complete, int C, long long F,
                                unsigned char* __restrict__ eligible, unsigned long long* __restrict__ n_eligible) {
  const long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (f >= F) return;
  int cnt = 0;
  for (int c = 0; c < C; ++c) cnt += complete[(long long)c * F + f];
  const bool e = cnt > 1;
  eligible[f] = e ? 1 : 0;
  if (e) atomicAdd(n_eligible, 1ull);
}

// one warp per (c,f) row of an eligible frame: err[c,f,n], mean over the finite ones
__global__ void frame_errors_kernel(const double* __restrict__ uvs, const double* __restrict__ obj,
                                    const double* __restrict__ x, const CamConst* __restrict__ cams,
                                    const unsigned char* __restrict__ eligible, int C, long long F, int N,
                                    double* __restrict__ err, double* __restrict__ mean_err,
                                    unsigned long long* __restrict__ n_finite) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const double qnan = nan("");
  for (long long row = warp; row < (long long)C * F; row += nwarps) {
    const int c = (int)(row / F);
    const long long f = row % F;
    double* e_out = err + row * N;
    if (!eligible[f]) {
      for (int n = lane; n < N; n += 32) e_out[n] = qnan;
      if (lane == 0) mean_err[row] = qnan;
      continue;
    }
    const CamConst& cam = cams[c];
    const Intr in{cam.fx, cam.fy, cam.cx, cam.cy, cam.k1, cam.k2};
    const double* ps = x + 12 * (long long)C + 6 * f;
    const double rho[3] = {ps[0], ps[1], ps[2]}, tau[3] = {ps[3], ps[4], ps[5]};
    double Rp[9], Rcf[9], tcf[3];
    rodrigues(rho, Rp);
    mat3_mul(cam.R, Rp, Rcf);
    mat3_vec(cam.R, tau, tcf);
    tcf[0] += cam.t[0]; tcf[1] += cam.t[1]; tcf[2] += cam.t[2];
    const double2* ob = reinterpret_cast<const double2*>(uvs) + row * N;
    double sum = 0.0;
    int cnt = 0;
    for (int n = lane; n < N; n += 32) {
      const double2 o = ob[n];
      double pu, pv;
      project(in, Rcf, tcf, obj[3 * n], obj[3 * n + 1], obj[3 * n + 2], pu, pv);
      const double du = o.x - pu, dv = o.y - pv;
      double e = sqrt(fma(du, du, dv * dv));
      if (e == e) { sum += e; ++cnt; } else e = qnan;   // canonical NaN: sorts after every finite value
      e_out[n] = e;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, off);
      cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    }
    if (lane == 0) {
      mean_err[row] = cnt ? sum / cnt : qnan;
      if (cnt) atomicAdd(n_finite, (unsigned long long)cnt);
    }
  }
}

// threshold (given, or 5 * median of the sorted finite errors) and the final frame mask
__global__ void select_kernel(const double* __restrict__ sorted_err, const unsigned long long* __restrict__ n_finite,
                              const double* __restrict__ mean_err, const unsigned char* __restrict__ eligible,
                              int C, long long F, double threshold_in, unsigned char* __restrict__ use,
                              double* __restrict__ stats /* [threshold, excluded] */) {
  __shared__ double s_thr;
  if (threadIdx.x == 0) {
    double thr = threshold_in;
    if (!(thr == thr)) {
      const unsigned long long n = *n_finite;
      if (n == 0) thr = nan("");
      else if (n & 1ull) thr = 5.0 * sorted_err[n / 2];
      else thr = 5.0 * (0.5 * (sorted_err[n / 2 - 1] + sorted_err[n / 2]));
    }
    s_thr = thr;
    if (blockIdx.x == 0) stats[0] = thr;
  }
  __syncthreads();
  const double thr = s_thr;
  const long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (f >= F) return;
  unsigned char u = 0;
  if (eligible[f]) {
    double worst = 0.0;   // nan_to_num(nanmax) : all-NaN -> 0
    bool any = false;
    for (int c = 0; c < C; ++c) {
      const double m = mean_err[(long long)c * F + f];
      if (m == m) { worst = any ? fmax(worst, m) : m; any = true; }
    }
    const bool excl = worst > thr;   // false when thr is NaN, like numpy
    u = excl ? 0 : 1;
    if (excl) atomicAdd(reinterpret_cast<unsigned long long*>(stats + 1), 1ull);
  }
  use[f] = u;
}

__global__ void gather_frames_kernel(const double2* __restrict__ src, int C, long long F, int N,
                                     const long long* __restrict__ idx, long long Fu, double2* __restrict__ dst) {
  const long long total = (long long)C * Fu * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const long long r = i / N;
    const long long j = r % Fu;
    const int c = (int)(r / Fu);
    dst[i] = src[((long long)c * F + idx[j]) * N + n];
  }
}

// Histogram of the next 8 key bits of the finite, non-negative values whose leading `prefix_bits`
// bits equal `prefix` (IEEE-754 bit patterns of non-negative doubles order like the values): one
// pass of an exact most-significant-digit radix selection.  Frame-sharded runs find the global
// nanmedian of the per-point errors with eight such passes and one 256-counter sum across ranks
// each, instead of gathering the error arrays.
__global__ void key_histogram_kernel(const double* __restrict__ vals, long long n, unsigned long long prefix,
                                     int prefix_bits, unsigned long long* __restrict__ hist) {
  __shared__ unsigned int s_hist[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const int shift = 64 - prefix_bits - 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = vals[i];
    if (!(v == v)) continue;
    const unsigned long long key = (unsigned long long)__double_as_longlong(v);
    if (prefix_bits == 0 || (key >> (64 - prefix_bits)) == prefix) atomicAdd(&s_hist[(key >> shift) & 0xffull], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
}

// The device's default stream-ordered memory pool keeps freed blocks (release threshold raised
// once per device), so the work arrays of the one-shot entry points cost no cudaMalloc / cudaFree
// after their first call.
int keep_async_pool(int device) {
  static bool pool_ready[64] = {};
  if (device >= 0 && device < 64 && !pool_ready[device]) {
    cudaMemPool_t pool;
    MCBA_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long keep = ~0ull;
    MCBA_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    pool_ready[device] = true;
  }
  return MCBA_OK;
}

}  // namespace mcba

using namespace mcba;

extern "C" int mcba_select_frames(int device, void* cuda_stream, const double* d_uvs, int C, int64_t F, int N,
                                  const double* d_obj, const double* d_x, double outlier_threshold,
                                  uint8_t* d_use, double* h_stats) {
  if (!d_uvs || !d_obj || !d_x || !d_use || !h_stats || C < 1 || F < 1 || N < 1) {
    set_error("mcba_select_frames: bad arguments");
    return MCBA_ERR_ARG;
  }
  MCBA_CUDA(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const long long rows = (long long)C * F, total = rows * N;
  if (total > 2147483647LL) {   // cub's radix sort takes an int item count
    set_error("mcba_select_frames: more than 2^31 - 1 (camera, frame, corner) slots in one call; shard the frames");
    return MCBA_ERR_ARG;
  }
  // One stream-ordered allocation carved into the work arrays.  The device's default memory pool
  // keeps the block between calls (release threshold raised once), so a repeated call pays no
  // cudaMalloc / cudaFree (those cost ~15 ms here for ~0.3 ms of kernels and a 2 ms sort).
  { int rc = keep_async_pool(device); if (rc) return rc; }
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, (const double*)nullptr, (double*)nullptr, (int)total, 0, 64, st);
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t o_err = 0, o_sorted = o_err + up(sizeof(double) * total), o_mean = o_sorted + up(sizeof(double) * total),
               o_stats = o_mean + up(sizeof(double) * rows), o_cnt = o_stats + 256, o_cams = o_cnt + 256,
               o_complete = o_cams + up(sizeof(CamConst) * C), o_elig = o_complete + up(rows), o_tmp = o_elig + up(F),
               ws_bytes = o_tmp + up(tmp_bytes ? tmp_bytes : 8);
  AsyncBlock block(st);   // released on every return
  MCBA_CUDA(cudaMallocAsync((void**)&block.p, ws_bytes, st));
  unsigned char* ws = block.p;
  double* d_err = reinterpret_cast<double*>(ws + o_err);
  double* d_sorted = reinterpret_cast<double*>(ws + o_sorted);
  double* d_mean = reinterpret_cast<double*>(ws + o_mean);
  double* d_stats = reinterpret_cast<double*>(ws + o_stats);
  unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(ws + o_cnt);
  CamConst* d_cams = reinterpret_cast<CamConst*>(ws + o_cams);
  unsigned char* d_complete = ws + o_complete;
  unsigned char* d_elig = ws + o_elig;
  void* d_tmp = ws + o_tmp;
  MCBA_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * 2, st));
  MCBA_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(double) * 4, st));
  const int blocks = 148 * 8;
  prep_cameras_frontend_kernel<<<(C + 31) / 32, 32, 0, st>>>(d_x, C, d_cams);
  complete_rows_kernel<<<blocks, 256, 0, st>>>(d_uvs, rows, N, d_complete);
  eligible_kernel<<<(int)((F + 255) / 256), 256, 0, st>>>(d_complete, C, F, d_elig, d_cnt);
  frame_errors_kernel<<<blocks, 256, 0, st>>>(d_uvs, d_obj, d_x, d_cams, d_elig, C, F, N, d_err, d_mean, d_cnt + 1);
  if (!(outlier_threshold == outlier_threshold))
    MCBA_CUDA(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_err, d_sorted, (int)total, 0, 64, st));
  select_kernel<<<(int)((F + 255) / 256), 256, 0, st>>>(d_sorted, d_cnt + 1, d_mean, d_elig, C, F, outlier_threshold,
                                                       d_use, d_stats);
  MCBA_CUDA(cudaGetLastError());
  double stats[2];
  unsigned long long cnt[2];
  MCBA_CUDA(cudaMemcpyAsync(stats, d_stats, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
  MCBA_CUDA(cudaMemcpyAsync(cnt, d_cnt, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, st));
  MCBA_CUDA(cudaStreamSynchronize(st));
  unsigned long long excluded;
  memcpy(&excluded, &stats[1], sizeof(excluded));
  h_stats[0] = stats[0];              // threshold used
  h_stats[1] = (double)cnt[0];        // eligible frames
  h_stats[2] = (double)excluded;      // excluded as outliers
  h_stats[3] = (double)cnt[1];        // finite error values
  return MCBA_OK;
}

// ---- the same front end in stages, for frame-sharded (multi-GPU) runs: every rank calls these on
// its own contiguous range of frames and only counters / histograms cross ranks
extern "C" int mcba_frame_errors(int device, void* cuda_stream, const double* d_uvs, int C, int64_t F, int N,
                                 const double* d_obj, const double* d_x, double* d_err, double* d_mean,
                                 uint8_t* d_elig, int64_t* h_counts) {
  if (!d_uvs || !d_obj || !d_x || !d_err || !d_mean || !d_elig || !h_counts || C < 1 || F < 1 || N < 1) {
    set_error("mcba_frame_errors: bad arguments");
    return MCBA_ERR_ARG;
  }
  MCBA_CUDA(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  { int rc = keep_async_pool(device); if (rc) return rc; }
  const long long rows = (long long)C * F;
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t o_cnt = 0, o_cams = 256, o_complete = o_cams + up(sizeof(CamConst) * C), ws_bytes = o_complete + up(rows);
  AsyncBlock block(st);
  MCBA_CUDA(cudaMallocAsync((void**)&block.p, ws_bytes, st));
  unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(block.p + o_cnt);
  CamConst* d_cams = reinterpret_cast<CamConst*>(block.p + o_cams);
  unsigned char* d_complete = block.p + o_complete;
  MCBA_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * 2, st));
  const int blocks = 148 * 8;
  prep_cameras_frontend_kernel<<<(C + 31) / 32, 32, 0, st>>>(d_x, C, d_cams);
  complete_rows_kernel<<<blocks, 256, 0, st>>>(d_uvs, rows, N, d_complete);
  eligible_kernel<<<(int)((F + 255) / 256), 256, 0, st>>>(d_complete, C, F, d_elig, d_cnt);
  frame_errors_kernel<<<blocks, 256, 0, st>>>(d_uvs, d_obj, d_x, d_cams, d_elig, C, F, N, d_err, d_mean, d_cnt + 1);
  MCBA_CUDA(cudaGetLastError());
  unsigned long long cnt[2];
  MCBA_CUDA(cudaMemcpyAsync(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, st));
  MCBA_CUDA(cudaStreamSynchronize(st));
  h_counts[0] = (int64_t)cnt[0];   // eligible frames
  h_counts[1] = (int64_t)cnt[1];   // finite error values
  return MCBA_OK;
}

extern "C" int mcba_key_histogram(int device, void* cuda_stream, const double* d_vals, int64_t n, uint64_t prefix,
                                  int prefix_bits, uint64_t* d_hist) {
  if (!d_vals || !d_hist || n < 0 || prefix_bits < 0 || prefix_bits > 56 || prefix_bits % 8) {
    set_error("mcba_key_histogram: bad arguments");
    return MCBA_ERR_ARG;
  }
  MCBA_CUDA(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  MCBA_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(uint64_t) * 256, st));
  if (n == 0) return MCBA_OK;
  const long long want = (n + 1023) / 1024;
  const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
  key_histogram_kernel<<<grid, 256, 0, st>>>(d_vals, n, (unsigned long long)prefix, prefix_bits,
                                             reinterpret_cast<unsigned long long*>(d_hist));
  MCBA_CUDA(cudaGetLastError());
  return MCBA_OK;
}

extern "C" int mcba_apply_threshold(int device, void* cuda_stream, const double* d_mean, const uint8_t* d_elig, int C,
                                    int64_t F, double threshold, uint8_t* d_use, int64_t* h_excluded) {
  if (!d_mean || !d_elig || !d_use || !h_excluded || C < 1 || F < 1) {
    set_error("mcba_apply_threshold: bad arguments");
    return MCBA_ERR_ARG;
  }
  MCBA_CUDA(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  { int rc = keep_async_pool(device); if (rc) return rc; }
  AsyncBlock block(st);
  MCBA_CUDA(cudaMallocAsync((void**)&block.p, 256, st));
  double* d_stats = reinterpret_cast<double*>(block.p);                              // [threshold, excluded]
  unsigned long long* d_zero = reinterpret_cast<unsigned long long*>(block.p + 64);   // "no finite values": a NaN threshold stays NaN
  MCBA_CUDA(cudaMemsetAsync(block.p, 0, 256, st));
  select_kernel<<<(int)((F + 255) / 256), 256, 0, st>>>(nullptr, d_zero, d_mean, d_elig, C, F, threshold, d_use, d_stats);
  MCBA_CUDA(cudaGetLastError());
  double stats[2];
  MCBA_CUDA(cudaMemcpyAsync(stats, d_stats, sizeof(stats), cudaMemcpyDeviceToHost, st));
  MCBA_CUDA(cudaStreamSynchronize(st));
  unsigned long long excluded;
  memcpy(&excluded, &stats[1], sizeof(excluded));
  *h_excluded = (int64_t)excluded;
  return MCBA_OK;
}

extern "C" int mcba_gather_frames(int device, void* cuda_stream, const double* d_uvs, int C, int64_t F, int N,
                                  const int64_t* d_idx, int64_t Fu, double* d_out) {
  if (!d_uvs || !d_idx || !d_out || C < 1 || F < 1 || N < 1 || Fu < 0) {
    set_error("mcba_gather_frames: bad arguments");
    return MCBA_ERR_ARG;
  }
  MCBA_CUDA(cudaSetDevice(device));
  if (Fu == 0) return MCBA_OK;
  const long long total = (long long)C * Fu * N;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  gather_frames_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(reinterpret_cast<const double2*>(d_uvs), C, F, N,
                                                                  reinterpret_cast<const long long*>(d_idx), Fu,
                                                                  reinterpret_cast<double2*>(d_out));
  MCBA_CUDA(cudaGetLastError());
  return MCBA_OK;
}

// K6: initialisation algebra that produces bundle_adjust's initial guess (SURVEY.md 8(f) row
// N2) -- the rigid-transform part of calibration.py, batched over frames:
//
//   pairwise transform   T_rel,f = T2_f T1_f^-1 on frames both cameras detected, vector form,
//                        median over frames per component            (calibration.py:116-143)
//   consensus poses      T_board->world,cf = T_world->cam,c^-1 T_board->cam,cf,
//                        nanmedian over cameras per component        (calibration.py:245-277)
//   rodrigues_inv / get_transformation_vector                        (geometry.py:38-65, 178-197)
//
// One thread per frame; 4x4 rigid matrices never exist (R, t pairs, closed-form inverse
// R^T, -R^T t instead of numpy's LU inverse).  Medians are exact: radix sort of each
// component over the common frames, in-register insertion sort over the (<= 32) cameras.
#include <cub/device/device_radix_sort.cuh>

#include "mcba_internal.h"

namespace mcba {

// geometry.py:52-65: r = a * theta / |a|,  a = (R21-R12, R02-R20, R10-R01), theta = arccos((tr R - 1)/2),
// |a| == 0 -> divide by 1.  Same formula, same branch; no clamping of the arccos argument.
__device__ __forceinline__ void rodrigues_inv(const double R[9], double r[3]) {
  const double a0 = R[7] - R[5], a1 = R[2] - R[6], a2 = R[3] - R[1];
  const double theta = acos(((R[0] + R[4] + R[8]) - 1.0) / 2.0);
  double n = sqrt(a0 * a0 + a1 * a1 + a2 * a2);
  if (n == 0.0) n = 1.0;
  r[0] = a0 * theta / n;
  r[1] = a1 * theta / n;
  r[2] = a2 * theta / n;
}

__device__ __forceinline__ bool row_finite6(const double* p) {
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 6; ++i) ok &= p[i] == p[i];
  return ok;
}

// out = A^-1 B for rigid (R, t) pairs:  R = Ra^T Rb,  t = Ra^T (tb - ta)
__device__ __forceinline__ void rigid_inv_mul(const double Ra[9], const double ta[3], const double Rb[9],
                                              const double tb[3], double R[9], double t[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) R[3 * i + j] = Ra[i] * Rb[j] + Ra[3 + i] * Rb[3 + j] + Ra[6 + i] * Rb[6 + j];
    t[i] = Ra[i] * (tb[0] - ta[0]) + Ra[3 + i] * (tb[1] - ta[1]) + Ra[6 + i] * (tb[2] - ta[2]);
  }
}

// keys[j][f] = component j of vec(T2_f T1_f^-1) on common frames, canonical NaN elsewhere
// (sorts after every finite value); *n_common counts the common frames.
__global__ void relative_transforms_kernel(const double* __restrict__ p1, const double* __restrict__ p2, long long F,
                                           double* __restrict__ keys, unsigned long long* __restrict__ n_common) {
  const long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  bool common = false;
  double v[6];
  if (f < F) {
    const double* a = p1 + 6 * f;
    const double* b = p2 + 6 * f;
    common = row_finite6(a) && row_finite6(b);
    if (common) {
      double R1[9], R2[9], R[9];
      rodrigues(a, R1);
      rodrigues(b, R2);
      // T2 T1^-1: R = R2 R1^T, t = t2 - R t1
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) R[3 * i + j] = R2[3 * i] * R1[3 * j] + R2[3 * i + 1] * R1[3 * j + 1] + R2[3 * i + 2] * R1[3 * j + 2];
      rodrigues_inv(R, v);
#pragma unroll
      for (int i = 0; i < 3; ++i) v[3 + i] = b[3 + i] - (R[3 * i] * a[3] + R[3 * i + 1] * a[4] + R[3 * i + 2] * a[5]);
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) v[i] = nan("");
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) keys[(long long)i * F + f] = v[i];
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, common);
  if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(n_common, (unsigned long long)__popc(ballot));
}

// np.median of each sorted component over the n common frames (NaN when there are none)
__global__ void column_median_kernel(const double* __restrict__ sorted, long long F,
                                     const unsigned long long* __restrict__ n_common, double* __restrict__ out) {
  const int j = threadIdx.x;
  if (j >= 6) return;
  const unsigned long long n = *n_common;
  const double* s = sorted + (long long)j * F;
  double m = nan("");
  // a NaN among the common values (arccos argument rounded above 1) sorts last and makes
  // np.median NaN as well
  if (n > 0 && s[n - 1] == s[n - 1]) m = (n & 1ull) ? s[n / 2] : (s[n / 2 - 1] + s[n / 2]) / 2.0;
  out[j] = m;
}

struct ExtConst {
  double R[9], t[3];
};

__global__ void prep_extrinsics_kernel(const double* __restrict__ ext, int C, ExtConst* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  ExtConst e;
  rodrigues(ext + 6 * c, e.R);
  e.t[0] = ext[6 * c + 3]; e.t[1] = ext[6 * c + 4]; e.t[2] = ext[6 * c + 5];
  out[c] = e;
}

// Thread per frame: every detecting camera's board pose mapped to world coordinates, then the
// nanmedian over cameras of each of the six components (NaN when no camera detected the board).
template <int kMaxC>
__global__ void consensus_poses_kernel(const double* __restrict__ poses /* (C,F,6) */, const ExtConst* __restrict__ ext,
                                       int C, long long F, double* __restrict__ out /* (F,6) */) {
  const long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (f >= F) return;
  double vals[6][kMaxC];
  int n = 0;
  for (int c = 0; c < C; ++c) {
    const double* p = poses + ((long long)c * F + f) * 6;
    if (!row_finite6(p)) continue;
    double Rb[9], R[9], t[3], v[3];
    rodrigues(p, Rb);
    rigid_inv_mul(ext[c].R, ext[c].t, Rb, p + 3, R, t);
    rodrigues_inv(R, v);
    // insertion into the six sorted columns (NaN-producing arccos arguments stay NaN and are
    // kept out of the order statistics, like np.nanmedian)
#pragma unroll
    for (int j = 0; j < 6; ++j) vals[j][n] = j < 3 ? v[j] : t[j - 3];
    ++n;
  }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    // drop NaNs, sort ascending
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const double x = vals[j][i];
      if (x == x) {
        int k = m++;
        while (k > 0 && vals[j][k - 1] > x) { vals[j][k] = vals[j][k - 1]; --k; }
        vals[j][k] = x;
      }
    }
    double med = nan("");
    if (m > 0) med = (m & 1) ? vals[j][m / 2] : (vals[j][m / 2 - 1] + vals[j][m / 2]) / 2.0;
    out[f * 6 + j] = med;
  }
}

// rotation matrices (P,3,3) -> vectors (P,3);  4x4 transforms (P,4,4) -> (P,6)
__global__ void rodrigues_inv_kernel(const double* __restrict__ M, long long P, int stride, int ld, double* __restrict__ out,
                                     int out_ld) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= P) return;
  const double* m = M + i * stride;
  double R[9], r[3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) R[3 * a + b] = m[a * ld + b];
  rodrigues_inv(R, r);
  double* o = out + i * out_ld;
  o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
  if (out_ld == 6) { o[3] = m[3]; o[4] = m[ld + 3]; o[5] = m[2 * ld + 3]; }
}

}  // namespace mcba

using namespace mcba;

extern "C" int mcba_pairwise_transform(int device, void* cuda_stream, const double* d_poses1, const double* d_poses2,
                                       int64_t F, double* h_transform, int64_t* h_n_common) {
  if (!d_poses1 || !d_poses2 || !h_transform || F < 1) {
    set_error("mcba_pairwise_transform: bad arguments");
    return MCBA_ERR_ARG;
  }
  if (F > 0x7fffffffLL) {
    set_error("mcba_pairwise_transform: more than 2^31 - 1 frames");
    return MCBA_ERR_ARG;
  }
  MCBA_CUDA(cudaSetDevice(device));
  { int rc = keep_async_pool(device); if (rc) return rc; }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, (const double*)nullptr, (double*)nullptr, (int)F, 0, 64, st);
  auto up = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t o_keys = 0, o_sorted = up(sizeof(double) * 6 * F), o_small = 2 * o_sorted, o_tmp = o_small + 256;
  AsyncBlock block(st);   // released on every return
  MCBA_CUDA(cudaMallocAsync((void**)&block.p, o_tmp + up(tmp_bytes ? tmp_bytes : 8), st));
  unsigned char* ws = block.p;
  double* keys = reinterpret_cast<double*>(ws + o_keys);
  double* sorted = reinterpret_cast<double*>(ws + o_sorted);
  double* d_out = reinterpret_cast<double*>(ws + o_small);
  unsigned long long* d_n = reinterpret_cast<unsigned long long*>(ws + o_small + 64);
  MCBA_CUDA(cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), st));
  relative_transforms_kernel<<<(int)((F + 255) / 256), 256, 0, st>>>(d_poses1, d_poses2, F, keys, d_n);
  for (int j = 0; j < 6; ++j)
    MCBA_CUDA(cub::DeviceRadixSort::SortKeys(ws + o_tmp, tmp_bytes, keys + (size_t)j * F, sorted + (size_t)j * F, (int)F, 0, 64, st));
  column_median_kernel<<<1, 32, 0, st>>>(sorted, F, d_n, d_out);
  MCBA_CUDA(cudaGetLastError());
  unsigned long long n = 0;
  MCBA_CUDA(cudaMemcpyAsync(h_transform, d_out, sizeof(double) * 6, cudaMemcpyDeviceToHost, st));
  MCBA_CUDA(cudaMemcpyAsync(&n, d_n, sizeof(n), cudaMemcpyDeviceToHost, st));
  MCBA_CUDA(cudaStreamSynchronize(st));
  if (h_n_common) *h_n_common = (int64_t)n;
  return MCBA_OK;
}

extern "C" int mcba_consensus_poses(int device, void* cuda_stream, const double* d_all_poses, const double* h_extrinsics,
                                    int C, int64_t F, double* d_poses) {
  if (!d_all_poses || !h_extrinsics || !d_poses || C < 1 || F < 1) {
    set_error("mcba_consensus_poses: bad arguments");
    return MCBA_ERR_ARG;
  }
  if (C > 32) {
    set_error("mcba_consensus_poses: at most 32 cameras are supported");
    return MCBA_ERR_ARG;
  }
  MCBA_CUDA(cudaSetDevice(device));
  { int rc = keep_async_pool(device); if (rc) return rc; }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  AsyncBlock block(st);   // released on every return
  const size_t ext_bytes = (sizeof(double) * 6 * C + 255) / 256 * 256;
  MCBA_CUDA(cudaMallocAsync((void**)&block.p, ext_bytes + sizeof(ExtConst) * C, st));
  unsigned char* ws = block.p;
  double* d_ext = reinterpret_cast<double*>(ws);
  ExtConst* d_E = reinterpret_cast<ExtConst*>(ws + ext_bytes);
  MCBA_CUDA(cudaMemcpyAsync(d_ext, h_extrinsics, sizeof(double) * 6 * C, cudaMemcpyHostToDevice, st));
  prep_extrinsics_kernel<<<1, 32, 0, st>>>(d_ext, C, d_E);
  const int grid = (int)((F + 127) / 128);
  if (C <= 8) consensus_poses_kernel<8><<<grid, 128, 0, st>>>(d_all_poses, d_E, C, F, d_poses);
  else if (C <= 16) consensus_poses_kernel<16><<<grid, 128, 0, st>>>(d_all_poses, d_E, C, F, d_poses);
  else consensus_poses_kernel<32><<<grid, 128, 0, st>>>(d_all_poses, d_E, C, F, d_poses);
  MCBA_CUDA(cudaGetLastError());
  MCBA_CUDA(cudaStreamSynchronize(st));   // h_extrinsics may be a temporary
  return MCBA_OK;
}

extern "C" int mcba_transformation_vectors(int device, void* cuda_stream, const double* d_matrices, int64_t P, int dim,
                                           double* d_vectors) {
  if (!d_matrices || !d_vectors || P < 0 || (dim != 3 && dim != 4)) {
    set_error("mcba_transformation_vectors: dim must be 3 (rotation matrices) or 4 (rigid transforms)");
    return MCBA_ERR_ARG;
  }
  MCBA_CUDA(cudaSetDevice(device));
  if (P == 0) return MCBA_OK;
  rodrigues_inv_kernel<<<(int)((P + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(d_matrices, P, dim * dim, dim, d_vectors,
                                                                                   dim == 4 ? 6 : 3);
  MCBA_CUDA(cudaGetLastError());
  return MCBA_OK;
}

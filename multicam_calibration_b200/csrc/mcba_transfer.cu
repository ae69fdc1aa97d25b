// Host <-> device transfers of caller-owned PAGEABLE numpy buffers at PCIe rate.
//
// The Python boundary (bundle_adjustment.py:195-327) hands over ordinary numpy arrays: 168 MB of
// observations at BASELINE configs[2], 1.8 GB at configs[3]; `residuals` returns 2 O doubles.  A
// plain cudaMemcpy from / to pageable memory is staged by the driver through one bounce buffer on
// one thread (10-15 GB/s, and a freshly allocated destination is page-faulted 4 KB at a time on
// that thread).  Here kWorkers host threads each own two pinned bounce buffers and a copy
// stream: while one buffer is in flight on the copy engine the thread fills (or drains) the
// other, so the host memcpy (and the page faults of a fresh destination) run kWorkers wide and
// overlap the DMA.  A source / destination that is already pinned goes down in one async copy.
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "mcba_internal.h"

namespace mcba {
namespace {

constexpr int kMaxWorkers = 32;
constexpr size_t kMaxChunk = 8u << 20;   // bytes per bounce buffer (allocation size)

// Workers and chunk size in use: tuned on the B200 box (scripts/transfer_timing.py), overridable for
// such sweeps through MCBA_XFER_WORKERS / MCBA_XFER_CHUNK_KB (read once).
int env_int(const char* name, int fallback, int lo, int hi) {
  const char* v = getenv(name);
  if (!v) return fallback;
  const int x = atoi(v);
  return x < lo ? lo : (x > hi ? hi : x);
}
const int kWorkers = env_int("MCBA_XFER_WORKERS", 8, 1, kMaxWorkers);
const size_t kChunk = (size_t)env_int("MCBA_XFER_CHUNK_KB", 4096, 64, (int)(kMaxChunk >> 10)) << 10;

struct Lane {
  unsigned char* buf[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  cudaStream_t stream = nullptr;
};

struct Staging {
  int device = -1;
  Lane lanes[kMaxWorkers];
  cudaEvent_t ready = nullptr;   // caller's stream -> worker streams
};

// one staging context per device (streams and events belong to a device): a process that drives
// several GPUs, or a rank that touched cuda:0 before selecting its own device, stages on each
constexpr int kMaxDevices = 64;
std::mutex g_mu;
Staging g_stages[kMaxDevices];

int ensure_staging(int device, Staging** out) {
  if (device < 0 || device >= kMaxDevices) {
    set_error("mcba_upload/mcba_download: device index out of range");
    return MCBA_ERR_ARG;
  }
  Staging& g_stage = g_stages[device];
  *out = &g_stage;
  if (g_stage.device == device) return MCBA_OK;
  for (int w = 0; w < kWorkers; ++w) {
    Lane& l = g_stage.lanes[w];
    for (int b = 0; b < 2; ++b) {
      MCBA_CUDA(cudaHostAlloc((void**)&l.buf[b], kChunk, cudaHostAllocDefault));
      MCBA_CUDA(cudaEventCreateWithFlags(&l.done[b], cudaEventDisableTiming));
    }
    MCBA_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
  }
  MCBA_CUDA(cudaEventCreateWithFlags(&g_stage.ready, cudaEventDisableTiming));
  g_stage.device = device;
  return MCBA_OK;
}

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// to_device: host -> device, else device -> host.  Chunk k belongs to worker k % kWorkers.
int staged_copy(int device, cudaStream_t stream, unsigned char* dev, unsigned char* host, size_t bytes, bool to_device) {
  std::lock_guard<std::mutex> lock(g_mu);
  MCBA_CUDA(cudaSetDevice(device));
  if (bytes == 0) return MCBA_OK;
  if (bytes < 2 * kChunk || is_pinned(host)) {
    MCBA_CUDA(cudaMemcpyAsync(to_device ? (void*)dev : (void*)host, to_device ? (void*)host : (void*)dev, bytes,
                              to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, stream));
    MCBA_CUDA(cudaStreamSynchronize(stream));
    return MCBA_OK;
  }
  Staging* stage_ptr = nullptr;
  int rc = ensure_staging(device, &stage_ptr);
  if (rc) return rc;
  Staging& g_stage = *stage_ptr;
  // everything queued on the caller's stream (the producer of a download, the previous user of an
  // upload target) precedes the worker streams
  MCBA_CUDA(cudaEventRecord(g_stage.ready, stream));
  const size_t n_chunks = (bytes + kChunk - 1) / kChunk;
  std::atomic<int> failed{0};
  auto work = [&](int w) {
    if (cudaSetDevice(device) != cudaSuccess) { failed = 1; return; }
    Lane& l = g_stage.lanes[w];
    if (cudaStreamWaitEvent(l.stream, g_stage.ready, 0) != cudaSuccess) { failed = 1; return; }
    size_t pending_off[2] = {0, 0}, pending_len[2] = {0, 0};
    int it = 0;
    for (size_t k = w; k < n_chunks; k += kWorkers, ++it) {
      const int b = it & 1;
      const size_t off = k * kChunk, len = bytes - off < kChunk ? bytes - off : kChunk;
      if (it >= 2) {
        if (cudaEventSynchronize(l.done[b]) != cudaSuccess) { failed = 1; return; }
        if (!to_device) std::memcpy(host + pending_off[b], l.buf[b], pending_len[b]);
      }
      if (to_device) {
        std::memcpy(l.buf[b], host + off, len);
        if (cudaMemcpyAsync(dev + off, l.buf[b], len, cudaMemcpyHostToDevice, l.stream) != cudaSuccess) failed = 1;
      } else {
        if (cudaMemcpyAsync(l.buf[b], dev + off, len, cudaMemcpyDeviceToHost, l.stream) != cudaSuccess) failed = 1;
        pending_off[b] = off;
        pending_len[b] = len;
      }
      if (cudaEventRecord(l.done[b], l.stream) != cudaSuccess) failed = 1;
      if (failed) return;
    }
    // drain: the last (up to) two buffers of this lane, oldest first
    for (int j = (it >= 2 ? it - 2 : 0); j < it; ++j) {
      const int b = j & 1;
      if (cudaEventSynchronize(l.done[b]) != cudaSuccess) { failed = 1; return; }
      if (!to_device) std::memcpy(host + pending_off[b], l.buf[b], pending_len[b]);
    }
  };
  std::vector<std::thread> pool;
  const int n_threads = (int)(n_chunks < (size_t)kWorkers ? n_chunks : (size_t)kWorkers);
  for (int w = 1; w < n_threads; ++w) pool.emplace_back(work, w);
  work(0);
  for (std::thread& t : pool) t.join();
  if (failed) {
    cudaError_t e = cudaGetLastError();
    set_error(std::string("staged host<->device copy failed: ") + cudaGetErrorString(e));
    return MCBA_ERR_CUDA;
  }
  // every worker has synchronised its own last event: the data is in place; order the caller's
  // stream after the worker streams anyway so that later kernels on it see the upload
  for (int w = 0; w < n_threads; ++w) {
    MCBA_CUDA(cudaEventRecord(g_stage.lanes[w].done[0], g_stage.lanes[w].stream));
    MCBA_CUDA(cudaStreamWaitEvent(stream, g_stage.lanes[w].done[0], 0));
  }
  return MCBA_OK;
}

}  // namespace
}  // namespace mcba

extern "C" {

int mcba_upload(int device, void* cuda_stream, void* d_dst, const void* h_src, size_t bytes) {
  if ((!d_dst || !h_src) && bytes) {
    mcba::set_error("mcba_upload: null pointer");
    return MCBA_ERR_ARG;
  }
  return mcba::staged_copy(device, (cudaStream_t)cuda_stream, (unsigned char*)d_dst, (unsigned char*)const_cast<void*>(h_src),
                           bytes, true);
}

int mcba_download(int device, void* cuda_stream, void* h_dst, const void* d_src, size_t bytes) {
  if ((!h_dst || !d_src) && bytes) {
    mcba::set_error("mcba_download: null pointer");
    return MCBA_ERR_ARG;
  }
  return mcba::staged_copy(device, (cudaStream_t)cuda_stream, (unsigned char*)const_cast<void*>(d_src), (unsigned char*)h_dst,
                           bytes, false);
}

}  // extern "C"

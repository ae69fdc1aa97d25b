// One-shot all-reduce of the packed reduced camera system over NVLink peer memory.
//
// The only exchange step of the multi-GPU path is the sum of a small buffer per evaluation
// ([S | b | g | diag | scalars]: 43 KB at 6 cameras, 297 KB at 16) plus 12 step scalars per LM
// trial.  At that size a library all-reduce is pure latency (measured 50 us at 8 ranks against a
// 0.38 ms evaluation).  Here every rank owns an exchange buffer that its peers map through CUDA
// IPC (one process per GPU, NVSwitch: every peer at full bandwidth):
//
//   push   each rank stores its buffer into slot [parity][rank] of EVERY rank's exchange buffer
//          (coalesced 16-byte stores over NVLink; remote writes are posted, nobody waits on a
//          read round trip), then, after a system-scope fence, publishes the call's epoch in the
//          flag word [parity][rank] of every rank;
//   sum    each rank waits until its own flag words of all ranks carry the epoch and adds the
//          slots in rank order: every rank computes bit-identical sums, so the replicated
//          Cholesky solve needs no broadcast.
//
// Two parities alternate between calls: a rank can start call k+1 while a slower peer still sums
// call k, and nobody can reach call k+2 before every peer has pushed call k+1, i.e. finished
// reading call k.  One kernel per call, a handful of CTAs (all resident); the last CTA to finish
// its pushes publishes the flags (fence + device counter).  A rank that waits longer than ~20 s
// traps instead of hanging the GPU.
#include <cstring>

#include "mcba_internal.h"

namespace mcba {

struct PeerParams {
  double* slots[kMaxRanks];              // peers' exchange buffers (own included), slot layout [2][nranks][cap]
  unsigned long long* flags[kMaxRanks];  // peers' flag words [2][kMaxRanks]
  int rank, nranks;
  long long cap;                         // doubles per slot
  unsigned long long epoch;
  int parity;
  unsigned int* counter;                 // local: CTAs that have finished pushing
};

__global__ void __launch_bounds__(512) peer_allreduce_kernel(const PeerParams p, double* __restrict__ buf, long long n) {
  __shared__ bool s_last;
  const long long per_cta = ((n + gridDim.x - 1) / gridDim.x + 1) & ~1ll;   // even: 16-byte stores
  const long long lo = blockIdx.x * per_cta, hi = lo + per_cta < n ? lo + per_cta : n;
  // ---- push this CTA's element range into slot [parity][rank] of every rank
  for (int s = 0; s < p.nranks; ++s) {
    double* dst = p.slots[(p.rank + s) % p.nranks] + ((size_t)p.parity * p.nranks + p.rank) * p.cap;
    for (long long i = lo + 2 * threadIdx.x; i < hi; i += 2 * blockDim.x) {
      if (i + 1 < hi) *reinterpret_cast<double2*>(dst + i) = *reinterpret_cast<const double2*>(buf + i);
      else dst[i] = buf[i];
    }
  }
  // the barrier orders the CTA's stores before thread 0, whose system-scope fence is cumulative:
  // one fence per CTA instead of one per thread
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = atomicAdd(p.counter, 1u) == gridDim.x - 1;
    if (s_last) {
      *p.counter = 0;
      __threadfence_system();   // acquire side of the counter: the other CTAs' pushes precede the flags
    }
  }
  __syncthreads();
  if (s_last) {   // every CTA of this rank has pushed (and fenced): publish the epoch everywhere
    if (threadIdx.x < p.nranks) {
      unsigned long long* f = p.flags[threadIdx.x] + (size_t)p.parity * kMaxRanks + p.rank;
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(p.epoch) : "memory");
    }
  }
  // ---- wait for every rank's push of this call
  if (threadIdx.x < p.nranks) {
    const unsigned long long* f = p.flags[p.rank] + (size_t)p.parity * kMaxRanks + threadIdx.x;
    unsigned long long v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v < p.epoch && clock64() - t0 > 40000000000ll) __trap();   // ~20 s: a peer is gone
    } while (v < p.epoch);
  }
  __syncthreads();
  // ---- sum the slots in rank order
  const double* mine = p.slots[p.rank] + (size_t)p.parity * p.nranks * p.cap;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    double acc = 0.0;
    for (int s = 0; s < p.nranks; ++s) acc += __ldcg(mine + (size_t)s * p.cap + i);
    buf[i] = acc;
  }
}

int peer_allreduce(mcba_handle* h, double* buf, long long n) {
  if (n > h->peer_cap) {
    set_error("peer all-reduce: buffer larger than the exchange slots");
    return MCBA_ERR_ARG;
  }
  PeerParams p;
  std::memset(&p, 0, sizeof(p));
  for (int s = 0; s < h->nranks; ++s) {
    p.slots[s] = h->peer_slots[s];
    p.flags[s] = h->peer_flags[s];
  }
  p.rank = h->rank;
  p.nranks = h->nranks;
  p.cap = h->peer_cap;
  p.epoch = ++h->peer_epoch;
  p.parity = (int)(h->peer_epoch & 1ull);
  p.counter = h->peer_counter;
  int grid = (int)((n + 4095) / 4096);
  if (grid < 1) grid = 1;
  if (grid > 32) grid = 32;
  peer_allreduce_kernel<<<grid, 512, 0, h->stream>>>(p, buf, n);
  h->launches++;
  MCBA_CUDA(cudaGetLastError());
  return MCBA_OK;
}

}  // namespace mcba

using namespace mcba;

extern "C" {

int mcba_comm_ipc_export(mcba_handle* h, int rank, int nranks, void* handle64) {
  if (!h || !handle64 || rank < 0 || rank >= nranks || nranks > kMaxRanks) {
    set_error("mcba_comm_ipc_export: bad arguments");
    return MCBA_ERR_ARG;
  }
  MCBA_CUDA(cudaSetDevice(h->device));
  if (!h->peer_block) {
    const long long cap = (h->L.redLen + 15) & ~15ll;
    const size_t flag_bytes = sizeof(unsigned long long) * 2 * kMaxRanks;
    const size_t bytes = flag_bytes + sizeof(double) * 2 * (size_t)nranks * cap;
    MCBA_CUDA(cudaMalloc(&h->peer_block, bytes));
    MCBA_CUDA(cudaMemset(h->peer_block, 0, bytes));
    MCBA_CUDA(cudaMalloc((void**)&h->peer_counter, sizeof(unsigned int)));
    MCBA_CUDA(cudaMemset(h->peer_counter, 0, sizeof(unsigned int)));
    MCBA_CUDA(cudaDeviceSynchronize());
    h->peer_cap = cap;
  }
  h->rank = rank;
  h->nranks = nranks;
  cudaIpcMemHandle_t ipc;
  MCBA_CUDA(cudaIpcGetMemHandle(&ipc, h->peer_block));
  static_assert(sizeof(ipc) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(handle64, &ipc, 64);
  return MCBA_OK;
}

int mcba_comm_ipc_open(mcba_handle* h, const void* handles64) {
  if (!h || !handles64 || !h->peer_block) {
    set_error("mcba_comm_ipc_open: call mcba_comm_ipc_export first");
    return MCBA_ERR_STATE;
  }
  MCBA_CUDA(cudaSetDevice(h->device));
  const size_t flag_bytes = sizeof(unsigned long long) * 2 * kMaxRanks;
  for (int s = 0; s < h->nranks; ++s) {
    void* base = h->peer_block;
    if (s != h->rank) {
      cudaIpcMemHandle_t ipc;
      std::memcpy(&ipc, (const unsigned char*)handles64 + 64 * (size_t)s, 64);
      MCBA_CUDA(cudaIpcOpenMemHandle(&base, ipc, cudaIpcMemLazyEnablePeerAccess));
      h->peer_mapped[s] = base;
    }
    h->peer_flags[s] = reinterpret_cast<unsigned long long*>(base);
    h->peer_slots[s] = reinterpret_cast<double*>((unsigned char*)base + flag_bytes);
  }
  h->peer_ready = true;
  return MCBA_OK;
}

int mcba_comm_ipc_enable(mcba_handle* h, int enable) {
  if (!h) return MCBA_ERR_ARG;
  if (enable && !h->peer_slots[0]) {
    set_error("mcba_comm_ipc_enable: the peer buffers are not mapped");
    return MCBA_ERR_STATE;
  }
  h->peer_ready = enable != 0;
  return MCBA_OK;
}

}  // extern "C"

// One-shot all-reduce of the packed reduced camera system over NVLink peer memory.
//
// The only exchange step of the multi-GPU path is the sum of a small buffer per evaluation
// ([S | b | g | diag | scalars]: 43 KB at 6 cameras, 297 KB at 16) plus 12 step scalars per LM
// trial.  At that size a library all-reduce is pure latency (measured 50 us at 8 ranks against a
// 0.38 ms evaluation).  Here every rank owns an exchange buffer that its peers map through CUDA
// IPC (one process per GPU, NVSwitch: every peer at full bandwidth):
//
//   push   each rank stores its buffer into slot [parity][rank] of EVERY rank's exchange buffer
//          (stores over NVLink; remote writes are posted, nobody waits on a read round trip),
//          then, after a system-scope fence, publishes the call's epoch in the flag words
//          [parity][rank][cta] of every rank;
//   sum    each rank waits until its own flag words of all ranks carry the epoch and adds the
//          slots in rank order: every rank computes bit-identical sums, so the replicated
//          Cholesky solve needs no broadcast.
//
// Two parities alternate between calls: a rank can start call k+1 while a slower peer still sums
// call k, and nobody can reach call k+2 before every peer has pushed call k+1, i.e. finished
// reading call k.  Flags are per (rank, CTA): CTA k of a rank exchanges only with CTA k of its peers
// (mcba_peer.cuh), so there is no grid-wide counter and the exchange can live inside the kernel that
// produces the data (finalize_kernel, k2_schur.cu).  A rank that waits longer than ~20 s traps
// instead of hanging the GPU.
#include <cstring>

#include "mcba_peer.cuh"

namespace mcba {

// General form (any small buffer: the LM loop's step scalars, mcba_cost): CTA k exchanges the
// element range k of the buffer.  The reduced camera system itself is exchanged by the finalize
// kernel (k2_schur.cu) with the same protocol, straight from the registers that computed it.
__global__ void __launch_bounds__(512) peer_allreduce_kernel(const PeerView p, double* __restrict__ buf, long long n) {
  const long long per_cta = ((n + gridDim.x - 1) / gridDim.x + 1) & ~1ll;   // even: 16-byte stores
  const long long lo = blockIdx.x * per_cta, hi = lo + per_cta < n ? lo + per_cta : n;
  for (int s = 0; s < p.nranks; ++s) {
    double* dst = peer_slot_of(p, (p.rank + s) % p.nranks);
    for (long long i = lo + 2 * threadIdx.x; i < hi; i += 2 * blockDim.x) {
      if (i + 1 < hi) *reinterpret_cast<double2*>(dst + i) = *reinterpret_cast<const double2*>(buf + i);
      else dst[i] = buf[i];
    }
  }
  peer_publish(p, blockIdx.x);
  peer_wait(p, blockIdx.x);
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    double acc = 0.0;
    for (int r = 0; r < p.nranks; ++r) acc += __ldcg(peer_slot_from(p, r) + i);
    buf[i] = acc;
  }
}

PeerView peer_next_call(mcba_handle* h) {
  PeerView p;
  std::memset(&p, 0, sizeof(p));
  for (int s = 0; s < h->nranks; ++s) {
    p.slots[s] = h->peer_slots[s];
    p.flags[s] = h->peer_flags[s];
  }
  p.rank = h->rank;
  p.nranks = h->nranks;
  p.nflag = h->peer_nflag;
  p.cap = h->peer_cap;
  p.epoch = ++h->peer_epoch;
  p.parity = (int)(h->peer_epoch & 1ull);
  return p;
}

int peer_allreduce(mcba_handle* h, double* buf, long long n) {
  if (n > h->peer_cap) {
    set_error("peer all-reduce: buffer larger than the exchange slots");
    return MCBA_ERR_ARG;
  }
  const PeerView p = peer_next_call(h);
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

static size_t peer_flag_bytes(int nranks, int nflag) {
  return (sizeof(unsigned long long) * 2 * (size_t)nranks * nflag + 255) & ~(size_t)255;
}

extern "C" {

int mcba_comm_ipc_export(mcba_handle* h, int rank, int nranks, void* handle64) {
  if (!h || !handle64 || rank < 0 || rank >= nranks || nranks > kMaxRanks) {
    set_error("mcba_comm_ipc_export: bad arguments");
    return MCBA_ERR_ARG;
  }
  MCBA_CUDA(cudaSetDevice(h->device));
  if (!h->peer_block) {
    const long long cap = (h->L.redLen + 15) & ~15ll;
    const int pairs = h->L.C * (h->L.C + 1) / 2 + 1;          // CTAs of the finalize kernel
    h->peer_nflag = pairs > 32 ? pairs : 32;                   // the general kernel uses up to 32 CTAs
    const size_t flag_bytes = peer_flag_bytes(nranks, h->peer_nflag);
    const size_t bytes = flag_bytes + sizeof(double) * 2 * (size_t)nranks * cap;
    MCBA_CUDA(cudaMalloc(&h->peer_block, bytes));
    MCBA_CUDA(cudaMemset(h->peer_block, 0, bytes));
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
  const size_t flag_bytes = peer_flag_bytes(h->nranks, h->peer_nflag);   // same camera count on every rank
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

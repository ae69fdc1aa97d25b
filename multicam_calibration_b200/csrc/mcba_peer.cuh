// Device side of the exchange over NVLink peer memory (see mcba_peer.cu): every rank owns an
// exchange block  [flags 2 x nranks x nflag | slots 2 x nranks x cap]  that its peers map through
// CUDA IPC.  A CTA that takes part in an exchange
//   1. stores its share of the rank's buffer into slot [parity][rank] of EVERY rank (posted remote
//      stores, nobody waits on a read round trip),
//   2. peer_publish(): after a system-scope fence writes the call's epoch into flag
//      [parity][rank][cta] of every rank,
//   3. peer_wait(): waits until its own flags [parity][r][cta] of all ranks r carry the epoch,
//   4. adds the slots in rank order (every rank gets bit-identical sums).
// CTA k of one rank only ever talks to CTA k of the other ranks: no grid-wide counter, no second
// launch.  Two parities alternate between calls; calls are stream-ordered on every rank, so a rank
// can be at most one call ahead of a peer and never overwrites a slot that is still being read.
#pragma once
#include "mcba_internal.h"

namespace mcba {

struct PeerView {
  double* slots[kMaxRanks];              // peers' exchange buffers (own included), [2][nranks][cap]
  unsigned long long* flags[kMaxRanks];  // peers' flag words [2][nranks][nflag]
  int rank, nranks, nflag, parity;
  long long cap;                         // doubles per slot
  unsigned long long epoch;
};

// this rank's slot of the current call in rank s's exchange block
__device__ __forceinline__ double* peer_slot_of(const PeerView& p, int s) {
  return p.slots[s] + ((size_t)p.parity * p.nranks + p.rank) * p.cap;
}
// the slot rank r filled for the current call in OUR exchange block
__device__ __forceinline__ const double* peer_slot_from(const PeerView& p, int r) {
  return p.slots[p.rank] + ((size_t)p.parity * p.nranks + r) * p.cap;
}

// Called by ALL threads of the CTA after its pushes.
__device__ __forceinline__ void peer_publish(const PeerView& p, int cta) {
  __syncthreads();   // orders the CTA's stores before thread 0, whose system-scope fence is cumulative
  if (threadIdx.x == 0) __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < p.nranks) {
    unsigned long long* f = p.flags[threadIdx.x] + ((size_t)p.parity * p.nranks + p.rank) * p.nflag + cta;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(p.epoch) : "memory");
  }
}

// Called by ALL threads of the CTA; returns when every rank's share for this CTA has arrived.
__device__ __forceinline__ void peer_wait(const PeerView& p, int cta) {
  if ((int)threadIdx.x < p.nranks) {
    const unsigned long long* f = p.flags[p.rank] + ((size_t)p.parity * p.nranks + threadIdx.x) * p.nflag + cta;
    unsigned long long v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v < p.epoch && clock64() - t0 > 40000000000ll) __trap();   // ~20 s: a peer is gone
    } while (v < p.epoch);
  }
  __syncthreads();
}

// host: the view of the NEXT call (advances the epoch); peers must make the same sequence of calls
PeerView peer_next_call(mcba_handle* h);

}  // namespace mcba

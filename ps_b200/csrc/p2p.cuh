/*
 * p2p.cuh — the sharded exchange over NVLink peer memory.
 *
 * net/PSRouterClient.java:60-151 fans a batch of keys out to the PS shards over gRPC and merges
 * the answers; on one B200 box every GPU maps every peer's mailbox slab (CUDA IPC) and the
 * kernels that PRODUCE a bucket store it straight into the consumer's HBM through NVSwitch:
 * the pack step and the transfer are one kernel, there is no collective call and no host
 * involvement — a whole sharded step is one CUDA graph per rank.
 *
 * Mailbox slab of a rank (double-buffered by step parity p = seq & 1):
 *   keys_in [p][R][cap]      u64   rank r's DE-DUPLICATED packed keys         (getList request)
 *   rows_in [p][R][cap][Dp]  f32   rows returned by owner r                  (getList response)
 *   grads_in[p][R][cap][Dp]  f32   per-key gradient SUMS pushed by rank r    (push)
 *   gcnt_in [p][R][cap]      u32   occurrences of the key in rank r's batch  (travels with the push: the owner needs
 *                                  the global count only when it applies the update)
 *   wide_in [p][R][NF]       i64   wide ids of rank r                        (replicated wide table)
 *   gsum_in [p][R][glen]     f32   dense gradient sums + loss + gbar of r    (PServer sync-mode sum)
 *   counts  [p][R]           i32   number of valid keys from rank r
 *   flags   [p][CH][R]       u32   step sequence number, written last with release.sys
 * There are no flag kernels: every thread of a PRODUCER kernel fences its peer stores at system scope, the block that
 * finishes last (a ticket) release-stores the flag into every consumer's slab (p2p_publish_last); a CONSUMER kernel
 * spins on its own flags in its prologue (p2p_wait_all, acquire.sys).  Every send precedes the matching wait in every
 * rank's program order on the same logical stream, and a waiting block depends on no other block of its own grid, so
 * there is no circular wait; ranks can drift by at most one step, which the parity double-buffering covers.
 */
#pragma once
#include "common.cuh"

namespace psb {

constexpr int kP2PMaxRanks = 8;
enum { CH_KEYS = 0, CH_ROWS = 1, CH_GRADS = 2, CH_WIDE = 3, CH_GSUM = 4, CH_SCAL = 5, CH_COUNT = 6 };

struct P2PState {                      /* lives in device memory; kernels read it, p2p_begin advances seq */
  int R, me, cap, Dp, NF, glen;
  unsigned char* peer[kP2PMaxRanks];   /* base of every rank's slab as mapped into THIS process */
  size_t off_keys, off_rows, off_grads, off_gcnt, off_wide, off_gsum, off_counts, off_flags, parity_stride;
  uint32_t seq;
  int32_t cursor[kP2PMaxRanks];
  uint32_t ticket[CH_COUNT];           /* blocks of the running producer kernel of each channel that have finished */
  int32_t overflow;
};

struct __align__(16) BatchSlot {       /* the sender's per-batch de-duplication table: key → bucket position, occurrences */
  unsigned long long key;
  uint32_t cnt;
  int32_t upos;                       /* owner * cap + position in the owner's bucket, -1 on overflow */
};

struct P2P {
  Ctx* ctx = nullptr;
  int R = 0, me = 0, cap = 0, Dp = 0, NF = 0, glen = 0;
  size_t slab_bytes = 0;
  unsigned char* slab = nullptr;       /* this rank's mailbox */
  void* peer_mapped[kP2PMaxRanks] = {};
  P2PState host{};
  P2PState* dev = nullptr;
  bool connected = false;

  void create(Ctx* c, int R_, int me_, int cap_, int Dp_, int NF_, int glen_, int64_t max_lookups);
  void get_handle(void* out64);                                  /* cudaIpcMemHandle_t of the slab */
  void connect(const void* all_handles /* R x 64 bytes, rank order */);
  void destroy();

  /* --- kernels (asynchronous on ctx->stream) --- */
  void begin();                                                                     /* seq += 1, cursors = 0 */
  /* sender-side de-duplication (what PSRouterClient's key→shard map does): unique keys get a bucket position,
   * every lookup remembers its batch slot; then {key, occurrences} of each unique key goes to its owner    */
  BatchSlot* bt = nullptr; uint32_t BT = 0; int32_t* lk_b = nullptr; float* gacc = nullptr; int64_t Lmax = 0;
  int32_t* ulist = nullptr;                                                         /* bucket position -> batch slot (its count travels with the push; the push clears it) */
  void route_send(const int64_t* E, int N, int F);                                  /* de-duplicate, reserve, store each key into its owner's keys_in; publishes CH_KEYS */
  void wait(int channel);                                                           /* one-warp consumer-side wait (before a large-grid consumer on a side stream) */
  void bcast(const void* src, size_t bytes, int channel);                           /* wide ids → every peer; publishes the channel */
  void unpack(int N, int F, int D, float* out, int ldo, const float* X, int Xn, int xoff);   /* rows_in (+ X) → concat buffer */
  void reduce_gsum(float* gsum);                                                    /* sum over ranks, fixed order */
  /* (the per-key sums are formed by EmbTable::scatter_rows into gacc) */
  void grad_send();                                                                 /* sums + counts → owners' grads_in / gcnt_in; publishes CH_GRADS */
  /* device addresses inside the LOCAL slab for the current parity are resolved in-kernel from seq */
  const P2PState* state() const { return dev; }
  bool overflowed();
};

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned char* p2p_region(const P2PState* st, int rank, size_t off) {
  return st->peer[rank] + (size_t)(st->seq & 1u) * st->parity_stride + off;
}
__device__ __forceinline__ void p2p_st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t p2p_ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
/* Producer side, called by EVERY thread of the grid once its stores into peer memory are issued.  The block barrier orders
 * the block's stores before thread 0, whose ONE system-scope fence (cumulative) orders them before its ticket — a fence per
 * thread made these small kernels 2-3x longer; the block that arrives last flags every peer with the step's sequence number
 * (CH_KEYS: after the key counts).  Returns true in that last block.                                                    */
__device__ __forceinline__ bool p2p_publish_last(P2PState* st, int channel, uint32_t nblocks) {
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const uint32_t t = atomicAdd(&st->ticket[channel], 1u);
    s_last = t == nblocks - 1u;
    if (s_last) st->ticket[channel] = 0u;
  }
  __syncthreads();
  if (s_last && (int)threadIdx.x < st->R) {
    const int r = threadIdx.x;
    if (channel == CH_KEYS) {
      const int c = min(*reinterpret_cast<volatile int32_t*>(&st->cursor[r]), st->cap);
      reinterpret_cast<volatile int32_t*>(p2p_region(st, r, st->off_counts))[st->me] = c;
    }
    __threadfence_system();
    p2p_st_release_sys(reinterpret_cast<uint32_t*>(p2p_region(st, r, st->off_flags)) + channel * kP2PMaxRanks + st->me, st->seq);
  }
  return s_last;
}
/* Consumer side, called by EVERY thread of a block before it reads the channel's mailbox */
__device__ __forceinline__ void p2p_wait_all(const P2PState* st, int channel) {
  if ((int)threadIdx.x < st->R) {
    const uint32_t* f = reinterpret_cast<const uint32_t*>(p2p_region(st, st->me, st->off_flags)) + channel * kP2PMaxRanks + threadIdx.x;
    const uint32_t seq = st->seq;
    while ((int32_t)(p2p_ld_acquire_sys(f) - seq) < 0) __nanosleep(20);
  }
  __syncthreads();
}
#endif

}  // namespace psb

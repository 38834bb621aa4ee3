/*
 * p2p.cuh — the sharded exchange over NVLink peer memory.
 *
 * net/PSRouterClient.java:60-151 fans a batch of keys out to the PS shards over gRPC and merges
 * the answers; on one B200 box every GPU maps every peer's mailbox slab (CUDA IPC) and the
 * kernels that PRODUCE a bucket store it straight into the consumer's HBM through NVSwitch:
 * the pack step and the transfer are one kernel, there is no collective call and no host
 * involvement — a whole sharded step is one CUDA graph per rank.
 *
 * Slab of a rank (double-buffered by step parity p = seq & 1):
 *   keys_in [p][R][cap]      u64   rank r's DE-DUPLICATED packed keys         (getList request, stored by r)
 *   rows_in [p][R][cap][Dp]  f32   rows returned by owner r                  (getList response, stored by r)
 *   gsums   [p][R][cap][Dp]  f32   THIS rank's per-key gradient sums FOR owner o at [o]  (push: formed here by the backward's
 *                                  scatter, READ by owner o over NVLink inside its update kernel — there is no send kernel)
 *   gcnt    [p][R][cap]      u32   occurrences of those keys in this rank's batch (read by the owner with the sums)
 *   wide_in [p][R][NF]       i64   wide ids of rank r                        (replicated wide table)
 *   gsum_in [p][R][glen]     f32   dense gradient sums + loss + gbar of r    (PServer sync-mode sum)
 *   counts  [p][R]           i32   number of valid keys from rank r
 *   flags   [p][CH][R]       u32   step sequence number, written last with release.sys
 * There are no flag kernels: thread 0 of every block of a PRODUCER kernel fences the block's stores at system scope, the block
 * that finishes last (a ticket) release-stores the flag into every consumer's slab (p2p_publish_last); a CONSUMER kernel
 * spins on its own flags in its prologue (p2p_wait_all, acquire.sys).  Every send precedes the matching wait in every
 * rank's program order on the same logical stream, and a waiting block depends on no other block of its own grid, so
 * there is no circular wait.  Ranks can drift by at most one step, which the parity double-buffering covers: a rank that has
 * seen every peer's CH_KEYS flag of step t knows that every peer is past its update of step t-1, so what the peers read or
 * wrote for step t-1 (this rank's gsums of the other parity among it) is free — the tidy kernel zeroes it then.
 * A sharded step has no begin kernel either: route_send works with seq + 1 and its last block publishes the new seq.
 */
#pragma once
#include "common.cuh"

namespace psb {

constexpr int kP2PMaxRanks = 8;
enum { CH_KEYS = 0, CH_ROWS = 1, CH_GRADS = 2, CH_WIDE = 3, CH_GSUM = 4, CH_SCAL = 5, CH_COUNT = 6 };

struct P2PState {                      /* lives in device memory; kernels read it, route_send (or p2p_begin: no embedding table) advances seq */
  int R, me, cap, Dp, NF, glen;
  unsigned char* peer[kP2PMaxRanks];   /* base of every rank's slab as mapped into THIS process */
  size_t off_keys, off_rows, off_grads, off_gcnt, off_wide, off_gsum, off_counts, off_flags, parity_stride;
  uint32_t seq;
  int32_t cursor[2][kP2PMaxRanks];     /* by parity: unique keys of this rank's batch per owner (final when route_send ends; zeroed by the NEXT step's tidy) */
  uint32_t ticket[CH_COUNT];           /* blocks of the running producer kernel of each channel that have finished */
  uint32_t tidy_ticket;
  uint32_t pub_seq[CH_COUNT];          /* the sequence number this rank last published on each channel: what a DEFERRED consumer (the owner update and
                                          the dense update of step t may run at the head of step t+1, beside its route_send, which moves seq on) works with */
  int32_t overflow;
  int32_t block_fence_sys;             /* 1: every block of a producer fences at system scope before its ticket (PS_P2P_BLOCK_FENCE_SYS=1); default 0, see p2p_publish_last */
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
  void begin();                                                                     /* seq += 1 (models without an embedding table: no route_send) */
  /* sender-side de-duplication (what PSRouterClient's key→shard map does): unique keys get a bucket position,
   * every lookup remembers its batch slot; then {key, occurrences} of each unique key goes to its owner    */
  BatchSlot* bt = nullptr; uint32_t BT = 0; int32_t* lk_b = nullptr; int64_t Lmax = 0;
  int32_t* ulist = nullptr;                                                         /* bucket position -> batch slot */
  void route_send(const int64_t* E, int N, int F);                                  /* seq + 1; de-duplicate, reserve, store each key into its owner's keys_in; publishes seq and CH_KEYS */
  void counts();                                                                    /* side stream, after route_send: gcnt[q] = occurrences of the q-th unique key */
  void tidy();                                                                      /* side stream, after the backward's scatter: clears the de-duplication table; zeroes the OTHER parity's gsums */
  void wait(int channel);                                                           /* one-warp consumer-side wait for the step this rank last published on the channel */
  void bcast(const void* src, size_t bytes, int channel);                           /* wide ids → every peer; publishes the channel */
  void unpack(int N, int F, int D, float* out, int ldo, const float* X, int Xn, int xoff);   /* rows_in (+ X) → concat buffer */
  void reduce_gsum(float* gsum);                                                    /* sum over ranks, fixed order */
  /* (the per-key sums are formed by EmbTable::scatter_rows in this rank's gsums region; its last block publishes CH_GRADS) */
  /* device addresses inside the LOCAL slab for the current parity are resolved in-kernel from seq */
  const P2PState* state() const { return dev; }
  bool overflowed();
};

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned char* p2p_region(const P2PState* st, int rank, size_t off) {
  return st->peer[rank] + (size_t)(st->seq & 1u) * st->parity_stride + off;
}
__device__ __forceinline__ unsigned char* p2p_region_of(const P2PState* st, int rank, size_t off, uint32_t seq) {
  return st->peer[rank] + (size_t)(seq & 1u) * st->parity_stride + off;
}
__device__ __forceinline__ void p2p_st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
/* loads from ANOTHER GPU's memory (after the acquire of its flag): system-scope relaxed, never served from a stale local line */
__device__ __forceinline__ float4 p2p_ld_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t p2p_ld_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t p2p_ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
/* Producer side, called by EVERY thread of the grid once its stores (into peer memory, or into local memory the peers will
 * read) are issued.  The block barrier orders the block's stores before thread 0, whose ONE device-scope fence (cumulative)
 * orders them before its ticket; the block that takes the last ticket has thereby observed every other block's stores, and
 * ITS system-scope fence + release store (cumulative again, PTX memory model: causality order is transitive across scopes)
 * publishes all of them to the peers with the step's sequence number.  A system-scope fence in every block is not needed for
 * that, and costs: system fences are served one after the other chip-wide — with a few hundred blocks that was 10-15 us per
 * producer kernel (r02_notes.md).  Returns true in the last block.                                                        */
__device__ __forceinline__ void p2p_block_fence(const P2PState* st) {
  if (st->block_fence_sys) __threadfence_system(); else __threadfence();
}
__device__ __forceinline__ bool p2p_publish_last(P2PState* st, int channel, uint32_t nblocks) {
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    p2p_block_fence(st);
    const uint32_t t = atomicAdd(&st->ticket[channel], 1u);
    s_last = t == nblocks - 1u;
    if (s_last) st->ticket[channel] = 0u;
  }
  __syncthreads();
  if (s_last && (int)threadIdx.x < st->R) {
    const int r = threadIdx.x;
    if (r == 0) st->pub_seq[channel] = st->seq;
    __threadfence_system();
    p2p_st_release_sys(reinterpret_cast<uint32_t*>(p2p_region(st, r, st->off_flags)) + channel * kP2PMaxRanks + st->me, st->seq);
  }
  return s_last;
}
/* Consumer side, called by EVERY thread of a block before it reads the channel's mailbox */
__device__ __forceinline__ void p2p_wait_all_seq(const P2PState* st, int channel, uint32_t seq) {
  if ((int)threadIdx.x < st->R) {
    const uint32_t* f = reinterpret_cast<const uint32_t*>(p2p_region_of(st, st->me, st->off_flags, seq)) + channel * kP2PMaxRanks + threadIdx.x;
    while ((int32_t)(p2p_ld_acquire_sys(f) - seq) < 0) __nanosleep(20);
  }
  __syncthreads();
}
__device__ __forceinline__ void p2p_wait_all(const P2PState* st, int channel) { p2p_wait_all_seq(st, channel, st->seq); }
#endif

}  // namespace psb

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
 *   keys_in [p][R][cap]      2xu64 {packed key, occurrences at the sender}: rank r's DE-DUPLICATED keys (getList request)
 *   rows_in [p][R][cap][Dp]  f32   rows returned by owner r                  (getList response)
 *   grads_in[p][R][cap][Dp]  f32   per-key gradient SUMS pushed by rank r    (push)
 *   wide_in [p][R][NF]       i64   wide ids of rank r                        (replicated wide table)
 *   gsum_in [p][R][glen]     f32   dense gradient sums + loss + gbar of r    (PServer sync-mode sum)
 *   counts  [p][R]           i32   number of valid keys from rank r
 *   flags   [p][CH][R]       u32   step sequence number, written last with release.sys
 * A producer kernel stores its payload into the consumer's slab; a one-warp publish kernel right after
 * it (kernel boundary = all stores complete) fences at system scope and release-stores the flag; the
 * consumer's stream runs a one-warp wait kernel (acquire.sys spin) before the kernels that read the mailbox.  Every send precedes the matching wait in every
 * rank's program order, so there is no circular wait; ranks can drift by at most one step, which
 * the parity double-buffering covers.
 */
#pragma once
#include "common.cuh"

namespace psb {

constexpr int kP2PMaxRanks = 8;
enum { CH_KEYS = 0, CH_ROWS = 1, CH_GRADS = 2, CH_WIDE = 3, CH_GSUM = 4, CH_COUNT = 5 };

struct P2PState {                      /* lives in device memory; kernels read it, p2p_begin advances seq */
  int R, me, cap, Dp, NF, glen;
  unsigned char* peer[kP2PMaxRanks];   /* base of every rank's slab as mapped into THIS process */
  size_t off_keys, off_rows, off_grads, off_wide, off_gsum, off_counts, off_flags, parity_stride;
  uint32_t seq;
  int32_t cursor[kP2PMaxRanks];
  uint32_t done[CH_COUNT];
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
  void dedup_route(const int64_t* E, int N, int F);
  void send_keys();
  void bcast(const void* src, size_t bytes, int channel);                           /* wide ids / gsum → every peer */
  void publish(int channel);                                                        /* flag every peer (after a producer kernel) */
  void publish_wait(int channel);                                                   /* publish + wait for all peers, one launch */
  void wait(int channel);
  void gather_send(const float* w, int D, const int32_t* lk_slot);                  /* rows → requesters' rows_in */
  void unpack(int N, int F, int D, float* out, int ldo, const float* X, int Xn, int xoff);   /* rows_in (+ X) → concat buffer */
  void reduce_gsum(float* gsum);                                                    /* sum over ranks, fixed order */
  /* per-lookup row gradients (ReLU mask applied) summed per unique key locally, then one sum per key to its owner */
  void grad_reduce(const float* delta, int ldd, const float* act, int lda, int N, int F, int D);
  void grad_send();
  /* device addresses inside the LOCAL slab for the current parity are resolved in-kernel from seq */
  const P2PState* state() const { return dev; }
  bool overflowed();
};

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned char* p2p_region(const P2PState* st, int rank, size_t off) {
  return st->peer[rank] + (size_t)(st->seq & 1u) * st->parity_stride + off;
}
#endif

}  // namespace psb

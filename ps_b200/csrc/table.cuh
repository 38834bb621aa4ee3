/*
 * table.cuh — the GPU-resident parameter tables behind store.KVStore for sparse keys.
 *
 * EmbTable: one open-addressing hash table for all F embedding fields of a layer
 * (keys "emF<j>.<id>" of layer/EmbeddingField.java:70, packed by ps_pack_key).  HBM layout:
 *
 *   slots[C]   16 B records {key u64, cnt u32, uidx u32}: two per 32 B sector, so the probe
 *              that finds the key also brings the per-batch bookkeeping into L2 at no extra DRAM cost.
 *              cnt  = occurrences of the key in the current batch (0 between batches)
 *              uidx = bit 31: the row has been written (kRowReady, permanent);
 *                     bits 0-30: 1 + the key's index in this batch's list of unique keys (0 between batches)
 *   rows[C][3*Dp]  array of records {w[Dp] | s1[Dp] | s2[Dp]} — weight, Adam M | Ftrl Z, Adam V | Ftrl N of one key
 *              side by side: the update touches ONE contiguous record (one DRAM page, one TLB entry) per key,
 *              the forward gather reads the first third of it.  Dp = D rounded to 4 floats: 16 B alignment.
 *   per-batch workspace (L = N*F lookups):
 *              lk_slot[F][N]  slot of every lookup, field-major (a warp works on one field of 32 consecutive samples)
 *              lk_mask[F][N][MW]  the ReLU mask of the gathered row, one bit per element: the backward never
 *                             re-reads the activations (EmbeddingField.java:91-93 needs only A > 0)
 *              uniq[U]        slots of this batch's unique keys in first-arrival order (U <= L)
 *              acc[U][Dp]     gradient accumulators, one row per unique key — compact, so the few MB a batch
 *                             touches stay L2-resident and the scatter-add never round-trips HBM
 *
 * WideTable: layer/LRLayer.java's 1x1 weights "wide.weights.<id>": 32 B records
 * {key, w, s1, s2} — one sector holds everything a probe, the forward sum and the update need.
 */
#pragma once
#include "common.cuh"
#include "updaters.cuh"

namespace psb { struct P2PState; }

namespace psb {

constexpr uint32_t kRowReady = 0x80000000u;
constexpr int kProbeLimit = 1 << 16;        /* linear-probe bound shared by every find / insert (lookup, put, checkpoint load) */

struct __align__(16) EmbSlot {
  unsigned long long key;
  uint32_t cnt;    /* occurrences of this key in the current batch (EmbeddingField.java:96 wgN) */
  uint32_t uidx;   /* kRowReady | (1 + index in this batch's unique list); see above */
};

/* EmbTable::counters */
enum { CNT_UNIQUE = 0,      /* (unused) */
       CNT_ERR = 1,         /* sticky error bits: 1 = an insert found the table full, 2 = a key id outside [0, 2^44) */
       CNT_ROWS = 2,        /* [2..3] u64 number of keys in the table */
       CNT_CURSOR = 4,      /* unique keys of the batch being processed: grows during the lookup kernel, final when it ends, read by the
                               tail (StepStatus.n_unique) and the update kernel, whose last block resets it */
       CNT_TICKET_FWD = 5, CNT_TICKET_UPD = 6, CNT_WORDS = 8 };

struct __align__(32) WideSlot {
  unsigned long long key;
  float w, s1, s2;
  uint32_t pad0;
  unsigned long long pad1;
};

struct EmbTable {
  Ctx* ctx = nullptr;
  int F = 0, D = 0, Dp = 0, tpl = 1;   /* tpl: lanes cooperating on one row (power of two >= Dp/4) */
  int rs = 0;                          /* floats between consecutive row records (3*Dp) */
  int MW = 1;                          /* mask words per lookup */
  int64_t C = 0;
  float maxv = 0.f;                    /* Xavier bound of EmbeddingField.java:40 */
  UpdaterDev upd;
  EmbSlot* slots = nullptr;
  float *rows = nullptr;               /* [C][3*Dp] */
  float *w = nullptr, *s1 = nullptr, *s2 = nullptr;   /* rows, rows + Dp, rows + 2*Dp: element (slot, d) of each lives at p[slot*rs + d] */
  /* per-batch workspace */
  int64_t Lcap = 0;
  int generation = 0;                  /* bumped whenever reserve() reallocates the workspace */
  int32_t* lk_slot = nullptr;
  uint32_t* lk_mask = nullptr;
  int32_t* uniq = nullptr;
  float* acc = nullptr;
  int32_t* chain = nullptr;            /* [Lcap] owner side of the sharded exchange: next entry of the same key (see LookupArgs::chain) */
  uint32_t* counters = nullptr;
  int64_t last_L = 0;
  int scatter_occ[2] = {1, 1};         /* resident scatter blocks per SM (unaligned / aligned instantiation) */
  int update_slab_occ = 1;             /* ... of the staged update kernel */
  int update_pull_occ = 1;             /* ... and of its owner-side (pulling) form */
  int scatter_slab_occ = 1;            /* ... of the staged scatter kernel */
  int lookup_occ = 2;                  /* resident blocks per SM of the gathering lookup kernel (sizes its persistent grid) */

  void create(Ctx* c, int F_, int D_, int64_t capacity, const ps_updater_spec& u, int64_t max_lookups);
  void destroy();
  void reserve(int64_t L);
  /* EmbeddingLayer.forward in ONE kernel: find-or-insert every (field, id) of the batch, count occurrences, claim each
   * key's place in the batch's unique list, and (out != null) gather relu(row) into out[n*ldo + j*D + d]
   * (EmbeddingField.java:66-78) while recording the ReLU mask.  ids: device pointer [N][F]; exactly one of ids_i64 /
   * ids_f32 non-null.  X != null: also copies the numeric features X[N][Xn] to columns [xoff, xoff+Xn) (ConcatLayer). */
  void lookup(const int64_t* ids_i64, const float* ids_f32, int N, float* out, int ldo, const float* X = nullptr, int Xn = 0, int xoff = 0);
  /* the same on already-packed keys (owner side of the key-hash sharded exchange): n lookups, one "field".
   * out != null: rows (ReLU applied) to out[n][Dp].  p2p: keys come from this step's keys_in mailbox and, with send_rows,
   * every row goes straight into the requester's rows_in mailbox over NVLink (PServer.getList).                      */
  void lookup_packed(const uint64_t* keys, int n, float* out, P2PState* p2p = nullptr, bool send_rows = false);
  /* requester side of the exchange: EmbeddingLayer.forward once the owners have answered — lookup t was resolved by the route kernel to
   * record batch_slots[lk_batch[t]] (its row's place in this step's rows_in mailbox); same gather, ReLU mask bits and ConcatLayer copy
   * as lookup(), no probing and no counting; waits for the owners' flags in-kernel */
  void gather_resolved(const void* batch_slots, const int32_t* lk_batch, P2PState* p2p, int N, float* out, int ldo, const float* X, int Xn, int xoff);
  /* owner side of the push over peer memory (n = R*cap, after lookup_packed on the same entries): the update kernel itself
   * reads every requester's gradient sum and count for each of this shard's keys out of the requester's slab (rank order),
   * applies the updater, resets the per-batch state; waits for the requesters' flags in-kernel */
  void update_pull(const P2PState* p2p, int n, int calls, const int* skip_flag);
  /* pre-summed scatter-add, then occurrence normalisation + updater step + per-batch reset (two launches, see table.cu).
   * The ReLU mask comes from the lookup's mask bits (use_mask), from `act` (non-null), or is taken as already applied. */
  void scatter_update(const float* delta, int ldd, const float* act, int lda, int N, int calls, const int* skip_flag, int F_eff = 0, bool use_mask = false);
  /* requester side of the sharded push: this batch's row gradients summed per unique key into this rank's slab (see table.cu) */
  void scatter_rows(const void* batch_slots, const int32_t* lk_batch, P2PState* p2p, const float* delta, int ldd, const float* act, int lda, int N);
  /* forget the batch without updating (predict path / early exit): cnt = 0 for touched slots */
  void clear_batch();
  void check_errors();                 /* syncs; throws PS_ERR_CAPACITY if an insert found the table full */
  int64_t size();
  /* host-driven row access (KVStore.get / put, PSClient.getList / updateList) */
  void get_rows(const int32_t* fields, const int64_t* ids, int n, float* w_out, float* s1_out, float* s2_out, int32_t* found);
  void put_rows(const int32_t* fields, const int64_t* ids, int n, float* w_io, int replace);
  /* PServer.push: one updater step per listed (existing) key with a host-supplied gradient [n][D] */
  void push_rows(const int32_t* fields, const int64_t* ids, int n, const float* grads, const ps_updater_spec& spec, int32_t* found);
};

struct WideTable {
  Ctx* ctx = nullptr;
  int64_t C = 0;
  WideSlot* slots = nullptr;
  uint32_t* counters = nullptr;        /* [0] error flag, [2..3] u64 key count */
  UpdaterDev upd;

  void create(Ctx* c, int64_t capacity, const ps_updater_spec& u);
  void destroy();
  /* z[n] = bias + sum_j w[W[n][j]]  in j order (LRLayer.java:70-84); inserts unseen keys with w = 0 */
  void forward(const int64_t* ids, int N, int F, const float* bias, float* z);
  void insert(const int64_t* ids, int n, const P2PState* p2p = nullptr);   /* create keys other replicas saw (multi-GPU) */
  /* LRLayer.backward pushes the SAME batch-mean delta to every key ever seen (LRLayer.java:110-117,
   * SURVEY quirk 7): sweep all occupied slots and apply the updater with g = *gbar.             */
  void update_all(const float* gbar, const int* skip_flag, float* bias /* {w,s1,s2} or null */, const ps_updater_spec* bias_upd);
  int64_t size();
  int get(int64_t id, float* w, float* s1, float* s2);   /* 0 = absent */
  void put(int64_t id, float w);
  int push(int64_t id, float g, const ps_updater_spec& spec);   /* PServer.push on one wide key; 0 when the key does not exist */
  void check_errors();
};

}  // namespace psb

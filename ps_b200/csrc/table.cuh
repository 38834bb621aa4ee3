/*
 * table.cuh — the GPU-resident parameter tables behind store.KVStore for sparse keys.
 *
 * EmbTable: one open-addressing hash table for all F embedding fields of a layer
 * (keys "emF<j>.<id>" of layer/EmbeddingField.java:70, packed by ps_pack_key).  HBM layout:
 *
 *   slots[C]   16 B records {key u64, cnt u32, first u32}: two per 32 B sector, so the probe
 *              that finds the key also brings the per-batch occurrence count and the index of the
 *              key's first lookup of the batch into L2 at no extra DRAM cost
 *   w[C][Dp], s1[C][Dp], s2[C][Dp]   row = slot index; SoA across {weight, Adam M | Ftrl Z,
 *              Adam V | Ftrl N} so the forward gather touches only w.  Dp = D rounded to 4
 *              floats: every row is 16 B aligned for 128-bit loads
 *   per-batch workspace (L = N*F lookups): lk_slot[F][N] (field-major: warps of the probe store, and warps
 *              of the backward kernels load, 128 contiguous bytes); acc[L][Dp] gradient accumulators — a
 *              key uses the row of its FIRST lookup in the batch (emb_probe leaves ~t of that lookup in the
 *              slot record by a fire-and-forget red.max), the update zeroes it again: nothing to number or
 *              reset, and the few MB a batch touches stay L2-resident, so the scatter-add never round-trips HBM
 *
 * WideTable: layer/LRLayer.java's 1x1 weights "wide.weights.<id>": 32 B records
 * {key, w, s1, s2} — one sector holds everything a probe, the forward sum and the update need.
 */
#pragma once
#include "common.cuh"
#include "updaters.cuh"

namespace psb { struct P2PState; }

namespace psb {

struct __align__(16) EmbSlot {
  unsigned long long key;
  uint32_t cnt;    /* occurrences of this key in the current batch (EmbeddingField.java:96 wgN) */
  uint32_t first;  /* ~t of the key's first lookup of the current batch (max over ~t = min over t; 0 between batches): that lookup
                      owns the key's accumulator row acc[t] and performs its update */
};

struct __align__(32) WideSlot {
  unsigned long long key;
  float w, s1, s2;
  uint32_t pad0;
  unsigned long long pad1;
};

struct EmbTable {
  Ctx* ctx = nullptr;
  int F = 0, D = 0, Dp = 0, tpl = 1;   /* tpl: lanes cooperating on one lookup (power of two >= Dp/4) */
  int64_t C = 0;
  float maxv = 0.f;                    /* Xavier bound of EmbeddingField.java:40 */
  UpdaterDev upd;
  EmbSlot* slots = nullptr;
  float *w = nullptr, *s1 = nullptr, *s2 = nullptr;
  /* per-batch workspace */
  int64_t Lcap = 0;
  int generation = 0;                  /* bumped whenever reserve() reallocates the workspace */
  int32_t* lk_slot = nullptr;
  float* acc = nullptr;
  uint32_t* counters = nullptr;        /* [0] monotonic count of updated (unique) keys, [1] error flag, [2..3] u64 row count */
  int64_t last_L = 0;
  int scatter_occ[2] = {1, 1};         /* resident scatter blocks per SM (unaligned / aligned instantiation) */

  void create(Ctx* c, int F_, int D_, int64_t capacity, const ps_updater_spec& u, int64_t max_lookups);
  void destroy();
  void reserve(int64_t L);
  /* find-or-insert every (field, id) of the batch, count occurrences, mark each key's first lookup.
   * ids: device pointer, [N][F]; exactly one of ids_i64 / ids_f32 non-null.                  */
  void probe(const int64_t* ids_i64, const float* ids_f32, int N);
  /* out[n*ldo + j*D + d] = relu(w[slot(n,j)][d])   (EmbeddingField.java:73-76)               */
  /* the same on already-packed keys (owner side of the key-hash sharded exchange): n lookups, one "field" */
  void probe_packed(const uint64_t* keys, int n, const P2PState* p2p = nullptr);   /* p2p: keys come from this step's mailbox */
  /* X != null: also copies the numeric features X[N][Xn] to columns [xoff, xoff+Xn) (ConcatLayer) */
  void gather(float* out, int ldo, int N, int F_eff = 0, const float* X = nullptr, int Xn = 0, int xoff = 0);
  /* pre-summed scatter-add, then occurrence normalisation + updater step + per-batch reset (two launches, see table.cu) */
  void scatter_update(const float* delta, int ldd, const float* act /* null: mask already applied */, int lda, int N, int calls,
                      const int* skip_flag, int F_eff = 0, const P2PState* p2p = nullptr /* delta = this step's grads_in mailbox */);
  /* forget the batch without updating (predict path / early exit): cnt = 0 for touched slots */
  void clear_batch();
  void check_errors();                 /* syncs; throws PS_ERR_CAPACITY if an insert found the table full */
  int64_t size();
  /* host-driven row access (KVStore.get / put, PSClient.getList / updateList) */
  void get_rows(const int32_t* fields, const int64_t* ids, int n, float* w_out, float* s1_out, float* s2_out, int32_t* found);
  void put_rows(const int32_t* fields, const int64_t* ids, int n, float* w_io, int replace);
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
  void check_errors();
};

}  // namespace psb

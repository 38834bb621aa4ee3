/*
 * table.cu — kernels of the GPU-resident embedding / wide tables (sm_100a).
 *
 * Path restated (reference, /root/reference/src/main/java/):
 *   probe   = EmbeddingField.checkExists + KVStore.get(key, init)   layer/EmbeddingField.java:49-54, store/KVStore.java:136-159,168-190
 *   gather  = EmbeddingField.forward + EmbeddingLayer.forward       layer/EmbeddingField.java:66-78, layer/EmbeddingLayer.java:25-48
 *   scatter = EmbeddingField.backward (x2) + KVStore.sum + KVStore.update + Updater.update
 *                                                                   layer/EmbeddingField.java:86-104, store/KVStore.java:192-200,240-268
 *   wide    = LRLayer.forward / backward                            layer/LRLayer.java:62-120
 *
 * All of it is HBM/L2-bound integer and copy work: no tensor cores, 128-bit accesses,
 * one thread group (Dp/4 lanes) per looked-up row.
 */
#include <algorithm>
#include <cmath>
#include <vector>

#include "p2p.cuh"
#include "table.cuh"

namespace psb {

static constexpr int kProbeLimit = 1 << 16;

/* ------------------------------------------------------------------ find / find-or-insert */
__device__ __forceinline__ int emb_find(const EmbSlot* slots, uint32_t C, unsigned long long key) {
  uint32_t slot = ps_bucket_of(key, C);
  const int limit = C < (uint32_t)kProbeLimit ? (int)C : kProbeLimit;
  for (int p = 0; p < limit; ++p) {
    const unsigned long long k = *reinterpret_cast<const volatile unsigned long long*>(&slots[slot].key);
    if (k == key) return (int)slot;
    if (k == PS_KEY_EMPTY) return -1;
    slot = slot + 1 == C ? 0 : slot + 1;
  }
  return -1;
}

/* returns the slot, or -1 when the table is full; *inserted tells the caller to initialise the row */
__device__ __forceinline__ int emb_find_or_insert(EmbSlot* slots, uint32_t C, unsigned long long key, bool* inserted) {
  uint32_t slot = ps_bucket_of(key, C);
  const int limit = C < (uint32_t)kProbeLimit ? (int)C : kProbeLimit;
  *inserted = false;
  for (int p = 0; p < limit; ++p) {
    const unsigned long long k = *reinterpret_cast<const volatile unsigned long long*>(&slots[slot].key);
    if (k == key) return (int)slot;
    if (k == PS_KEY_EMPTY) {
      const unsigned long long old = atomicCAS(&slots[slot].key, (unsigned long long)PS_KEY_EMPTY, key);
      if (old == PS_KEY_EMPTY) { *inserted = true; return (int)slot; }
      if (old == key) return (int)slot;
    }
    slot = slot + 1 == C ? 0 : slot + 1;
  }
  return -1;
}

/* find-or-insert that starts from an already loaded first probe (`k0` = the key found in the home bucket) */
__device__ __forceinline__ int emb_find_or_insert_from(EmbSlot* slots, uint32_t C, unsigned long long key, uint32_t slot, unsigned long long k,
                                                       bool* inserted) {
  const int limit = C < (uint32_t)kProbeLimit ? (int)C : kProbeLimit;
  *inserted = false;
  for (int p = 0; p < limit; ++p) {
    if (p > 0) k = *reinterpret_cast<const volatile unsigned long long*>(&slots[slot].key);
    if (k == key) return (int)slot;
    if (k == PS_KEY_EMPTY) {
      const unsigned long long old = atomicCAS(&slots[slot].key, (unsigned long long)PS_KEY_EMPTY, key);
      if (old == PS_KEY_EMPTY) { *inserted = true; return (int)slot; }
      if (old == key) return (int)slot;
    }
    slot = slot + 1 == C ? 0 : slot + 1;
  }
  return -1;
}

/* Key resolution of one batch.  Row creation follows KVStore.create (KVStore.java:168-190): the creating thread draws the
 * row from the deterministic initialiser of ps_spec.h; optimiser state stays at the zero the arena was allocated with
 * (AdamUpdater.initMandV, :76-84).
 *
 * Work decomposition: a block owns 32*SG consecutive samples and ALL F fields of them; a warp task is (field j, 32
 * consecutive samples), so
 *   - the 32 lanes of a warp probe the SAME field: duplicates of a hot key (a low-cardinality field) meet in one warp
 *     and are counted with one L2 reduction per warp, not 32;
 *   - the 8 warps of a block read neighbouring fields of the same samples at the same time: every 32 B sector of the
 *     [N][F] id matrix is fetched from HBM once;
 *   - lk_slot is FIELD-major ([F][N], index t = j*N + n): a warp stores 128 contiguous bytes, and the backward kernels,
 *     which walk the same order, read it back coalesced.
 * Up to 4 tasks of a warp are in flight at once (ids, then the home buckets, are loaded for all of them before any is used).
 * Per-batch bookkeeping in the slot record needs no returned atomic: `cnt += occurrences` and `first = max(first, ~t)`
 * are fire-and-forget reductions; the lookup with the smallest t owns the key's accumulator row (acc[t]).
 * F == 0: `ids` already holds packed keys (the owner side of the sharded exchange); p2p: they come from this step's mailbox. */
template <class IdT>
__global__ void __launch_bounds__(256) emb_probe_kernel(EmbSlot* __restrict__ slots, uint32_t C, float* __restrict__ w, int Dp, int D,
                                                        const IdT* __restrict__ ids, int N, int F, int SG, uint64_t seed, float maxv,
                                                        int32_t* __restrict__ lk_slot,
                                                        uint32_t* __restrict__ counters, const P2PState* __restrict__ p2p) {
  constexpr int R = 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_launch_dependents();                     /* the gather's blocks may be scheduled; they wait for this grid before reading */
  const int Fe = F > 0 ? F : 1;
  const int tasks = Fe * SG;
  const long n0 = (long)blockIdx.x * (32 * SG);
  for (int task0 = warp; task0 < tasks; task0 += 8 * R) {
    unsigned long long key[R], k0[R];
    uint32_t bucket[R], add[R];
    long t[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int task = task0 + 8 * r;
      key[r] = PS_KEY_EMPTY; add[r] = 1u; t[r] = -1;
      if (task < tasks) {
        const int j = task % Fe, sg = task / Fe;
        const long n = n0 + sg * 32 + lane;
        if (n < N) {
          t[r] = (long)j * N + n;
          if (p2p != nullptr) {                  /* owner side of the peer-memory exchange: this step's keys_in mailbox */
            const int src = (int)(n / p2p->cap), idx = (int)(n - (long)src * p2p->cap);
            if (idx < reinterpret_cast<const int32_t*>(p2p_region(p2p, p2p->me, p2p->off_counts))[src]) {
              /* entries are {key, occurrences at the sender}: senders de-duplicate their batch (PSRouterClient sends a key once) */
              const ulonglong2 e = reinterpret_cast<const ulonglong2*>(p2p_region(p2p, p2p->me, p2p->off_keys))[n];
              key[r] = e.x;
              add[r] = (uint32_t)e.y + (1u << 24);   /* low 24 bits: occurrences; high 8 bits: entries */
            }
          } else if (F > 0) {
            key[r] = ps_pack_key((uint32_t)j, (uint64_t)(int64_t)ids[n * F + j]);
          } else {
            key[r] = (unsigned long long)ids[n];
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      bucket[r] = 0u; k0[r] = PS_KEY_EMPTY;
      if (key[r] != PS_KEY_EMPTY) {              /* EMPTY marks padding in the fixed-capacity sharded exchange */
        bucket[r] = ps_bucket_of(key[r], C);
        k0[r] = *reinterpret_cast<const volatile unsigned long long*>(&slots[bucket[r]].key);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (task0 + 8 * r >= tasks) break;         /* warp-uniform */
      int slot = -1;
      if (key[r] != PS_KEY_EMPTY) {
        bool inserted;
        slot = emb_find_or_insert_from(slots, C, key[r], bucket[r], k0[r], &inserted);
        if (slot < 0) counters[1] = 1u;
        else if (inserted) {
          float* row = w + (size_t)slot * Dp;
          for (int d = 0; d < D; ++d) row[d] = ps_init_value(seed, key[r], (uint32_t)d, maxv);
          atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2), 1ull);
        }
      }
      if (t[r] >= 0) lk_slot[t[r]] = slot;
      /* warp-aggregated bookkeeping: lanes holding the same slot reduce once (the lowest lane has the smallest t) */
      const unsigned peers = __match_any_sync(0xffffffffu, slot >= 0 ? slot : (-1 - lane));
      /* every lookup counts 1 unless it is a pre-counted entry of the peer-memory exchange (a reduction over a partial
       * mask runs once per distinct key in the warp: not worth it for a popcount) */
      const uint32_t total_add = p2p == nullptr ? (uint32_t)__popc(peers) : __reduce_add_sync(peers, add[r]);
      if (slot >= 0 && (__ffs(peers) - 1) == lane) {
        red_add_u32(&slots[slot].cnt, total_add);
        red_max_u32(&slots[slot].first, ~(uint32_t)t[r]);
      }
    }
  }
}

/* TPL lanes per lookup, each moving one 16 B chunk of the row: consecutive lanes write
 * consecutive addresses of the (F*D) x N output (fields of one sample are adjacent), so the
 * stores of a warp coalesce into full 128 B lines; ReLU (EmbeddingField.java:75) is fused.     */
template <int TPL, bool ALIGNED>
__global__ void __launch_bounds__(256) emb_gather_kernel(const float* __restrict__ w, int Dp, int D, const int32_t* __restrict__ lk_slot,
                                                         int L, int F, float* __restrict__ out, int ldo, const float* __restrict__ X, int Xn,
                                                         int xoff, int N) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  pdl_launch_dependents();                     /* fc0's forward GEMM may set itself up (it waits for this grid before reading) */
  pdl_wait();                                  /* launched as a programmatic dependent of the probe */
  if (g >= (long)L * TPL) {                    /* ConcatLayer.forward (ConcatLayer.java:30-37): numeric features next to the embeddings */
    const long i = g - (long)L * TPL;
    if (X != nullptr && i < (long)N * Xn) { const int n = (int)(i / Xn), x = (int)(i - (long)n * Xn); out[(size_t)n * ldo + xoff + x] = X[i]; }
    return;
  }
  const int l = (int)(g / TPL), part = (int)(g % TPL);
  if (part * 4 >= D) return;
  const int n = l / F, j = l - n * F;
  const int slot = lk_slot[(size_t)j * N + n];   /* field-major (see emb_probe_kernel) */
  if (slot < 0) return;
  float4 v = __ldg(reinterpret_cast<const float4*>(w + (size_t)slot * Dp + part * 4));
  v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  float* o = out + (size_t)n * ldo + j * D + part * 4;
  if (ALIGNED) {
    st_f4(o, v);
  } else {
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) if (part * 4 + i < D) o[i] = e[i];
  }
}

__device__ __forceinline__ float4 shfl_f4(float4 v, int src) {
  float4 r;
  r.x = __shfl_sync(0xffffffffu, v.x, src); r.y = __shfl_sync(0xffffffffu, v.y, src);
  r.z = __shfl_sync(0xffffffffu, v.z, src); r.w = __shfl_sync(0xffffffffu, v.w, src);
  return r;
}

/* Effective gradient the reference ends up applying for a key with n occurrences and S = sum of
 * its per-occurrence gradients.  calls == 2 (what DNN/WideDeepNN do): pass 1
 * stores S/n by reference in KVStore.sum, pass 2 adds the n gradients again, divides by 2n, the
 * aliased sum doubles it and KVStore.update halves it: ((S/n) + S) / (2n).  calls == 1: S/n.
 * EXACT: the two IEEE divisions of the Java code.  Otherwise: multiplications by 1/n and 1/(2n), rounded once per key
 * (<= 1.5 ulp off the exact quotient each).                                                      */
struct GeffScale { float n, n2, rn, rn2; int calls; };
template <bool EXACT>
__device__ __forceinline__ GeffScale make_geff(uint32_t n, int calls) {
  GeffScale g;
  g.n = (float)n; g.n2 = (float)(2u * n); g.calls = calls;
  g.rn = EXACT ? 0.f : __frcp_rn(g.n);
  g.rn2 = 0.5f * g.rn;
  return g;
}
template <bool EXACT>
__device__ __forceinline__ float emb_geff(float S, const GeffScale& g) {
  if (S == 0.0f) return S;
  const float q = EXACT ? __fdiv_rn(S, g.n) : __fmul_rn(S, g.rn);
  if (g.calls == 1) return q;
  return EXACT ? __fdiv_rn(__fadd_rn(q, S), g.n2) : __fmul_rn(__fadd_rn(q, S), g.rn2);
}

/* Sparse backward = two launches on one stream, the second a programmatic dependent of the first:
 *   emb_scatter_kernel  g_k = delta[:,k] * (A[:,k] > 0)  (EmbeddingField.java:91-93).  A block owns SB consecutive samples
 *            and all F fields of them; a warp task is (field j, 32/TPL consecutive samples), so duplicates of a key meet
 *            in the same warp and block, and the 8 warps read neighbouring columns of the same delta / act rows at the
 *            same time (whole DRAM pages).  Three levels of pre-summation keep a hot key (a low-cardinality field:
 *            thousands of occurrences of one row) from serialising in L2:  (1) reduce-by-key tree over the lanes of a
 *            warp that __match_any_sync groups;  (2) keys frequent enough to recur inside one block's samples (the
 *            probe left the batch count in the slot record; threshold = max(PS_HOT_MIN, 2N/SB)) are summed in a
 *            per-block shared-memory table and leave the block ONCE;  (3) everything else goes out as one
 *            red.global.add.v4.f32 per key and 16 B chunk into the key's accumulator row (L2-resident, indexed by the
 *            work index of the key's first lookup).
 *   emb_update_kernel   a warp scans 32 work indices, ballots the ones that own an accumulator row (the key's first
 *            lookup) and deals their 16 B chunks over ALL its lanes: read S back from L2, form g_eff, run the Adam /
 *            Ftrl / SGD step on w, s1, s2 in place and reset the per-batch state (acc, cnt, first) — KVStore.sum +
 *            update + clear (KVStore.java:192-200,240-277).  Launched with programmatic stream serialisation: its
 *            scan and the w/s1/s2 loads of its first round run while the scatter kernel drains; only the accumulator
 *            read sits behind griddepcontrol.wait.                                                                  */
static constexpr int kHotMin = 8;        /* default occurrences in the batch from which a key is pre-summed per block (PS_HOT_MIN) */
static constexpr int kHotBits = 6;
static constexpr int kHotEntries = 1 << kHotBits;   /* per-block hot-key table (open addressing, 4 probes) */

template <int TPL, int CPL, int PASSES, bool ALIGNED>
__global__ void __launch_bounds__(256) emb_scatter_kernel(const EmbSlot* __restrict__ slots, int Dp, int D, const int32_t* __restrict__ lk_slot, int N,
                                                          int F, int SB, const float* __restrict__ delta, int ldd, const float* __restrict__ act, int lda,
                                                          float* __restrict__ acc, const int* __restrict__ skip_flag,
                                                          const P2PState* __restrict__ p2p, uint32_t hot_min) {
  constexpr int GPW = 32 / TPL;                  /* lookups (lane groups) per warp task */
  constexpr int ROWF = TPL * CPL * 4;            /* floats of a (padded) row */
  __shared__ float hot_acc[kHotEntries][ROWF];
  __shared__ int hot_slot[kHotEntries];
  __shared__ uint32_t hot_row[kHotEntries];
  pdl_launch_dependents();                       /* the update kernel may start its scan now (it waits before reading acc) */
  if (skip_flag != nullptr && *skip_flag != 0) return;   /* DNN.java:58-63 early exit: nothing is pushed */
  if (p2p != nullptr) delta = reinterpret_cast<const float*>(p2p_region(p2p, p2p->me, p2p->off_grads));   /* this step's grads_in mailbox */
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int part = lane % TPL, grp = lane / TPL;
  const int c0 = part * CPL * 4;                 /* first float of this lane's chunks */
  /* lanes holding the same `part` of their lookups: bits at multiples of TPL, shifted by part */
  unsigned part_lanes = 0u;
#pragma unroll
  for (int g = 0; g < GPW; ++g) part_lanes |= 1u << (g * TPL);
  part_lanes <<= part;

  for (int i = threadIdx.x; i < kHotEntries; i += 256) hot_slot[i] = -1;
  for (int i = threadIdx.x; i < kHotEntries * ROWF; i += 256) (&hot_acc[0][0])[i] = 0.f;
  __syncthreads();

  /* the block's tile: SB consecutive samples x all F fields; a warp task = (field j, GPW consecutive samples); the 8 warps
   * work on neighbouring fields of the same samples at the same time, so the rows of delta / act are read as whole
   * DRAM pages although every lookup only needs D of their columns */
  const int tasks = F * (SB / GPW);
  /* persistent blocks (one wave of them, see launch_scatter) stride over the tiles: no tail wave, and the hot table keeps
   * summing across all tiles of the block */
  for (long n0 = (long)blockIdx.x * SB; n0 < N; n0 += (long)gridDim.x * SB)
  for (int task0 = warp; task0 < tasks; task0 += 8 * PASSES) {
    /* ---- every load of PASSES tasks is issued before anything is consumed ---- */
    int slot[PASSES];
    float4 gk[PASSES][CPL];
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      const int task = task0 + 8 * p;
      slot[p] = -1;
#pragma unroll
      for (int c = 0; c < CPL; ++c) gk[p][c] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int j = task % F;
      const long n = n0 + (long)(task / F) * GPW + grp;
      if (task < tasks && n < N) {
        slot[p] = lk_slot[(long)j * N + n];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const int cc = c0 + 4 * c;
          if (cc < D) {
            const size_t od = (size_t)n * ldd + j * D + cc, oa = (size_t)n * lda + j * D + cc;
            float dv[4] = {0.f, 0.f, 0.f, 0.f}, av[4] = {0.f, 0.f, 0.f, 0.f};
            if (ALIGNED) {
              const float4 d4 = ld_f4(delta + od);
              const float4 a4 = act ? ld_f4(act + oa) : make_float4(1.f, 1.f, 1.f, 1.f);   /* act == null: the mask was applied by the sender */
              dv[0] = d4.x; dv[1] = d4.y; dv[2] = d4.z; dv[3] = d4.w;
              av[0] = a4.x; av[1] = a4.y; av[2] = a4.z; av[3] = a4.w;
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) if (cc + i < D) { dv[i] = delta[od + i]; av[i] = act ? act[oa + i] : 1.f; }
            }
            /* Relu.backward: dy *= (y > 0 ? 1 : 0)  (activations/Relu.java:14-19) */
            gk[p][c].x = __fmul_rn(dv[0], av[0] > 0.f ? 1.f : 0.f); gk[p][c].y = __fmul_rn(dv[1], av[1] > 0.f ? 1.f : 0.f);
            gk[p][c].z = __fmul_rn(dv[2], av[2] > 0.f ? 1.f : 0.f); gk[p][c].w = __fmul_rn(dv[3], av[3] > 0.f ? 1.f : 0.f);
          }
        }
      }
    }
    uint32_t cnt[PASSES], row[PASSES];
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      cnt[p] = 0u; row[p] = 0u;
      if (slot[p] >= 0) { const uint4 m = *reinterpret_cast<const uint4*>(&slots[slot[p]]); cnt[p] = m.z; row[p] = ~m.w; }   /* acc row = the key's first lookup */
    }
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      if (task0 + 8 * p >= tasks) break;           /* warp-uniform */
      const bool valid = slot[p] >= 0;
      /* ---- (1) reduce-by-key inside the warp: rank r of a key's lanes adds rank r+s, s = 1, 2, 4, ... ---- */
      const unsigned pmask = __match_any_sync(0xffffffffu, valid ? slot[p] : (-1 - lane)) & part_lanes;
      const int npeer = __popc(pmask);
      const int rank = __popc(pmask & ((1u << lane) - 1u));
      const int maxn = __reduce_max_sync(0xffffffffu, npeer);
      for (int s = 1; s < maxn; s <<= 1) {
        const bool has = rank + s < npeer;
        const int partner = has ? (int)__fns(pmask, (unsigned)lane, s + 1) : lane;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const float4 o = shfl_f4(gk[p][c], partner);
          if (has) { gk[p][c].x += o.x; gk[p][c].y += o.y; gk[p][c].z += o.z; gk[p][c].w += o.w; }
        }
      }
      if (!valid || rank != 0) continue;
      /* ---- (2) hot keys: sum inside the block (the peer-memory exchange packs {entries << 24 | occurrences}: never hot) ---- */
      int e = -1;
      if (p2p == nullptr && cnt[p] >= hot_min) {
        uint32_t h = ((uint32_t)slot[p] * 2654435761u) >> (32 - kHotBits);
#pragma unroll 1
        for (int t = 0; t < 4; ++t) {
          const int old = atomicCAS(&hot_slot[h], -1, slot[p]);
          if (old == -1) hot_row[h] = row[p];
          if (old == -1 || old == slot[p]) { e = (int)h; break; }
          h = (h + 1u) & (uint32_t)(kHotEntries - 1);
        }
      }
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        const int cc = c0 + 4 * c;
        if (cc >= D) continue;
        if (e >= 0) {
          float* a = &hot_acc[e][cc];
          if (gk[p][c].x != 0.f) atomicAdd(a + 0, gk[p][c].x);
          if (gk[p][c].y != 0.f) atomicAdd(a + 1, gk[p][c].y);
          if (gk[p][c].z != 0.f) atomicAdd(a + 2, gk[p][c].z);
          if (gk[p][c].w != 0.f) atomicAdd(a + 3, gk[p][c].w);
        } else {
          red_add_f4(acc + (size_t)row[p] * Dp + cc, gk[p][c]);          /* ---- (3) ---- */
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kHotEntries * (ROWF / 4); i += 256) {
    const int e = i / (ROWF / 4), cc = (i % (ROWF / 4)) * 4;
    if (hot_slot[e] < 0 || cc >= D) continue;
    const float4 v = *reinterpret_cast<const float4*>(&hot_acc[e][cc]);
    if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) red_add_f4(acc + (size_t)hot_row[e] * Dp + cc, v);
  }
}

/* TPK lanes per key (4 floats each).  A WARP scans 32 consecutive work indices, ballots the ones that own an accumulator
 * row and deals the owners' (key, 16 B chunk) items round-robin over its 32 lanes, IPL items per lane and round with every
 * load of a round issued before any is consumed; no block-level synchronisation.  EXACT: see updaters.cuh.              */
template <int TPK, bool EXACT>
__global__ void __launch_bounds__(256, 3) emb_update_kernel(EmbSlot* __restrict__ slots, float* __restrict__ w, float* __restrict__ s1, float* __restrict__ s2,
                                                         int Dp, int D, const int32_t* __restrict__ lk_slot, long L, float* __restrict__ acc,
                                                         UpdaterDev upd, int calls, const int* __restrict__ skip_flag, int packed_cnt,
                                                         uint32_t* __restrict__ counters) {
  constexpr int IPL = TPK < 2 ? 1 : 2;             /* items per lane and round */
  const int lane = threadIdx.x & 31;
  const long lk = (long)blockIdx.x * 256 + threadIdx.x;
  int slot = -1; uint32_t cnt = 0u, first = 0u;
  if (lk < L) {
    slot = lk_slot[lk];
    if (slot >= 0) { const uint4 m = *reinterpret_cast<const uint4*>(&slots[slot]); cnt = m.z; first = m.w; }
  }
  const bool owner = slot >= 0 && first == ~(uint32_t)lk;          /* the key's first lookup of the batch */
  const unsigned owners = __ballot_sync(0xffffffffu, owner);
  const int items = __popc(owners) * TPK;
  if (lane == 0 && owners != 0u) red_add_u32(&counters[0], (uint32_t)__popc(owners));   /* statistics (StepStatus.n_unique), monotonic */
  const bool skip = skip_flag != nullptr && *skip_flag != 0;

  int kslot[IPL]; uint32_t kcnt[IPL]; long krow[IPL]; bool act[IPL];
  float4 wv[IPL], m1[IPL], m2[IPL];
  auto load_round = [&](int base) {
#pragma unroll
    for (int q = 0; q < IPL; ++q) {
      const int i = base + q * 32 + lane;
      const bool in = i < items;
      const int src = in ? (int)__fns(owners, 0u, i / TPK + 1) : 0;   /* lane of the (i / TPK)-th owner */
      kslot[q] = __shfl_sync(0xffffffffu, slot, src);
      kcnt[q] = __shfl_sync(0xffffffffu, cnt, src);
      krow[q] = (long)(lk - lane + src);                              /* the owner's work index = its accumulator row */
      const int cc = (i % TPK) * 4;
      act[q] = in && cc < D;
      wv[q] = m1[q] = m2[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (act[q] && !skip) {
        const size_t o = (size_t)kslot[q] * Dp + cc;
        wv[q] = ld_f4(w + o);
        if (upd.kind != PS_UPD_SIMPLE) { m1[q] = ld_f4(s1 + o); m2[q] = ld_f4(s2 + o); }
      }
    }
  };
  /* ---- the rows of the first round are requested before the scatter kernel is known to be complete ---- */
  load_round(0);
  pdl_wait();
  for (int base = 0; base < items; base += 32 * IPL) {
    if (base > 0) load_round(base);
    float4 S[IPL];
#pragma unroll
    for (int q = 0; q < IPL; ++q) {
      const int cc = ((base + q * 32 + lane) % TPK) * 4;
      S[q] = (act[q] && !skip) ? __ldcg(reinterpret_cast<const float4*>(acc + (size_t)krow[q] * Dp + cc)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int q = 0; q < IPL; ++q) {
      const int cc = ((base + q * 32 + lane) % TPK) * 4;
      /* peer-memory exchange: senders pre-reduce, cnt packs {entries << 24 | occurrences} */
      const uint32_t n_occ = packed_cnt ? (kcnt[q] & 0xFFFFFFu) : kcnt[q];
      const GeffScale gs = make_geff<EXACT>(n_occ > 0u ? n_occ : 1u, calls);
      bool do_upd = !skip;
      if (upd.kind == PS_UPD_FTRL) {                  /* FtrlUpdater.java:52: `if (dw.get(0) == 0) return w` */
        const float S0 = __shfl_sync(0xffffffffu, S[q].x, lane - (lane % TPK));
        do_upd = do_upd && emb_geff<EXACT>(S0, gs) != 0.0f;
      }
      if (!act[q]) continue;
      if (!skip) {
        const size_t o = (size_t)kslot[q] * Dp + cc;
        if (do_upd) {
          apply_elem<EXACT>(upd, wv[q].x, m1[q].x, m2[q].x, emb_geff<EXACT>(S[q].x, gs));
          apply_elem<EXACT>(upd, wv[q].y, m1[q].y, m2[q].y, emb_geff<EXACT>(S[q].y, gs));
          apply_elem<EXACT>(upd, wv[q].z, m1[q].z, m2[q].z, emb_geff<EXACT>(S[q].z, gs));
          apply_elem<EXACT>(upd, wv[q].w, m1[q].w, m2[q].w, emb_geff<EXACT>(S[q].w, gs));
          st_f4(w + o, wv[q]);
          if (upd.kind != PS_UPD_SIMPLE) { st_f4(s1 + o, m1[q]); st_f4(s2 + o, m2[q]); }
        }
        st_f4(acc + (size_t)krow[q] * Dp + cc, make_float4(0.f, 0.f, 0.f, 0.f));
      }
      /* KVStore.clear (also after the early exit: the batch is forgotten): {cnt, first} = 0 in one 8 B store */
      if (cc == 0) *reinterpret_cast<unsigned long long*>(&slots[kslot[q]].cnt) = 0ull;
    }
  }
}

__global__ void emb_clear_batch_kernel(EmbSlot* __restrict__ slots, const int32_t* __restrict__ lk_slot, int L) {
  for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < L; l += gridDim.x * blockDim.x) {
    const int slot = lk_slot[l];
    if (slot >= 0) *reinterpret_cast<unsigned long long*>(&slots[slot].cnt) = 0ull;       /* {cnt, first}; duplicates store the same zero */
  }
}

/* host-driven row access: thread per key */
__global__ void emb_get_rows_kernel(const EmbSlot* __restrict__ slots, uint32_t C, const float* __restrict__ w, const float* __restrict__ s1,
                                    const float* __restrict__ s2, int Dp, int D, const int32_t* __restrict__ fields,
                                    const int64_t* __restrict__ ids, int n, float* __restrict__ wo, float* __restrict__ s1o,
                                    float* __restrict__ s2o, int32_t* __restrict__ found) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int slot = emb_find(slots, C, ps_pack_key((uint32_t)fields[i], (uint64_t)ids[i]));
  found[i] = slot >= 0;
  for (int d = 0; d < D; ++d) {
    const size_t o = (size_t)(slot < 0 ? 0 : slot) * Dp + d;
    wo[(size_t)i * D + d] = slot < 0 ? 0.f : w[o];
    if (s1o) s1o[(size_t)i * D + d] = slot < 0 ? 0.f : s1[o];
    if (s2o) s2o[(size_t)i * D + d] = slot < 0 ? 0.f : s2[o];
  }
}

/* KVStore.put (replace) / PServer.upsertList with replace=false (net/PServer.java:144-162):
 * insert-if-absent; the caller's buffer receives the winning row.                              */
__global__ void emb_put_rows_kernel(EmbSlot* __restrict__ slots, uint32_t C, float* __restrict__ w, int Dp, int D,
                                    const int32_t* __restrict__ fields, const int64_t* __restrict__ ids, int n, float* __restrict__ wio,
                                    int replace, uint32_t* __restrict__ counters) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool inserted;
  const int slot = emb_find_or_insert(slots, C, ps_pack_key((uint32_t)fields[i], (uint64_t)ids[i]), &inserted);
  if (slot < 0) { counters[1] = 1u; return; }
  if (inserted) atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2), 1ull);
  float* row = w + (size_t)slot * Dp;
  if (inserted || replace) { for (int d = 0; d < D; ++d) row[d] = wio[(size_t)i * D + d]; }
  else { for (int d = 0; d < D; ++d) wio[(size_t)i * D + d] = row[d]; }
}

/* ------------------------------------------------------------------ EmbTable host side */
static int pow2_ge(int x) { int p = 1; while (p < x) p <<= 1; return p; }

void EmbTable::create(Ctx* c, int F_, int D_, int64_t capacity, const ps_updater_spec& u, int64_t max_lookups) {
  PS_REQUIRE(F_ > 0 && D_ > 0 && D_ <= 128, PS_ERR_ARG, "embedding: need F > 0 and 0 < D <= 128");
  PS_REQUIRE(capacity > 0 && capacity < (1ll << 31), PS_ERR_ARG, "embedding: capacity must be in (0, 2^31)");
  ctx = c; F = F_; D = D_; Dp = round_up(D_, 4); tpl = pow2_ge(Dp / 4); C = capacity;
  maxv = (float)(4 * (std::sqrt(6.0) / std::sqrt((double)(1 + D_))));   /* EmbeddingField.java:40 with in=1,out=D (EmbeddingLayer.java:52) */
  upd = make_updater_dev(u);
  slots = dmalloc_zero<EmbSlot>((size_t)C, ctx->stream);
  w = dmalloc_zero<float>((size_t)C * Dp, ctx->stream);
  s1 = dmalloc_zero<float>((size_t)C * Dp, ctx->stream);
  s2 = dmalloc_zero<float>((size_t)C * Dp, ctx->stream);
  counters = dmalloc_zero<uint32_t>(4, ctx->stream);
  reserve(max_lookups > 0 ? max_lookups : 1);
  scatter_update(nullptr, 0, nullptr, 0, 0, 2, nullptr);   /* fills scatter_occ (sizes the scatter's persistent grid) */
}

void EmbTable::reserve(int64_t L) {
  if (L <= Lcap) return;
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  dfree(lk_slot); dfree(acc);
  Lcap = L; ++generation;                      /* captured graphs that hold the old workspace pointers are stale now */
  lk_slot = dmalloc<int32_t>((size_t)L);
  acc = dmalloc_zero<float>((size_t)L * Dp, ctx->stream);   /* one accumulator row per lookup index; a batch uses those of its keys' first lookups */
}

void EmbTable::destroy() {
  dfree(slots); dfree(w); dfree(s1); dfree(s2); dfree(counters);
  dfree(lk_slot); dfree(acc);
  slots = nullptr; w = s1 = s2 = nullptr;
}

/* samples per block = 32 * SG with SG chosen so that a block's F * SG warp tasks keep its 8 warps busy */
static int probe_sg(int F) { return F >= 8 ? 1 : (8 + F - 1) / F; }

void EmbTable::probe(const int64_t* ids_i64, const float* ids_f32, int N) {
  const int64_t L = (int64_t)N * F;
  PS_REQUIRE(L <= Lcap, PS_ERR_ARG, "embedding: batch larger than the reserved workspace");
  last_L = L;
  const int SG = probe_sg(F);
  const int grid = ceil_div(N, 32 * SG);
  if (ids_i64)
    emb_probe_kernel<int64_t><<<grid, 256, 0, ctx->stream>>>(slots, (uint32_t)C, w, Dp, D, ids_i64, N, F, SG, ctx->seed, maxv, lk_slot, counters, nullptr);
  else
    emb_probe_kernel<float><<<grid, 256, 0, ctx->stream>>>(slots, (uint32_t)C, w, Dp, D, ids_f32, N, F, SG, ctx->seed, maxv, lk_slot, counters, nullptr);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

void EmbTable::probe_packed(const uint64_t* keys, int n, const P2PState* p2p) {
  reserve(n);
  last_L = n;
  if (n <= 0) return;
  const int SG = probe_sg(1);
  emb_probe_kernel<unsigned long long><<<ceil_div(n, 32 * SG), 256, 0, ctx->stream>>>(slots, (uint32_t)C, w, Dp, D, reinterpret_cast<const unsigned long long*>(keys), n, 0, SG,
                                                                                      ctx->seed, maxv, lk_slot, counters, p2p);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

template <int TPL>
static void launch_gather(EmbTable& t, float* out, int ldo, int N, int F, const float* X, int Xn, int xoff) {
  const long L = (long)N * F;
  const bool aligned = (t.D % 4 == 0) && (ldo % 4 == 0) && ((uintptr_t)out % 16 == 0);
  const int grid = ceil_div(L * TPL + (X ? (long)N * Xn : 0), 256);
  if (aligned) launch_pdl(t.ctx, emb_gather_kernel<TPL, true>, dim3(grid), dim3(256), t.w, t.Dp, t.D, t.lk_slot, (int)L, F, out, ldo, X, Xn, xoff, N);
  else launch_pdl(t.ctx, emb_gather_kernel<TPL, false>, dim3(grid), dim3(256), t.w, t.Dp, t.D, t.lk_slot, (int)L, F, out, ldo, X, Xn, xoff, N);
}

void EmbTable::gather(float* out, int ldo, int N, int F_eff, const float* X, int Xn, int xoff) {
  const int Fe = F_eff > 0 ? F_eff : F;
  PS_REQUIRE((int64_t)N * Fe == last_L, PS_ERR_STATE, "embedding: gather without a matching probe");
  switch (tpl) {
    case 1: launch_gather<1>(*this, out, ldo, N, Fe, X, Xn, xoff); break;
    case 2: launch_gather<2>(*this, out, ldo, N, Fe, X, Xn, xoff); break;
    case 4: launch_gather<4>(*this, out, ldo, N, Fe, X, Xn, xoff); break;
    case 8: launch_gather<8>(*this, out, ldo, N, Fe, X, Xn, xoff); break;
    case 16: launch_gather<16>(*this, out, ldo, N, Fe, X, Xn, xoff); break;
    default: launch_gather<32>(*this, out, ldo, N, Fe, X, Xn, xoff); break;
  }
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

template <int TPK>
static void launch_update(EmbTable& t, int N, int F, int calls, const int* skip, int packed_cnt) {
  const long L = (long)N * F;
  if (t.ctx->exact_updaters)
    launch_pdl(t.ctx, emb_update_kernel<TPK, true>, dim3(ceil_div(L, 256)), dim3(256), t.slots, t.w, t.s1, t.s2, t.Dp, t.D, (const int32_t*)t.lk_slot, L,
               t.acc, t.upd, calls, skip, packed_cnt, t.counters);
  else
    launch_pdl(t.ctx, emb_update_kernel<TPK, false>, dim3(ceil_div(L, 256)), dim3(256), t.slots, t.w, t.s1, t.s2, t.Dp, t.D, (const int32_t*)t.lk_slot, L,
               t.acc, t.upd, calls, skip, packed_cnt, t.counters);
}

template <int TPL, int CPL>
static void launch_scatter(EmbTable& t, const float* delta, int ldd, const float* act, int lda, int N, int F, int calls, const int* skip,
                           const P2PState* p2p) {
  constexpr int PASSES = TPL >= 4 ? 4 : TPL;     /* warp tasks in flight per warp */
  constexpr int GPW = 32 / TPL;
  const long L = (long)N * F;
  const bool aligned = (t.D % 4 == 0) && (ldd % 4 == 0) && (lda % 4 == 0) && ((uintptr_t)delta % 16 == 0) && ((uintptr_t)act % 16 == 0);
  /* samples per tile: small enough that there are >= 4 tiles per resident block (balance), at most 32 */
  if (N == 0) {                                  /* EmbTable::create: resident blocks per SM of the two instantiations (not inside a capture) */
    PS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t.scatter_occ[1], emb_scatter_kernel<TPL, CPL, PASSES, true>, 256, 0));
    PS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t.scatter_occ[0], emb_scatter_kernel<TPL, CPL, PASSES, false>, 256, 0));
    return;
  }
  const int resident = t.ctx->num_sms * std::max(1, t.scatter_occ[aligned ? 1 : 0]);
  int SB = GPW;
  while (SB * 2 <= 32 && ceil_div(N, SB * 2) >= 4 * resident) SB *= 2;
  const int grid = std::min(ceil_div(N, SB), resident);
  /* a key is pre-summed per block when it is frequent enough to recur among the samples one block sees */
  const long per_block = (long)SB * ceil_div(ceil_div(N, SB), grid);
  const uint32_t hot_min = t.ctx->hot_min == 0xFFFFFFFFu ? 0xFFFFFFFFu : std::max<uint32_t>(t.ctx->hot_min, (uint32_t)(2L * N / per_block));
  (void)L;
  if (aligned)
    emb_scatter_kernel<TPL, CPL, PASSES, true><<<grid, 256, 0, t.ctx->stream>>>(t.slots, t.Dp, t.D, t.lk_slot, N, F, SB, delta, ldd, act, lda, t.acc, skip, p2p, hot_min);
  else
    emb_scatter_kernel<TPL, CPL, PASSES, false><<<grid, 256, 0, t.ctx->stream>>>(t.slots, t.Dp, t.D, t.lk_slot, N, F, SB, delta, ldd, act, lda, t.acc, skip, p2p, hot_min);
  PS_LAUNCH_CHECK();
  switch (t.tpl) {                              /* the update spends 4 floats per lane whatever the scatter's chunking was */
    case 1: launch_update<1>(t, N, F, calls, skip, p2p != nullptr ? 1 : 0); break;
    case 2: launch_update<2>(t, N, F, calls, skip, p2p != nullptr ? 1 : 0); break;
    case 4: launch_update<4>(t, N, F, calls, skip, p2p != nullptr ? 1 : 0); break;
    case 8: launch_update<8>(t, N, F, calls, skip, p2p != nullptr ? 1 : 0); break;
    case 16: launch_update<16>(t, N, F, calls, skip, p2p != nullptr ? 1 : 0); break;
    default: launch_update<32>(t, N, F, calls, skip, p2p != nullptr ? 1 : 0); break;
  }
  t.ctx->launches += 2;
}

void EmbTable::scatter_update(const float* delta, int ldd, const float* act, int lda, int N, int calls, const int* skip_flag, int F_eff,
                              const P2PState* p2p) {
  const int Fe = F_eff > 0 ? F_eff : F;
  if (N > 0) {                                  /* N == 0: occupancy query from create() */
    PS_REQUIRE((int64_t)N * Fe == last_L, PS_ERR_STATE, "embedding: backward without a matching forward");
    PS_REQUIRE(calls == 1 || calls == 2, PS_ERR_ARG, "embedding: backward calls must be 1 or 2");
  }
  if (Dp % 8 == 0) {                            /* two 16 B chunks per lane: half the threads, twice the bytes in flight per thread */
    switch (pow2_ge(Dp / 8)) {
      case 1: launch_scatter<1, 2>(*this, delta, ldd, act, lda, N, Fe, calls, skip_flag, p2p); break;
      case 2: launch_scatter<2, 2>(*this, delta, ldd, act, lda, N, Fe, calls, skip_flag, p2p); break;
      case 4: launch_scatter<4, 2>(*this, delta, ldd, act, lda, N, Fe, calls, skip_flag, p2p); break;
      case 8: launch_scatter<8, 2>(*this, delta, ldd, act, lda, N, Fe, calls, skip_flag, p2p); break;
      default: launch_scatter<16, 2>(*this, delta, ldd, act, lda, N, Fe, calls, skip_flag, p2p); break;
    }
  } else {
    switch (tpl) {
      case 1: launch_scatter<1, 1>(*this, delta, ldd, act, lda, N, Fe, calls, skip_flag, p2p); break;
      case 2: launch_scatter<2, 1>(*this, delta, ldd, act, lda, N, Fe, calls, skip_flag, p2p); break;
      case 4: launch_scatter<4, 1>(*this, delta, ldd, act, lda, N, Fe, calls, skip_flag, p2p); break;
      case 8: launch_scatter<8, 1>(*this, delta, ldd, act, lda, N, Fe, calls, skip_flag, p2p); break;
      case 16: launch_scatter<16, 1>(*this, delta, ldd, act, lda, N, Fe, calls, skip_flag, p2p); break;
      default: launch_scatter<32, 1>(*this, delta, ldd, act, lda, N, Fe, calls, skip_flag, p2p); break;
    }
  }
  if (N == 0) return;
  PS_LAUNCH_CHECK();
  last_L = 0;
}

void EmbTable::clear_batch() {
  if (last_L <= 0) return;
  emb_clear_batch_kernel<<<std::min<long>(ceil_div(last_L, 256), (long)ctx->num_sms * 4), 256, 0, ctx->stream>>>(slots, lk_slot, (int)last_L);
  PS_LAUNCH_CHECK();
  ctx->launches++;
  last_L = 0;
}

void EmbTable::check_errors() {
  uint32_t h[4];
  PS_CUDA(cudaMemcpyAsync(h, counters, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  PS_REQUIRE(h[1] == 0, PS_ERR_CAPACITY, "embedding table is full: raise capacity");
}

int64_t EmbTable::size() {
  uint32_t h[4];
  PS_CUDA(cudaMemcpyAsync(h, counters, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  return (int64_t)(((uint64_t)h[3] << 32) | h[2]);
}

void EmbTable::get_rows(const int32_t* fields, const int64_t* ids, int n, float* w_out, float* s1_out, float* s2_out, int32_t* found) {
  if (n <= 0) return;
  cudaStream_t st = ctx->stream;
  int32_t* d_f = dmalloc<int32_t>(n); int64_t* d_i = dmalloc<int64_t>(n); int32_t* d_found = dmalloc<int32_t>(n);
  float* d_w = dmalloc<float>((size_t)n * D);
  float* d_s1 = s1_out ? dmalloc<float>((size_t)n * D) : nullptr;
  float* d_s2 = s2_out ? dmalloc<float>((size_t)n * D) : nullptr;
  PS_CUDA(cudaMemcpyAsync(d_f, fields, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
  PS_CUDA(cudaMemcpyAsync(d_i, ids, sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
  emb_get_rows_kernel<<<ceil_div(n, 128), 128, 0, st>>>(slots, (uint32_t)C, w, s1, s2, Dp, D, d_f, d_i, n, d_w, d_s1, d_s2, d_found);
  PS_LAUNCH_CHECK();
  ctx->launches++;
  PS_CUDA(cudaMemcpyAsync(w_out, d_w, sizeof(float) * n * D, cudaMemcpyDeviceToHost, st));
  if (s1_out) PS_CUDA(cudaMemcpyAsync(s1_out, d_s1, sizeof(float) * n * D, cudaMemcpyDeviceToHost, st));
  if (s2_out) PS_CUDA(cudaMemcpyAsync(s2_out, d_s2, sizeof(float) * n * D, cudaMemcpyDeviceToHost, st));
  PS_CUDA(cudaMemcpyAsync(found, d_found, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  PS_CUDA(cudaStreamSynchronize(st));
  dfree(d_f); dfree(d_i); dfree(d_found); dfree(d_w); dfree(d_s1); dfree(d_s2);
}

void EmbTable::put_rows(const int32_t* fields, const int64_t* ids, int n, float* w_io, int replace) {
  if (n <= 0) return;
  cudaStream_t st = ctx->stream;
  int32_t* d_f = dmalloc<int32_t>(n); int64_t* d_i = dmalloc<int64_t>(n);
  float* d_w = dmalloc<float>((size_t)n * D);
  PS_CUDA(cudaMemcpyAsync(d_f, fields, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
  PS_CUDA(cudaMemcpyAsync(d_i, ids, sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
  PS_CUDA(cudaMemcpyAsync(d_w, w_io, sizeof(float) * n * D, cudaMemcpyHostToDevice, st));
  emb_put_rows_kernel<<<ceil_div(n, 128), 128, 0, st>>>(slots, (uint32_t)C, w, Dp, D, d_f, d_i, n, d_w, replace, counters);
  PS_LAUNCH_CHECK();
  ctx->launches++;
  PS_CUDA(cudaMemcpyAsync(w_io, d_w, sizeof(float) * n * D, cudaMemcpyDeviceToHost, st));
  PS_CUDA(cudaStreamSynchronize(st));
  dfree(d_f); dfree(d_i); dfree(d_w);
  check_errors();
}

/* ------------------------------------------------------------------ WideTable */
__device__ __forceinline__ int wide_find_or_insert(WideSlot* slots, uint32_t C, unsigned long long key, bool insert, bool* inserted) {
  uint32_t slot = ps_bucket_of(key, C);
  const int limit = C < (uint32_t)kProbeLimit ? (int)C : kProbeLimit;
  *inserted = false;
  for (int p = 0; p < limit; ++p) {
    const unsigned long long k = *reinterpret_cast<const volatile unsigned long long*>(&slots[slot].key);
    if (k == key) return (int)slot;
    if (k == PS_KEY_EMPTY) {
      if (!insert) return -1;
      const unsigned long long old = atomicCAS(&slots[slot].key, (unsigned long long)PS_KEY_EMPTY, key);
      if (old == PS_KEY_EMPTY) { *inserted = true; return (int)slot; }
      if (old == key) return (int)slot;
    }
    slot = slot + 1 == C ? 0 : slot + 1;
  }
  return -1;
}

/* One warp per sample: lane j resolves "wide.weights.<W[j,n]>" (created as zeros(1) on first
 * touch, LRLayer.java:40-44,78), lane 0 then adds the F weights in j order exactly like the
 * reference's `sumW += wi.get(0)` loop (:76-81) and adds the bias (:84).                      */
__global__ void __launch_bounds__(256) wide_forward_kernel(WideSlot* __restrict__ slots, uint32_t C, const int64_t* __restrict__ ids, int N, int F,
                                                           const float* __restrict__ bias, float* __restrict__ z, uint32_t* __restrict__ counters) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= N) return;
  float sum = 0.0f;
  for (int j0 = 0; j0 < F; j0 += 32) {
    const int j = j0 + lane;
    float v = 0.0f;
    if (j < F) {
      bool inserted;
      const int slot = wide_find_or_insert(slots, C, ps_pack_key(0u, (uint64_t)ids[(size_t)warp * F + j]), true, &inserted);
      if (slot < 0) counters[0] = 1u;
      else {
        if (inserted) atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2), 1ull);
        v = *reinterpret_cast<const volatile float*>(&slots[slot].w);
      }
    }
    const int cnt = min(32, F - j0);
    for (int k = 0; k < cnt; ++k) sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, v, k));
  }
  if (lane == 0) z[warp] = __fadd_rn(sum, bias[0]);
}

__global__ void __launch_bounds__(256) wide_update_all_kernel(WideSlot* __restrict__ slots, uint32_t C, UpdaterDev upd, const float* __restrict__ gbar,
                                                              const int* __restrict__ skip_flag, float* __restrict__ bias, UpdaterDev bias_upd) {
  if (skip_flag != nullptr && *skip_flag != 0) return;
  const float g = *gbar;
  if (bias != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && !(bias_upd.kind == PS_UPD_FTRL && g == 0.0f)) {
    float w = bias[0], a = bias[1], b = bias[2];       /* "wide.bias" gets the same gbar (LRLayer.java:112-113) */
    apply_elem(bias_upd, w, a, b, g);
    bias[0] = w; bias[1] = a; bias[2] = b;
  }
  if (upd.kind == PS_UPD_FTRL && g == 0.0f) return;   /* FtrlUpdater.java:52 */
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < C; s += gridDim.x * blockDim.x) {
    WideSlot r = slots[s];
    if (r.key == PS_KEY_EMPTY) continue;
    apply_elem(upd, r.w, r.s1, r.s2, g);
    slots[s].w = r.w; slots[s].s1 = r.s1; slots[s].s2 = r.s2;
  }
}

/* replicas of the wide table must know every key any rank has seen (LRLayer.weights never shrinks) */
__global__ void __launch_bounds__(256) wide_insert_kernel(WideSlot* __restrict__ slots, uint32_t C, const int64_t* __restrict__ ids, int n,
                                                          uint32_t* __restrict__ counters, const P2PState* __restrict__ p2p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (p2p != nullptr) ids = reinterpret_cast<const int64_t*>(p2p_region(p2p, p2p->me, p2p->off_wide));   /* this step's wide_in mailbox */
  bool inserted;
  const int slot = wide_find_or_insert(slots, C, ps_pack_key(0u, (uint64_t)ids[i]), true, &inserted);
  if (slot < 0) counters[0] = 1u;
  else if (inserted) atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2), 1ull);
}

__global__ void wide_get_kernel(WideSlot* slots, uint32_t C, int64_t id, float* out) {
  bool ins;
  const int slot = wide_find_or_insert(slots, C, ps_pack_key(0u, (uint64_t)id), false, &ins);
  out[0] = slot >= 0 ? 1.f : 0.f;
  if (slot >= 0) { out[1] = slots[slot].w; out[2] = slots[slot].s1; out[3] = slots[slot].s2; }
}
__global__ void wide_put_kernel(WideSlot* slots, uint32_t C, int64_t id, float wv, uint32_t* counters) {
  bool ins;
  const int slot = wide_find_or_insert(slots, C, ps_pack_key(0u, (uint64_t)id), true, &ins);
  if (slot < 0) { counters[0] = 1u; return; }
  if (ins) atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2), 1ull);
  slots[slot].w = wv;
}

void WideTable::create(Ctx* c, int64_t capacity, const ps_updater_spec& u) {
  PS_REQUIRE(capacity > 0 && capacity < (1ll << 31), PS_ERR_ARG, "wide: capacity must be in (0, 2^31)");
  ctx = c; C = capacity; upd = make_updater_dev(u);
  slots = dmalloc_zero<WideSlot>((size_t)C, ctx->stream);
  counters = dmalloc_zero<uint32_t>(4, ctx->stream);
}
void WideTable::destroy() { dfree(slots); dfree(counters); slots = nullptr; counters = nullptr; }

void WideTable::forward(const int64_t* ids, int N, int F, const float* bias, float* z) {
  wide_forward_kernel<<<ceil_div((long)N * 32, 256), 256, 0, ctx->stream>>>(slots, (uint32_t)C, ids, N, F, bias, z, counters);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}
void WideTable::insert(const int64_t* ids, int n, const P2PState* p2p) {
  if (n <= 0) return;
  wide_insert_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(slots, (uint32_t)C, ids, n, counters, p2p);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}
void WideTable::update_all(const float* gbar, const int* skip_flag, float* bias, const ps_updater_spec* bias_upd) {
  const int grid = std::min<long>(ceil_div(C, 256), (long)ctx->num_sms * 8);
  wide_update_all_kernel<<<grid, 256, 0, ctx->stream>>>(slots, (uint32_t)C, upd, gbar, skip_flag, bias, bias_upd ? make_updater_dev(*bias_upd) : upd);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}
int64_t WideTable::size() {
  uint32_t h[4];
  PS_CUDA(cudaMemcpyAsync(h, counters, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  return (int64_t)(((uint64_t)h[3] << 32) | h[2]);
}
void WideTable::check_errors() {
  uint32_t h[4];
  PS_CUDA(cudaMemcpyAsync(h, counters, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  PS_REQUIRE(h[0] == 0, PS_ERR_CAPACITY, "wide table is full: raise capacity");
}
int WideTable::get(int64_t id, float* wv, float* s1v, float* s2v) {
  float* d = dmalloc_zero<float>(4, ctx->stream);
  wide_get_kernel<<<1, 1, 0, ctx->stream>>>(slots, (uint32_t)C, id, d);
  ctx->launches++;
  float h[4];
  PS_CUDA(cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  dfree(d);
  if (h[0] == 0.f) return 0;
  if (wv) *wv = h[1]; if (s1v) *s1v = h[2]; if (s2v) *s2v = h[3];
  return 1;
}
void WideTable::put(int64_t id, float wv) {
  wide_put_kernel<<<1, 1, 0, ctx->stream>>>(slots, (uint32_t)C, id, wv, counters);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

}  // namespace psb

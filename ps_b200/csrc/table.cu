/*
 * table.cu — kernels of the GPU-resident embedding / wide tables (sm_100a).
 *
 * Path restated (reference, /root/reference/src/main/java/):
 *   lookup  = EmbeddingField.checkExists + KVStore.get(key, init)   layer/EmbeddingField.java:49-54, store/KVStore.java:136-159,168-190
 *           + EmbeddingField.forward + EmbeddingLayer.forward       layer/EmbeddingField.java:66-78, layer/EmbeddingLayer.java:25-48
 *   scatter = EmbeddingField.backward (x2) + KVStore.sum + KVStore.update + Updater.update
 *                                                                   layer/EmbeddingField.java:86-104, store/KVStore.java:192-200,240-268
 *   wide    = LRLayer.forward / backward                            layer/LRLayer.java:62-120
 *
 * All of it is HBM/L2-bound integer and copy work: no tensor cores, 128-bit accesses,
 * one thread group (Dp/4 lanes) per looked-up row; rows that several lookups of a warp share are
 * staged ONCE in shared memory by the TMA engine (cp.async.bulk + mbarrier).
 */
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "p2p.cuh"
#include "table.cuh"

namespace psb {


/* ------------------------------------------------------------------ find / find-or-insert */
__device__ __forceinline__ ulonglong2 ld_slot(const EmbSlot* p) {   /* {key, cnt | uidx << 32} in one 16 B transaction */
  ulonglong2 r;
  asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
  return r;
}

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

/* find-or-insert that starts from an already loaded home-bucket record.  *ready: the record that matched carried
 * kRowReady, i.e. the row had been written (and fenced) before this thread saw the key — the row may be loaded.
 * Otherwise (inserted here, or the key was published by a thread that may still be writing the row) the caller takes
 * the row from the deterministic initialiser instead of from memory: same bits, no waiting.                       */
__device__ __forceinline__ int emb_resolve(EmbSlot* slots, uint32_t C, unsigned long long key, uint32_t slot, ulonglong2 rec, bool* inserted,
                                           bool* ready) {
  const int limit = C < (uint32_t)kProbeLimit ? (int)C : kProbeLimit;
  *inserted = false; *ready = false;
  for (int p = 0; p < limit; ++p) {
    if (p > 0) rec = ld_slot(&slots[slot]);
    if (rec.x == key) { *ready = ((uint32_t)(rec.y >> 32) & kRowReady) != 0u; return (int)slot; }
    if (rec.x == PS_KEY_EMPTY) {
      const unsigned long long old = atomicCAS(&slots[slot].key, (unsigned long long)PS_KEY_EMPTY, key);
      if (old == PS_KEY_EMPTY) { *inserted = true; return (int)slot; }
      if (old == key) return (int)slot;
    }
    slot = slot + 1 == C ? 0 : slot + 1;
  }
  return -1;
}

/* ---- TMA (bulk async copy) + mbarrier wrappers for the hot-row staging ---- */
__device__ __forceinline__ uint32_t tb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tb_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tb_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tb_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
/* global -> shared bulk copy executed by the TMA unit; completion is signalled on the mbarrier as transferred bytes */
__device__ __forceinline__ void tb_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

/* 16 B asynchronous global -> shared copy (LDGSTS, L2 only): no register is held while it is in flight */
__device__ __forceinline__ void tb_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

/* true in exactly one block of the grid: the one whose threads arrive last.  Call from all threads.
 * system: fence the block's stores at system scope before the ticket (only for the A/B switch of p2p_publish_last: the last
 * block's own system-scope fence is cumulative over everything the tickets made it observe). */
__device__ __forceinline__ bool last_block_done(uint32_t* ticket, uint32_t nblocks, bool system = false) {
  __shared__ bool s_last;
  __syncthreads();                               /* the block's writes happen-before thread 0's (cumulative) fence */
  if (threadIdx.x == 0) {
    if (system) __threadfence_system(); else __threadfence();
    const uint32_t t = atomicAdd(ticket, 1u);
    s_last = t == nblocks - 1u;
    if (s_last) *ticket = 0u;
  }
  __syncthreads();
  return s_last;
}

struct LookupArgs {
  EmbSlot* slots; uint32_t C;
  float* rows; int rs, Dp, D;
  const void* ids; int N, F;             /* F == 0: ids holds packed keys (the owner side of the sharded exchange) */
  uint64_t seed; float maxv;
  int32_t* lk_slot; uint32_t* lk_mask; int MW;
  int32_t* uniq; uint32_t* counters;
  P2PState* p2p; int send_rows;          /* keys from this step's keys_in mailbox (after waiting for CH_KEYS); rows to the requesters'
                                            rows_in mailboxes, CH_ROWS published by the last block */
  float* out; int ldo;
  const float* X; int Xn, xoff;
  /* requester side of the exchange ("pre-resolved"): lookup t reads record pre_recs[pre_lk[t]] = {key, count, row index | -1} of the
   * per-batch de-duplication table instead of probing; rows come from this step's rows_in mailbox (after waiting for CH_ROWS) */
  const EmbSlot* pre_recs; const int32_t* pre_lk;
  /* owner side of the exchange: instead of counting, every entry (requester, position) links itself into its key's chain — the
   * slot's batch counter holds 1 + the chain's head entry, chain[entry] the next one (0 ends it).  The update kernel walks it to
   * find the gradient sums the requesters hold for the key (at most one entry per requester: they de-duplicate).            */
  int32_t* chain;
  int task_blocks, hot_tma, hot_share;   /* hot_share: lookups of one warp task that must share a row before the TMA unit fetches it (1: every row) */
};

static size_t lookup_smem_bytes(int Dp) { return 128 + (size_t)8 * 32 * Dp * sizeof(float); }
static constexpr int kHotShare = 4;                 /* lookups of one warp task that must share a row before it is staged by TMA */
static constexpr int kHotRows = 32 / kHotShare;     /* so at most this many staged rows per warp */

__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src) {
  const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
  return ((unsigned long long)hi << 32) | lo;
}

/* EmbeddingLayer.forward as ONE kernel.  A warp task is (field j, 32 consecutive samples); the warps of a persistent grid
 * stride over the tasks, neighbouring warps taking neighbouring fields of the same samples (the 32 B sectors of the [N][F]
 * id matrix are shared inside a block).  A warp's loop is software-pipelined three tasks deep so that the dependent loads of
 * a lookup (id -> slot record -> row) of DIFFERENT tasks are in flight together:
 *   stage A (task i+2)  load the ids
 *   stage B (task i+1)  pack the keys, hash, load the home-bucket slot records
 *   stage C (task i)    1. resolve   lane <-> sample: find-or-insert the key (row creation follows KVStore.create,
 *                          KVStore.java:168-190: the creating thread draws the row from the deterministic initialiser of
 *                          ps_spec.h, optimiser state stays at the arena's zero, AdamUpdater.java:76-84); lk_slot[j][n] = slot.
 *                       2. count     duplicates of a key inside the warp (a low-cardinality field) elect one lane, which adds
 *                          the group's occurrences to the slot's batch counter with ONE returning atomic; the group that
 *                          finds the counter at zero owns the key for this batch: it appends the slot to the batch's
 *                          unique list (one cursor atomic per warp) and leaves 1 + that index in the slot record — the
 *                          accumulator row of the backward.  Both atomics return while the row loads are in flight.
 *                       3. gather    TPL lanes per row move relu(row) as 128-bit chunks to out[n][j*D ..]
 *                          (EmbeddingField.java:73-76, EmbeddingLayer.java:36-46) and record the mask bits.  Rows wanted by
 *                          >= kHotShare lookups of the task are fetched ONCE into shared memory by the TMA unit
 *                          (cp.async.bulk, mbarrier completion) and read from there; the others stream from L2/HBM.
 * The cursor of the unique list is final when the kernel ends: it is the update kernel's bound and StepStatus.n_unique. */
template <class IdT, bool GATHER, int TPL, bool ALIGNED>
__global__ void __launch_bounds__(256) emb_lookup_kernel(const __grid_constant__ LookupArgs a) {
  extern __shared__ __align__(128) unsigned char lookup_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  /* owner side of the exchange, launched as a programmatic dependent of route_send: that grid must be complete (it publishes the
   * step's sequence number last).  Only then are this kernel's own dependents released, so whatever follows also runs after it. */
  if (a.p2p != nullptr && a.pre_lk == nullptr) pdl_wait();
  pdl_launch_dependents();
  if ((int)blockIdx.x >= a.task_blocks) {          /* ConcatLayer.forward (ConcatLayer.java:30-37): numeric features next to the embeddings */
    const long base = (long)((int)blockIdx.x - a.task_blocks) * 1024 + threadIdx.x;
    const long total = (long)a.N * a.Xn;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long i = base + k * 256;
      if (i < total) { const int n = (int)(i / a.Xn), x = (int)(i - (long)n * a.Xn); a.out[(size_t)n * a.ldo + a.xoff + x] = a.X[i]; }
    }
    return;
  }
  const bool pre = a.pre_lk != nullptr;
  if (a.p2p != nullptr) p2p_wait_all(a.p2p, pre ? CH_ROWS : CH_KEYS);   /* the owners' rows / every requester's keys (and their count) have landed */
  const float* __restrict__ rowsp = pre ? reinterpret_cast<const float*>(p2p_region(a.p2p, a.p2p->me, a.p2p->off_rows)) : a.rows;
  const int rsp = pre ? a.Dp : a.rs;
  const IdT* __restrict__ ids = (a.p2p != nullptr && !pre) ? reinterpret_cast<const IdT*>(p2p_region(a.p2p, a.p2p->me, a.p2p->off_keys)) : static_cast<const IdT*>(a.ids);
  const int32_t* __restrict__ pcounts = (a.p2p != nullptr && !pre) ? reinterpret_cast<const int32_t*>(p2p_region(a.p2p, a.p2p->me, a.p2p->off_counts)) : nullptr;
  const int Fe = a.F > 0 ? a.F : 1;
  const long ntasks = (long)((a.N + 31) / 32) * Fe;
  const long W = (long)a.task_blocks * 8;          /* warps striding over the tasks */
  const long w0 = (long)blockIdx.x * 8 + warp;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(lookup_smem);
  float* stage = reinterpret_cast<float*>(lookup_smem + 128) + (size_t)warp * 32 * a.Dp;   /* the warp's slab: one row per lookup of a task */
  uint32_t hot_parity = 0u;
  if (GATHER) {
    if (lane == 0) tb_mbar_init(tb_smem_u32(&mbar[warp]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
  }
  /* pipeline registers: raw ids of task i+2 (loaded in the previous iteration), key / home bucket / its record of task i+1 */
  IdT raw_n = IdT(0); bool rv_n = false;
  unsigned long long key_c = PS_KEY_EMPTY; uint32_t bucket_c = 0u; ulonglong2 rec_c = make_ulonglong2(0ull, 0ull);

#pragma unroll 1
  for (long it = -2;; ++it) {
    const long tC = w0 + it * W, tB = tC + W, tA = tB + W;
    if (it >= 0 && tC >= ntasks) break;
    /* ---- stage A: ids of task tA ---- */
    IdT raw_a = IdT(0); bool rv_a = false;
    if (tA < ntasks) {
      const int sg = (int)(tA / Fe), j = (int)(tA - (long)sg * Fe);
      const long n = (long)sg * 32 + lane;
      if (n < a.N) {
        if (pre) { const int b = a.pre_lk[(long)j * a.N + n]; raw_a = (IdT)b; rv_a = b >= 0; }
        else if (a.p2p != nullptr) {               /* owner side of the peer-memory exchange: entry n = (requester n / cap, its idx-th key) */
          const int src = (int)(n / a.p2p->cap), idx = (int)(n - (long)src * a.p2p->cap);
          if (idx < pcounts[src]) { raw_a = ids[n]; rv_a = true; }
        } else if (a.F > 0) { raw_a = ids[n * a.F + j]; rv_a = true; }
        else { raw_a = ids[n]; rv_a = true; }
      }
    }
    /* ---- stage B: key, hash and home-bucket record of task tB (its ids were requested one iteration ago) ---- */
    unsigned long long key_b = PS_KEY_EMPTY; uint32_t bucket_b = 0u; ulonglong2 rec_b = make_ulonglong2(0ull, 0ull);
    if (pre) {
      if (it >= -1 && tB < ntasks && rv_n) { key_b = 1ull; bucket_b = (uint32_t)(int64_t)raw_n; rec_b = ld_slot(&a.pre_recs[bucket_b]); }
    } else if (it >= -1 && tB < ntasks && rv_n) {
      const int j = (int)(tB % Fe);
      key_b = a.F > 0 ? ps_pack_key((uint32_t)j, (uint64_t)(int64_t)raw_n) : (unsigned long long)raw_n;   /* EMPTY marks padding in the fixed-capacity exchange */
      /* ids outside [0, 2^44) would silently alias another key's row (ps_pack_key masks): refuse the batch instead */
      if (a.F > 0 && (uint64_t)(int64_t)raw_n > PS_KEY_ID_MASK) { atomicOr(&a.counters[CNT_ERR], 2u); key_b = PS_KEY_EMPTY; }
      if (key_b != PS_KEY_EMPTY) { bucket_b = ps_bucket_of(key_b, a.C); rec_b = ld_slot(&a.slots[bucket_b]); }
    }
    /* ---- stage C: task tC ---- */
    if (it >= 0) {
      const int sg = (int)(tC / Fe), j = (int)(tC - (long)sg * Fe);
      const long nbase = (long)sg * 32;
      const long n = nbase + lane;
      const bool in = n < a.N;
      const unsigned long long key = key_c;
      /* 1. resolve */
      int slot = -1;
      bool ready = false;
      if (pre) {
        if (key != PS_KEY_EMPTY) { slot = (int)(uint32_t)(rec_c.y >> 32); ready = slot >= 0; }   /* the row's place in rows_in (-1: its bucket overflowed) */
      } else if (key != PS_KEY_EMPTY) {
        bool inserted;
        slot = emb_resolve(a.slots, a.C, key, bucket_c, rec_c, &inserted, &ready);
        if (slot < 0) atomicOr(&a.counters[CNT_ERR], 1u);   /* table full: the tail turns this into the step's skip flag — nothing is updated */
        else if (inserted) {
          float* row = a.rows + (size_t)slot * a.rs;
          for (int d = 0; d < a.D; ++d) row[d] = ps_init_value(a.seed, key, (uint32_t)d, a.maxv);
          __threadfence();                            /* the row is visible before anybody can see kRowReady */
          atomicOr(&a.slots[slot].uidx, kRowReady);
          atomicAdd(reinterpret_cast<unsigned long long*>(a.counters + CNT_ROWS), 1ull);
        }
      }
      if (in && !pre) a.lk_slot[(long)j * a.N + n] = slot;
      /* 2. count: every lookup counts 1 (owner side of the exchange: every requester's entry; the occurrence counts arrive with the push) */
      const unsigned peers = __match_any_sync(0xffffffffu, slot >= 0 ? slot : (-1 - lane));
      const int leader = __ffs(peers) - 1;
      uint32_t old = 1u, ubase = 0u;
      const bool counts_itself = a.chain != nullptr || leader == lane;
      if (slot >= 0 && !pre) {
        if (a.chain != nullptr) { old = atomicExch(&a.slots[slot].cnt, (uint32_t)n + 1u); a.chain[n] = (int32_t)old; }
        else if (leader == lane) old = atomicAdd(&a.slots[slot].cnt, (uint32_t)__popc(peers));
      }
      unsigned omask = 0u;
      bool is_owner = false;
      /* the group that found the batch counter at zero owns the key: claim its place in the unique list (consumes `old`) */
      auto claim = [&]() {
        is_owner = slot >= 0 && counts_itself && old == 0u;
        omask = __ballot_sync(0xffffffffu, is_owner);
        if (omask != 0u && lane == __ffs(omask) - 1) ubase = atomicAdd(&a.counters[CNT_CURSOR], (uint32_t)__popc(omask));
      };
      /* 3. gather: every DISTINCT row of the task is staged once in the warp's shared-memory slab — rows several lookups of the
       * task share (a low-cardinality field) by the TMA unit (cp.async.bulk, mbarrier completion), the others by 16 B
       * asynchronous copies (LDGSTS) — so a whole task's rows are in flight at once without holding a register; the slab is
       * then written out with ReLU and the mask bits.  Rows created by this very kernel come from the initialiser's bits. */
      if (GATHER) {
        constexpr int GPW = 32 / TPL;                 /* rows per pass */
        constexpr int NP = TPL;                       /* passes over the task's 32 rows */
        constexpr int MSH = TPL < 8 ? TPL : 8;        /* lanes whose mask nibbles share one 32-bit word */
        const int part = lane % TPL, grp = lane / TPL;
        const bool lane_on = part * 4 < a.Dp;
        const bool fetch = slot >= 0 && leader == lane;               /* one lane per distinct row */
        const bool hot = fetch && ready && a.hot_tma && __popc(peers) >= a.hot_share;
        const unsigned hmask = __ballot_sync(0xffffffffu, hot);
        /* the slab was read with ordinary loads by the previous task: order those before the async-proxy writes below */
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (hmask != 0u) {                            /* warp-uniform */
          const uint32_t bar = tb_smem_u32(&mbar[warp]);
          if (lane == 0) tb_mbar_expect_tx(bar, (uint32_t)__popc(hmask) * (uint32_t)a.Dp * 4u);
          __syncwarp();
          if (hot) tb_bulk_g2s(tb_smem_u32(stage + (size_t)lane * a.Dp), rowsp + (size_t)slot * rsp, (uint32_t)a.Dp * 4u, bar);
        }
        const int code = !fetch ? 0 : hot ? 1 : ready ? 2 : 3;       /* how row `lane` reaches the slab: - | TMA | LDGSTS | initialiser */
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const int r = p * GPW + grp;
          const int rs_ = __shfl_sync(0xffffffffu, slot, r);
          const int rc_ = __shfl_sync(0xffffffffu, code, r);
          if (rc_ == 2 && lane_on) tb_cp_async16(tb_smem_u32(stage + (size_t)r * a.Dp + part * 4), rowsp + (size_t)rs_ * rsp + part * 4);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (code == 3) {                              /* created by this kernel (here or elsewhere): the initialiser's bits, not memory */
          float* srow = stage + (size_t)lane * a.Dp;
          for (int d = 0; d < a.Dp; ++d) srow[d] = d < a.D ? ps_init_value(a.seed, key, (uint32_t)d, a.maxv) : 0.f;
        }
        claim();                                      /* the count atomic has returned by now; the cursor atomic flies with the copies */
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (hmask != 0u) { tb_mbar_wait(tb_smem_u32(&mbar[warp]), hot_parity); hot_parity ^= 1u; }
        __syncwarp();
#pragma unroll 4
        for (int p = 0; p < NP; ++p) {
          const int r = p * GPW + grp;
          const long nr = nbase + r;
          const int rs_ = __shfl_sync(0xffffffffu, slot, r);
          const int rl_ = __shfl_sync(0xffffffffu, leader, r);      /* the lane whose slab row holds row r's key */
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rs_ >= 0 && lane_on) v = *reinterpret_cast<const float4*>(stage + (size_t)rl_ * a.Dp + part * 4);
          /* EmbeddingField.java:75: relu in place; the mask bit is what Relu.backward will ask for (Relu.java:14-19) */
          uint32_t m = (v.x > 0.f ? 1u : 0u) | (v.y > 0.f ? 2u : 0u) | (v.z > 0.f ? 4u : 0u) | (v.w > 0.f ? 8u : 0u);
          v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
          if (a.lk_mask != nullptr) {
            m <<= (part & 7) * 4;
#pragma unroll
            for (int o = 1; o < MSH; o <<= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
            if (nr < a.N && (part & 7) == 0 && lane_on) a.lk_mask[((size_t)j * a.N + nr) * a.MW + (part >> 3)] = m;
          }
          if (nr >= a.N || !lane_on) continue;
          float* o;
          if (a.F > 0) o = a.out + (size_t)nr * a.ldo + j * a.D + part * 4;
          else if (a.send_rows) {                     /* PServer.getList response: straight into the requester's rows_in[me][idx] */
            if (rs_ < 0) continue;
            const int src = (int)(nr / a.p2p->cap), idx = (int)(nr - (long)src * a.p2p->cap);
            o = reinterpret_cast<float*>(p2p_region(a.p2p, src, a.p2p->off_rows)) + ((size_t)a.p2p->me * a.p2p->cap + idx) * a.Dp + part * 4;
          } else o = a.out + (size_t)nr * a.ldo + part * 4;
          if (ALIGNED || a.F == 0) {
            st_f4(o, v);
          } else {
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) if (part * 4 + i < a.D) o[i] = e[i];
          }
        }
      } else {
        claim();
      }
      /* 2b. the owners take consecutive places after the warp's cursor reservation */
      if (omask != 0u) {
        ubase = __shfl_sync(0xffffffffu, ubase, __ffs(omask) - 1);
        if (is_owner) {
          const uint32_t u = ubase + (uint32_t)__popc(omask & ((1u << lane) - 1u));
          a.uniq[u] = slot;
          atomicOr(&a.slots[slot].uidx, u + 1u);
        }
      }
    }
    /* ---- rotate the pipeline ---- */
    key_c = key_b; bucket_c = bucket_b; rec_c = rec_b;
    raw_n = raw_a; rv_n = rv_a;
  }
  /* owner side of the exchange only: the block that finishes last flags every requester (PServer.getList answered) */
  if (a.send_rows && last_block_done(&a.counters[CNT_TICKET_FWD], (uint32_t)a.task_blocks, a.p2p->block_fence_sys != 0) && (int)threadIdx.x < a.p2p->R) {
    __threadfence_system();
    p2p_st_release_sys(reinterpret_cast<uint32_t*>(p2p_region(a.p2p, threadIdx.x, a.p2p->off_flags)) + CH_ROWS * kP2PMaxRanks + a.p2p->me, a.p2p->seq);
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
 *   emb_scatter_kernel  g_k = delta[:,k] * (A[:,k] > 0)  (EmbeddingField.java:91-93; the mask comes from the bits the lookup
 *            recorded, so the activations are not read again).  A block owns SB consecutive samples
 *            and all F fields of them; a warp task is (field j, 32/TPL consecutive samples), so duplicates of a key meet
 *            in the same warp and block, and the 8 warps read neighbouring columns of the same delta rows at the
 *            same time (whole DRAM pages).  Three levels of pre-summation keep a hot key (a low-cardinality field:
 *            thousands of occurrences of one row) from serialising in L2:  (1) reduce-by-key tree over the lanes of a
 *            warp that __match_any_sync groups;  (2) keys frequent enough to recur inside one block's samples (the
 *            lookup left the batch count in the slot record; threshold = max(PS_HOT_MIN, 2N/SB)) are summed in a
 *            per-block shared-memory table and leave the block ONCE;  (3) everything else goes out as one
 *            red.global.add.v4.f32 per key and 16 B chunk into the key's accumulator row acc[uidx] (compact, L2-resident).
 *   emb_update_kernel   walks the batch's unique list: 32/(Dp/4) keys per warp, one 16 B chunk per lane: read S back from
 *            L2, form g_eff, run the Adam / Ftrl / SGD step on the key's {w | s1 | s2} record in place and reset the
 *            per-batch state (acc, cnt, uidx) — KVStore.sum + update + clear (KVStore.java:192-200,240-277).  Launched
 *            with programmatic stream serialisation: the list, the counts and the records of its first round are
 *            requested while the scatter kernel drains; only the accumulator read sits behind griddepcontrol.wait. */
static constexpr int kHotMin = 8;        /* default occurrences in the batch from which a key is pre-summed per block (PS_HOT_MIN) */
static constexpr int kHotBits = 6;
static constexpr int kHotEntries = 1 << kHotBits;   /* per-block hot-key table (open addressing, 4 probes) */

template <int TPL, int CPL, int PASSES, bool ALIGNED>
__global__ void __launch_bounds__(256) emb_scatter_kernel(const EmbSlot* __restrict__ slots, int Dp, int D, const int32_t* __restrict__ lk_slot,
                                                          const uint32_t* __restrict__ lk_mask, int MW, int N,
                                                          int F, int SB, const float* __restrict__ delta, int ldd, const float* __restrict__ act, int lda,
                                                          float* __restrict__ acc, const int* __restrict__ skip_flag,
                                                          int raw_row, uint32_t hot_min, P2PState* pub) {
  constexpr int GPW = 32 / TPL;                  /* lookups (lane groups) per warp task */
  /* requester side of the sharded push (skip_flag is null there): the sums go to this step's gsums region of this rank's slab,
   * where the owners read them; the block that finishes last flags them (CH_GRADS) */
  if (pub != nullptr) acc = reinterpret_cast<float*>(p2p_region(pub, pub->me, pub->off_grads));
  constexpr int ROWF = TPL * CPL * 4;            /* floats of a (padded) row */
  __shared__ float hot_acc[kHotEntries][ROWF];
  __shared__ int hot_slot[kHotEntries];
  __shared__ uint32_t hot_row[kHotEntries];
  pdl_launch_dependents();                       /* the update kernel may start its prefetch now (it waits before reading acc) */
  if (skip_flag != nullptr && *skip_flag != 0) return;   /* DNN.java:58-63 early exit: nothing is pushed */
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
  pdl_wait();                                    /* (as a programmatic dependent of the last dgrad: delta is complete from here on) */
  __syncthreads();

  /* the block's tile: SB consecutive samples x all F fields; a warp task = (field j, GPW consecutive samples); the 8 warps
   * work on neighbouring fields of the same samples at the same time, so the rows of delta are read as whole
   * DRAM pages although every lookup only needs D of their columns */
  const int tasks = F * (SB / GPW);
  /* persistent blocks (one wave of them, see launch_scatter) stride over the tiles: no tail wave, and the hot table keeps
   * summing across all tiles of the block */
  for (long n0 = (long)blockIdx.x * SB; n0 < N; n0 += (long)gridDim.x * SB)
  for (int task0 = warp; task0 < tasks; task0 += 8 * PASSES) {
    /* ---- every load of PASSES tasks is issued before anything is consumed ---- */
    int slot[PASSES];
    float4 gk[PASSES][CPL];
    uint32_t mw[PASSES];
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      const int task = task0 + 8 * p;
      slot[p] = -1; mw[p] = 0xFFFFFFFFu;
#pragma unroll
      for (int c = 0; c < CPL; ++c) gk[p][c] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int j = task % F;
      const long n = n0 + (long)(task / F) * GPW + grp;
      if (task < tasks && n < N) {
        slot[p] = lk_slot[(long)j * N + n];
        if (lk_mask != nullptr && c0 < D) mw[p] = lk_mask[((size_t)j * N + n) * MW + (c0 >> 5)] >> (c0 & 31);   /* this lane's 4*CPL bits */
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const int cc = c0 + 4 * c;
          if (cc < D) {
            const size_t od = (size_t)n * ldd + j * D + cc, oa = (size_t)n * lda + j * D + cc;
            float dv[4] = {0.f, 0.f, 0.f, 0.f}, av[4] = {1.f, 1.f, 1.f, 1.f};
            if (ALIGNED) {
              const float4 d4 = ld_f4(delta + od);
              dv[0] = d4.x; dv[1] = d4.y; dv[2] = d4.z; dv[3] = d4.w;
              if (act != nullptr) { const float4 a4 = ld_f4(act + oa); av[0] = a4.x; av[1] = a4.y; av[2] = a4.z; av[3] = a4.w; }
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) if (cc + i < D) { dv[i] = delta[od + i]; if (act != nullptr) av[i] = act[oa + i]; }
            }
            gk[p][c] = make_float4(dv[0], dv[1], dv[2], dv[3]);
            if (act != nullptr) {
              /* Relu.backward: dy *= (y > 0 ? 1 : 0)  (activations/Relu.java:14-19) */
              gk[p][c].x = __fmul_rn(dv[0], av[0] > 0.f ? 1.f : 0.f); gk[p][c].y = __fmul_rn(dv[1], av[1] > 0.f ? 1.f : 0.f);
              gk[p][c].z = __fmul_rn(dv[2], av[2] > 0.f ? 1.f : 0.f); gk[p][c].w = __fmul_rn(dv[3], av[3] > 0.f ? 1.f : 0.f);
            }
          }
        }
      }
    }
    uint32_t cnt[PASSES], row[PASSES];
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      cnt[p] = 0u; row[p] = 0u;
      if (slot[p] >= 0) {                          /* the key's batch count and accumulator row */
        const uint4 m = *reinterpret_cast<const uint4*>(&slots[slot[p]]);
        cnt[p] = m.z;
        if (raw_row) { row[p] = m.w; if ((int)m.w < 0) slot[p] = -1; }       /* requester side of the sharded exchange: BatchSlot {key, cnt, bucket position | -1} */
        else row[p] = (m.w & ~kRowReady) - 1u;
      }
      if (lk_mask != nullptr) {                    /* the same multiplication by 0 / 1, the factor taken from the recorded mask bits */
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const uint32_t b = mw[p] >> (4 * c);
          gk[p][c].x = __fmul_rn(gk[p][c].x, (b & 1u) ? 1.f : 0.f); gk[p][c].y = __fmul_rn(gk[p][c].y, (b & 2u) ? 1.f : 0.f);
          gk[p][c].z = __fmul_rn(gk[p][c].z, (b & 4u) ? 1.f : 0.f); gk[p][c].w = __fmul_rn(gk[p][c].w, (b & 8u) ? 1.f : 0.f);
        }
      }
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
      /* ---- (2) hot keys: sum inside the block ---- */
      int e = -1;
      if (cnt[p] >= hot_min) {
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
  if (pub != nullptr) p2p_publish_last(pub, CH_GRADS, gridDim.x);
}

/* The scatter for the common case (16 B aligned rows, ReLU mask from the lookup's bits): same three levels of pre-summation, but
 * a warp task is (field j, 32 consecutive samples) and the task's 32 delta rows are STAGED in the warp's shared-memory slab by
 * 16 B asynchronous copies (LDGSTS) issued before anything else — the whole task is in flight at once without holding
 * registers, beside the lk_slot / mask / slot-record loads.  Then, in shared memory: (1) every row is masked in place,
 * (2) rows whose key another lane of the task leads are added into the leader's row, (3) leader rows leave the warp — into the
 * block's hot-key table when the key is frequent in the batch, else as one red.global.add.v4.f32 per 16 B chunk into acc[uidx]. */
template <int TPL>
__global__ void __launch_bounds__(256) emb_scatter_slab_kernel(const EmbSlot* __restrict__ slots, int Dp, int D, const int32_t* __restrict__ lk_slot,
                                                               const uint32_t* __restrict__ lk_mask, int MW, int N, int F,
                                                               const float* __restrict__ delta, int ldd, float* __restrict__ acc,
                                                               const int* __restrict__ skip_flag, int raw_row, uint32_t hot_min, P2PState* pub) {
  constexpr int GPW = 32 / TPL, NP = TPL;
  if (pub != nullptr) acc = reinterpret_cast<float*>(p2p_region(pub, pub->me, pub->off_grads));   /* see emb_scatter_kernel */
  extern __shared__ __align__(128) unsigned char scatter_smem[];
  int* hot_slot = reinterpret_cast<int*>(scatter_smem);                                  /* [kHotEntries] */
  uint32_t* hot_row = reinterpret_cast<uint32_t*>(scatter_smem + 4 * kHotEntries);       /* [kHotEntries] */
  uint32_t* maskw_all = reinterpret_cast<uint32_t*>(scatter_smem + 8 * kHotEntries);     /* [8 warps][32][4] */
  float* hot_acc = reinterpret_cast<float*>(scatter_smem + 8 * kHotEntries + 8 * 32 * 4 * 4);   /* [kHotEntries][Dp] */
  float* slab_all = hot_acc + (size_t)kHotEntries * Dp;                                  /* [8 warps][32][Dp] */
  pdl_launch_dependents();                       /* the update kernel may start its prefetch now (it waits before reading acc) */
  if (skip_flag != nullptr && *skip_flag != 0) return;   /* DNN.java:58-63 early exit: nothing is pushed */
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int part = lane % TPL, grp = lane / TPL;
  const bool lane_on = part * 4 < D;
  float* slab = slab_all + (size_t)warp * 32 * Dp;
  uint32_t* maskw = maskw_all + warp * 32 * 4;
  for (int i = threadIdx.x; i < kHotEntries; i += 256) hot_slot[i] = -1;
  for (int i = threadIdx.x; i < kHotEntries * Dp; i += 256) hot_acc[i] = 0.f;
  pdl_wait();                                    /* (as a programmatic dependent of the last dgrad: delta is complete from here on) */
  __syncthreads();

  const long ntasks = (long)((N + 31) / 32) * F;
  for (long task = (long)blockIdx.x * 8 + warp; task < ntasks; task += (long)gridDim.x * 8) {
    const int sg = (int)(task / F), j = (int)(task - (long)sg * F);
    const long nbase = (long)sg * 32, n = nbase + lane;
    const bool in = n < N;
    /* the task's delta rows: addresses are plain arithmetic, so the copies go out first */
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int r = p * GPW + grp;
      if (nbase + r < N && lane_on) tb_cp_async16(tb_smem_u32(slab + (size_t)r * Dp + part * 4), delta + (size_t)(nbase + r) * ldd + j * D + part * 4);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const long t = (long)j * N + n;
    int slot = in ? lk_slot[t] : -1;
    if (in) {
#pragma unroll 4
      for (int w = 0; w < MW; ++w) maskw[lane * 4 + w] = lk_mask[(size_t)t * MW + w];
    }
    uint32_t cnt = 0u, row = 0u;
    if (slot >= 0) {
      const uint4 m = *reinterpret_cast<const uint4*>(&slots[slot]);
      cnt = m.z;
      if (raw_row) { row = m.w; if ((int)m.w < 0) slot = -1; }         /* requester side of the exchange: BatchSlot {key, cnt, bucket position | -1} */
      else row = (m.w & ~kRowReady) - 1u;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, slot >= 0 ? slot : (-1 - lane));
    const int leader = __ffs(peers) - 1;
    /* keys frequent in the batch are summed in the block's table and leave the block once */
    int e = -1;
    if (slot >= 0 && leader == lane && cnt >= hot_min) {
      uint32_t h = ((uint32_t)slot * 2654435761u) >> (32 - kHotBits);
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        const int old = atomicCAS(&hot_slot[h], -1, slot);
        if (old == -1) hot_row[h] = row;
        if (old == -1 || old == slot) { e = (int)h; break; }
        h = (h + 1u) & (uint32_t)(kHotEntries - 1);
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    /* (1) Relu.backward in place: dy *= (y > 0 ? 1 : 0)  (activations/Relu.java:14-19), the factor from the recorded mask bits */
#pragma unroll 4
    for (int p = 0; p < NP; ++p) {
      const int r = p * GPW + grp;
      if (nbase + r < N && lane_on) {
        float4 v = *reinterpret_cast<float4*>(slab + (size_t)r * Dp + part * 4);
        const uint32_t b = maskw[r * 4 + (part >> 3)] >> ((part & 7) * 4);
        v.x = __fmul_rn(v.x, (b & 1u) ? 1.f : 0.f); v.y = __fmul_rn(v.y, (b & 2u) ? 1.f : 0.f);
        v.z = __fmul_rn(v.z, (b & 4u) ? 1.f : 0.f); v.w = __fmul_rn(v.w, (b & 8u) ? 1.f : 0.f);
        *reinterpret_cast<float4*>(slab + (size_t)r * Dp + part * 4) = v;
      }
    }
    __syncwarp();
    /* (2) reduce-by-key inside the task: a row whose key another lane leads is added into the leader's row */
    if (__any_sync(0xffffffffu, slot >= 0 && leader != lane)) {
#pragma unroll 4
      for (int p = 0; p < NP; ++p) {
        const int r = p * GPW + grp;
        const int sr = __shfl_sync(0xffffffffu, slot, r), lr = __shfl_sync(0xffffffffu, leader, r);
        if (sr >= 0 && lr != r && lane_on) {
          const float4 v = *reinterpret_cast<const float4*>(slab + (size_t)r * Dp + part * 4);
          float* a = slab + (size_t)lr * Dp + part * 4;
          if (v.x != 0.f) atomicAdd(a + 0, v.x);
          if (v.y != 0.f) atomicAdd(a + 1, v.y);
          if (v.z != 0.f) atomicAdd(a + 2, v.z);
          if (v.w != 0.f) atomicAdd(a + 3, v.w);
        }
      }
      __syncwarp();
    }
    /* (3) leader rows leave the warp */
#pragma unroll 4
    for (int p = 0; p < NP; ++p) {
      const int r = p * GPW + grp;
      const int sr = __shfl_sync(0xffffffffu, slot, r), lr = __shfl_sync(0xffffffffu, leader, r);
      const int er = __shfl_sync(0xffffffffu, e, r);
      const uint32_t rr = __shfl_sync(0xffffffffu, row, r);
      if (sr < 0 || lr != r || !lane_on) continue;
      const float4 v = *reinterpret_cast<const float4*>(slab + (size_t)r * Dp + part * 4);
      if (er >= 0) {
        float* a = hot_acc + (size_t)er * Dp + part * 4;
        if (v.x != 0.f) atomicAdd(a + 0, v.x);
        if (v.y != 0.f) atomicAdd(a + 1, v.y);
        if (v.z != 0.f) atomicAdd(a + 2, v.z);
        if (v.w != 0.f) atomicAdd(a + 3, v.w);
      } else {
        red_add_f4(acc + (size_t)rr * Dp + part * 4, v);
      }
    }
    __syncwarp();                                /* the slab and the mask words are rewritten by the next task */
  }
  __syncthreads();
  const int CH = Dp >> 2;
  for (int i = threadIdx.x; i < kHotEntries * CH; i += 256) {
    const int he = i / CH, cc = (i - he * CH) * 4;
    if (hot_slot[he] < 0 || cc >= D) continue;
    const float4 v = *reinterpret_cast<const float4*>(hot_acc + (size_t)he * Dp + cc);
    if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) red_add_f4(acc + (size_t)hot_row[he] * Dp + cc, v);
  }
  if (pub != nullptr) p2p_publish_last(pub, CH_GRADS, gridDim.x);
}

/* KVStore.update + clear for the batch's unique keys.  A warp takes KPW = 32 / (Dp/4) consecutive entries of the unique
 * list per round, lane = (key, 16 B chunk); the chunk-0 lane of a key reads its batch count once and resets the slot's
 * per-batch fields when the key is done.  EXACT: see updaters.cuh.                                                   */
template <bool EXACT>
__global__ void __launch_bounds__(256) emb_update_kernel(EmbSlot* __restrict__ slots, float* __restrict__ rows, int rs, int Dp, int D,
                                                         const int32_t* __restrict__ uniq, float* __restrict__ acc, UpdaterDev upd, int calls,
                                                         const int* __restrict__ skip_flag, uint32_t* __restrict__ counters) {
  const int lane = threadIdx.x & 31;
  const int CH = Dp >> 2;
  const int KPW = 32 / CH;
  const int kl = lane / CH, ch = lane - kl * CH, cc = ch * 4;
  const bool lane_on = kl < KPW;
  const uint32_t U = counters[CNT_CURSOR];                   /* final since the lookup kernel ended; reset below by the last block */
  const bool skip = skip_flag != nullptr && *skip_flag != 0; /* written by the tail kernel, long before the scatter */
  const long wstride = (long)gridDim.x * 8;
  long wi = (long)blockIdx.x * 8 + (threadIdx.x >> 5);

  int slot = -1; uint32_t cnt = 0u;
  float4 wv, m1, m2;
  auto prefetch = [&](long w) {
    const long u = w * KPW + kl;
    slot = -1; cnt = 0u;
    wv = m1 = m2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane_on && u < (long)U) {
      slot = uniq[u];
      if (ch == 0) cnt = *reinterpret_cast<const volatile uint32_t*>(&slots[slot].cnt);
      if (!skip) {
        const float* r = rows + (size_t)slot * rs + cc;
        wv = ld_f4(r);
        if (upd.kind != PS_UPD_SIMPLE) { m1 = ld_f4(r + Dp); m2 = ld_f4(r + 2 * Dp); }
      }
    }
  };
  /* ---- the records of the first round are requested before the scatter kernel is known to be complete ---- */
  bool have = wi * KPW < (long)U;
  if (have) prefetch(wi);
  pdl_wait();
  while (have) {
    const long u = wi * KPW + kl;
    const bool on = slot >= 0;
    float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on && !skip) S = __ldcg(reinterpret_cast<const float4*>(acc + (size_t)u * Dp + cc));
    const int src0 = lane - ch;                               /* the key's chunk-0 lane */
    const uint32_t n_occ = __shfl_sync(0xffffffffu, cnt, src0);
    const float S0 = __shfl_sync(0xffffffffu, S.x, src0);
    const GeffScale gs = make_geff<EXACT>(n_occ > 0u ? n_occ : 1u, calls);
    bool do_upd = !skip;
    if (upd.kind == PS_UPD_FTRL) do_upd = do_upd && emb_geff<EXACT>(S0, gs) != 0.0f;   /* FtrlUpdater.java:52: `if (dw.get(0) == 0) return w` */
    if (on) {
      if (!skip) {
        float* r = rows + (size_t)slot * rs + cc;
        if (do_upd) {
          apply_elem<EXACT>(upd, wv.x, m1.x, m2.x, emb_geff<EXACT>(S.x, gs));
          apply_elem<EXACT>(upd, wv.y, m1.y, m2.y, emb_geff<EXACT>(S.y, gs));
          apply_elem<EXACT>(upd, wv.z, m1.z, m2.z, emb_geff<EXACT>(S.z, gs));
          apply_elem<EXACT>(upd, wv.w, m1.w, m2.w, emb_geff<EXACT>(S.w, gs));
          st_f4(r, wv);
          if (upd.kind != PS_UPD_SIMPLE) { st_f4(r + Dp, m1); st_f4(r + 2 * Dp, m2); }
        }
        st_f4(acc + (size_t)u * Dp + cc, make_float4(0.f, 0.f, 0.f, 0.f));
      }
      /* KVStore.clear (also after the early exit: the batch is forgotten): {cnt, uidx} = {0, ready} in one 8 B store */
      if (ch == 0) *reinterpret_cast<unsigned long long*>(&slots[slot].cnt) = (unsigned long long)kRowReady << 32;
    }
    wi += wstride;
    have = wi * KPW < (long)U;
    if (have) prefetch(wi);
  }
  if (last_block_done(&counters[CNT_TICKET_UPD], gridDim.x) && threadIdx.x == 0) counters[CNT_CURSOR] = 0u;
}

/* The same update with the TMA unit doing the memory work.  A warp takes kUpdKeys consecutive entries of the unique list per
 * round: lane k fetches key k's whole record {w | s1 | s2} (ONE 12*Dp-byte cp.async.bulk — contiguous, one DRAM page) and its
 * accumulator row into the warp's shared-memory slab, the 32 lanes run the updater on the slab, and lane k writes the record back
 * (and the zeroed accumulator row) with one bulk store each.  No register holds data in flight; a round moves 8 KB per warp at D = 64. */
static constexpr int kUpdKeys = 8;
__device__ __forceinline__ void tb_bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
/* PULL — the owner side of the sharded push (PServer.push in sync mode, PServer.java:164-195: the pushes of all workers are summed,
 * one update): there is no accumulator row.  The lookup linked every (requester, position) entry that asked for the key into the
 * key's chain (LookupArgs::chain); lane k walks key k's chain (<= R hops, local), then the warp reads the requesters' gradient
 * sums for the key straight out of THEIR slabs over NVLink — rank order, so the sum is the same on every run — and the occurrence
 * counts beside them.  The kernel waits for every requester's CH_GRADS flag once its first records are on the way.          */
static constexpr int kUpdSlabOff = 512 + 8 * kUpdKeys * kP2PMaxRanks * 4;
template <bool EXACT, bool PULL>
__global__ void __launch_bounds__(256) emb_update_slab_kernel(EmbSlot* __restrict__ slots, float* __restrict__ rows, int rs, int Dp, int D,
                                                              const int32_t* __restrict__ uniq, float* __restrict__ acc, UpdaterDev upd, int calls,
                                                              const int* __restrict__ skip_flag, uint32_t* __restrict__ counters,
                                                              const int32_t* __restrict__ chain, const P2PState* __restrict__ p2p) {
  extern __shared__ __align__(128) unsigned char update_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(update_smem);
  uint32_t* kcnt = reinterpret_cast<uint32_t*>(update_smem + 128) + warp * kUpdKeys;             /* occurrence count of the round's keys */
  int* kpos = reinterpret_cast<int*>(update_smem + 512) + warp * kUpdKeys * kP2PMaxRanks;        /* PULL: [key][requester] position in the requester's bucket, -1: none */
  float* slab = reinterpret_cast<float*>(update_smem + kUpdSlabOff) + (size_t)warp * (kUpdKeys * 4 + 1) * Dp;   /* [key][w | s1 | s2 | S], then one row of zeros */
  float* zero_row = slab + (size_t)kUpdKeys * 4 * Dp;
  const uint32_t bar = tb_smem_u32(&mbar[warp]);
  for (int i = lane; i < Dp; i += 32) zero_row[i] = 0.f;
  if (lane == 0) tb_mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int CH = Dp >> 2;
  const uint32_t U = counters[CNT_CURSOR];                   /* final since the lookup kernel ended; reset below by the last block */
  const bool skip = skip_flag != nullptr && *skip_flag != 0; /* written by the tail kernel, long before the scatter */
  const long rounds = ((long)U + kUpdKeys - 1) / kUpdKeys;
  const long wstride = (long)gridDim.x * 8;
  long wi = (long)blockIdx.x * 8 + warp;
  uint32_t parity = 0u;
  const uint32_t rec_bytes = 12u * (uint32_t)Dp, row_bytes = 4u * (uint32_t)Dp;

  int slot = -1; uint32_t cnt = 0u;
  auto fetch_records = [&](long w) {               /* lane k: key k of round w — its slot, its batch count, its record on the way */
    const long u = w * kUpdKeys + lane;
    slot = -1; cnt = 0u;
    if (lane < kUpdKeys && u < (long)U) {
      slot = uniq[u];
      cnt = *reinterpret_cast<const volatile uint32_t*>(&slots[slot].cnt);
    }
    const unsigned vm = __ballot_sync(0xffffffffu, slot >= 0);
    if (!skip && vm != 0u) {
      if (lane == 0) tb_mbar_expect_tx(bar, (uint32_t)__popc(vm) * (PULL ? rec_bytes : rec_bytes + row_bytes));
      __syncwarp();
      if (slot >= 0) tb_bulk_g2s(tb_smem_u32(slab + (size_t)lane * 4 * Dp), rows + (size_t)slot * rs, rec_bytes, bar);
    }
  };
  bool have = wi < rounds;
  if (have) fetch_records(wi);                     /* before the scatter kernel is known to be complete */
  /* the step this rank last published its own sums for — when the update is deferred to the head of the next step, that step's
   * route_send may already have moved seq on */
  const uint32_t pseq = PULL ? p2p->pub_seq[CH_GRADS] : 0u;
  if (PULL) p2p_wait_all_seq(p2p, CH_GRADS, pseq);   /* every requester's sums (and counts) of that step are final */
  else pdl_wait();
  while (have) {
    const long u = wi * kUpdKeys + lane;
    if (!skip) {
      const int nk = __popc(__ballot_sync(0xffffffffu, slot >= 0));
      if (PULL) {
        const int R = p2p->R, cap = p2p->cap, me = p2p->me;
        if (lane < kUpdKeys) {                     /* `cnt` is the head of key `lane`'s chain: 1 + entry, entry = requester * cap + position */
          int* sp = kpos + lane * kP2PMaxRanks;
#pragma unroll
          for (int r = 0; r < kP2PMaxRanks; ++r) sp[r] = -1;
          uint32_t e = cnt;
          for (int hop = 0; e != 0u && hop < kP2PMaxRanks; ++hop) {
            const int entry = (int)e - 1, src = entry / cap;
            sp[src] = entry - src * cap;
            e = (uint32_t)chain[entry];
          }
        }
        __syncwarp();
        for (int i = lane; i < nk * CH; i += 32) {
          const int k = i / CH, cc = (i - k * CH) * 4;
          float4 v[kP2PMaxRanks];
#pragma unroll
          for (int r = 0; r < kP2PMaxRanks; ++r) {   /* every requester's load is issued before the first is consumed */
            const int pos = r < R ? kpos[k * kP2PMaxRanks + r] : -1;
            v[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pos >= 0) v[r] = p2p_ld_sys_f4(reinterpret_cast<const float*>(p2p_region_of(p2p, r, p2p->off_grads, pseq)) + ((size_t)me * cap + pos) * Dp + cc);
          }
          float4 S = v[0];
#pragma unroll
          for (int r = 1; r < kP2PMaxRanks; ++r) { S.x += v[r].x; S.y += v[r].y; S.z += v[r].z; S.w += v[r].w; }
          *reinterpret_cast<float4*>(slab + (size_t)k * 4 * Dp + 3 * Dp + cc) = S;
        }
        if (lane < kUpdKeys) {
          uint32_t n = 0u;
          if (slot >= 0)
            for (int r = 0; r < R; ++r) {
              const int pos = kpos[lane * kP2PMaxRanks + r];
              if (pos >= 0) n += p2p_ld_sys_u32(reinterpret_cast<const uint32_t*>(p2p_region_of(p2p, r, p2p->off_gcnt, pseq)) + (size_t)me * cap + pos);
            }
          kcnt[lane] = n;
        }
      } else {
        if (slot >= 0) tb_bulk_g2s(tb_smem_u32(slab + (size_t)lane * 4 * Dp + 3 * Dp), acc + (size_t)u * Dp, row_bytes, bar);
        if (lane < kUpdKeys) kcnt[lane] = cnt;
      }
      if (nk != 0) { tb_mbar_wait(bar, parity); parity ^= 1u; }
      __syncwarp();
      for (int i = lane; i < nk * CH; i += 32) {
        const int k = i / CH, cc = (i - k * CH) * 4;
        float* r = slab + (size_t)k * 4 * Dp;
        const uint32_t n_occ = kcnt[k];
        const GeffScale gs = make_geff<EXACT>(n_occ > 0u ? n_occ : 1u, calls);
        float4 wv = *reinterpret_cast<float4*>(r + cc), m1 = *reinterpret_cast<float4*>(r + Dp + cc), m2 = *reinterpret_cast<float4*>(r + 2 * Dp + cc);
        const float4 S = *reinterpret_cast<float4*>(r + 3 * Dp + cc);
        bool do_upd = true;
        if (upd.kind == PS_UPD_FTRL) do_upd = emb_geff<EXACT>(r[3 * Dp], gs) != 0.0f;   /* FtrlUpdater.java:52: `if (dw.get(0) == 0) return w` */
        if (do_upd) {
          apply_elem<EXACT>(upd, wv.x, m1.x, m2.x, emb_geff<EXACT>(S.x, gs));
          apply_elem<EXACT>(upd, wv.y, m1.y, m2.y, emb_geff<EXACT>(S.y, gs));
          apply_elem<EXACT>(upd, wv.z, m1.z, m2.z, emb_geff<EXACT>(S.z, gs));
          apply_elem<EXACT>(upd, wv.w, m1.w, m2.w, emb_geff<EXACT>(S.w, gs));
        }
        *reinterpret_cast<float4*>(r + cc) = wv;    /* S stays as it is (another lane may still need S[0]): the accumulator is zeroed from zero_row */
        if (upd.kind != PS_UPD_SIMPLE) { *reinterpret_cast<float4*>(r + Dp + cc) = m1; *reinterpret_cast<float4*>(r + 2 * Dp + cc) = m2; }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   /* the slab's new contents are visible to the TMA unit */
      __syncwarp();
      if (slot >= 0) {
        tb_bulk_s2g(rows + (size_t)slot * rs, tb_smem_u32(slab + (size_t)lane * 4 * Dp), rec_bytes);
        if (!PULL) tb_bulk_s2g(acc + (size_t)u * Dp, tb_smem_u32(zero_row), row_bytes);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    /* KVStore.clear (also after the early exit: the batch is forgotten): {cnt, uidx} = {0, ready} in one 8 B store */
    if (slot >= 0) *reinterpret_cast<unsigned long long*>(&slots[slot].cnt) = (unsigned long long)kRowReady << 32;
    if (!skip) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   /* the slab may be refilled */
    __syncwarp();
    wi += wstride;
    have = wi < rounds;
    if (have) fetch_records(wi);
  }
  if (!skip) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (last_block_done(&counters[CNT_TICKET_UPD], gridDim.x) && threadIdx.x == 0) counters[CNT_CURSOR] = 0u;
}

/* forget the batch: the slots of its unique list get {cnt, uidx} = {0, ready}; the cursor restarts */
__global__ void __launch_bounds__(256) emb_clear_batch_kernel(EmbSlot* __restrict__ slots, const int32_t* __restrict__ uniq, uint32_t* __restrict__ counters) {
  const uint32_t U = counters[CNT_CURSOR];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < U; i += gridDim.x * blockDim.x)
    *reinterpret_cast<unsigned long long*>(&slots[uniq[i]].cnt) = (unsigned long long)kRowReady << 32;
  if (last_block_done(&counters[CNT_TICKET_UPD], gridDim.x) && threadIdx.x == 0) counters[CNT_CURSOR] = 0u;
}

/* host-driven row access: thread per key */
__global__ void emb_get_rows_kernel(const EmbSlot* __restrict__ slots, uint32_t C, const float* __restrict__ rows, int rs, int Dp, int D,
                                    const int32_t* __restrict__ fields, const int64_t* __restrict__ ids, int n, float* __restrict__ wo,
                                    float* __restrict__ s1o, float* __restrict__ s2o, int32_t* __restrict__ found) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int slot = emb_find(slots, C, ps_pack_key((uint32_t)fields[i], (uint64_t)ids[i]));
  found[i] = slot >= 0;
  for (int d = 0; d < D; ++d) {
    const size_t o = (size_t)(slot < 0 ? 0 : slot) * rs + d;
    wo[(size_t)i * D + d] = slot < 0 ? 0.f : rows[o];
    if (s1o) s1o[(size_t)i * D + d] = slot < 0 ? 0.f : rows[o + Dp];
    if (s2o) s2o[(size_t)i * D + d] = slot < 0 ? 0.f : rows[o + 2 * Dp];
  }
}

/* PServer.push → KVStore.update(updater, key) (PServer.java:164-184, KVStore.java:202-208): one updater step on an existing row with
 * the gradient a (legacy gRPC) worker pushed; thread per key, the exact updater forms; found[i] = 0 and nothing happens for a missing key */
__global__ void emb_push_rows_kernel(const EmbSlot* __restrict__ slots, uint32_t C, float* __restrict__ rows, int rs, int Dp, int D,
                                     const int32_t* __restrict__ fields, const int64_t* __restrict__ ids, int n, const float* __restrict__ g,
                                     UpdaterDev upd, int32_t* __restrict__ found) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int slot = emb_find(slots, C, ps_pack_key((uint32_t)fields[i], (uint64_t)ids[i]));
  found[i] = slot >= 0;
  if (slot < 0) return;
  if (upd.kind == PS_UPD_FTRL && g[(size_t)i * D] == 0.0f) return;        /* FtrlUpdater.java:52 */
  float* row = rows + (size_t)slot * rs;
  for (int d = 0; d < D; ++d) apply_elem<true>(upd, row[d], row[Dp + d], row[2 * Dp + d], g[(size_t)i * D + d]);
}

/* KVStore.put (replace) / PServer.upsertList with replace=false (net/PServer.java:144-162):
 * insert-if-absent; the caller's buffer receives the winning row.                              */
__global__ void emb_put_rows_kernel(EmbSlot* __restrict__ slots, uint32_t C, float* __restrict__ rows, int rs, int D,
                                    const int32_t* __restrict__ fields, const int64_t* __restrict__ ids, int n, float* __restrict__ wio,
                                    int replace, uint32_t* __restrict__ counters) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool inserted;
  const int slot = emb_find_or_insert(slots, C, ps_pack_key((uint32_t)fields[i], (uint64_t)ids[i]), &inserted);
  if (slot < 0) { atomicOr(&counters[CNT_ERR], 1u); return; }
  if (inserted) atomicAdd(reinterpret_cast<unsigned long long*>(counters + CNT_ROWS), 1ull);
  float* row = rows + (size_t)slot * rs;
  if (inserted || replace) { for (int d = 0; d < D; ++d) row[d] = wio[(size_t)i * D + d]; }
  else { for (int d = 0; d < D; ++d) wio[(size_t)i * D + d] = row[d]; }
  if (inserted) { __threadfence(); atomicOr(&slots[slot].uidx, kRowReady); }
}

/* ------------------------------------------------------------------ EmbTable host side */
static int pow2_ge(int x) { int p = 1; while (p < x) p <<= 1; return p; }
static void launch_update(EmbTable& t, long L, int calls, const int* skip, const P2PState* pull);

/* the slab of every warp (32 rows) + the mbarriers: 64 KB at D = 64 — more than the default 48 KB, so every gathering
 * instantiation opts in (once per table, outside any stream capture) */
template <class IdT, int TPL>
static void lookup_allow_smem(size_t smem) {
  PS_CUDA(cudaFuncSetAttribute(emb_lookup_kernel<IdT, true, TPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PS_CUDA(cudaFuncSetAttribute(emb_lookup_kernel<IdT, true, TPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
}
template <int TPL>
static int lookup_occupancy(size_t smem) {
  lookup_allow_smem<int64_t, TPL>(smem); lookup_allow_smem<float, TPL>(smem); lookup_allow_smem<unsigned long long, TPL>(smem);
  int occ = 1;
  PS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, emb_lookup_kernel<int64_t, true, TPL, true>, 256, smem));
  return occ;
}
static void query_lookup_occupancy(EmbTable& t) {      /* not inside a capture: called from create() */
  const size_t smem = lookup_smem_bytes(t.Dp);
  switch (t.tpl) {
    case 1: t.lookup_occ = lookup_occupancy<1>(smem); break;
    case 2: t.lookup_occ = lookup_occupancy<2>(smem); break;
    case 4: t.lookup_occ = lookup_occupancy<4>(smem); break;
    case 8: t.lookup_occ = lookup_occupancy<8>(smem); break;
    case 16: t.lookup_occ = lookup_occupancy<16>(smem); break;
    default: t.lookup_occ = lookup_occupancy<32>(smem); break;
  }
}

void EmbTable::create(Ctx* c, int F_, int D_, int64_t capacity, const ps_updater_spec& u, int64_t max_lookups) {
  PS_REQUIRE(F_ > 0 && D_ > 0 && D_ <= 128, PS_ERR_ARG, "embedding: need F > 0 and 0 < D <= 128");
  PS_REQUIRE(capacity > 0 && capacity < (1ll << 31), PS_ERR_ARG, "embedding: capacity must be in (0, 2^31)");
  ctx = c; F = F_; D = D_; Dp = round_up(D_, 4); tpl = pow2_ge(Dp / 4); C = capacity;
  rs = 3 * Dp; MW = std::max(1, tpl / 8);
  maxv = (float)(4 * (std::sqrt(6.0) / std::sqrt((double)(1 + D_))));   /* EmbeddingField.java:40 with in=1,out=D (EmbeddingLayer.java:52) */
  upd = make_updater_dev(u);
  slots = dmalloc_zero<EmbSlot>((size_t)C, ctx->stream);
  rows = dmalloc_zero<float>((size_t)C * rs, ctx->stream);
  w = rows; s1 = rows + Dp; s2 = rows + 2 * Dp;
  counters = dmalloc_zero<uint32_t>(CNT_WORDS, ctx->stream);
  reserve(max_lookups > 0 ? max_lookups : 1);
  scatter_update(nullptr, 0, nullptr, 0, 0, 2, nullptr);   /* fills scatter_occ (sizes the scatter's persistent grid) */
  launch_update(*this, 0, 2, nullptr, nullptr);             /* opt-in shared memory + occupancy of the staged update kernel */
  query_lookup_occupancy(*this);
}

void EmbTable::reserve(int64_t L) {
  if (L <= Lcap) return;
  if (last_L > 0) clear_batch();                /* a forward that was never followed by backward still owns counts in the OLD workspace's list */
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  dfree(lk_slot); dfree(lk_mask); dfree(uniq); dfree(acc); dfree(chain);
  Lcap = L; ++generation;                      /* captured graphs that hold the old workspace pointers are stale now */
  lk_slot = dmalloc<int32_t>((size_t)L);
  lk_mask = dmalloc<uint32_t>((size_t)L * MW);
  uniq = dmalloc<int32_t>((size_t)L);
  acc = dmalloc_zero<float>((size_t)L * Dp, ctx->stream);   /* one accumulator row per unique key of a batch (<= L) */
  chain = dmalloc_zero<int32_t>((size_t)L, ctx->stream);
}

void EmbTable::destroy() {
  dfree(slots); dfree(rows); dfree(counters);
  dfree(lk_slot); dfree(lk_mask); dfree(uniq); dfree(acc); dfree(chain);
  chain = nullptr; slots = nullptr; rows = w = s1 = s2 = nullptr; lk_slot = nullptr; lk_mask = nullptr; uniq = nullptr; acc = nullptr; counters = nullptr;
}

template <class IdT, int TPL>
static void launch_lookup_t(EmbTable& t, LookupArgs& a, bool gather, bool aligned, int grid, size_t smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = gather ? smem : 0; cfg.stream = t.ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (a.p2p != nullptr && t.ctx->pdl && t.ctx->pdl_exchange) ? 1 : 0;   /* the exchange's lookups order themselves by flags */
  if (!gather) PS_CUDA(cudaLaunchKernelEx(&cfg, emb_lookup_kernel<IdT, false, 1, true>, a));
  else if (aligned) PS_CUDA(cudaLaunchKernelEx(&cfg, emb_lookup_kernel<IdT, true, TPL, true>, a));
  else PS_CUDA(cudaLaunchKernelEx(&cfg, emb_lookup_kernel<IdT, true, TPL, false>, a));
}

template <class IdT>
static void launch_lookup(EmbTable& t, LookupArgs& a, bool gather) {
  const long ntasks = (long)((a.N + 31) / 32) * (a.F > 0 ? a.F : 1);
  /* persistent: one wave of blocks (their warps stride over the tasks three deep, see the kernel); small batches get a warp per task */
  const int resident = t.ctx->num_sms * std::max(1, gather ? t.lookup_occ : 8);
  a.task_blocks = (int)std::min<long>(ceil_div(ntasks, 8), resident);
  const int xblocks = (gather && a.X != nullptr) ? ceil_div((long)a.N * a.Xn, 1024) : 0;
  a.hot_tma = (gather && t.ctx->hot_tma) ? 1 : 0;
  a.hot_share = t.ctx->hot_share;
  const bool aligned = a.F == 0 || ((t.D % 4 == 0) && (a.ldo % 4 == 0) && ((uintptr_t)a.out % 16 == 0));
  const size_t smem = gather ? lookup_smem_bytes(t.Dp) : 0;
  const int grid = a.task_blocks + xblocks;
  switch (t.tpl) {
    case 1: launch_lookup_t<IdT, 1>(t, a, gather, aligned, grid, smem); break;
    case 2: launch_lookup_t<IdT, 2>(t, a, gather, aligned, grid, smem); break;
    case 4: launch_lookup_t<IdT, 4>(t, a, gather, aligned, grid, smem); break;
    case 8: launch_lookup_t<IdT, 8>(t, a, gather, aligned, grid, smem); break;
    case 16: launch_lookup_t<IdT, 16>(t, a, gather, aligned, grid, smem); break;
    default: launch_lookup_t<IdT, 32>(t, a, gather, aligned, grid, smem); break;
  }
  PS_LAUNCH_CHECK();
  t.ctx->launches++;
}

static LookupArgs base_args(EmbTable& t) {
  LookupArgs a{};
  a.slots = t.slots; a.C = (uint32_t)t.C; a.rows = t.rows; a.rs = t.rs; a.Dp = t.Dp; a.D = t.D;
  a.seed = t.ctx->seed; a.maxv = t.maxv;
  a.lk_slot = t.lk_slot; a.lk_mask = nullptr; a.MW = t.MW; a.uniq = t.uniq; a.counters = t.counters;
  return a;
}

void EmbTable::lookup(const int64_t* ids_i64, const float* ids_f32, int N, float* out, int ldo, const float* X, int Xn, int xoff) {
  const int64_t L = (int64_t)N * F;
  PS_REQUIRE(L <= Lcap, PS_ERR_ARG, "embedding: batch larger than the reserved workspace");
  if (last_L > 0) clear_batch();               /* a forward that was never followed by backward: drop its counts */
  last_L = L;
  LookupArgs a = base_args(*this);
  a.N = N; a.F = F; a.out = out; a.ldo = ldo; a.X = out ? X : nullptr; a.Xn = Xn; a.xoff = xoff;
  a.lk_mask = out ? lk_mask : nullptr;
  if (ids_i64) { a.ids = ids_i64; launch_lookup<int64_t>(*this, a, out != nullptr); }
  else { a.ids = ids_f32; launch_lookup<float>(*this, a, out != nullptr); }
}

void EmbTable::lookup_packed(const uint64_t* keys, int n, float* out, P2PState* p2p, bool send_rows) {
  if (last_L > 0) clear_batch();
  reserve(n);
  last_L = n;
  if (n <= 0) return;
  LookupArgs a = base_args(*this);
  a.N = n; a.F = 0; a.ids = keys; a.out = out; a.ldo = Dp; a.p2p = p2p; a.send_rows = send_rows ? 1 : 0;
  a.chain = (p2p != nullptr && send_rows) ? chain : nullptr;    /* the owner's update will pull the requesters' sums along the chains */
  launch_lookup<unsigned long long>(*this, a, out != nullptr || send_rows);
}

/* the scatter launch alone.  recs / lk / accp: the table's own slot records, lk_slot and accumulator rows — or, on the requester
 * side of the sharded exchange, the per-batch de-duplication table (same 16 B record shape), its lookup index and the local
 * per-key gradient sums (raw_row = 1: the record's last word is the row index itself).                                    */
struct ScatterJob {
  const EmbSlot* recs; const int32_t* lk; const uint32_t* mask; float* accp;
  const float* delta; int ldd; const float* act; int lda; int N, F; const int* skip; int raw_row;
  P2PState* pub;                                 /* requester side of the sharded push: sums into this rank's slab, CH_GRADS published by the last block */
};
void EmbTable::gather_resolved(const void* batch_slots, const int32_t* lk_batch, P2PState* p2p, int N, float* out, int ldo, const float* X, int Xn, int xoff) {
  PS_REQUIRE((int64_t)N * F <= Lcap, PS_ERR_ARG, "embedding: batch larger than the reserved workspace");
  LookupArgs a = base_args(*this);
  a.N = N; a.F = F; a.out = out; a.ldo = ldo; a.X = X; a.Xn = Xn; a.xoff = xoff;
  a.lk_mask = lk_mask; a.p2p = p2p; a.pre_recs = reinterpret_cast<const EmbSlot*>(batch_slots); a.pre_lk = lk_batch; a.ids = nullptr;
  launch_lookup<int64_t>(*this, a, true);
}

template <int TPL, int CPL>
static void launch_scatter(EmbTable& t, const ScatterJob& j) {
  constexpr int PASSES = TPL >= 4 ? 4 : TPL;     /* warp tasks in flight per warp */
  constexpr int GPW = 32 / TPL;
  const bool aligned = (t.D % 4 == 0) && (j.ldd % 4 == 0) && (j.lda % 4 == 0) && ((uintptr_t)j.delta % 16 == 0) && ((uintptr_t)j.act % 16 == 0);
  if (j.N == 0) {                                /* EmbTable::create: resident blocks per SM of the two instantiations (not inside a capture) */
    PS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t.scatter_occ[1], emb_scatter_kernel<TPL, CPL, PASSES, true>, 256, 0));
    PS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t.scatter_occ[0], emb_scatter_kernel<TPL, CPL, PASSES, false>, 256, 0));
    return;
  }
  const int N = j.N;
  const int resident = t.ctx->num_sms * std::max(1, t.scatter_occ[aligned ? 1 : 0]);
  /* samples per tile: small enough that there are >= 4 tiles per resident block (balance), at most 32 */
  int SB = GPW;
  while (SB * 2 <= 32 && ceil_div(N, SB * 2) >= 4 * resident) SB *= 2;
  const int grid = std::min(ceil_div(N, SB), resident);
  /* a key is pre-summed per block when it is frequent enough to recur among the samples one block sees */
  const long per_block = (long)SB * ceil_div(ceil_div(N, SB), grid);
  const uint32_t hot_min = t.ctx->hot_min == 0xFFFFFFFFu ? 0xFFFFFFFFu : std::max<uint32_t>(t.ctx->hot_min, (uint32_t)(2L * N / per_block));
  const bool dep = t.ctx->pdl_exchange != 0;
  if (aligned)
    launch_dep(t.ctx, dep, emb_scatter_kernel<TPL, CPL, PASSES, true>, dim3(grid), dim3(256), 0, j.recs, t.Dp, t.D, j.lk, j.mask, t.MW, N, j.F, SB, j.delta, j.ldd, j.act, j.lda, j.accp, j.skip, j.raw_row, hot_min, j.pub);
  else
    launch_dep(t.ctx, dep, emb_scatter_kernel<TPL, CPL, PASSES, false>, dim3(grid), dim3(256), 0, j.recs, t.Dp, t.D, j.lk, j.mask, t.MW, N, j.F, SB, j.delta, j.ldd, j.act, j.lda, j.accp, j.skip, j.raw_row, hot_min, j.pub);
  PS_LAUNCH_CHECK();
  t.ctx->launches++;
}

static size_t scatter_slab_smem(int Dp) { return 8 * kHotEntries + 8 * 32 * 4 * 4 + (size_t)kHotEntries * Dp * 4 + (size_t)8 * 32 * Dp * 4; }
template <int TPL>
static void launch_scatter_slab(EmbTable& t, const ScatterJob& j) {
  const size_t smem = scatter_slab_smem(t.Dp);
  if (j.N == 0) {                                /* EmbTable::create (not inside a capture): opt in to the slab's shared memory, size the persistent grid */
    PS_CUDA(cudaFuncSetAttribute(emb_scatter_slab_kernel<TPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t.scatter_slab_occ, emb_scatter_slab_kernel<TPL>, 256, smem));
    return;
  }
  const long ntasks = (long)((j.N + 31) / 32) * j.F;
  const int grid = (int)std::max<long>(1, std::min<long>(ceil_div(ntasks, 8), (long)t.ctx->num_sms * std::max(1, t.scatter_slab_occ)));
  const long per_block = 32L * ceil_div(ntasks, (long)grid * 8) * 8 / std::max(1, j.F) + 32;      /* samples of one field a block sees */
  const uint32_t hot_min = t.ctx->hot_min == 0xFFFFFFFFu ? 0xFFFFFFFFu : std::max<uint32_t>(t.ctx->hot_min, (uint32_t)(2L * j.N / per_block));
  launch_dep(t.ctx, t.ctx->pdl_exchange != 0, emb_scatter_slab_kernel<TPL>, dim3(grid), dim3(256), smem, j.recs, t.Dp, t.D, j.lk, j.mask, t.MW, j.N, j.F, j.delta, j.ldd, j.accp, j.skip, j.raw_row, hot_min, j.pub);
  PS_LAUNCH_CHECK();
  t.ctx->launches++;
}

static void dispatch_scatter(EmbTable& t, const ScatterJob& j) {
  /* the staged form: 16 B aligned rows, the lookup's mask bits, the table's own records */
  const bool slab_ok = t.ctx->scatter_slab && t.D % 4 == 0 && (j.N == 0 || (j.mask != nullptr && j.act == nullptr && j.ldd % 4 == 0 && (uintptr_t)j.delta % 16 == 0));
  if (slab_ok) {
    switch (t.tpl) {
      case 1: launch_scatter_slab<1>(t, j); break;
      case 2: launch_scatter_slab<2>(t, j); break;
      case 4: launch_scatter_slab<4>(t, j); break;
      case 8: launch_scatter_slab<8>(t, j); break;
      case 16: launch_scatter_slab<16>(t, j); break;
      default: launch_scatter_slab<32>(t, j); break;
    }
    if (j.N != 0) return;                        /* N == 0: also query the general kernel below */
  }
  if (t.Dp % 8 == 0) {                          /* two 16 B chunks per lane: half the threads, twice the bytes in flight per thread */
    switch (pow2_ge(t.Dp / 8)) {
      case 1: launch_scatter<1, 2>(t, j); break;
      case 2: launch_scatter<2, 2>(t, j); break;
      case 4: launch_scatter<4, 2>(t, j); break;
      case 8: launch_scatter<8, 2>(t, j); break;
      default: launch_scatter<16, 2>(t, j); break;
    }
  } else {
    switch (t.tpl) {
      case 1: launch_scatter<1, 1>(t, j); break;
      case 2: launch_scatter<2, 1>(t, j); break;
      case 4: launch_scatter<4, 1>(t, j); break;
      case 8: launch_scatter<8, 1>(t, j); break;
      case 16: launch_scatter<16, 1>(t, j); break;
      default: launch_scatter<32, 1>(t, j); break;
    }
  }
}

/* the update walks the unique list (<= L entries, how many is only known on the device): enough warps for one round at the
 * typical unique fraction, a grid-stride loop beyond; a programmatic dependent of the scatter launched just before it */
static size_t update_slab_smem(int Dp) { return kUpdSlabOff + (size_t)8 * (kUpdKeys * 4 + 1) * Dp * sizeof(float); }
static void launch_update(EmbTable& t, long L, int calls, const int* skip, const P2PState* pull) {
  if (t.ctx->update_slab || pull != nullptr) {   /* the owner side of the sharded push exists in the staged form only */
    const size_t smem = update_slab_smem(t.Dp);
    if (L == 0) {                                /* EmbTable::create (not inside a capture) */
      PS_CUDA(cudaFuncSetAttribute(emb_update_slab_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      PS_CUDA(cudaFuncSetAttribute(emb_update_slab_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      PS_CUDA(cudaFuncSetAttribute(emb_update_slab_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      PS_CUDA(cudaFuncSetAttribute(emb_update_slab_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      PS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t.update_slab_occ, emb_update_slab_kernel<false, false>, 256, smem));
      PS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t.update_pull_occ, emb_update_slab_kernel<false, true>, 256, smem));
      return;
    }
    if (pull != nullptr) {                       /* an ordinary launch: the kernel's own flag wait orders it behind the requesters */
      const long rounds = ceil_div(L, kUpdKeys);
      const int ugrid = (int)std::max<long>(1, std::min<long>(ceil_div(rounds, 8), (long)t.ctx->num_sms * std::max(1, t.update_pull_occ)));
      /* (optionally a programmatic dependent of the scatter: it may start while that drains — it reads nothing of it before the flags) */
      const bool dep = t.ctx->pdl_exchange != 0;
      if (t.ctx->exact_updaters)
        launch_dep(t.ctx, dep, emb_update_slab_kernel<true, true>, dim3(ugrid), dim3(256), smem, t.slots, t.rows, t.rs, t.Dp, t.D, (const int32_t*)t.uniq, (float*)nullptr, t.upd, calls, skip, t.counters, (const int32_t*)t.chain, pull);
      else
        launch_dep(t.ctx, dep, emb_update_slab_kernel<false, true>, dim3(ugrid), dim3(256), smem, t.slots, t.rows, t.rs, t.Dp, t.D, (const int32_t*)t.uniq, (float*)nullptr, t.upd, calls, skip, t.counters, (const int32_t*)t.chain, pull);
      PS_LAUNCH_CHECK();
      t.ctx->launches++;
      return;
    }
    const long rounds = ceil_div(L, kUpdKeys);
    const int ugrid = (int)std::max<long>(1, std::min<long>(ceil_div(rounds, 8), (long)t.ctx->num_sms * std::max(1, t.update_slab_occ)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ugrid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = t.ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = t.ctx->pdl ? 1 : 0;
    if (t.ctx->exact_updaters)
      PS_CUDA(cudaLaunchKernelEx(&cfg, emb_update_slab_kernel<true, false>, t.slots, t.rows, t.rs, t.Dp, t.D, (const int32_t*)t.uniq, t.acc, t.upd, calls, skip, t.counters, (const int32_t*)nullptr, (const P2PState*)nullptr));
    else
      PS_CUDA(cudaLaunchKernelEx(&cfg, emb_update_slab_kernel<false, false>, t.slots, t.rows, t.rs, t.Dp, t.D, (const int32_t*)t.uniq, t.acc, t.upd, calls, skip, t.counters, (const int32_t*)nullptr, (const P2PState*)nullptr));
    t.ctx->launches++;
    return;
  }
  if (L == 0) return;
  const int KPW = 32 / (t.Dp / 4);
  const int ugrid = (int)std::max<long>(1, std::min<long>(ceil_div(ceil_div(L, KPW), 8), (long)t.ctx->num_sms * 16));
  if (t.ctx->exact_updaters)
    launch_pdl(t.ctx, emb_update_kernel<true>, dim3(ugrid), dim3(256), t.slots, t.rows, t.rs, t.Dp, t.D, (const int32_t*)t.uniq, t.acc, t.upd, calls, skip, t.counters);
  else
    launch_pdl(t.ctx, emb_update_kernel<false>, dim3(ugrid), dim3(256), t.slots, t.rows, t.rs, t.Dp, t.D, (const int32_t*)t.uniq, t.acc, t.upd, calls, skip, t.counters);
  t.ctx->launches++;
}

/* requester side of the push: per-lookup row gradients (ReLU mask from `act`) summed per unique key of THIS rank's batch into
 * the gsums region of this rank's slab (row = bucket position), with the same three levels of pre-summation as the local
 * backward — a hot key's thousands of occurrences leave a block once instead of serialising on one L2 line.  The kernel's last
 * block flags the owners (CH_GRADS): they read the sums from here, nothing is sent                                        */
void EmbTable::scatter_rows(const void* batch_slots, const int32_t* lk_batch, P2PState* p2p, const float* delta, int ldd, const float* act, int lda, int N) {
  /* act == null: the mask bits gather_resolved recorded for this batch */
  ScatterJob j{reinterpret_cast<const EmbSlot*>(batch_slots), lk_batch, act ? nullptr : lk_mask, nullptr, delta, ldd, act, lda, N, F, nullptr, 1, p2p};
  dispatch_scatter(*this, j);
}

/* owner side of the push + psUpdate: see emb_update_slab_kernel<.., PULL> */
void EmbTable::update_pull(const P2PState* p2p, int n, int calls, const int* skip_flag) {
  PS_REQUIRE((int64_t)n == last_L, PS_ERR_STATE, "embedding: push without a matching lookup");
  PS_REQUIRE(calls == 1 || calls == 2, PS_ERR_ARG, "embedding: backward calls must be 1 or 2");
  launch_update(*this, n, calls, skip_flag, p2p);
  last_L = 0;
}

void EmbTable::scatter_update(const float* delta, int ldd, const float* act, int lda, int N, int calls, const int* skip_flag, int F_eff, bool use_mask) {
  const int Fe = F_eff > 0 ? F_eff : F;
  if (N > 0) {                                  /* N == 0: occupancy query from create() */
    PS_REQUIRE((int64_t)N * Fe == last_L, PS_ERR_STATE, "embedding: backward without a matching forward");
    PS_REQUIRE(calls == 1 || calls == 2, PS_ERR_ARG, "embedding: backward calls must be 1 or 2");
  }
  ScatterJob j{slots, lk_slot, use_mask ? lk_mask : nullptr, acc, delta, ldd, use_mask ? nullptr : act, lda, N, Fe, skip_flag, 0, nullptr};
  dispatch_scatter(*this, j);
  if (N == 0) return;
  launch_update(*this, (long)N * Fe, calls, skip_flag, nullptr);
  last_L = 0;
}

void EmbTable::clear_batch() {
  if (last_L <= 0) return;
  emb_clear_batch_kernel<<<std::min<long>(ceil_div(last_L, 256), (long)ctx->num_sms * 4), 256, 0, ctx->stream>>>(slots, uniq, counters);
  PS_LAUNCH_CHECK();
  ctx->launches++;
  last_L = 0;
}

void EmbTable::check_errors() {
  uint32_t h[4];
  PS_CUDA(cudaMemcpyAsync(h, counters, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  PS_REQUIRE((h[CNT_ERR] & 2u) == 0, PS_ERR_ARG, "embedding id outside [0, 2^44): the batch was refused");
  PS_REQUIRE(h[CNT_ERR] == 0, PS_ERR_CAPACITY, "embedding table is full: raise capacity");
}

int64_t EmbTable::size() {
  uint32_t h[4];
  PS_CUDA(cudaMemcpyAsync(h, counters, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  return (int64_t)(((uint64_t)h[3] << 32) | h[2]);
}

void EmbTable::get_rows(const int32_t* fields, const int64_t* ids, int n, float* w_out, float* s1_out, float* s2_out, int32_t* found) {
  if (n <= 0) return;
  for (int i = 0; i < n; ++i)                    /* a field outside [0, F) would alias another namespace (or the EMPTY sentinel) */
    PS_REQUIRE(fields[i] >= 0 && fields[i] < F && ids[i] >= 0 && ids[i] <= (int64_t)PS_KEY_ID_MASK, PS_ERR_ARG, "embedding: key outside the (field, id) domain");
  cudaStream_t st = ctx->stream;
  int32_t* d_f = dmalloc<int32_t>(n); int64_t* d_i = dmalloc<int64_t>(n); int32_t* d_found = dmalloc<int32_t>(n);
  float* d_w = dmalloc<float>((size_t)n * D);
  float* d_s1 = s1_out ? dmalloc<float>((size_t)n * D) : nullptr;
  float* d_s2 = s2_out ? dmalloc<float>((size_t)n * D) : nullptr;
  PS_CUDA(cudaMemcpyAsync(d_f, fields, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
  PS_CUDA(cudaMemcpyAsync(d_i, ids, sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
  emb_get_rows_kernel<<<ceil_div(n, 128), 128, 0, st>>>(slots, (uint32_t)C, rows, rs, Dp, D, d_f, d_i, n, d_w, d_s1, d_s2, d_found);
  PS_LAUNCH_CHECK();
  ctx->launches++;
  PS_CUDA(cudaMemcpyAsync(w_out, d_w, sizeof(float) * n * D, cudaMemcpyDeviceToHost, st));
  if (s1_out) PS_CUDA(cudaMemcpyAsync(s1_out, d_s1, sizeof(float) * n * D, cudaMemcpyDeviceToHost, st));
  if (s2_out) PS_CUDA(cudaMemcpyAsync(s2_out, d_s2, sizeof(float) * n * D, cudaMemcpyDeviceToHost, st));
  PS_CUDA(cudaMemcpyAsync(found, d_found, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  PS_CUDA(cudaStreamSynchronize(st));
  dfree(d_f); dfree(d_i); dfree(d_found); dfree(d_w); dfree(d_s1); dfree(d_s2);
}

void EmbTable::push_rows(const int32_t* fields, const int64_t* ids, int n, const float* grads, const ps_updater_spec& spec, int32_t* found) {
  if (n <= 0) return;
  for (int i = 0; i < n; ++i)
    PS_REQUIRE(fields[i] >= 0 && fields[i] < F && ids[i] >= 0 && ids[i] <= (int64_t)PS_KEY_ID_MASK, PS_ERR_ARG, "embedding: key outside the (field, id) domain");
  cudaStream_t st = ctx->stream;
  int32_t* d_f = dmalloc<int32_t>(n); int64_t* d_i = dmalloc<int64_t>(n); int32_t* d_found = dmalloc<int32_t>(n);
  float* d_g = dmalloc<float>((size_t)n * D);
  PS_CUDA(cudaMemcpyAsync(d_f, fields, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
  PS_CUDA(cudaMemcpyAsync(d_i, ids, sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
  PS_CUDA(cudaMemcpyAsync(d_g, grads, sizeof(float) * n * D, cudaMemcpyHostToDevice, st));
  emb_push_rows_kernel<<<ceil_div(n, 128), 128, 0, st>>>(slots, (uint32_t)C, rows, rs, Dp, D, d_f, d_i, n, d_g, make_updater_dev(spec), d_found);
  PS_LAUNCH_CHECK();
  ctx->launches++;
  PS_CUDA(cudaMemcpyAsync(found, d_found, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
  PS_CUDA(cudaStreamSynchronize(st));
  dfree(d_f); dfree(d_i); dfree(d_found); dfree(d_g);
}

void EmbTable::put_rows(const int32_t* fields, const int64_t* ids, int n, float* w_io, int replace) {
  if (n <= 0) return;
  for (int i = 0; i < n; ++i)
    PS_REQUIRE(fields[i] >= 0 && fields[i] < F && ids[i] >= 0 && ids[i] <= (int64_t)PS_KEY_ID_MASK, PS_ERR_ARG, "embedding: key outside the (field, id) domain");
  cudaStream_t st = ctx->stream;
  int32_t* d_f = dmalloc<int32_t>(n); int64_t* d_i = dmalloc<int64_t>(n);
  float* d_w = dmalloc<float>((size_t)n * D);
  PS_CUDA(cudaMemcpyAsync(d_f, fields, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
  PS_CUDA(cudaMemcpyAsync(d_i, ids, sizeof(int64_t) * n, cudaMemcpyHostToDevice, st));
  PS_CUDA(cudaMemcpyAsync(d_w, w_io, sizeof(float) * n * D, cudaMemcpyHostToDevice, st));
  emb_put_rows_kernel<<<ceil_div(n, 128), 128, 0, st>>>(slots, (uint32_t)C, rows, rs, D, d_f, d_i, n, d_w, replace, counters);
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
  if (p2p != nullptr) {                          /* this step's wide_in mailbox, once every replica's ids have landed */
    p2p_wait_all(p2p, CH_WIDE);
    ids = reinterpret_cast<const int64_t*>(p2p_region(p2p, p2p->me, p2p->off_wide));
  }
  if (i >= n) return;
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
__global__ void wide_push_kernel(WideSlot* slots, uint32_t C, int64_t id, float g, UpdaterDev upd, int* found) {
  bool ins;
  const int slot = wide_find_or_insert(slots, C, ps_pack_key(0u, (uint64_t)id), false, &ins);
  *found = slot >= 0;
  if (slot < 0 || (upd.kind == PS_UPD_FTRL && g == 0.0f)) return;
  apply_elem<true>(upd, slots[slot].w, slots[slot].s1, slots[slot].s2, g);
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
int WideTable::push(int64_t id, float g, const ps_updater_spec& spec) {
  int* d = dmalloc_zero<int>(1, ctx->stream);
  wide_push_kernel<<<1, 1, 0, ctx->stream>>>(slots, (uint32_t)C, id, g, make_updater_dev(spec), d);
  PS_LAUNCH_CHECK();
  ctx->launches++;
  int h = 0;
  PS_CUDA(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  dfree(d);
  return h;
}
void WideTable::put(int64_t id, float wv) {
  wide_put_kernel<<<1, 1, 0, ctx->stream>>>(slots, (uint32_t)C, id, wv, counters);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

}  // namespace psb

/*
 * p2p.cu — producer-side kernels that store straight into peer HBM over NVLink, and the
 * consumer-side wait / unpack / reduce kernels (see p2p.cuh for the protocol).
 */
#include "p2p.cuh"

#include <algorithm>
#include <cstdlib>

namespace psb {

namespace {

__global__ void p2p_begin_kernel(P2PState* st) {
  if (threadIdx.x == 0) st->seq += 1u;
}

/* PSRouterClient.getList, request side (PSRouterClient.java:60-68): the router sends each key of the batch ONCE to
 * its shard.  Every lookup finds-or-inserts its key in the per-batch table and counts itself; the first occurrence
 * reserves a position in the owner's bucket (one global atomic per (block, owner)) and stores the key straight into
 * the owner's keys_in[me][pos] over NVLink.  The occurrence count is only complete when the kernel ends (p2p_counts_kernel
 * copies it next to the gradient sums).  This is the FIRST kernel of a sharded step: it works with seq + 1 throughout, and
 * the block that finishes last makes that the step's sequence number before it flags the owners — every other block has
 * read the old value by then (its ticket comes after), every later kernel of the step reads the new one.               */
__global__ void __launch_bounds__(256) p2p_route_send_kernel(P2PState* st, BatchSlot* __restrict__ bt, uint32_t BT, const int64_t* __restrict__ E, int L, int F,
                                                             int32_t* __restrict__ lk_b, int32_t* __restrict__ ulist) {
  const int lane = threadIdx.x & 31;
  const int R = st->R, cap = st->cap;
  const int N = L / F;
  pdl_launch_dependents();                       /* the owner-side lookup may be scheduled: it waits for this grid's completion before it reads seq */
  const uint32_t seq = *reinterpret_cast<const volatile uint32_t*>(&st->seq) + 1u;
  int32_t* cursor = st->cursor[seq & 1u];
  __shared__ int s_cnt[kP2PMaxRanks], s_base[kP2PMaxRanks];
  __shared__ bool s_last;
  /* a capped grid striding over tiles of 256 lookups (the publish at the end costs one system fence + one ticket per block) */
  for (int t0 = blockIdx.x * 256; t0 < L; t0 += gridDim.x * 256) {
    const int t = t0 + threadIdx.x;                           /* field-major: a warp works on one field, consecutive samples */
    int b = -1, owner = 0;
    unsigned long long key = PS_KEY_EMPTY;
    if (t < L) {
      const int j = t / N;
      key = ps_pack_key((uint32_t)j, (uint64_t)E[(size_t)(t - j * N) * F + j]);
      owner = (int)ps_owner_of(key, (uint32_t)R);
    }
    /* duplicates inside the warp (a hot key fills whole warps) are resolved by ONE lane: one probe, one count update */
    const unsigned peers = __match_any_sync(0xffffffffu, key != PS_KEY_EMPTY ? key : (unsigned long long)lane);
    const int leader = __ffs(peers) - 1;
    bool first = false;
    if (key != PS_KEY_EMPTY && lane == leader) {
      uint32_t s = (uint32_t)(ps_mix64(key) >> 20) & (BT - 1u);
      for (uint32_t p = 0; p < BT; ++p) {
        const unsigned long long k = *reinterpret_cast<volatile unsigned long long*>(&bt[s].key);
        if (k == key) { b = (int)s; break; }
        if (k == PS_KEY_EMPTY) {
          const unsigned long long old = atomicCAS(&bt[s].key, (unsigned long long)PS_KEY_EMPTY, key);
          if (old == PS_KEY_EMPTY || old == key) { b = (int)s; break; }
        }
        s = (s + 1u) & (BT - 1u);
      }
      if (b >= 0) first = atomicAdd(&bt[b].cnt, (uint32_t)__popc(peers)) == 0u;
    }
    b = __shfl_sync(0xffffffffu, b, leader);
    if (t < L) lk_b[t] = b;                      /* field-major like EmbTable::lk_slot: the backward's scatter walks it the same way */
    if (threadIdx.x < kP2PMaxRanks) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    int rank_in_block = 0;
    if (first) rank_in_block = atomicAdd(&s_cnt[owner], 1);
    __syncthreads();
    if (threadIdx.x < R) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], s_cnt[threadIdx.x]) : 0;
    __syncthreads();
    if (first) {
      const int pos = s_base[owner] + rank_in_block;
      if (pos < cap) {
        bt[b].upos = owner * cap + pos;
        ulist[owner * cap + pos] = b;
        reinterpret_cast<unsigned long long*>(p2p_region_of(st, owner, st->off_keys, seq))[(size_t)st->me * cap + pos] = key;
      } else { bt[b].upos = -1; st->overflow = 1; }
    }
    __syncthreads();                             /* s_cnt / s_base are rewritten by the next tile */
  }
  /* p2p_publish_last with the sequence number this kernel introduces */
  __syncthreads();
  if (threadIdx.x == 0) {
    p2p_block_fence(st);
    const uint32_t tk = atomicAdd(&st->ticket[CH_KEYS], 1u);
    s_last = tk == gridDim.x - 1u;
    if (s_last) { st->ticket[CH_KEYS] = 0u; st->seq = seq; st->pub_seq[CH_KEYS] = seq; }
  }
  __syncthreads();
  if (s_last && (int)threadIdx.x < R) {
    const int r = threadIdx.x;
    const int c = min(*reinterpret_cast<volatile int32_t*>(&cursor[r]), cap);
    reinterpret_cast<volatile int32_t*>(p2p_region_of(st, r, st->off_counts, seq))[st->me] = c;
    __threadfence_system();
    p2p_st_release_sys(reinterpret_cast<uint32_t*>(p2p_region_of(st, r, st->off_flags, seq)) + CH_KEYS * kP2PMaxRanks + st->me, seq);
  }
}

/* the occurrence count of every unique key of the batch, next to its gradient sum (the owner reads both): side stream, after route_send */
__global__ void __launch_bounds__(256) p2p_counts_kernel(const P2PState* st, const BatchSlot* __restrict__ bt, const int32_t* __restrict__ ulist) {
  const int cap = st->cap, R = st->R;
  const int32_t* cursor = st->cursor[st->seq & 1u];
  uint32_t* gc = reinterpret_cast<uint32_t*>(p2p_region(st, st->me, st->off_gcnt));
  for (int owner = 0; owner < R; ++owner) {
    const int valid = min(cursor[owner], cap);
    for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < valid; pos += gridDim.x * blockDim.x) {
      const int q = owner * cap + pos;
      gc[q] = bt[ulist[q]].cnt;
    }
  }
}

/* side stream, after the backward's scatter (the last reader of the de-duplication table): (1) clears exactly the table entries
 * this batch used; (2) zeroes the gradient sums of the OTHER parity — the owners read them during their previous step's update,
 * which they have all left (this rank has seen their CH_KEYS of the current step) — and then that parity's cursors.       */
__global__ void __launch_bounds__(256) p2p_tidy_kernel(P2PState* st, BatchSlot* __restrict__ bt, const int32_t* __restrict__ ulist) {
  const int cap = st->cap, R = st->R, tpl = st->Dp >> 2, Dp = st->Dp;
  const uint32_t seq = st->seq;
  const int32_t* cur = st->cursor[seq & 1u];
  int32_t* old = st->cursor[(seq + 1u) & 1u];
  float* gs = reinterpret_cast<float*>(p2p_region_of(st, st->me, st->off_grads, seq + 1u));
  const long g0 = (long)blockIdx.x * blockDim.x + threadIdx.x, gstride = (long)gridDim.x * blockDim.x;
  for (int owner = 0; owner < R; ++owner) {
    const int valid = min(cur[owner], cap);
    for (long pos = g0; pos < valid; pos += gstride)
      *reinterpret_cast<uint4*>(&bt[ulist[(long)owner * cap + pos]]) = make_uint4(0u, 0u, 0u, 0u);
    const long zv = (long)min(*reinterpret_cast<volatile int32_t*>(&old[owner]), cap) * tpl;
    float* base = gs + (size_t)owner * cap * Dp;
    for (long g = g0; g < zv; g += gstride) st_f4(base + g * 4, make_float4(0.f, 0.f, 0.f, 0.f));
  }
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t tk = atomicAdd(&st->tidy_ticket, 1u);
    s_last = tk == gridDim.x - 1u;
    if (s_last) st->tidy_ticket = 0u;
  }
  __syncthreads();
  if (s_last && (int)threadIdx.x < kP2PMaxRanks) old[threadIdx.x] = 0;
}

/* all-gather by stores: this rank's `bytes` go to slot `me` of the channel's region on every rank */
__global__ void __launch_bounds__(256) p2p_bcast_kernel(P2PState* st, const uint4* __restrict__ src, size_t n16, size_t off, int channel) {
  const int R = st->R, me = st->me;
  const size_t total = n16 * (size_t)R;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / n16);
    const size_t c = i - (size_t)r * n16;
    reinterpret_cast<uint4*>(p2p_region(st, r, off))[(size_t)me * n16 + c] = src[c];
  }
  p2p_publish_last(st, channel, gridDim.x);
}

template <bool VEC>
__global__ void __launch_bounds__(256) p2p_unpack_kernel(const P2PState* st, const BatchSlot* __restrict__ bt, const int32_t* __restrict__ lk_b, int L, int F, int D,
                                                         float* __restrict__ out, int ldo, const float* __restrict__ X, int Xn, int xoff, int N) {
  const int Dp = st->Dp;
  p2p_wait_all(st, CH_ROWS);                   /* every owner's gather has landed in rows_in */
  const float* rows = reinterpret_cast<const float*>(p2p_region(st, st->me, st->off_rows));
  long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long emb_work = VEC ? (long)L * (Dp >> 2) : (long)L * Dp;
  if (g >= emb_work) {                         /* ConcatLayer: the numeric features next to the embeddings */
    g -= emb_work;
    if (X != nullptr && g < (long)N * Xn) { const int n = (int)(g / Xn), x = (int)(g - (long)n * Xn); out[(size_t)n * ldo + xoff + x] = X[g]; }
    return;
  }
  if (VEC) {                                   /* Dp/4 lanes per lookup, 128-bit moves */
    const int tpl = Dp >> 2;
    const long l = g / tpl;
    const int part = (int)(g - l * tpl);
    if (l >= L) return;
    const int n = (int)(l / F), j = (int)(l - (long)n * F);
    const int b = lk_b[(long)j * N + n];
    const int pos = b >= 0 ? bt[b].upos : -1;
    const float4 v = pos >= 0 ? ld_f4(rows + (size_t)pos * Dp + part * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    st_f4(out + (size_t)n * ldo + j * D + part * 4, v);
  } else {
    const long l = g / Dp;
    const int d = (int)(g - l * Dp);
    if (l >= L || d >= D) return;
    const int n = (int)(l / F), j = (int)(l - (long)n * F);
    const int b = lk_b[(long)j * N + n];
    const int pos = b >= 0 ? bt[b].upos : -1;
    out[(size_t)n * ldo + j * D + d] = pos >= 0 ? rows[(size_t)pos * Dp + d] : 0.f;
  }
}

/* PServer sync mode sums the pushes of all workers (PServer.java:164-195): every rank adds the R
 * mailboxes in rank order, so all replicas compute bit-identical dense updates                  */
__global__ void __launch_bounds__(256) p2p_reduce_kernel(const P2PState* st, float* __restrict__ gsum) {
  const int glen = st->glen, R = st->R;
  const float* in = reinterpret_cast<const float*>(p2p_region(st, st->me, st->off_gsum));
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= glen) return;
  float s = 0.f;
  for (int r = 0; r < R; ++r) s = __fadd_rn(s, in[(size_t)r * glen + i]);
  gsum[i] = s;
}

/* KVStore.update → client.push per key (KVStore.java:257-260): EmbTable::scatter_rows (table.cu) sums the per-lookup row gradients
 * (ReLU mask of EmbeddingField.java:91-93) per unique key of this rank's batch into the gsums region of this rank's slab and flags
 * the owners; an owner's update kernel reads the sums of every requester straight out of their slabs (EmbTable::update_pull). */
size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

void P2P::create(Ctx* c, int R_, int me_, int cap_, int Dp_, int NF_, int glen_, int64_t max_lookups) {
  PS_REQUIRE(R_ >= 1 && R_ <= kP2PMaxRanks && me_ >= 0 && me_ < R_ && cap_ > 0, PS_ERR_ARG, "p2p: bad rank layout");
  ctx = c; R = R_; me = me_; cap = cap_; Dp = Dp_; NF = NF_; glen = (glen_ + 3) / 4 * 4;
  size_t off = 0;
  host = P2PState{};
  { const char* e = std::getenv("PS_P2P_BLOCK_FENCE_SYS"); host.block_fence_sys = (e && std::atoi(e) != 0) ? 1 : 0; }
  host.R = R; host.me = me; host.cap = cap; host.Dp = Dp; host.NF = NF; host.glen = glen;
  host.off_keys = off; off = align_up(off + (size_t)R * cap * 8, 256);
  host.off_rows = off; off = align_up(off + (size_t)R * cap * Dp * 4, 256);
  host.off_grads = off; off = align_up(off + (size_t)R * cap * Dp * 4, 256);
  host.off_gcnt = off; off = align_up(off + (size_t)R * cap * 4, 256);
  host.off_wide = off; off = align_up(off + (size_t)R * NF * 8, 256);
  host.off_gsum = off; off = align_up(off + (size_t)R * glen * 4, 256);
  host.off_counts = off; off = align_up(off + (size_t)kP2PMaxRanks * 4, 256);
  host.off_flags = off; off = align_up(off + (size_t)CH_COUNT * kP2PMaxRanks * 4, 256);
  host.parity_stride = off;
  slab_bytes = 2 * off;
  PS_CUDA(cudaMalloc(&slab, slab_bytes));
  PS_CUDA(cudaMemsetAsync(slab, 0, slab_bytes, ctx->stream));
  dev = dmalloc_zero<P2PState>(1, ctx->stream);
  Lmax = max_lookups;
  BT = 1024;
  while ((int64_t)BT < 2 * Lmax) BT <<= 1;
  bt = dmalloc_zero<BatchSlot>(BT, ctx->stream);
  lk_b = dmalloc<int32_t>((size_t)std::max<int64_t>(Lmax, 1));
  ulist = dmalloc_zero<int32_t>((size_t)R * cap, ctx->stream);
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
}

void P2P::get_handle(void* out64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  PS_CUDA(cudaIpcGetMemHandle(&h, slab));
  std::memcpy(out64, &h, 64);
}

void P2P::connect(const void* all_handles) {
  for (int r = 0; r < R; ++r) {
    if (r == me) { host.peer[r] = slab; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const unsigned char*>(all_handles) + 64 * r, 64);
    void* p = nullptr;
    PS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    peer_mapped[r] = p;
    host.peer[r] = static_cast<unsigned char*>(p);
  }
  PS_CUDA(cudaMemcpyAsync(dev, &host, sizeof(P2PState), cudaMemcpyHostToDevice, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  connected = true;
}

void P2P::destroy() {
  for (int r = 0; r < R; ++r) if (peer_mapped[r]) { cudaIpcCloseMemHandle(peer_mapped[r]); peer_mapped[r] = nullptr; }
  dfree(slab); dfree(dev); dfree(bt); dfree(lk_b); dfree(ulist);
  slab = nullptr; dev = nullptr; bt = nullptr; lk_b = nullptr; ulist = nullptr; connected = false;
}

#define P2P_LAUNCHED() do { PS_LAUNCH_CHECK(); ctx->launches++; } while (0)

/* a one-warp wait on side streams whose consumer kernel has a large grid: 900 spinning blocks would hold every SM's thread slots
 * against the main stream's kernels, one warp holds none */
__global__ void p2p_wait_kernel(const P2PState* st, int channel) { p2p_wait_all_seq(st, channel, st->pub_seq[channel]); }
void P2P::wait(int channel) { p2p_wait_kernel<<<1, 32, 0, ctx->stream>>>(dev, channel); P2P_LAUNCHED(); }

void P2P::begin() { p2p_begin_kernel<<<1, 32, 0, ctx->stream>>>(dev); P2P_LAUNCHED(); }

void P2P::route_send(const int64_t* E, int N, int F) {
  const int L = N * F;
  PS_REQUIRE(L <= Lmax, PS_ERR_ARG, "p2p: batch larger than the de-duplication table");
  p2p_route_send_kernel<<<std::min(ceil_div(L, 256), ctx->num_sms * 4), 256, 0, ctx->stream>>>(dev, bt, BT, E, L, F, lk_b, ulist);
  P2P_LAUNCHED();
}

void P2P::bcast(const void* src, size_t bytes, int channel) {
  PS_REQUIRE(bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0, PS_ERR_ARG, "p2p bcast: 16-byte granularity");
  PS_REQUIRE(channel == CH_WIDE ? bytes <= (size_t)NF * 8 : bytes <= (size_t)glen * 4, PS_ERR_ARG, "p2p bcast: payload larger than the mailbox");
  const size_t n16 = bytes / 16;
  const int grid = (int)std::min<size_t>((n16 * R + 255) / 256, (size_t)ctx->num_sms * 4);
  p2p_bcast_kernel<<<std::max(grid, 1), 256, 0, ctx->stream>>>(dev, static_cast<const uint4*>(src), n16, channel == CH_WIDE ? host.off_wide : host.off_gsum, channel);
  P2P_LAUNCHED();
}

void P2P::unpack(int N, int F, int D, float* out, int ldo, const float* X, int Xn, int xoff) {
  const bool vec = D % 4 == 0 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const long xw = X ? (long)N * Xn : 0;
  if (vec) p2p_unpack_kernel<true><<<ceil_div((long)N * F * (Dp / 4) + xw, 256), 256, 0, ctx->stream>>>(dev, bt, lk_b, N * F, F, D, out, ldo, X, Xn, xoff, N);
  else p2p_unpack_kernel<false><<<ceil_div((long)N * F * Dp + xw, 256), 256, 0, ctx->stream>>>(dev, bt, lk_b, N * F, F, D, out, ldo, X, Xn, xoff, N);
  P2P_LAUNCHED();
}

void P2P::reduce_gsum(float* gsum) {
  p2p_reduce_kernel<<<ceil_div(glen, 256), 256, 0, ctx->stream>>>(dev, gsum);
  P2P_LAUNCHED();
}

void P2P::counts() {
  p2p_counts_kernel<<<std::max(1, std::min(ceil_div(cap, 256), ctx->num_sms * 2)), 256, 0, ctx->stream>>>(dev, bt, ulist);
  P2P_LAUNCHED();
}

void P2P::tidy() {
  const long total = (long)cap * (Dp / 4);
  p2p_tidy_kernel<<<(int)std::max<long>(1, std::min<long>(ceil_div(total, 256), (long)ctx->num_sms * 4)), 256, 0, ctx->stream>>>(dev, bt, ulist);
  P2P_LAUNCHED();
}

bool P2P::overflowed() {
  P2PState h;
  PS_CUDA(cudaMemcpyAsync(&h, dev, sizeof(P2PState), cudaMemcpyDeviceToHost, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  return h.overflow != 0;
}

}  // namespace psb

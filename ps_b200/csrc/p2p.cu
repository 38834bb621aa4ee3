/*
 * p2p.cu — producer-side kernels that store straight into peer HBM over NVLink, and the
 * consumer-side wait / unpack / reduce kernels (see p2p.cuh for the protocol).
 */
#include "p2p.cuh"

#include <algorithm>

namespace psb {

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

/* Runs as its own one-warp kernel right after a producer kernel: the kernel boundary has completed every
 * store of the producer (including those that crossed NVLink), so ONE system-scope fence and one release
 * store per peer publish the payload — no per-block fences inside the bandwidth-bound producers.       */
__global__ void p2p_publish_kernel(P2PState* st, int channel) {
  const int r = threadIdx.x;
  if (r >= st->R) return;
  if (channel == CH_KEYS) {
    const int c = min(*reinterpret_cast<volatile int32_t*>(&st->cursor[r]), st->cap);
    reinterpret_cast<volatile int32_t*>(p2p_region(st, r, st->off_counts))[st->me] = c;
  }
  __threadfence_system();
  uint32_t* f = reinterpret_cast<uint32_t*>(p2p_region(st, r, st->off_flags)) + channel * kP2PMaxRanks + st->me;
  st_release_sys(f, st->seq);
}

__global__ void p2p_begin_kernel(P2PState* st) {
  if (threadIdx.x == 0) st->seq += 1u;
  if (threadIdx.x < kP2PMaxRanks) st->cursor[threadIdx.x] = 0;
}

/* PSRouterClient.getList, request side (PSRouterClient.java:60-68): bucket by owner and store each key
 * directly into the owner's keys_in[me][pos]; the last block publishes the per-owner counts and flags. */
__global__ void __launch_bounds__(256) p2p_route_send_kernel(P2PState* st, const int64_t* __restrict__ E, int L, int F, int32_t* __restrict__ send_pos) {
  /* field-major work order (t = j*N + n): consecutive bucket positions then hold the same field for consecutive
   * samples, so the OWNER's probe and scatter kernels can collapse a hot key warp-wide (one atomic per 32) */
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int R = st->R, cap = st->cap, me = st->me;
  const bool valid = t < L;
  int l = 0;
  unsigned long long key = 0;
  int owner = -1 - lane;
  if (valid) {
    const int N = L / F, j = t / N;
    l = (t - j * N) * F + j;
    key = ps_pack_key((uint32_t)j, (uint64_t)E[l]);
    owner = (int)ps_owner_of(key, (uint32_t)R);
  }
  /* bucket positions: warp-aggregated counts into shared memory, then ONE global atomic per (block, owner) */
  __shared__ int s_cnt[kP2PMaxRanks], s_base[kP2PMaxRanks];
  if (threadIdx.x < kP2PMaxRanks) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const unsigned peers = __match_any_sync(0xffffffffu, owner);
  const int leader = __ffs(peers) - 1;
  int rank_in_block = 0;
  if (valid && lane == leader) rank_in_block = atomicAdd(&s_cnt[owner], __popc(peers));
  rank_in_block = __shfl_sync(0xffffffffu, rank_in_block, leader) + __popc(peers & ((1u << lane) - 1u));
  __syncthreads();
  if (threadIdx.x < R) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(&st->cursor[threadIdx.x], s_cnt[threadIdx.x]) : 0;
  __syncthreads();
  if (valid) {
    const int pos = s_base[owner] + rank_in_block;
    if (pos < cap) {
      reinterpret_cast<unsigned long long*>(p2p_region(st, owner, st->off_keys))[(size_t)me * cap + pos] = key;
      send_pos[l] = owner * cap + pos;
    } else {
      send_pos[l] = -1;
      st->overflow = 1;
    }
  }
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
}

/* publish this rank's payload of `channel` and wait for everybody else's, in one launch */
__global__ void p2p_publish_wait_kernel(P2PState* st, int channel) {
  const int r = threadIdx.x;
  if (r < st->R) {
    if (channel == CH_KEYS) {
      const int c = min(*reinterpret_cast<volatile int32_t*>(&st->cursor[r]), st->cap);
      reinterpret_cast<volatile int32_t*>(p2p_region(st, r, st->off_counts))[st->me] = c;
    }
    __threadfence_system();
    const uint32_t seq = st->seq;
    st_release_sys(reinterpret_cast<uint32_t*>(p2p_region(st, r, st->off_flags)) + channel * kP2PMaxRanks + st->me, seq);
    const uint32_t* f = reinterpret_cast<const uint32_t*>(p2p_region(st, st->me, st->off_flags)) + channel * kP2PMaxRanks + r;
    while ((int32_t)(ld_acquire_sys(f) - seq) < 0) __nanosleep(32);
  }
  __threadfence_system();
}

__global__ void p2p_wait_kernel(const P2PState* st, int channel) {
  const int r = threadIdx.x;
  if (r < st->R) {
    const uint32_t* f = reinterpret_cast<const uint32_t*>(p2p_region(st, st->me, st->off_flags)) + channel * kP2PMaxRanks + r;
    const uint32_t seq = st->seq;
    while ((int32_t)(ld_acquire_sys(f) - seq) < 0) __nanosleep(64);
  }
  __threadfence_system();
}

/* PServer.getList, response side (PServer.java:102-117): the owner's gather writes each row (ReLU applied,
 * EmbeddingField.java:75) into the REQUESTER's rows_in[me][idx] — gather and transfer are one kernel.   */
__global__ void __launch_bounds__(256) p2p_gather_send_kernel(P2PState* st, const float* __restrict__ w, int D, const int32_t* __restrict__ lk_slot) {
  const int cap = st->cap, Dp = st->Dp, me = st->me;
  const int tpl = Dp >> 2;
  const long total = (long)st->R * cap * tpl;
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < total) {
    const int q = (int)(g / tpl), part = (int)(g - (long)q * tpl);
    const int slot = lk_slot[q];
    if (slot >= 0) {
      const int src = q / cap, idx = q - src * cap;
      float4 v = __ldg(reinterpret_cast<const float4*>(w + (size_t)slot * Dp + part * 4));
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      float* dst = reinterpret_cast<float*>(p2p_region(st, src, st->off_rows)) + ((size_t)me * cap + idx) * Dp + part * 4;
      st_f4(dst, v);
    }
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256) p2p_unpack_kernel(const P2PState* st, const int32_t* __restrict__ send_pos, int L, int F, int D,
                                                         float* __restrict__ out, int ldo, const float* __restrict__ X, int Xn, int xoff, int N) {
  const int Dp = st->Dp;
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
    const int pos = send_pos[l];
    const float4 v = pos >= 0 ? ld_f4(rows + (size_t)pos * Dp + part * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    st_f4(out + (size_t)n * ldo + j * D + part * 4, v);
  } else {
    const long l = g / Dp;
    const int d = (int)(g - l * Dp);
    if (l >= L || d >= D) return;
    const int n = (int)(l / F), j = (int)(l - (long)n * F);
    const int pos = send_pos[l];
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

/* KVStore.update → client.push per key (KVStore.java:257-260): per-lookup row gradient (ReLU mask of
 * EmbeddingField.java:91-93 applied here) stored into the owner's grads_in[me][pos]                 */
template <bool VEC>
__global__ void __launch_bounds__(256) p2p_pack_send_kernel(P2PState* st, const float* __restrict__ delta, int ldd, const float* __restrict__ act, int lda,
                                                            const int32_t* __restrict__ send_pos, int L, int F, int D) {
  const int Dp = st->Dp, cap = st->cap, me = st->me;
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    const int tpl = Dp >> 2;
    const long l = g / tpl;
    const int part = (int)(g - l * tpl);
    if (l < L) {
      const int pos = send_pos[l];
      if (pos >= 0) {
        const int n = (int)(l / F), j = (int)(l - (long)n * F);
        const float4 dv = ld_f4(delta + (size_t)n * ldd + j * D + part * 4), av = ld_f4(act + (size_t)n * lda + j * D + part * 4);
        float4 v;
        v.x = __fmul_rn(dv.x, av.x > 0.f ? 1.f : 0.f); v.y = __fmul_rn(dv.y, av.y > 0.f ? 1.f : 0.f);
        v.z = __fmul_rn(dv.z, av.z > 0.f ? 1.f : 0.f); v.w = __fmul_rn(dv.w, av.w > 0.f ? 1.f : 0.f);
        const int owner = pos / cap, idx = pos - owner * cap;
        st_f4(reinterpret_cast<float*>(p2p_region(st, owner, st->off_grads)) + ((size_t)me * cap + idx) * Dp + part * 4, v);
      }
    }
    return;
  }
  const long l = g / Dp;
  const int d = (int)(g - l * Dp);
  if (l < L) {
    const int pos = send_pos[l];
    if (pos >= 0) {
      const int n = (int)(l / F), j = (int)(l - (long)n * F);
      float v = 0.f;
      if (d < D) v = __fmul_rn(delta[(size_t)n * ldd + j * D + d], act[(size_t)n * lda + j * D + d] > 0.f ? 1.f : 0.f);
      const int owner = pos / cap, idx = pos - owner * cap;
      reinterpret_cast<float*>(p2p_region(st, owner, st->off_grads))[((size_t)me * cap + idx) * Dp + d] = v;
    }
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

void P2P::create(Ctx* c, int R_, int me_, int cap_, int Dp_, int NF_, int glen_) {
  PS_REQUIRE(R_ >= 1 && R_ <= kP2PMaxRanks && me_ >= 0 && me_ < R_ && cap_ > 0, PS_ERR_ARG, "p2p: bad rank layout");
  ctx = c; R = R_; me = me_; cap = cap_; Dp = Dp_; NF = NF_; glen = (glen_ + 3) / 4 * 4;
  size_t off = 0;
  host = P2PState{};
  host.R = R; host.me = me; host.cap = cap; host.Dp = Dp; host.NF = NF; host.glen = glen;
  host.off_keys = off; off = align_up(off + (size_t)R * cap * 8, 256);
  host.off_rows = off; off = align_up(off + (size_t)R * cap * Dp * 4, 256);
  host.off_grads = off; off = align_up(off + (size_t)R * cap * Dp * 4, 256);
  host.off_wide = off; off = align_up(off + (size_t)R * NF * 8, 256);
  host.off_gsum = off; off = align_up(off + (size_t)R * glen * 4, 256);
  host.off_counts = off; off = align_up(off + (size_t)kP2PMaxRanks * 4, 256);
  host.off_flags = off; off = align_up(off + (size_t)CH_COUNT * kP2PMaxRanks * 4, 256);
  host.parity_stride = off;
  slab_bytes = 2 * off;
  PS_CUDA(cudaMalloc(&slab, slab_bytes));
  PS_CUDA(cudaMemsetAsync(slab, 0, slab_bytes, ctx->stream));
  dev = dmalloc_zero<P2PState>(1, ctx->stream);
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
  dfree(slab); dfree(dev);
  slab = nullptr; dev = nullptr; connected = false;
}

#define P2P_LAUNCHED() do { PS_LAUNCH_CHECK(); ctx->launches++; } while (0)

void P2P::publish(int channel) { p2p_publish_kernel<<<1, 32, 0, ctx->stream>>>(dev, channel); P2P_LAUNCHED(); }

void P2P::begin() { p2p_begin_kernel<<<1, 32, 0, ctx->stream>>>(dev); P2P_LAUNCHED(); }

void P2P::route_send(const int64_t* E, int N, int F, int32_t* send_pos) {
  const int L = N * F;
  p2p_route_send_kernel<<<ceil_div(L, 256), 256, 0, ctx->stream>>>(dev, E, L, F, send_pos);
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

void P2P::publish_wait(int channel) { p2p_publish_wait_kernel<<<1, 32, 0, ctx->stream>>>(dev, channel); P2P_LAUNCHED(); }
void P2P::wait(int channel) { p2p_wait_kernel<<<1, 32, 0, ctx->stream>>>(dev, channel); P2P_LAUNCHED(); }

void P2P::gather_send(const float* w, int D, const int32_t* lk_slot) {
  const long total = (long)R * cap * (Dp / 4);
  p2p_gather_send_kernel<<<ceil_div(total, 256), 256, 0, ctx->stream>>>(dev, w, D, lk_slot);
  P2P_LAUNCHED();
}

void P2P::unpack(const int32_t* send_pos, int N, int F, int D, float* out, int ldo, const float* X, int Xn, int xoff) {
  const bool vec = D % 4 == 0 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const long xw = X ? (long)N * Xn : 0;
  if (vec) p2p_unpack_kernel<true><<<ceil_div((long)N * F * (Dp / 4) + xw, 256), 256, 0, ctx->stream>>>(dev, send_pos, N * F, F, D, out, ldo, X, Xn, xoff, N);
  else p2p_unpack_kernel<false><<<ceil_div((long)N * F * Dp + xw, 256), 256, 0, ctx->stream>>>(dev, send_pos, N * F, F, D, out, ldo, X, Xn, xoff, N);
  P2P_LAUNCHED();
}

void P2P::reduce_gsum(float* gsum) {
  p2p_reduce_kernel<<<ceil_div(glen, 256), 256, 0, ctx->stream>>>(dev, gsum);
  P2P_LAUNCHED();
}

void P2P::pack_send(const float* delta, int ldd, const float* act, int lda, const int32_t* send_pos, int N, int F, int D) {
  const bool vec = D % 4 == 0 && ldd % 4 == 0 && lda % 4 == 0 && (reinterpret_cast<uintptr_t>(delta) & 15) == 0 && (reinterpret_cast<uintptr_t>(act) & 15) == 0;
  if (vec) p2p_pack_send_kernel<true><<<ceil_div((long)N * F * (Dp / 4), 256), 256, 0, ctx->stream>>>(dev, delta, ldd, act, lda, send_pos, N * F, F, D);
  else p2p_pack_send_kernel<false><<<ceil_div((long)N * F * Dp, 256), 256, 0, ctx->stream>>>(dev, delta, ldd, act, lda, send_pos, N * F, F, D);
  P2P_LAUNCHED();
}

bool P2P::overflowed() {
  P2PState h;
  PS_CUDA(cudaMemcpyAsync(&h, dev, sizeof(P2PState), cudaMemcpyDeviceToHost, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  return h.overflow != 0;
}

}  // namespace psb

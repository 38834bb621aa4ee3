/*
 * dense.cu — dense parameter kernels and the network tail (sm_100a).
 *
 * Reference restated (/root/reference/src/main/java/):
 *   dense_update      store/KVStore.java:240-268 (g = sum / cnt) + update/*.java, for FcLayer keys
 *   tail_binary       layer/AddLayer.java:33-61, activations/Sigmoid.java:9-21,
 *                     loss/CrossEntropy.java:10-28, model/DNN.java:47-63
 *   tail_softmax      activations/Softmax.java:21-67, loss/SoftmaxLoss.java:9-28
 */
#include "dense.cuh"
#include "gemm.cuh"
#include "p2p.cuh"

#include <algorithm>

namespace psb {

/* ------------------------------------------------------------------ init */
/* MatrixUtil.rand(out, in, max) (util/MatrixUtil.java:62-74) with the counter-based draw of
 * ps_spec.h; element j of the reference's column-major out x in matrix is j = o + out*i.     */
__global__ void dense_init_kernel(float* __restrict__ W, int out, int in, int ldw, float* __restrict__ Wt, int ldwt, uint64_t seed, uint64_t key, float maxv) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)out * in) return;
  const int o = (int)(idx / in), i = (int)(idx - (long)o * in);
  const float v = ps_init_value(seed, key, (uint32_t)(o + out * i), maxv);
  W[(size_t)o * ldw + i] = v;
  if (Wt) Wt[(size_t)i * ldwt + o] = v;
}
void dense_init(Ctx* ctx, float* W, int out, int in, int ldw, float* Wt, int ldwt, uint64_t key, float maxv) {
  dense_init_kernel<<<ceil_div((long)out * in, 256), 256, 0, ctx->stream>>>(W, out, in, ldw, Wt, ldwt, ctx->seed, key, maxv);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

__device__ __forceinline__ float tf32_residual(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__global__ void split_lo_kernel(const float* __restrict__ x, float* __restrict__ lo, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) lo[i] = tf32_residual(x[i]);
}
void split_lo(Ctx* ctx, const float* x, float* lo, size_t n) {
  if (!x || !lo || n == 0) return;
  split_lo_kernel<<<(int)std::min<size_t>((n + 255) / 256, 1184), 256, 0, ctx->stream>>>(x, lo, n);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

__global__ void fill_column_kernel(float* buf, int ld, int col, int rows, float value) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) buf[(size_t)r * ld + col] = value;
}
void fill_column(Ctx* ctx, float* buf, int ld, int col, int rows, float value) {
  fill_column_kernel<<<ceil_div(rows, 256), 256, 0, ctx->stream>>>(buf, ld, col, rows, value);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

/* 32x32 shared-memory tiles: coalesced 128 B reads along the input rows and 128 B writes along the output rows */
__global__ void __launch_bounds__(256) transpose_copy_kernel(const float* __restrict__ in, int ldi, float* __restrict__ out, int ldo, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int r = r0 + ty + k, c = c0 + tx;
    tile[ty + k][tx] = (r < rows && c < cols) ? in[(size_t)r * ldi + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int c = c0 + ty + k, r = r0 + tx;
    if (r < rows && c < cols) out[(size_t)c * ldo + r] = tile[tx][ty + k];
  }
}
void transpose_copy(Ctx* ctx, const float* in, int ldi, float* out, int ldo, int rows, int cols) {
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  transpose_copy_kernel<<<grid, 256, 0, ctx->stream>>>(in, ldi, out, ldo, rows, cols);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

/* ------------------------------------------------------------------ fused dense update */
/* One work item per (o, c), c in [0, in]: c < in is weight (o, c), c == in is bias o.  The
 * gradient is the sum of the wgrad partial slabs divided by N: FcLayer.java:103 rowMeans and
 * :105 divi(N); KVStore.update's own division is by sumCnt = 1 (thread = 1).  Ftrl's early
 * return looks at element 0 of the key's gradient (FtrlUpdater.java:52).                     */
__device__ __forceinline__ void publish(StepStatus* st, const uint32_t* emb_counters, const uint32_t* wide_counters, StepStatus* host) {
  StepStatus s = *st;
  s.emb_err = emb_counters ? emb_counters[1] : 0u;
  /* n_unique was copied from the lookup's cursor by the tail kernel: final long before any publish, untouched by the update's reset */
  s.wide_err = wide_counters ? wide_counters[0] : 0u;
  *host = s;
  __threadfence_system();
}

__global__ void __launch_bounds__(256) dense_update_kernel(const __grid_constant__ DenseUpdateArgs a, StepStatus* __restrict__ st,
                                                           const uint32_t* __restrict__ emb_counters, const uint32_t* __restrict__ wide_counters,
                                                           StepStatus* __restrict__ host, const P2PState* __restrict__ p2p) {
  if (host != nullptr && blockIdx.x == 0 && threadIdx.x == 0) publish(st, emb_counters, wide_counters, host);   /* status is final before the updates */
  if (st->skip) return;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.total) return;
  int li = 0;
  while (li + 1 < a.n_layers && idx >= a.l[li + 1].first) ++li;
  const DenseLayerDesc& L = a.l[li];
  const long r = idx - L.first;
  const int cols = L.in + 1;
  const int o = (int)(r / cols), c = (int)(r - (long)o * cols);
  const float Nf = (float)a.N;
  /* single GPU: the wgrad partial slabs; sharded over peer memory: the R ranks' flat gradient buffers in this
   * step's gsum_in mailbox, added in rank order (every replica computes the same bits)                        */
  const float* G = L.G;
  size_t slab = L.slab;
  int nsplit = L.nsplit, ldg = L.ldg;
  if (p2p != nullptr) {
    G = reinterpret_cast<const float*>(p2p_region_of(p2p, p2p->me, p2p->off_gsum, p2p->pub_seq[CH_GSUM])) + L.first;   /* (the step this rank last sent its sums for) */
    slab = (size_t)p2p->glen; nsplit = p2p->R; ldg = cols;
  }
  float g = 0.0f;
  for (int z = 0; z < nsplit; ++z) g = __fadd_rn(g, G[(size_t)z * slab + (size_t)o * ldg + c]);
  g = __fdiv_rn(g, Nf);
  const bool is_bias = c == L.in;
  const UpdaterDev& u = is_bias ? L.updB : L.updW;
  if (u.kind == PS_UPD_FTRL) {
    float g0 = 0.0f;
    const size_t o0 = is_bias ? (size_t)L.in : 0;
    for (int z = 0; z < nsplit; ++z) g0 = __fadd_rn(g0, G[(size_t)z * slab + o0]);
    if (__fdiv_rn(g0, Nf) == 0.0f) return;
  }
  if (is_bias) {
    float w = L.bias[o], m1 = L.sb1[o], m2 = L.sb2[o];
    apply_elem(u, w, m1, m2, g);
    L.bias[o] = w; L.sb1[o] = m1; L.sb2[o] = m2;
  } else {
    const size_t off = (size_t)o * L.ldw + c;
    float w = L.W[off], m1 = L.sW1[off], m2 = L.sW2[off];
    apply_elem(u, w, m1, m2, g);
    L.W[off] = w; L.sW1[off] = m1; L.sW2[off] = m2;
    if (L.Wt) L.Wt[(size_t)c * L.ldwt + o] = w;
    if (L.Wlo) { const float lo = tf32_residual(w); L.Wlo[off] = lo; if (L.Wtlo) L.Wtlo[(size_t)c * L.ldwt + o] = lo; }
  }
}
void dense_update(Ctx* ctx, const DenseUpdateArgs& a, StepStatus* st, const uint32_t* emb_counters, const uint32_t* wide_counters,
                  StepStatus* host_mapped, const P2PState* p2p) {
  dense_update_kernel<<<ceil_div(a.total, 256), 256, 0, ctx->stream>>>(a, st, emb_counters, wide_counters, host_mapped, p2p);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

__global__ void __launch_bounds__(256) dense_reduce_kernel(const __grid_constant__ DenseUpdateArgs a, const StepStatus* __restrict__ st,
                                                           float* __restrict__ gsum, const uint32_t* __restrict__ emb_counters) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  /* third scalar: this rank's embedding shard is full — summed over ranks it makes EVERY replica skip the step */
  if (idx == 0) { gsum[a.total] = st->loss; gsum[a.total + 1] = st->gbar; gsum[a.total + 2] = (emb_counters != nullptr && emb_counters[1] != 0u) ? 1.0f : 0.0f; }
  if (idx >= a.total) return;
  int li = 0;
  while (li + 1 < a.n_layers && idx >= a.l[li + 1].first) ++li;
  const DenseLayerDesc& L = a.l[li];
  const long r = idx - L.first;
  const int cols = L.in + 1;
  const int o = (int)(r / cols), c = (int)(r - (long)o * cols);
  float g = 0.0f;
  for (int z = 0; z < L.nsplit; ++z) g = __fadd_rn(g, L.G[(size_t)z * L.slab + (size_t)o * L.ldg + c]);
  gsum[idx] = g;
}
void dense_reduce(Ctx* ctx, const DenseUpdateArgs& a, const StepStatus* st, float* gsum, const uint32_t* emb_counters) {
  dense_reduce_kernel<<<ceil_div(a.total + 1, 256), 256, 0, ctx->stream>>>(a, st, gsum, emb_counters);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

/* dense_reduce + all-gather by stores: every rank's slot `me` of gsum_in receives this rank's sums */
__global__ void __launch_bounds__(256) dense_reduce_send_kernel(const __grid_constant__ DenseUpdateArgs a, const StepStatus* __restrict__ st,
                                                                P2PState* __restrict__ p2p, const uint32_t* __restrict__ emb_counters) {
  const int R = p2p->R, me = p2p->me, glen = p2p->glen;
  /* (the scalars [loss, gbar, table-full] at the tail of the vector travelled earlier, right after the tail kernel: scalars_send) */
  /* a capped grid striding over the parameters: the publish below costs one system fence + one ticket per block */
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < a.total; idx += (long)gridDim.x * blockDim.x) {
    int li = 0;
    while (li + 1 < a.n_layers && idx >= a.l[li + 1].first) ++li;
    const DenseLayerDesc& L = a.l[li];
    const long r0 = idx - L.first;
    const int cols = L.in + 1;
    const int o = (int)(r0 / cols), c = (int)(r0 - (long)o * cols);
    float g = 0.0f;
    for (int z = 0; z < L.nsplit; ++z) g = __fadd_rn(g, L.G[(size_t)z * L.slab + (size_t)o * L.ldg + c]);
    for (int r = 0; r < R; ++r) reinterpret_cast<float*>(p2p_region(p2p, r, p2p->off_gsum))[(size_t)me * glen + idx] = g;
  }
  p2p_publish_last(p2p, CH_GSUM, gridDim.x);     /* the last block flags every replica: this rank's sums are in its gsum_in */
}
void dense_reduce_send(Ctx* ctx, const DenseUpdateArgs& a, const StepStatus* st, P2PState* p2p, const uint32_t* emb_counters) {
  dense_reduce_send_kernel<<<std::min(ceil_div(a.total + 1, 256), ctx->num_sms * 4), 256, 0, ctx->stream>>>(a, st, p2p, emb_counters);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

/* Right after the tail kernel: this rank's [loss, gbar, table-full] into the tail of its slot of every replica's gsum_in mailbox, on a
 * channel of its own — the global early-exit flag (and LRLayer's gbar) is known long before the weight gradients are, so the owner-side
 * embedding update does not wait for the dense gradient exchange */
__global__ void scalars_send_kernel(const StepStatus* __restrict__ st, P2PState* __restrict__ p2p, long total, const uint32_t* __restrict__ emb_counters) {
  const int r = threadIdx.x;
  if (r < p2p->R) {
    float* dst = reinterpret_cast<float*>(p2p_region(p2p, r, p2p->off_gsum)) + (size_t)p2p->me * p2p->glen;
    dst[total] = st->loss; dst[total + 1] = st->gbar;
    dst[total + 2] = (emb_counters != nullptr && emb_counters[1] != 0u) ? 1.0f : 0.0f;
  }
  p2p_publish_last(p2p, CH_SCAL, 1);
}
void scalars_send(Ctx* ctx, const StepStatus* st, P2PState* p2p, long total, const uint32_t* emb_counters) {
  scalars_send_kernel<<<1, 32, 0, ctx->stream>>>(st, p2p, total, emb_counters);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

/* global loss / gbar / early-exit flag from the R ranks' [loss, gbar] in the gsum_in mailbox (rank order) */
__global__ void shard_finish_scalars_p2p_kernel(StepStatus* st, const P2PState* p2p, long total) {
  p2p_wait_all(p2p, CH_SCAL);                    /* every replica's scalars have landed */
  if (threadIdx.x != 0) return;
  const float* in = reinterpret_cast<const float*>(p2p_region(p2p, p2p->me, p2p->off_gsum));
  float l = 0.f, g = 0.f, full = 0.f;
  for (int r = 0; r < p2p->R; ++r) {
    l = __fadd_rn(l, in[(size_t)r * p2p->glen + total]); g = __fadd_rn(g, in[(size_t)r * p2p->glen + total + 1]);
    full += in[(size_t)r * p2p->glen + total + 2];
  }
  const float loss = __fdiv_rn(l, (float)p2p->R);
  st->loss = loss;
  st->gbar = __fdiv_rn(g, (float)p2p->R);
  st->skip = (loss <= 0.01f || isnan(loss) || full != 0.f) ? 1 : 0;
}
void shard_finish_scalars_p2p(Ctx* ctx, StepStatus* st, const P2PState* p2p, long total) {
  shard_finish_scalars_p2p_kernel<<<1, 32, 0, ctx->stream>>>(st, p2p, total);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

__global__ void shard_finish_scalars_kernel(StepStatus* st, const float* tail, int R) {
  const float loss = __fdiv_rn(tail[0], (float)R);
  st->loss = loss;
  st->gbar = __fdiv_rn(tail[1], (float)R);
  st->skip = (loss <= 0.01f || isnan(loss) || tail[2] != 0.f) ? 1 : 0;      /* tail[2]: some rank's embedding shard is full */
}
void shard_finish_scalars(Ctx* ctx, StepStatus* st, const float* gsum_tail, int R) {
  shard_finish_scalars_kernel<<<1, 1, 0, ctx->stream>>>(st, gsum_tail, R);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

__global__ void updater_apply_kernel(UpdaterDev u, float* w, float* s1, float* s2, const float* g, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (u.kind == PS_UPD_FTRL && g[0] == 0.0f) return;
  float wv = w[i], a = s1[i], b = s2[i];
  apply_elem(u, wv, a, b, g[i]);
  w[i] = wv; s1[i] = a; s2[i] = b;
}
void updater_apply(Ctx* ctx, const UpdaterDev& u, float* w, float* s1, float* s2, const float* g, int n) {
  updater_apply_kernel<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(u, w, s1, s2, g, n);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

/* ------------------------------------------------------------------ tails */
constexpr int kTailThreads = 256;
constexpr int kTailMaxBlocks = 1024;

__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int t = threadIdx.x;
  sh[t] = v;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (t < s) sh[t] = __fadd_rn(sh[t], sh[t + s]);
    __syncthreads();
  }
  const float r = sh[0];
  __syncthreads();
  return r;
}

/* Each block reduces its samples, the block that takes the last ticket adds the per-block
 * partials in block order (a fixed tree: same bits every run) and writes the step status.
 * ws: [0, 2*kTailMaxBlocks) partial sums, then one u32 ticket that the finisher resets.       */
__device__ __forceinline__ void tail_finish(float loss_part, float d_part, int N, float* ws, StepStatus* st, float* sh, const uint32_t* emb_counters,
                                            bool ext_loss = false, float ext_loss_value = 0.f) {
  const float bl = block_sum(loss_part, sh), bd = block_sum(d_part, sh);
  __shared__ bool is_last;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(ws + 2 * kTailMaxBlocks);
  if (threadIdx.x == 0) {
    ws[2 * blockIdx.x] = bl; ws[2 * blockIdx.x + 1] = bd;
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  float l = 0.0f, d = 0.0f;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) { l = __fadd_rn(l, __ldcg(ws + 2 * b)); d = __fadd_rn(d, __ldcg(ws + 2 * b + 1)); }
  l = block_sum(l, sh); d = block_sum(d, sh);
  if (threadIdx.x == 0) {
    const float loss = ext_loss ? ext_loss_value : __fdiv_rn(l, (float)N);   /* ext_loss: the caller's own Loss.forward produced it (DNN.java:47) */
    st->loss = loss;
    st->gbar = __fdiv_rn(d, (float)N);
    /* DNN.java:58, CrossEntropy.slim — and a full embedding table: keys of this batch could not be created, so NOTHING of the
     * step is applied (backward, dense / wide / embedding updates all honour this flag); collect() reports PS_ERR_CAPACITY */
    const bool table_full = emb_counters != nullptr && emb_counters[1] != 0u;
    const bool bad_input = st->pad != 0u;        /* submit_text: a line of the batch could not be parsed — the batch is dropped (DataSet.java:96-98) */
    const bool slim = !ext_loss && (loss <= 0.01f || isnan(loss));   /* with the loss in the caller, the early exit of DNN.java:58-63 is the caller's too */
    st->skip = (slim || table_full || bad_input) ? 1 : 0;
    st->n_unique = emb_counters != nullptr ? emb_counters[4] : 0u;           /* EmbTable CNT_CURSOR: unique keys of this batch */
    st->seq += 1u;
    *ticket = 0u;
  }
}

/* Per-sample arithmetic follows the Java expressions operation by operation (double exp/log,
 * float elsewhere); the two batch sums (loss, rowMeans of delta) are tree reductions, so they
 * agree with the reference's running float sums to ~1e-7 relative, not bit for bit.           */
__global__ void __launch_bounds__(kTailThreads) tail_binary_kernel(int N, const float* __restrict__ zdeep, int ldz, const float* __restrict__ zwide,
                                                                   const float* __restrict__ Y, float* __restrict__ p_out, int ldp,
                                                                   float* __restrict__ d_out, int ldd, float* __restrict__ dt_out, int train,
                                                                   StepStatus* __restrict__ st, float* __restrict__ ws, const uint32_t* __restrict__ emb_counters) {
  __shared__ float sh[kTailThreads];
  float loss_part = 0.0f, d_part = 0.0f;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    float z = zdeep[(size_t)n * ldz];
    if (zwide) z = __fadd_rn(z, zwide[n]);                                   /* AddLayer.java:36 */
    const float p = sigmoid_clipped(z);                                      /* Sigmoid.java:11 */
    p_out[(size_t)n * ldp] = p;
    if (train) {
      const float l = Y[n];
      const float omp = __fsub_rn(1.0f, p);
      loss_part = __fadd_rn(loss_part, (float)((double)(-l) * log((double)p) - ((double)__fsub_rn(1.0f, l) * log((double)omp))));   /* CrossEntropy.java:15 */
      float d = __fdiv_rn(__fsub_rn(p, l), __fmul_rn(p, omp));                /* CrossEntropy.java:25 */
      d = __fmul_rn(d, __fmul_rn(p, omp));                                   /* Sigmoid.java:18 */
      d_out[(size_t)n * ldd] = d;
      if (dt_out) dt_out[n] = d;
      d_part = __fadd_rn(d_part, d);
    }
  }
  if (!train) return;
  tail_finish(loss_part, d_part, N, ws, st, sh, emb_counters);
}
/* The last layer's half of the reverse loop when the loss lives in the caller (DNN.java:47-49,64): delta_top = loss.backward(P, Y)
 * arrives from the host; FcLayer.backward's first act is activation.backward (FcLayer.java:100-102) = Sigmoid.backward
 * (Sigmoid.java:16-21: dy * y * (1 - y)); LRLayer.backward needs rowMeans of the result (LRLayer.java:110).               */
__global__ void __launch_bounds__(kTailThreads) tail_binary_from_delta_kernel(int N, const float* __restrict__ p, int ldp, const float* __restrict__ dtop,
                                                                              float* __restrict__ d_out, int ldd, float* __restrict__ dt_out, float loss,
                                                                              StepStatus* __restrict__ st, float* __restrict__ ws,
                                                                              const uint32_t* __restrict__ emb_counters) {
  __shared__ float sh[kTailThreads];
  float d_part = 0.0f;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    const float y = p[(size_t)n * ldp];
    const float d = __fmul_rn(dtop[n], __fmul_rn(y, __fsub_rn(1.0f, y)));
    d_out[(size_t)n * ldd] = d;
    if (dt_out) dt_out[n] = d;
    d_part = __fadd_rn(d_part, d);
  }
  tail_finish(0.0f, d_part, N, ws, st, sh, emb_counters, true, loss);
}

static int tail_blocks(int N) { return std::max(1, std::min(kTailMaxBlocks, ceil_div(N, kTailThreads))); }
void tail_binary(Ctx* ctx, int N, const float* zdeep, int ldz, const float* zwide, const float* Y, float* p_out, int ldp, float* d_out, int ldd,
                 float* dt_out, int train, StepStatus* st, float* ws, const uint32_t* emb_counters) {
  tail_binary_kernel<<<tail_blocks(N), kTailThreads, 0, ctx->stream>>>(N, zdeep, ldz, zwide, Y, p_out, ldp, d_out, ldd, dt_out, train, st, ws, emb_counters);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

__global__ void __launch_bounds__(kTailThreads) tail_softmax_kernel(int N, int C, float* __restrict__ Z, int ldz, const float* __restrict__ Y,
                                                                    float* __restrict__ d_out, int ldd, float* __restrict__ dt_out, int ldt, int train,
                                                                    StepStatus* __restrict__ st, float* __restrict__ ws, const uint32_t* __restrict__ emb_counters) {
  __shared__ float sh[kTailThreads];
  float loss_part = 0.0f;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    float* z = Z + (size_t)n * ldz;
    float mx = -INFINITY;
    for (int i = 0; i < C; ++i) { const float x = __fdiv_rn(z[i], 10000.0f); z[i] = x; mx = fmaxf(mx, x); }      /* Softmax.java:22-24 */
    float s = 0.0f;
    for (int i = 0; i < C; ++i) { const float e = (float)exp((double)__fsub_rn(z[i], mx)); z[i] = e; s = __fadd_rn(s, e); }   /* :26-33 */
    for (int i = 0; i < C; ++i) {
      float v = __fdiv_rn(z[i], s);
      if (v == 0.0f) v = 0.001f; else if (v == 1.0f) v = 0.999f;             /* :36-40 */
      z[i] = v;
    }
    if (train) {
      const int hot = (int)Y[n];                                             /* SoftmaxLoss.java:12 */
      const float ph = z[hot];
      loss_part = __fadd_rn(loss_part, (float)(-log((double)ph)));
      const float dy = __fdiv_rn(-1.0f, ph);                                 /* SoftmaxLoss.java:26 */
      float* d = d_out + (size_t)n * ldd;
      for (int k = 0; k < C; ++k) {                                          /* Softmax.java:51-63, one-hot dy */
        const float yk = z[k];
        const float t = (k == hot) ? __fmul_rn(yk, __fsub_rn(1.0f, yk)) : __fmul_rn(-ph, yk);
        d[k] = __fmul_rn(t, dy);
        if (dt_out) dt_out[(size_t)k * ldt + n] = d[k];
      }
    }
  }
  if (!train) return;
  tail_finish(loss_part, 0.0f, N, ws, st, sh, emb_counters);
}
void tail_binary_from_delta(Ctx* ctx, int N, const float* p, int ldp, const float* dtop, float* d_out, int ldd, float* dt_out, float loss,
                            StepStatus* st, float* ws, const uint32_t* emb_counters) {
  tail_binary_from_delta_kernel<<<tail_blocks(N), kTailThreads, 0, ctx->stream>>>(N, p, ldp, dtop, d_out, ldd, dt_out, loss, st, ws, emb_counters);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

/* FullConnectedNN with the loss in the caller: delta_top[N][C] = SoftmaxLoss.backward(P, Y) (or any dy) -> Softmax.backward
 * (Softmax.java:45-67) operation by operation, including its `delta[k] *= d` INSIDE the class loop (for a one-hot dy — what
 * SoftmaxLoss produces — this is the usual Jacobian product; for several non-zero entries the reference's own arithmetic is kept) */
__global__ void __launch_bounds__(kTailThreads) tail_softmax_from_delta_kernel(int N, int C, const float* __restrict__ P, int ldp, const float* __restrict__ dtop,
                                                                               float* __restrict__ d_out, int ldd, float* __restrict__ dt_out, int ldt, float loss,
                                                                               StepStatus* __restrict__ st, float* __restrict__ ws) {
  __shared__ float sh[kTailThreads];
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    const float* y = P + (size_t)n * ldp;
    float* d = d_out + (size_t)n * ldd;
    for (int k = 0; k < C; ++k) d[k] = 0.0f;
    for (int j = 0; j < C; ++j) {
      const float dj = dtop[(size_t)n * C + j];
      if (dj == 0.0f) continue;
      for (int k = 0; k < C; ++k) {
        const float t = (j == k) ? __fmul_rn(y[k], __fsub_rn(1.0f, y[k])) : __fmul_rn(-y[j], y[k]);
        d[k] = __fmul_rn(__fadd_rn(d[k], t), dj);
      }
    }
    if (dt_out) for (int k = 0; k < C; ++k) dt_out[(size_t)k * ldt + n] = d[k];
  }
  tail_finish(0.0f, 0.0f, N, ws, st, sh, nullptr, true, loss);
}
void tail_softmax_from_delta(Ctx* ctx, int N, int C, const float* P, int ldp, const float* dtop, float* d_out, int ldd, float* dt_out, int ldt, float loss,
                             StepStatus* st, float* ws) {
  tail_softmax_from_delta_kernel<<<tail_blocks(N), kTailThreads, 0, ctx->stream>>>(N, C, P, ldp, dtop, d_out, ldd, dt_out, ldt, loss, st, ws);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

void tail_softmax(Ctx* ctx, int N, int C, float* Z, int ldz, const float* Y, float* d_out, int ldd, float* dt_out, int ldt, int train,
                  StepStatus* st, float* ws) {
  tail_softmax_kernel<<<tail_blocks(N), kTailThreads, 0, ctx->stream>>>(N, C, Z, ldz, Y, d_out, ldd, dt_out, ldt, train, st, ws, nullptr);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

/* ------------------------------------------------------------------ the 1-unit top FcLayer on CUDA cores
 * With out = 1 (FcLayer.build's last layer in DNN / WideDeepNN, FcLayer.java:53-70) the three
 * contractions are a GEMV, a rank-1 outer product and a weighted column sum: memory-trivial, no
 * tensor-core shape.  Fusing the forward GEMV with the tail takes two launches off the critical path. */
__global__ void __launch_bounds__(kTailThreads) fc1_forward_tail_kernel(int N, int in, const float* __restrict__ A, int lda, const float* __restrict__ w,
                                                                        const float* __restrict__ bias, const float* __restrict__ zwide,
                                                                        const float* __restrict__ Y, float* __restrict__ z_out, int ldz,
                                                                        float* __restrict__ p_out, int ldp, float* __restrict__ d_out, int ldd,
                                                                        float* __restrict__ dt_out, int train, StepStatus* __restrict__ st,
                                                                        float* __restrict__ ws, const uint32_t* __restrict__ emb_counters) {
  __shared__ float sh[kTailThreads];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const bool vec = (in % 4 == 0) && (lda % 4 == 0);
  float loss_part = 0.0f, d_part = 0.0f;
  /* one warp per sample: a cooperative 128-bit dot product (one L2 round trip), then lane 0 runs the
   * scalar tail; 8 samples per block keep thousands of warps in flight instead of a serial chain */
  for (int n = warp; n < N; n += nwarps) {
    const float* a = A + (size_t)n * lda;
    float acc = 0.0f;
    if (vec) {
      for (int c = lane * 4; c < in; c += 128) {
        const float4 x = ld_f4(a + c), y = __ldg(reinterpret_cast<const float4*>(w + c));
        acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
      }
    } else {
      for (int c = lane; c < in; c += 32) acc = fmaf(a[c], __ldg(w + c), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      float z = __fadd_rn(acc, bias[0]);                                       /* FcLayer.java:76-77 */
      z_out[(size_t)n * ldz] = z;
      if (zwide) z = __fadd_rn(z, zwide[n]);                                   /* AddLayer.java:36 */
      const float p = sigmoid_clipped(z);                                      /* Sigmoid.java:11 */
      p_out[(size_t)n * ldp] = p;
      if (train) {
        const float l = Y[n];
        const float omp = __fsub_rn(1.0f, p);
        loss_part = __fadd_rn(loss_part, (float)((double)(-l) * log((double)p) - ((double)__fsub_rn(1.0f, l) * log((double)omp))));
        float d = __fdiv_rn(__fsub_rn(p, l), __fmul_rn(p, omp));                /* CrossEntropy.java:25 */
        d = __fmul_rn(d, __fmul_rn(p, omp));                                   /* Sigmoid.java:18 */
        d_out[(size_t)n * ldd] = d;
        if (dt_out) dt_out[n] = d;
        d_part = __fadd_rn(d_part, d);
      }
    }
  }
  if (!train) return;
  tail_finish(loss_part, d_part, N, ws, st, sh, emb_counters);
}
void fc1_forward_tail(Ctx* ctx, int N, int in, const float* A, int lda, const float* w, const float* bias, const float* zwide, const float* Y,
                      float* z_out, int ldz, float* p_out, int ldp, float* d_out, int ldd, float* dt_out, int train, StepStatus* st, float* ws,
                      const uint32_t* emb_counters) {
  const int blocks = std::max(1, std::min(kTailMaxBlocks, ceil_div((long)N * 32, kTailThreads)));
  fc1_forward_tail_kernel<<<blocks, kTailThreads, 0, ctx->stream>>>(N, in, A, lda, w, bias, zwide, Y, z_out, ldz, p_out, ldp, d_out, ldd, dt_out, train, st, ws, emb_counters);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

/* dX[b][i] = d[b] * w[i] times the activation derivative below; 4 elements per thread.  Work items
 * [0, N*in/4): row-major output; [N*in/4, ...): the transposed copy (TF32 path), read from the
 * transposed activation so both halves are fully coalesced.                                       */
template <bool VEC>
__global__ void __launch_bounds__(256) fc1_dgrad_kernel(int N, int in, const float* __restrict__ d, int ldd, const float* __restrict__ dT,
                                                        const float* __restrict__ w, int act_below, const float* __restrict__ Y, int ldy,
                                                        float* __restrict__ dX, int ldx, const float* __restrict__ Yt, int ldyt,
                                                        float* __restrict__ dXt, int ldxt) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    const int in4 = in >> 2, n4 = (N + 3) >> 2;
    const long rowwork = (long)N * in4;
    if (g < rowwork) {
      const int n = (int)(g / in4), i = (int)(g - (long)n * in4) << 2;
      const float dv = d[(size_t)n * ldd];
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + i));
      float4 y = make_float4(1.f, 1.f, 1.f, 1.f);
      if (act_below != PS_ACT_NONE) y = ld_f4(Y + (size_t)n * ldy + i);
      float4 v;
      v.x = act_backward(act_below, __fmul_rn(wv.x, dv), y.x); v.y = act_backward(act_below, __fmul_rn(wv.y, dv), y.y);
      v.z = act_backward(act_below, __fmul_rn(wv.z, dv), y.z); v.w = act_backward(act_below, __fmul_rn(wv.w, dv), y.w);
      st_f4(dX + (size_t)n * ldx + i, v);
    } else if (dXt != nullptr && g < rowwork + (long)in * n4) {
      const long h = g - rowwork;
      const int i = (int)(h / n4), n = (int)(h - (long)i * n4) << 2;
      const float wi = __ldg(w + i);
      float dv[4], yv[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
      for (int k = 0; k < 4; ++k) dv[k] = n + k < N ? (dT ? dT[n + k] : d[(size_t)(n + k) * ldd]) : 0.f;
      if (act_below != PS_ACT_NONE) {
#pragma unroll
        for (int k = 0; k < 4; ++k) if (n + k < N) yv[k] = Yt[(size_t)i * ldyt + n + k];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) if (n + k < N) dXt[(size_t)i * ldxt + n + k] = act_backward(act_below, __fmul_rn(wi, dv[k]), yv[k]);
    }
  } else {
    const long total = (long)N * in;
    if (g < total) {
      const int n = (int)(g / in), i = (int)(g - (long)n * in);
      const float v = __fmul_rn(w[i], d[(size_t)n * ldd]);
      dX[(size_t)n * ldx + i] = act_backward(act_below, v, act_below != PS_ACT_NONE ? Y[(size_t)n * ldy + i] : 1.f);
    } else if (dXt != nullptr && g < 2 * total) {
      const long h = g - total;
      const int i = (int)(h / N), n = (int)(h - (long)i * N);
      const float v = __fmul_rn(w[i], d[(size_t)n * ldd]);
      dXt[(size_t)i * ldxt + n] = act_backward(act_below, v, act_below != PS_ACT_NONE ? Yt[(size_t)i * ldyt + n] : 1.f);
    }
  }
}
void fc1_dgrad(Ctx* ctx, int N, int in, const float* d, int ldd, const float* dT, const float* w, int act_below, const float* Y, int ldy, float* dX,
               int ldx, const float* Yt, int ldyt, float* dXt, int ldxt) {
  const bool vec = (in % 4 == 0) && (ldy % 4 == 0) && (ldx % 4 == 0);
  if (vec) {
    const long total = (long)N * (in / 4) + (dXt ? (long)in * ((N + 3) / 4) : 0);
    fc1_dgrad_kernel<true><<<ceil_div(total, 256), 256, 0, ctx->stream>>>(N, in, d, ldd, dT, w, act_below, Y, ldy, dX, ldx, Yt, ldyt, dXt, ldxt);
  } else {
    const long total = (long)N * in * (dXt ? 2 : 1);
    fc1_dgrad_kernel<false><<<ceil_div(total, 256), 256, 0, ctx->stream>>>(N, in, d, ldd, dT, w, act_below, Y, ldy, dX, ldx, Yt, ldyt, dXt, ldxt);
  }
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

/* G[z][c] = sum over the z-th batch chunk of d[b] * [A | 1][b][c]; 32 columns x 8 row lanes per block */
__global__ void __launch_bounds__(256) fc1_wgrad_kernel(int N, int cols, const float* __restrict__ d, int ldd, const float* __restrict__ A, int lda,
                                                        float* __restrict__ G, size_t slab, int chunk) {
  __shared__ float sh[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int b0 = blockIdx.y * chunk, b1 = min(N, b0 + chunk);
  float acc = 0.0f;
  if (c < cols)
    for (int b = b0 + ry; b < b1; b += 8) acc = fmaf(d[(size_t)b * ldd], A[(size_t)b * lda + c], acc);
  sh[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float s = 0.0f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += sh[r][cx];
    G[(size_t)blockIdx.y * slab + c] = s;
  }
}
void fc1_wgrad(Ctx* ctx, int N, int in, const float* d, int ldd, const float* A, int lda, float* G, size_t slab, int nsplit) {
  const int chunk = ceil_div(N, nsplit);
  dim3 grid(ceil_div(in + 1, 32), nsplit);
  fc1_wgrad_kernel<<<grid, 256, 0, ctx->stream>>>(N, in + 1, d, ldd, A, lda, G, slab, chunk);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

__global__ void publish_status_kernel(StepStatus* st, const uint32_t* emb_counters, const uint32_t* wide_counters, StepStatus* host) {
  publish(st, emb_counters, wide_counters, host);
}
void publish_status(Ctx* ctx, StepStatus* st, const uint32_t* emb_counters, const uint32_t* wide_counters, StepStatus* host_mapped) {
  publish_status_kernel<<<1, 1, 0, ctx->stream>>>(st, emb_counters, wide_counters, host_mapped);
  PS_LAUNCH_CHECK();
  ctx->launches++;
}

}  // namespace psb

/*
 * fcop.cu — layer.FcLayer as a standalone operator (FcLayer.java:34-115), for callers that drive the reference's layer
 * list themselves instead of handing the whole step to ps_model_train_step:
 *   forward   Z = W * A_prev + b 1^T ; A = act(Z)                                   FcLayer.java:74-91
 *   backward  delta <- act'(delta) ; db = rowMeans(delta) ; dW = delta * A_prev^T / N ; delta_prev = W^T * delta
 *             (db, dW are what the reference hands to KVStore.sum)                   FcLayer.java:93-110
 *   update    KVStore.update(updaters) + clear() for "<name>.weights" / "<name>.bias"   KVStore.java:240-277
 * Same kernels as the fused step (gemm_tc.cu tcgen05 / gemm_simt.cu FFMA, dense_update); host matrices are jblas
 * column-major (features x N), device buffers batch-major with the constant-1 column that yields db in the wgrad.
 */
#include "fcop.cuh"

namespace psb {

__global__ void __launch_bounds__(256) act_backward_kernel(int act, float* __restrict__ d, int ldd, const float* __restrict__ y, int ldy, int N, int out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * out) return;
  const int n = (int)(i / out), o = (int)(i - (long)n * out);
  d[(size_t)n * ldd + o] = act_backward(act, d[(size_t)n * ldd + o], y[(size_t)n * ldy + o]);
}

void FcOp::create(Ctx* c, const std::string& name, int in, int out, int act, const ps_updater_spec& upd, int max_batch) {
  PS_REQUIRE(in > 0 && out > 0 && max_batch > 0, PS_ERR_ARG, "fc: bad dims");
  PS_REQUIRE(act == PS_ACT_NONE || act == PS_ACT_RELU || act == PS_ACT_SIGMOID, PS_ERR_ARG, "fc: activation must be none, relu or sigmoid");
  ctx = c; Bmax = max_batch;
  f.create(ctx, name, in, out, act, upd, 8);
  ldA = round_up(in + 1, 8); ldZ = round_up(out + 1, 8); ldt = round_up(Bmax, 4);
  cudaStream_t s = ctx->stream;
  A = dmalloc_zero<float>((size_t)Bmax * ldA, s);
  At = dmalloc_zero<float>((size_t)(in + 1) * ldt, s);
  Z = dmalloc_zero<float>((size_t)Bmax * ldZ, s);
  dl = dmalloc_zero<float>((size_t)Bmax * ldZ, s);
  dlT = dmalloc_zero<float>((size_t)out * ldt, s);
  dX = dmalloc_zero<float>((size_t)Bmax * ldA, s);
  st = dmalloc_zero<StepStatus>(1, s);
  fill_column(ctx, A, ldA, in, Bmax, 1.0f);                       /* [A | 1]: column `in` of the wgrad is the bias-gradient sum */
  fill_column(ctx, At + (size_t)in * ldt, 1, 0, ldt, 1.0f);
  fc_tf32_init();
  PS_CUDA(cudaStreamSynchronize(s));
}

void FcOp::destroy() {
  if (!ctx) return;
  cudaStreamSynchronize(ctx->stream);
  f.destroy();
  dfree(A); dfree(At); dfree(Z); dfree(dl); dfree(dlT); dfree(dX); dfree(st);
  ctx = nullptr;
}

void FcOp::forward(const float* A_host, int N, float* out_host) {
  PS_REQUIRE(A_host && out_host && N > 0 && N <= Bmax, PS_ERR_ARG, "fc forward: bad argument");
  cudaStream_t s = ctx->stream;
  const bool fp32 = ctx->fc_precision == PS_FC_FP32;
  PS_CUDA(cudaMemcpy2DAsync(A, sizeof(float) * ldA, A_host, sizeof(float) * f.in, sizeof(float) * f.in, N, cudaMemcpyHostToDevice, s));
  FcFwdArgs a{};
  a.B = N; a.in = f.in; a.out = f.out; a.A = A; a.lda = ldA; a.W = f.W; a.ldw = f.ldw; a.Wlo = f.Wlo; a.bias = f.bias; a.act = f.act;
  a.Z = Z; a.ldz = ldZ; a.Zt = nullptr; a.ldzt = ldt;
  if (fp32) fc_forward_fp32(ctx, a); else fc_forward_tf32(ctx, a);
  PS_CUDA(cudaMemcpy2DAsync(out_host, sizeof(float) * f.out, Z, sizeof(float) * ldZ, sizeof(float) * f.out, N, cudaMemcpyDeviceToHost, s));
  PS_CUDA(cudaStreamSynchronize(s));
  lastN = N; has_grad = false;
}

void FcOp::backward(const float* delta_host, int N, float* dprev_host) {
  PS_REQUIRE(delta_host && N > 0, PS_ERR_ARG, "fc backward: bad argument");
  PS_REQUIRE(N == lastN, PS_ERR_STATE, "fc backward: no matching forward");
  PS_REQUIRE(!has_grad, PS_ERR_STATE, "fc backward: a gradient is already pending (call ps_fc_update first)");
  cudaStream_t s = ctx->stream;
  const bool fp32 = ctx->fc_precision == PS_FC_FP32;
  PS_CUDA(cudaMemcpy2DAsync(dl, sizeof(float) * ldZ, delta_host, sizeof(float) * f.out, sizeof(float) * f.out, N, cudaMemcpyHostToDevice, s));
  if (f.act != PS_ACT_NONE) {                                   /* delta = activation.backward(delta, Z, A)  (FcLayer.java:100-102) */
    act_backward_kernel<<<ceil_div((long)N * f.out, 256), 256, 0, s>>>(f.act, dl, ldZ, Z, ldZ, N, f.out);
    PS_LAUNCH_CHECK();
    ctx->launches++;
  }
  if (!fp32) {                                                  /* K-major operands of the tcgen05 wgrad */
    transpose_copy(ctx, A, ldA, At, ldt, N, f.in);
    transpose_copy(ctx, dl, ldZ, dlT, ldt, N, f.out);
  }
  FcWgradArgs g{};
  g.B = N; g.in = f.in; g.out = f.out; g.dl = dl; g.ldd = ldZ; g.A = A; g.lda = ldA; g.dlT = dlT; g.AT = At; g.ldt = ldt;
  g.G = f.G; g.ldg = f.ldw; g.slab = (size_t)f.out * f.ldw; g.nsplit = f.nsplit;
  if (fp32) fc_wgrad_fp32(ctx, g); else fc_wgrad_tf32(ctx, g);
  FcDgradArgs d{};
  d.B = N; d.in = f.in; d.out = f.out; d.dl = dl; d.ldd = ldZ; d.W = f.W; d.ldw = f.ldw; d.Wt = f.Wt; d.ldwt = f.ldwt; d.Wtlo = f.Wtlo;
  d.act_below = PS_ACT_NONE; d.Y = A; d.ldy = ldA; d.Yt = At; d.ldyt = ldt; d.n_cols = f.in; d.dX = dX; d.ldx = ldA; d.dXt = nullptr; d.ldxt = ldt;
  if (fp32) fc_dgrad_fp32(ctx, d); else fc_dgrad_tf32(ctx, d);
  if (dprev_host) PS_CUDA(cudaMemcpy2DAsync(dprev_host, sizeof(float) * f.in, dX, sizeof(float) * ldA, sizeof(float) * f.in, N, cudaMemcpyDeviceToHost, s));
  PS_CUDA(cudaStreamSynchronize(s));
  has_grad = true;
}

void FcOp::update() {
  PS_REQUIRE(has_grad, PS_ERR_STATE, "fc update: no pending gradient");
  DenseUpdateArgs u{};
  u.n_layers = 1; u.N = lastN;
  DenseLayerDesc& q = u.l[0];
  q.W = f.W; q.Wt = f.Wt; q.Wlo = f.Wlo; q.Wtlo = f.Wtlo; q.bias = f.bias; q.sW1 = f.sW1; q.sW2 = f.sW2; q.sb1 = f.sb1; q.sb2 = f.sb2;
  q.G = f.G; q.slab = (size_t)f.out * f.ldw; q.nsplit = f.nsplit; q.out = f.out; q.in = f.in; q.ldw = f.ldw; q.ldwt = f.ldwt; q.ldg = f.ldw;
  q.updW = make_updater_dev(f.updW); q.updB = make_updater_dev(f.updB);
  q.first = 0;
  u.total = (long)f.out * (f.in + 1);
  dense_update(ctx, u, st, nullptr, nullptr, nullptr);
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  has_grad = false;
}

/* the gradients the reference hands to KVStore.sum: dW (out x in, column-major) = delta * A_prev^T / N, db = rowMeans(delta) */
void FcOp::gradients(float* dW_host, float* db_host) {
  PS_REQUIRE(has_grad, PS_ERR_STATE, "fc gradients: no pending gradient");
  const size_t slab = (size_t)f.out * f.ldw;
  std::vector<float> tmp(slab * f.nsplit);
  PS_CUDA(cudaMemcpyAsync(tmp.data(), f.G, sizeof(float) * tmp.size(), cudaMemcpyDeviceToHost, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int o = 0; o < f.out; ++o)
    for (int c = 0; c <= f.in; ++c) {
      float sum = 0.f;
      for (int k = 0; k < f.nsplit; ++k) sum += tmp[k * slab + (size_t)o * f.ldw + c];   /* same slab order as the update kernel */
      const float mean = sum / (float)lastN;
      if (c < f.in) { if (dW_host) dW_host[(size_t)o + (size_t)f.out * c] = mean; }
      else if (db_host) db_host[o] = mean;
    }
}

/* KVStore.get / put of this layer's keys: weights as jblas out x in column-major (index(o, i) = o + out * i) */
void FcOp::get(int which, std::vector<float>& out) {
  cudaStream_t s = ctx->stream;
  if (which == 1) {
    out.resize(f.out);
    PS_CUDA(cudaMemcpyAsync(out.data(), f.bias, sizeof(float) * f.out, cudaMemcpyDeviceToHost, s));
    PS_CUDA(cudaStreamSynchronize(s));
    return;
  }
  std::vector<float> tmp((size_t)f.out * f.ldw);
  PS_CUDA(cudaMemcpyAsync(tmp.data(), f.W, sizeof(float) * tmp.size(), cudaMemcpyDeviceToHost, s));
  PS_CUDA(cudaStreamSynchronize(s));
  out.resize((size_t)f.out * f.in);
  for (int i = 0; i < f.in; ++i)
    for (int o = 0; o < f.out; ++o) out[(size_t)o + (size_t)f.out * i] = tmp[(size_t)o * f.ldw + i];
}

void FcOp::put(int which, const float* in, int n) {
  cudaStream_t s = ctx->stream;
  if (which == 1) {
    PS_REQUIRE(n == f.out, PS_ERR_ARG, "fc put: bias length mismatch");
    PS_CUDA(cudaMemcpyAsync(f.bias, in, sizeof(float) * n, cudaMemcpyHostToDevice, s));
    PS_CUDA(cudaStreamSynchronize(s));
    return;
  }
  PS_REQUIRE(n == f.out * f.in, PS_ERR_ARG, "fc put: weight length mismatch");
  std::vector<float> tmp((size_t)f.out * f.ldw, 0.f), tmpt((size_t)f.in * f.ldwt, 0.f);
  for (int i = 0; i < f.in; ++i)
    for (int o = 0; o < f.out; ++o) {
      const float v = in[(size_t)o + (size_t)f.out * i];
      tmp[(size_t)o * f.ldw + i] = v; tmpt[(size_t)i * f.ldwt + o] = v;
    }
  PS_CUDA(cudaMemcpyAsync(f.W, tmp.data(), sizeof(float) * tmp.size(), cudaMemcpyHostToDevice, s));
  PS_CUDA(cudaMemcpyAsync(f.Wt, tmpt.data(), sizeof(float) * tmpt.size(), cudaMemcpyHostToDevice, s));
  f.refresh_lo(ctx);
  PS_CUDA(cudaStreamSynchronize(s));
}

}  // namespace psb

/*
 * model.cu — model.DNN / model.WideDeepNN / model.FullConnectedNN wired to the GPU store and
 * stepped the way train.Trainer steps them with thread = 1 (TrainerThread.java:29-39,
 * Trainer.java:70-101): pullWeights → forward loop → loss → reverse loop → KVStore.update →
 * KVStore.clear, as ONE ordered sequence of kernels on one stream.
 *
 * Layer order of the reference and where it went:
 *   EmbeddingLayer.forward + ConcatLayer.forward   → EmbTable::probe + gather straight into the
 *        concat buffer act[0] = [emb | X | 1]       (EmbeddingLayer.java:25-48, ConcatLayer.java:30-37)
 *   FcLayer.forward ×L                             → fc_forward_*        (FcLayer.java:74-91)
 *   LRLayer.forward                                → WideTable::forward  (LRLayer.java:62-98)
 *   AddLayer + Sigmoid + CrossEntropy fwd/bwd      → tail_binary         (AddLayer.java:33-61, CrossEntropy.java:10-28)
 *   FcLayer.backward ×L                            → fc_wgrad_* + fc_dgrad_* (FcLayer.java:93-110)
 *   EmbeddingLayer.backward ×2 + KVStore.update    → EmbTable::scatter_update
 *   LRLayer.backward + KVStore.update              → WideTable::update_all + wide_bias_update
 *   KVStore.update for fc keys                     → dense_update
 */
#include "model.cuh"

#include <algorithm>
#include <cmath>
#include <exception>
#include <cstddef>
#include <cstdlib>

namespace psb {

static float xavier(int in, int out) { /* FcLayer.java:39,46 / EmbeddingField.java:40 */
  return (float)(4 * (std::sqrt(6.0) / std::sqrt((double)(in + out))));
}

void FcLayer::create(Ctx* ctx, const std::string& nm, int in_, int out_, int act_, const ps_updater_spec& u, int nsplit_) {
  name = nm; in = in_; out = out_; act = act_; nsplit = nsplit_;
  ldw = round_up(in + 1, 8); ldwt = round_up(out, 8);
  updW = u; updB = u;
  cudaStream_t s = ctx->stream;
  W = dmalloc_zero<float>((size_t)out * ldw, s);
  Wt = dmalloc_zero<float>((size_t)in * ldwt, s);
  Wlo = dmalloc_zero<float>((size_t)out * ldw, s);
  Wtlo = dmalloc_zero<float>((size_t)in * ldwt, s);
  bias = dmalloc_zero<float>(out, s);
  sW1 = dmalloc_zero<float>((size_t)out * ldw, s); sW2 = dmalloc_zero<float>((size_t)out * ldw, s);
  sb1 = dmalloc_zero<float>(out, s); sb2 = dmalloc_zero<float>(out, s);
  G = dmalloc_zero<float>((size_t)nsplit * out * ldw, s);
  /* FcLayer.pullWeights (FcLayer.java:112-115): KVStore.get(name + ".weights", initW) creates on first use */
  dense_init(ctx, W, out, in, ldw, Wt, ldwt, ps_name_key((name + ".weights").c_str()), xavier(in, out));
  dense_init(ctx, bias, out, 1, 1, nullptr, 0, ps_name_key((name + ".bias").c_str()), xavier(in, 1));
  refresh_lo(ctx);
}
void FcLayer::refresh_lo(Ctx* ctx) {
  split_lo(ctx, W, Wlo, (size_t)out * ldw);
  split_lo(ctx, Wt, Wtlo, (size_t)in * ldwt);
}
void FcLayer::destroy() {
  dfree(W); dfree(Wt); dfree(Wlo); dfree(Wtlo); dfree(bias); dfree(sW1); dfree(sW2); dfree(sb1); dfree(sb2); dfree(G);
  W = Wt = Wlo = Wtlo = bias = sW1 = sW2 = sb1 = sb2 = G = nullptr;
}

static ps_updater_spec mk_spec(int kind, float a, float b, float c, float d) {
  ps_updater_spec s; s.kind = kind; s.p[0] = a; s.p[1] = b; s.p[2] = c; s.p[3] = d; return s;
}

void Model::create(Ctx* c, int kind_, int F_, int D_, int Xn_, const int32_t* fc_dims, int n_fc, int64_t emb_capacity,
                   const ps_updater_spec* emb_updater, int max_batch) {
  PS_REQUIRE(kind_ >= PS_MODEL_DNN && kind_ <= PS_MODEL_FCNN, PS_ERR_ARG, "model: unknown kind");
  PS_REQUIRE(n_fc >= 1 && n_fc <= kMaxDenseLayers && max_batch > 0 && Xn_ > 0, PS_ERR_ARG, "model: bad dims");
  ctx = c; kind = kind_; Xn = Xn_; Bmax = max_batch; L = n_fc;
  has_emb = kind != PS_MODEL_FCNN; has_wide = kind == PS_MODEL_WIDEDEEP;
  F = has_emb ? F_ : 0; D = has_emb ? D_ : 0;
  if (has_emb) PS_REQUIRE(F > 0 && D > 0, PS_ERR_ARG, "model: need F > 0 and D > 0");
  if (kind != PS_MODEL_FCNN) PS_REQUIRE(fc_dims[n_fc - 1] == 1, PS_ERR_ARG, "model: DNN/WideDeepNN end in a 1-unit layer");
  upd_default = mk_spec(PS_UPD_ADAM, (float)0.005, (float)0.9, (float)0.999, (float)std::pow(10.0, -8));   /* DNN.java:95 */
  upd_wide = mk_spec(PS_UPD_FTRL, 0.005f, 1.0f, 0.001f, 0.001f);                                           /* WideDeepNN.java:109 */
  cudaStream_t s = ctx->stream;

  width.assign(L + 1, 0); ld.assign(L + 1, 0);
  width[0] = has_emb ? F * D + Xn : Xn;
  for (int l = 0; l < L; ++l) width[l + 1] = fc_dims[l];
  act.assign(L + 1, nullptr); delta.assign(L + 1, nullptr); act_t.assign(L + 1, nullptr); delta_t.assign(L + 1, nullptr);
  ldt = round_up(Bmax, 4);
  for (int l = 0; l <= L; ++l) {
    ld[l] = round_up(width[l] + 1, 8);
    act[l] = dmalloc_zero<float>((size_t)Bmax * ld[l], s);
    delta[l] = dmalloc_zero<float>((size_t)Bmax * ld[l], s);
    act_t[l] = dmalloc_zero<float>((size_t)(width[l] + 1) * ldt, s);
    delta_t[l] = dmalloc_zero<float>((size_t)width[l] * ldt, s);
    if (l < L) {                     /* the constant-1 column (row of the transposed copy) that yields db in wgrad */
      fill_column(ctx, act[l], ld[l], width[l], Bmax, 1.0f);
      fill_column(ctx, act_t[l] + (size_t)width[l] * ldt, 1, 0, ldt, 1.0f);
    }
  }
  fcs.resize(L);
  for (int l = 0; l < L; ++l)   /* FcLayer.build (FcLayer.java:53-70): ReLU inside, the top activation belongs to the tail */
    fcs[l].create(ctx, "fc" + std::to_string(l), width[l], width[l + 1], l + 1 == L ? PS_ACT_NONE : PS_ACT_RELU, upd_default, 8);

  if (has_emb) emb.create(ctx, F, D, emb_capacity, emb_updater ? *emb_updater : upd_default, (int64_t)Bmax * F);
  if (has_wide) {
    wide.create(ctx, 1 << 18, upd_wide);                   /* CTR.wideSize = 100000 hashed ids (CTR.java:36) */
    wide_bias = dmalloc_zero<float>(4, s);                 /* LRLayer ctor: zeros(1) (LRLayer.java:45-52) */
    wide_z = dmalloc_zero<float>(Bmax, s);
    P = dmalloc_zero<float>(Bmax, s);
  }
  st_dev = dmalloc_zero<StepStatus>(1, s);
  tail_ws = dmalloc_zero<float>(kTailWorkspaceFloats, s);
  ev_pool.resize(48);
  for (auto& e : ev_pool) PS_CUDA(cudaEventCreate(&e));
  for (auto& a : aux) PS_CUDA(cudaStreamCreateWithPriority(&a, cudaStreamNonBlocking, ctx->prio_side));
  for (auto& e : sync_ev) PS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  fc_tf32_init();
  for (auto& S : stage) {
    if (has_emb) S.E = dmalloc<int64_t>((size_t)Bmax * F);
    if (has_wide) S.W = dmalloc<int64_t>((size_t)Bmax * F);
    S.X = dmalloc<float>((size_t)Bmax * Xn);
    S.Y = dmalloc<float>(Bmax);
    PS_CUDA(cudaHostAlloc(&S.st_host, sizeof(StepStatus), cudaHostAllocMapped));
    std::memset(S.st_host, 0, sizeof(StepStatus));
    PS_CUDA(cudaEventCreateWithFlags(&S.h2d_done, cudaEventDisableTiming));
    PS_CUDA(cudaEventCreateWithFlags(&S.step_done, cudaEventDisableTiming));
  }
  PS_CUDA(cudaStreamSynchronize(s));
}

void Model::destroy() {
  if (!ctx) return;
  cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->copy_stream);
  for (auto& f : fcs) f.destroy();
  for (auto p : act) dfree(p);
  for (auto p : delta) dfree(p);
  for (auto p : act_t) dfree(p);
  for (auto p : delta_t) dfree(p);
  if (has_emb) emb.destroy();
  if (has_wide) { wide.destroy(); dfree(wide_bias); dfree(wide_z); dfree(P); }
  dfree(st_dev); dfree(tail_ws); dfree(gsum); dfree(send_pos); dfree(dtop_stage); dtop_stage = nullptr;
  if (p2p.slab) p2p.destroy();
  for (auto& g : graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
  graphs.clear();
  for (auto& S : stage) {
    dfree(S.E); dfree(S.W); dfree(S.X); dfree(S.Y); dfree(S.text); dfree(S.text_ws); dfree(S.text_status);
    if (S.st_host) cudaFreeHost(S.st_host);
    if (S.h2d_done) cudaEventDestroy(S.h2d_done);
    if (S.step_done) cudaEventDestroy(S.step_done);
  }
  for (auto e : ev_pool) cudaEventDestroy(e);
  ev_pool.clear();
  for (auto& a : aux) if (a) { cudaStreamSynchronize(a); cudaStreamDestroy(a); a = nullptr; }
  for (auto& e : sync_ev) if (e) { cudaEventDestroy(e); e = nullptr; }
  ctx = nullptr;
}

void Model::mark(const char* phase) {
  if (!profile || ev_n >= (int)ev_pool.size()) return;
  PS_CUDA(cudaEventRecordWithFlags(ev_pool[ev_n], ctx->stream, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
  ev_n++;
  ev_names.push_back(phase);
}
void Model::finish_profile() {
  if (!profile || phase_names.size() < 2) return;
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  phase_ms.clear();
  for (size_t i = 1; i < phase_names.size(); ++i) {
    float ms = 0.f;
    PS_CUDA(cudaEventElapsedTime(&ms, ev_pool[i - 1], ev_pool[i]));
    phase_ms.push_back(ms);
  }
}

static const int* skip_ptr(const StepStatus* st) {
  return reinterpret_cast<const int*>(reinterpret_cast<const char*>(st) + offsetof(StepStatus, skip));
}
static const float* gbar_ptr(const StepStatus* st) {
  return reinterpret_cast<const float*>(reinterpret_cast<const char*>(st) + offsetof(StepStatus, gbar));
}

void Model::run_step(const int64_t* E, const float* X, const int64_t* W, const float* Y, int N, bool train, StepStatus* publish_to, int mode) {
  PS_REQUIRE(N > 0 && N <= Bmax, PS_ERR_ARG, "model: batch size must be in [1, max_batch]");
  if (mode != 1) flush_deferred();
  const int consume_N = (mode == 1 && deferred_pending) ? deferred_N : 0;
  const bool defer = mode == 1 && (ctx->p2p_defer == 1 || (ctx->p2p_defer == 2 && (!has_emb || (size_t)N * F * emb.Dp * sizeof(float) < ((size_t)16 << 20))));
  pending_forward_N = 0;                         /* a forward whose backward never came is forgotten: drop its counts NOW — a cached */
  if (has_emb && emb.last_L > 0 && consume_N == 0) emb.clear_batch();   /* graph was captured from a clean state and would not */
  struct AfterStep {                             /* host-side state a REPLAYED graph cannot set */
    Model* m; int mode, N; bool defer;
    ~AfterStep() {
      if (mode != 1 || std::uncaught_exceptions() > 0) return;
      m->deferred_pending = defer; m->deferred_N = N;
      if (m->has_emb) m->emb.last_L = defer ? (int64_t)m->p2p.R * m->p2p.cap : 0;
    }
  } after{this, mode, N, defer};
  if (!use_graph) {
    ev_n = 0; ev_names.clear();
    if (mode == 1) { p2p_step(E, X, W, Y, N, consume_N, defer); if (publish_to) publish_status(ctx, st_dev, has_emb ? emb.counters : nullptr, has_wide ? wide.counters : nullptr, publish_to); }
    else step_device(E, X, W, Y, N, train, publish_to);
    phase_names = ev_names;
    return;
  }
  const auto key = std::make_tuple((const void*)E, (const void*)X, (const void*)W, (const void*)Y, N,
                                   (train ? 1 : 0) | (profile ? 2 : 0) | (mode << 2) | (ctx->exact_updaters << 6) | ((defer ? 1 : 0) << 7) | (consume_N << 8),
                                   (const void*)publish_to, ctx->fc_precision);
  if (has_emb && emb.generation != emb_generation) {     /* the embedding workspace was reallocated (p2p_init, a larger shard lookup) */
    for (auto& g : graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    graphs.clear();
    emb_generation = emb.generation;
  }
  auto it = graphs.find(key);
  if (it == graphs.end()) {
    if (graphs.size() >= 256) {
      for (auto& g : graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
      graphs.clear();
    }
    const long l0 = ctx->launches;
    cudaGraph_t graph = nullptr;
    PS_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    capturing = true; ev_n = 0; ev_names.clear();
    try {
      if (mode == 1) { p2p_step(E, X, W, Y, N, consume_N, defer); if (publish_to) publish_status(ctx, st_dev, has_emb ? emb.counters : nullptr, has_wide ? wide.counters : nullptr, publish_to); }
      else step_device(E, X, W, Y, N, train, publish_to);
    }
    catch (...) { capturing = false; cudaStreamEndCapture(ctx->stream, &graph); if (graph) cudaGraphDestroy(graph); throw; }
    capturing = false;
    PS_CUDA(cudaStreamEndCapture(ctx->stream, &graph));
    GraphEntry ge;
    ge.names = ev_names;
    ge.kernels = ctx->launches - l0;
    ctx->launches = l0;
    PS_CUDA(cudaGraphInstantiate(&ge.exec, graph, 0));
    PS_CUDA(cudaGraphDestroy(graph));
    it = graphs.emplace(key, ge).first;
  }
  PS_CUDA(cudaGraphLaunch(it->second.exec, ctx->stream));
  ctx->launches += it->second.kernels;
  phase_names = it->second.names;
}

/* ------------------------------------------------------------------ per-layer pieces */
void Model::fwd_layer(int l, int N) {                /* FcLayer.forward (FcLayer.java:74-91) */
  if (l == L - 1 && top_is_unit()) return;           /* the 1-unit top layer is fused into the tail */
  const bool fp32 = ctx->fc_precision == PS_FC_FP32;
  FcFwdArgs a{};
  a.B = N; a.in = fcs[l].in; a.out = fcs[l].out;
  a.A = act[l]; a.lda = ld[l]; a.W = fcs[l].W; a.ldw = fcs[l].ldw; a.Wlo = fcs[l].Wlo; a.bias = fcs[l].bias; a.act = fcs[l].act;
  a.Z = act[l + 1]; a.ldz = ld[l + 1];
  a.Zt = (!fp32 && l + 1 < L) ? act_t[l + 1] : nullptr; a.ldzt = ldt;
  if (fp32) fc_forward_fp32(ctx, a); else fc_forward_tf32(ctx, a);
}

void Model::run_tail(const float* Y, int N, bool train) {
  const bool fp32 = ctx->fc_precision == PS_FC_FP32;
  if (kind == PS_MODEL_FCNN) {
    tail_softmax(ctx, N, width[L], act[L], ld[L], Y, delta[L], ld[L], fp32 ? nullptr : delta_t[L], ldt, train ? 1 : 0, st_dev, tail_ws);
  } else if (top_is_unit()) {
    const FcLayer& f = fcs[L - 1];
    fc1_forward_tail(ctx, N, f.in, act[L - 1], ld[L - 1], f.W, f.bias, has_wide ? wide_z : nullptr, Y, act[L], ld[L], has_wide ? P : act[L],
                     has_wide ? 1 : ld[L], delta[L], ld[L], fp32 ? nullptr : delta_t[L], train ? 1 : 0, st_dev, tail_ws, has_emb ? emb.counters : nullptr);
  } else {
    tail_binary(ctx, N, act[L], ld[L], has_wide ? wide_z : nullptr, Y, has_wide ? P : act[L], has_wide ? 1 : ld[L], delta[L], ld[L],
                fp32 ? nullptr : delta_t[L], train ? 1 : 0, st_dev, tail_ws, has_emb ? emb.counters : nullptr);
  }
}

void Model::wgrad_layer(int l, int N) {              /* FcLayer.backward: db, dW (FcLayer.java:103-106) */
  const bool fp32 = ctx->fc_precision == PS_FC_FP32;
  const FcLayer& f = fcs[l];
  if (l == L - 1 && top_is_unit()) {
    fc1_wgrad(ctx, N, f.in, delta[L], ld[L], act[l], ld[l], f.G, (size_t)f.out * f.ldw, f.nsplit);
    return;
  }
  FcWgradArgs g{};
  g.B = N; g.in = f.in; g.out = f.out;
  g.dl = delta[l + 1]; g.ldd = ld[l + 1]; g.A = act[l]; g.lda = ld[l];
  g.dlT = delta_t[l + 1]; g.AT = act_t[l]; g.ldt = ldt;
  g.G = f.G; g.ldg = f.ldw; g.slab = (size_t)f.out * f.ldw; g.nsplit = f.nsplit;
  if (fp32) fc_wgrad_fp32(ctx, g); else fc_wgrad_tf32(ctx, g);
}

void Model::dgrad_layer(int l, int N) {              /* FcLayer.backward: delta = W^T delta (FcLayer.java:108) + the derivative below */
  const bool fp32 = ctx->fc_precision == PS_FC_FP32;
  const FcLayer& f = fcs[l];
  const int act_below = l > 0 ? fcs[l - 1].act : PS_ACT_NONE;
  if (l == L - 1 && top_is_unit()) {
    fc1_dgrad(ctx, N, f.in, delta[L], ld[L], fp32 ? nullptr : delta_t[L], f.W, act_below, act[l], ld[l], delta[l], ld[l], act_t[l], ldt, (!fp32 && l > 0) ? delta_t[l] : nullptr, ldt);
    return;
  }
  FcDgradArgs d{};
  d.B = N; d.in = f.in; d.out = f.out;
  d.dl = delta[l + 1]; d.ldd = ld[l + 1]; d.W = f.W; d.ldw = f.ldw; d.Wt = f.Wt; d.ldwt = f.ldwt; d.Wtlo = f.Wtlo;
  d.act_below = act_below; d.Y = act[l]; d.ldy = ld[l]; d.Yt = act_t[l]; d.ldyt = ldt;
  d.n_cols = f.in; d.dX = delta[l]; d.ldx = ld[l];
  d.dXt = (!fp32 && l > 0) ? delta_t[l] : nullptr; d.ldxt = ldt;
  if (fp32) fc_dgrad_fp32(ctx, d); else fc_dgrad_tf32(ctx, d);
}

namespace {
struct StreamScope {   /* every launch helper reads ctx->stream: run a few of them on a side stream */
  Ctx* c; cudaStream_t saved;
  StreamScope(Ctx* ctx, cudaStream_t s) : c(ctx), saved(ctx->stream) { c->stream = s; }
  ~StreamScope() { c->stream = saved; }
};
}  // namespace

/* `to` waits for everything enqueued on `from` so far */
void Model::fork(cudaStream_t from, cudaStream_t to) {
  cudaEvent_t e = sync_ev[sync_n++ % 24];
  PS_CUDA(cudaEventRecord(e, from));
  PS_CUDA(cudaStreamWaitEvent(to, e, 0));
}

void Model::forward_backward(const int64_t* W, const int64_t* W_all, int n_all, const float* Y, int N, bool train, bool wide_update_now) {
  forward_layers(N, train);
  run_tail(Y, N, train);
  mark("tail");
  if (train && p2p_scalars_now) {                /* sharded step over peer memory: the global loss / gbar / early-exit flag, exchanged NOW on side stream 2 */
    cudaStream_t s = ctx->stream, s2 = aux[1];
    fork(s, s2);
    StreamScope sc(ctx, s2);
    const DenseUpdateArgs u = dense_args(N * p2p.R);
    scalars_send(ctx, st_dev, p2p.dev, u.total, has_emb ? emb.counters : nullptr);
    shard_finish_scalars_p2p(ctx, st_dev, p2p.state(), u.total);
    if (has_wide) wide.update_all(gbar_ptr(st_dev), skip_ptr(st_dev), wide_bias, &upd_wide);   /* LRLayer.backward needs only the global gbar */
  }
  if (train) backward_layers(N, wide_update_now);
}

/* the forward loop of DNN.train / predict (DNN.java:44-46) after the input layers: FcLayer.forward x L; the wide branch the caller
 * started on side stream 1 joins before the tail */
void Model::forward_layers(int N, bool train) {
  cudaStream_t s = ctx->stream, s1 = aux[0], s2 = aux[1];
  const bool fp32 = ctx->fc_precision == PS_FC_FP32;
  fork(s, s2);
  if (!fp32 && train) { StreamScope sc(ctx, s2); transpose_copy(ctx, act[0], ld[0], act_t[0], ldt, N, width[0]); }   /* only wgrad0 needs it */
  for (int l = 0; l < L; ++l) {
    fwd_layer(l, N);
    mark(("fc_fwd" + std::to_string(l)).c_str());
  }
  fork(s1, s);                                   /* wide_z */
}

/* the reverse loop of DNN.train (DNN.java:64-68) given delta[L] and the step status (gbar, skip) */
void Model::backward_layers(int N, bool wide_update_now) {
  cudaStream_t s = ctx->stream, s1 = aux[0], s2 = aux[1];
  /* ---- backward (DNN.java:64-68): the dgrad chain is the critical path; each wgrad runs beside it ---- */
  if (has_wide && wide_update_now) {
    fork(s, s2);
    StreamScope sc(ctx, s2);
    wide.update_all(gbar_ptr(st_dev), skip_ptr(st_dev), wide_bias, &upd_wide);
  }
  const bool fp32 = ctx->fc_precision == PS_FC_FP32;
  if (fp32 || !ctx->group_wgrad) {               /* each wgrad beside its layer's dgrad */
    for (int l = L - 1; l >= 0; --l) {
      fork(s, s1);                               /* delta[l+1] (tail or dgrad(l+1)) is ready */
      if (l == 0) fork(s2, s1);                  /* act_t[0] */
      dgrad_layer(l, N);                         /* captured first: the critical chain's node precedes its sibling */
      { StreamScope sc(ctx, s1); wgrad_layer(l, N); }
      mark(("fc_dgrad" + std::to_string(l)).c_str());
    }
    return;
  }
  /* tensor-core modes: the dgrad chain runs alone (a 128-CTA GEMM shares its SMs with nobody), then EVERY layer's weight gradient
   * goes out as one grouped launch on side stream 1 — beside the embedding scatter / update that follows on the main stream */
  for (int l = L - 1; l >= 0; --l) {
    if (l == L - 1 && top_is_unit()) { fork(s, s1); StreamScope sc(ctx, s1); wgrad_layer(l, N); }   /* the 1-unit top layer's CUDA-core wgrad needs only the tail */
    dgrad_layer(l, N);
    mark(("fc_dgrad" + std::to_string(l)).c_str());
  }
  fork(s, s1); fork(s2, s1);                     /* every delta; act_t[0] */
  {
    StreamScope sc(ctx, s1);
    FcWgradArgs ga[kMaxDenseLayers];
    int n = 0;
    for (int l = L - 1; l >= 0; --l) {
      if (l == L - 1 && top_is_unit()) continue;
      const FcLayer& f = fcs[l];
      FcWgradArgs& g = ga[n++];
      g = FcWgradArgs{};
      g.B = N; g.in = f.in; g.out = f.out;
      g.dl = delta[l + 1]; g.ldd = ld[l + 1]; g.A = act[l]; g.lda = ld[l];
      g.dlT = delta_t[l + 1]; g.AT = act_t[l]; g.ldt = ldt;
      g.G = f.G; g.ldg = f.ldw; g.slab = (size_t)f.out * f.ldw; g.nsplit = f.nsplit;
    }
    if (!fc_wgrad_grouped_tf32(ctx, ga, n))
      for (int i = 0; i < n; ++i) fc_wgrad_tf32(ctx, ga[i]);
  }
}

void Model::step_device(const int64_t* E, const float* X, const int64_t* W, const float* Y, int N, bool train, StepStatus* publish_to) {
  PS_REQUIRE(N > 0 && N <= Bmax, PS_ERR_ARG, "model: batch size must be in [1, max_batch]");
  cudaStream_t s = ctx->stream, s1 = aux[0], s2 = aux[1];
  mark("begin");
  /* ---- forward (DNN.java:44-46) ---- */
  fork(s, s1);
  if (has_wide) { StreamScope sc(ctx, s1); wide.forward(W, N, F, wide_bias, wide_z); }   /* LRLayer.forward beside the deep branch */
  if (has_emb) {
    emb.lookup(E, nullptr, N, act[0], ld[0], X, Xn, F * D);   /* EmbeddingLayer.forward + ConcatLayer.forward: one kernel */
    mark("emb_lookup");
  } else {
    PS_CUDA(cudaMemcpy2DAsync(act[0], sizeof(float) * ld[0], X, sizeof(float) * Xn, sizeof(float) * Xn, N, cudaMemcpyDeviceToDevice, s));
  }
  forward_backward(W, nullptr, 0, Y, N, train, true);
  if (!train) {
    if (has_emb) emb.clear_batch();
    if (publish_to) publish_status(ctx, st_dev, has_emb ? emb.counters : nullptr, has_wide ? wide.counters : nullptr, publish_to);
    fork(s2, s);
    return;
  }
  /* ---- KVStore.update + clear (Trainer.java:93,95) ---- */
  fork(s, s1);                                   /* every dgrad has read W / Wt: the dense update may overwrite them */
  {
    StreamScope sc(ctx, s1);
    const DenseUpdateArgs u = dense_args(N);
    dense_update(ctx, u, st_dev, has_emb ? emb.counters : nullptr, has_wide ? wide.counters : nullptr, publish_to);
  }
  if (has_emb) { emb.scatter_update(delta[0], ld[0], nullptr, 0, N, 2, skip_ptr(st_dev), 0, true); mark("emb_bwd_update"); }
  fork(s1, s);
  fork(s2, s);
  mark("end");
}

/* ------------------------------------------------------------------ sharded step pieces */
DenseUpdateArgs Model::dense_args(int N) {
  DenseUpdateArgs u{};
  u.n_layers = L; u.N = N;
  long first = 0;
  for (int l = 0; l < L; ++l) {
    DenseLayerDesc& q = u.l[l];
    const FcLayer& f = fcs[l];
    q.W = f.W; q.Wt = f.Wt; q.Wlo = f.Wlo; q.Wtlo = f.Wtlo; q.bias = f.bias; q.sW1 = f.sW1; q.sW2 = f.sW2; q.sb1 = f.sb1; q.sb2 = f.sb2;
    q.G = f.G; q.slab = (size_t)f.out * f.ldw; q.nsplit = f.nsplit; q.out = f.out; q.in = f.in; q.ldw = f.ldw; q.ldwt = f.ldwt; q.ldg = f.ldw;
    q.updW = make_updater_dev(f.updW); q.updB = make_updater_dev(f.updB);
    q.first = first; first += (long)f.out * (f.in + 1);
  }
  u.total = first;
  return u;
}

void Model::shard_emb_lookup(const uint64_t* keys, int n, float* rows_out) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  PS_REQUIRE(has_emb, PS_ERR_STATE, "model has no embedding layer");
  emb.lookup_packed(keys, n, rows_out);
}

void Model::shard_unpack_rows(const float* rows, const int32_t* send_pos, int N) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  PS_REQUIRE(has_emb && N > 0 && N <= Bmax, PS_ERR_ARG, "shard_unpack_rows: bad batch");
  shard_unpack(ctx, rows, send_pos, N, F, D, emb.Dp, act[0], ld[0]);
}

/* everything between the two exchanges: concat, wide branch, FcLayer forward, tail, FcLayer backward,
 * then the per-rank gradient sums and scalars go into ONE flat buffer for the all-reduce.      */
void Model::shard_dense_step(const float* X, const int64_t* W_local, const int64_t* W_all, int n_all, const float* Y, int N) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  PS_REQUIRE(N > 0 && N <= Bmax, PS_ERR_ARG, "model: batch size must be in [1, max_batch]");
  cudaStream_t s = ctx->stream, s1 = aux[0], s2 = aux[1];
  fork(s, s1);
  if (has_wide) {
    StreamScope sc(ctx, s1);
    wide.insert(W_all, n_all);                     /* the union of every replica's keys */
    wide.forward(W_local, N, F, wide_bias, wide_z);
  }
  if (has_emb) PS_CUDA(cudaMemcpy2DAsync(act[0] + F * D, sizeof(float) * ld[0], X, sizeof(float) * Xn, sizeof(float) * Xn, N, cudaMemcpyDeviceToDevice, s));
  else PS_CUDA(cudaMemcpy2DAsync(act[0], sizeof(float) * ld[0], X, sizeof(float) * Xn, sizeof(float) * Xn, N, cudaMemcpyDeviceToDevice, s));
  forward_backward(W_local, W_all, n_all, Y, N, true, false);
  fork(s1, s);                                     /* all wgrads */
  fork(s2, s);
  const DenseUpdateArgs u = dense_args(N);
  if (!gsum) { gsum_len = u.total + 3; gsum = dmalloc_zero<float>((size_t)gsum_len, s); }
  dense_reduce(ctx, u, st_dev, gsum, has_emb ? emb.counters : nullptr);
  last_N = N; last_train = true;
}

void Model::shard_pack(const int32_t* send_pos, int N, float* grads_send) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  PS_REQUIRE(has_emb, PS_ERR_STATE, "model has no embedding layer");
  shard_pack_grads(ctx, delta[0], ld[0], act[0], ld[0], send_pos, N, F, D, emb.Dp, grads_send);
}

/* gsum now holds the SUM over R ranks of per-rank batch sums / means: every replica applies the
 * same global-batch update — N-GPU result == 1-GPU result on the concatenated batch */
void Model::shard_finish(int N_global, int R) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  PS_REQUIRE(gsum != nullptr, PS_ERR_STATE, "shard_finish before shard_dense_step");
  DenseUpdateArgs u = dense_args(N_global);
  shard_finish_scalars(ctx, st_dev, gsum + u.total, R);
  if (has_wide) wide.update_all(gbar_ptr(st_dev), skip_ptr(st_dev), wide_bias, &upd_wide);
  for (int l = 0; l < L; ++l) { u.l[l].G = gsum + u.l[l].first; u.l[l].slab = 0; u.l[l].nsplit = 1; u.l[l].ldg = fcs[l].in + 1; }
  dense_update(ctx, u, st_dev, has_emb ? emb.counters : nullptr, has_wide ? wide.counters : nullptr, nullptr);
}

void Model::shard_emb_apply(const float* grads_recv, int n) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  PS_REQUIRE(has_emb, PS_ERR_STATE, "model has no embedding layer");
  if (n > 0) emb.scatter_update(grads_recv, emb.Dp, nullptr, emb.Dp, n, 2, skip_ptr(st_dev), 1);
  else emb.last_L = 0;
}

/* ------------------------------------------------------------------ sharded step over peer memory */
void Model::p2p_init(int R, int rank, int cap, void* handle_out64) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  PS_REQUIRE(!p2p.slab, PS_ERR_STATE, "p2p already initialised");
  const DenseUpdateArgs u = dense_args(1);
  const long glen = (u.total + 3 + 3) / 4 * 4;
  if (!gsum) { gsum_len = glen; gsum = dmalloc_zero<float>((size_t)gsum_len, ctx->stream); }
  PS_REQUIRE(gsum_len >= glen, PS_ERR_STATE, "gradient buffer was created before p2p_init with a smaller size");
  p2p.create(ctx, R, rank, has_emb ? cap : 1, has_emb ? emb.Dp : 4, std::max(2, Bmax * std::max(F, 1)), (int)glen, (int64_t)Bmax * std::max(F, 1));
  if (has_emb) emb.reserve((int64_t)R * cap);
  p2p.get_handle(handle_out64);
}

void Model::p2p_connect(const void* all_handles) { p2p.connect(all_handles); }

/* one Trainer step of the R-rank group on the concatenated batch, this rank's share: every exchange is a
 * store into the consumer's mailbox by the kernel that produced the data (see p2p.cuh)              */
void Model::flush_deferred() {
  if (!deferred_pending) return;
  deferred_pending = false;
  const int R = p2p.R;
  if (has_emb) emb.update_pull(p2p.state(), R * p2p.cap, 2, skip_ptr(st_dev));
  const DenseUpdateArgs u = dense_args(deferred_N * R);
  p2p.wait(CH_GSUM);
  dense_update(ctx, u, st_dev, has_emb ? emb.counters : nullptr, has_wide ? wide.counters : nullptr, nullptr, p2p.state());
}

void Model::p2p_step(const int64_t* E, const float* X, const int64_t* W, const float* Y, int N, int consume_N, bool defer) {
  PS_REQUIRE(p2p.connected, PS_ERR_STATE, "p2p_step before p2p_connect");
  PS_REQUIRE(N > 0 && N <= Bmax, PS_ERR_ARG, "model: batch size must be in [1, max_batch]");
  cudaStream_t s = ctx->stream, s1 = aux[0], s2 = aux[1];
  const int R = p2p.R, cap = p2p.cap;
  mark("begin");
  cudaEvent_t dense_done = nullptr;
  if (consume_N > 0) {
    /* the previous sharded step's owner-side embedding update (PServer.push + psUpdate) and dense update, beside this step's route_send:
     * they wait for the flags of THAT step (P2PState::pub_seq) and read nothing this step's first kernels write */
    fork(s, s2);
    if (has_emb) { StreamScope sc(ctx, s2); emb.update_pull(p2p.state(), R * cap, 2, skip_ptr(st_dev)); }
    fork(s, s1);
    {
      StreamScope sc(ctx, s1);
      const DenseUpdateArgs u = dense_args(consume_N * R);
      p2p.wait(CH_GSUM);                                                        /* every replica's sums have landed */
      dense_update(ctx, u, st_dev, has_emb ? emb.counters : nullptr, has_wide ? wide.counters : nullptr, nullptr, p2p.state());
      dense_done = sync_ev[sync_n++ % 24];
      PS_CUDA(cudaEventRecord(dense_done, s1));                                 /* (waited for just before the first FcLayer: the wide branch follows on s1) */
    }
  }
  /* every exchange below is ONE producer kernel that flags its consumers when its last block ends and ONE consumer kernel
   * that waits for the flags in its prologue: no flag kernels, no send kernels on the way back, no host */
  if (has_emb) {
    p2p.route_send(E, N, F);                                                    /* PSRouterClient.getList: the step's sequence number; each key of the batch once, stored into its owner's mailbox */
    mark("route_send");
  } else {
    p2p.begin();
  }
  fork(s, s1);
  if (has_wide) {                                                               /* wide branch beside the embedding exchange */
    PS_REQUIRE(((size_t)N * F * 8) % 16 == 0, PS_ERR_ARG, "p2p: N*F must be even");
    StreamScope sc(ctx, s1);
    p2p.bcast(W, (size_t)N * F * 8, CH_WIDE);                                   /* flags the peers itself */
    wide.insert(nullptr, R * N * F, p2p.state());                               /* the union of every replica's keys (waits for their flags) */
    wide.forward(W, N, F, wide_bias, wide_z);
  }
  if (has_emb) {
    fork(s, s2);
    { StreamScope sc(ctx, s2); p2p.counts(); }                                  /* the keys' occurrence counts next to where their gradient sums will be */
    if (consume_N > 0) fork(s2, s);                                             /* the previous step's rows are updated, its batch forgotten */
    emb.lookup_packed(nullptr, R * cap, nullptr, p2p.dev, true);                /* PServer.getList on the owner: find-or-insert, rows stored straight into the requesters' mailboxes */
    mark("owner_lookup");
    if (D % 4 == 0) emb.gather_resolved(p2p.bt, p2p.lk_b, p2p.dev, N, act[0], ld[0], X, Xn, F * D);   /* rows_in -> concat buffer (+ mask bits, ConcatLayer) */
    else p2p.unpack(N, F, D, act[0], ld[0], X, Xn, F * D);
    mark("gather");
  } else {
    PS_CUDA(cudaMemcpy2DAsync(act[0], sizeof(float) * ld[0], X, sizeof(float) * Xn, sizeof(float) * Xn, N, cudaMemcpyDeviceToDevice, s));
  }
  if (dense_done != nullptr) PS_CUDA(cudaStreamWaitEvent(s, dense_done, 0));    /* the previous step's dense update */
  p2p_scalars_now = true;
  try { forward_backward(W, nullptr, 0, Y, N, true, false); }                   /* main: delta[0]; side 1: the wgrads; side 2: the global scalars + wide update */
  catch (...) { p2p_scalars_now = false; throw; }
  p2p_scalars_now = false;
  /* side stream 1, behind the weight gradients: PServer sync mode for the dense keys — every rank's gradient sums into every mailbox,
   * then the dense update (which waits for the replicas' flags itself), all beside the row-gradient push on the main stream        */
  fork(s2, s1);                                                                 /* the global skip flag */
  fork(s, s1);                                                                  /* every dgrad has read W / Wt: the dense update may overwrite them */
  {
    StreamScope sc(ctx, s1);
    const DenseUpdateArgs u = dense_args(N * R);
    dense_reduce_send(ctx, u, st_dev, p2p.dev, has_emb ? emb.counters : nullptr);
    if (!defer) {
      p2p.wait(CH_GSUM);                                                        /* every replica's sums have landed */
      dense_update(ctx, u, st_dev, has_emb ? emb.counters : nullptr, has_wide ? wide.counters : nullptr, nullptr, p2p.state());
    }
  }
  /* s2 carries, in order: the counts, the global scalars + wide update (forward_backward) */
  fork(s2, s);                                                                  /* the counts are in place; the global skip flag — NOT the weight gradients */
  if (has_emb) {
    emb.scatter_rows(p2p.bt, p2p.lk_b, p2p.dev, delta[0], ld[0], D % 4 == 0 ? nullptr : act[0], ld[0], N);   /* client.push: one gradient sum per unique key of this rank's batch, left in this rank's slab; flags the owners */
    mark("scatter_rows");
    fork(s, s2);
    { StreamScope sc(ctx, s2); p2p.tidy(); }                                    /* beside the update: forget the batch's de-duplication, zero the other parity's sums */
    if (!defer) {
      emb.update_pull(p2p.state(), R * cap, 2, skip_ptr(st_dev));               /* PServer.push (sync mode) + psUpdate on the owner: reads the requesters' sums over NVLink */
      mark("owner_update");
    }
  }
  fork(s1, s);                                                                  /* the dense update */
  fork(s2, s);                                                                  /* the tidy */
  mark("end");
  last_N = N; last_train = true;
}

void Model::kernel_times(const int64_t* const* E_ring, int n_ring, int N, int reps, float* out) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  PS_REQUIRE(has_emb && n_ring > 0 && reps > 0 && N > 0 && N <= Bmax, PS_ERR_ARG, "kernel_times: bad argument");
  PS_REQUIRE(in_flight == 0, PS_ERR_STATE, "kernel_times: steps in flight");
  cudaStream_t s = ctx->stream;
  cudaEvent_t e0, e1;
  PS_CUDA(cudaEventCreate(&e0)); PS_CUDA(cudaEventCreate(&e1));
  float total[4] = {0, 0, 0, 0};
  for (int variant = 0; variant < 4; ++variant) {
    /* 0: resolve-only lookup + clear   1: lookup (resolve + gather) + clear   2: lookup + scatter_update   3: resolve-only + clear + clear */
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    const long l0 = ctx->launches;
    PS_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    try {
      for (int r = 0; r < reps; ++r) {
        emb.lookup(E_ring[r % n_ring], nullptr, N, (variant == 1 || variant == 2) ? act[0] : nullptr, ld[0]);
        if (variant == 2) emb.scatter_update(delta[0], ld[0], nullptr, 0, N, 2, nullptr, 0, true);
        else emb.clear_batch();
        if (variant == 3) emb.clear_batch();
      }
    } catch (...) { cudaStreamEndCapture(s, &graph); if (graph) cudaGraphDestroy(graph); throw; }
    PS_CUDA(cudaStreamEndCapture(s, &graph));
    const long per_graph = ctx->launches - l0;
    PS_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    PS_CUDA(cudaGraphLaunch(exec, s));                       /* warm-up */
    PS_CUDA(cudaEventRecord(e0, s));
    PS_CUDA(cudaGraphLaunch(exec, s));
    PS_CUDA(cudaEventRecord(e1, s));
    PS_CUDA(cudaStreamSynchronize(s));
    ctx->launches += per_graph;
    float ms = 0.f;
    PS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    total[variant] = 1e3f * ms / (float)reps;
    cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  const float clear = total[3] - total[0];
  out[0] = total[0] - clear;                 /* key resolution alone (the lookup kernel without its gather) */
  out[1] = total[1] - clear;                 /* the lookup kernel: key resolution + gather + mask, what EmbeddingLayer.forward costs */
  out[2] = total[2] - total[1] + clear;      /* scatter_update */
  out[3] = clear;
  emb.check_errors();
}

void Model::gemm_times(int N, int reps, float* out) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  PS_REQUIRE(reps > 0 && N > 0 && N <= Bmax, PS_ERR_ARG, "gemm_times: bad argument");
  PS_REQUIRE(in_flight == 0, PS_ERR_STATE, "gemm_times: steps in flight");
  cudaStream_t s = ctx->stream;
  cudaEvent_t e0, e1;
  PS_CUDA(cudaEventCreate(&e0)); PS_CUDA(cudaEventCreate(&e1));
  for (int l = 0; l < L; ++l)
    for (int which = 0; which < 3; ++which) {
      cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
      const long l0 = ctx->launches;
      PS_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      try {
        for (int r = 0; r < reps; ++r) {
          if (which == 0) { if (l == L - 1 && top_is_unit()) run_tail(stage[0].Y, N, true); else fwd_layer(l, N); }
          else if (which == 1) dgrad_layer(l, N);
          else wgrad_layer(l, N);
        }
      } catch (...) { cudaStreamEndCapture(s, &graph); if (graph) cudaGraphDestroy(graph); throw; }
      PS_CUDA(cudaStreamEndCapture(s, &graph));
      const long per_graph = ctx->launches - l0;
      PS_CUDA(cudaGraphInstantiate(&exec, graph, 0));
      PS_CUDA(cudaGraphLaunch(exec, s));
      PS_CUDA(cudaEventRecord(e0, s));
      PS_CUDA(cudaGraphLaunch(exec, s));
      PS_CUDA(cudaEventRecord(e1, s));
      PS_CUDA(cudaStreamSynchronize(s));
      ctx->launches += per_graph;
      float ms = 0.f;
      PS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      out[3 * l + which] = 1e3f * ms / (float)reps;
      cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
    }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
}

void Model::submit(const HostBatch& b, int mode) {
  PS_REQUIRE(in_flight < kStages, PS_ERR_STATE, "model: four steps already in flight; collect first");
  PS_REQUIRE(b.N > 0 && b.N <= Bmax, PS_ERR_ARG, "model: batch size must be in [1, max_batch]");
  PS_REQUIRE(b.X && b.Y && (!has_emb || b.E) && (!has_wide || b.W), PS_ERR_ARG, "model: missing input matrix");
  Stage& S = stage[next_stage];
  cudaStream_t cs = ctx->copy_stream;
  const size_t N = (size_t)b.N;
  if (pad_dirty) {                               /* a text batch may have left its bad-line count in the status: this batch is not text */
    PS_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(st_dev) + offsetof(StepStatus, pad), 0, sizeof(uint32_t), ctx->stream));
    pad_dirty = false;
  }
  if (has_emb) PS_CUDA(cudaMemcpyAsync(S.E, b.E, sizeof(int64_t) * N * F, cudaMemcpyHostToDevice, cs));
  if (has_wide) PS_CUDA(cudaMemcpyAsync(S.W, b.W, sizeof(int64_t) * N * F, cudaMemcpyHostToDevice, cs));
  PS_CUDA(cudaMemcpyAsync(S.X, b.X, sizeof(float) * N * Xn, cudaMemcpyHostToDevice, cs));
  PS_CUDA(cudaMemcpyAsync(S.Y, b.Y, sizeof(float) * N, cudaMemcpyHostToDevice, cs));
  PS_CUDA(cudaEventRecord(S.h2d_done, cs));
  PS_CUDA(cudaStreamWaitEvent(ctx->stream, S.h2d_done, 0));
  S.N = b.N;
  run_step(S.E, S.X, S.W, S.Y, b.N, true, S.st_host, mode);
  PS_CUDA(cudaEventRecord(S.step_done, ctx->stream));
  S.busy = true;
  next_stage = (next_stage + 1) % kStages; in_flight++;
  last_N = b.N; last_train = true;
}

float Model::collect() {
  PS_REQUIRE(in_flight > 0, PS_ERR_STATE, "model: nothing in flight");
  Stage& S = stage[oldest_stage];
  PS_CUDA(cudaEventSynchronize(S.step_done));
  last_status = *S.st_host;
  S.busy = false;
  oldest_stage = (oldest_stage + 1) % kStages; in_flight--;
  if (profile && in_flight == 0) finish_profile();
  PS_REQUIRE((last_status.emb_err & 2u) == 0, PS_ERR_ARG, "embedding id outside [0, 2^44): the batch was refused, nothing was applied");
  PS_REQUIRE(last_status.emb_err == 0, PS_ERR_CAPACITY, "embedding table is full: raise emb_capacity");
  PS_REQUIRE(last_status.wide_err == 0, PS_ERR_CAPACITY, "wide table is full");
  return last_status.loss;
}

float Model::read_loss() {
  PS_REQUIRE(in_flight == 0, PS_ERR_STATE, "model: host steps in flight");
  publish_status(ctx, st_dev, has_emb ? emb.counters : nullptr, has_wide ? wide.counters : nullptr, stage[0].st_host);
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  last_status = *stage[0].st_host;
  if (profile) finish_profile();
  PS_REQUIRE((last_status.emb_err & 2u) == 0, PS_ERR_ARG, "embedding id outside [0, 2^44): the batch was refused, nothing was applied");
  PS_REQUIRE(last_status.emb_err == 0, PS_ERR_CAPACITY, "embedding table is full: raise emb_capacity");
  PS_REQUIRE(last_status.wide_err == 0, PS_ERR_CAPACITY, "wide table is full");
  return last_status.loss;
}

/* ------------------------------------------------------------------ the reference's own train loop, call by call
 * DNN.train / WideDeepNN.train run: forward loop -> loss.forward / loss.backward IN JAVA (DNN.java:47-49) -> setDelta on the
 * last layer -> reverse loop (DNN.java:64-68); Trainer then calls KVStore.update + clear (Trainer.java:93,95).  forward_host is
 * the forward loop (the batch stays pending: occurrence counts, ReLU masks, activations), backward_update_host takes the delta
 * Java's loss.backward produced — dLoss/dP, BEFORE the output activation's derivative, exactly what setDelta receives — and runs
 * the reverse loop and the update.  No label ever crosses the boundary.                                                    */
void Model::forward_host(const HostBatch& b, float* P_out) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  PS_REQUIRE(in_flight == 0, PS_ERR_STATE, "model: steps in flight; collect first");
  PS_REQUIRE(b.N > 0 && b.N <= Bmax, PS_ERR_ARG, "model: batch size must be in [1, max_batch]");
  PS_REQUIRE(b.X && P_out && (!has_emb || b.E) && (!has_wide || b.W), PS_ERR_ARG, "model: missing input matrix");
  Stage& S = stage[0];
  cudaStream_t s = ctx->stream, s1 = aux[0], s2 = aux[1];
  const size_t N = (size_t)b.N;
  if (has_emb) PS_CUDA(cudaMemcpyAsync(S.E, b.E, sizeof(int64_t) * N * F, cudaMemcpyHostToDevice, s));
  if (has_wide) PS_CUDA(cudaMemcpyAsync(S.W, b.W, sizeof(int64_t) * N * F, cudaMemcpyHostToDevice, s));
  PS_CUDA(cudaMemcpyAsync(S.X, b.X, sizeof(float) * N * Xn, cudaMemcpyHostToDevice, s));
  fork(s, s1);
  if (has_wide) { StreamScope sc(ctx, s1); wide.forward(S.W, b.N, F, wide_bias, wide_z); }
  if (has_emb) emb.lookup(S.E, nullptr, b.N, act[0], ld[0], S.X, Xn, F * D);
  else PS_CUDA(cudaMemcpy2DAsync(act[0], sizeof(float) * ld[0], S.X, sizeof(float) * Xn, sizeof(float) * Xn, N, cudaMemcpyDeviceToDevice, s));
  forward_layers(b.N, true);
  run_tail(nullptr, b.N, false);                 /* AddLayer + Sigmoid (or Softmax) only: P */
  if (kind == PS_MODEL_FCNN)                     /* FullConnectedNN: P is C x N (FullConnectedNN.java:47) */
    PS_CUDA(cudaMemcpy2DAsync(P_out, sizeof(float) * width[L], act[L], sizeof(float) * ld[L], sizeof(float) * width[L], N, cudaMemcpyDeviceToHost, s));
  else if (has_wide) PS_CUDA(cudaMemcpyAsync(P_out, P, sizeof(float) * N, cudaMemcpyDeviceToHost, s));
  else PS_CUDA(cudaMemcpy2DAsync(P_out, sizeof(float), act[L], sizeof(float) * ld[L], sizeof(float), N, cudaMemcpyDeviceToHost, s));
  fork(s2, s);
  PS_CUDA(cudaStreamSynchronize(s));
  pending_forward_N = b.N;
  last_N = b.N; last_train = false;
}

void Model::backward_update_host(const float* delta_top, int N, float loss) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  PS_REQUIRE(pending_forward_N > 0 && N == pending_forward_N, PS_ERR_STATE, "backward_update without a matching forward");
  PS_REQUIRE(delta_top != nullptr, PS_ERR_ARG, "backward_update: null delta");
  Stage& S = stage[0];
  cudaStream_t s = ctx->stream, s1 = aux[0], s2 = aux[1];
  const bool fp32 = ctx->fc_precision == PS_FC_FP32;
  if (kind == PS_MODEL_FCNN) {
    /* delta_top is C x N: staged in delta[L]'s transposed twin's neighbour — the (unused) top delta_t buffer is [C][ldt], large enough */
    const int Cn = width[L];
    if (!dtop_stage) dtop_stage = dmalloc<float>((size_t)Bmax * Cn);
    PS_CUDA(cudaMemcpyAsync(dtop_stage, delta_top, sizeof(float) * (size_t)N * Cn, cudaMemcpyHostToDevice, s));
    tail_softmax_from_delta(ctx, N, Cn, act[L], ld[L], dtop_stage, delta[L], ld[L], fp32 ? nullptr : delta_t[L], ldt, loss, st_dev, tail_ws);
  } else {
    PS_CUDA(cudaMemcpyAsync(S.Y, delta_top, sizeof(float) * (size_t)N, cudaMemcpyHostToDevice, s));   /* the label buffer doubles as the delta staging */
    /* last layer's activation.backward (FcLayer.java:100-102 with Sigmoid.java:16-21) + rowMeans for LRLayer; loss as Java computed it */
    tail_binary_from_delta(ctx, N, has_wide ? P : act[L], has_wide ? 1 : ld[L], S.Y, delta[L], ld[L], fp32 ? nullptr : delta_t[L], loss, st_dev, tail_ws,
                           has_emb ? emb.counters : nullptr);
  }
  backward_layers(N, true);
  fork(s, s1);
  {
    StreamScope sc(ctx, s1);
    const DenseUpdateArgs u = dense_args(N);
    dense_update(ctx, u, st_dev, has_emb ? emb.counters : nullptr, has_wide ? wide.counters : nullptr, S.st_host);
  }
  if (has_emb) emb.scatter_update(delta[0], ld[0], nullptr, 0, N, 2, skip_ptr(st_dev), 0, true);
  fork(s1, s);
  fork(s2, s);
  PS_CUDA(cudaStreamSynchronize(s));
  pending_forward_N = 0;
  last_status = *S.st_host;
  last_N = N; last_train = true;
  PS_REQUIRE((last_status.emb_err & 2u) == 0, PS_ERR_ARG, "embedding id outside [0, 2^44): the batch was refused, nothing was applied");
  PS_REQUIRE(last_status.emb_err == 0, PS_ERR_CAPACITY, "embedding table is full: raise emb_capacity");
  PS_REQUIRE(last_status.wide_err == 0, PS_ERR_CAPACITY, "wide table is full");
}

/* DataSet.next + Trainer.train in one submission (DataSet.java:77-100, CTR.parseFeature CTR.java:47-68): `len` bytes of libsvm
 * text holding exactly N complete lines are copied to the device as they are, parsed there (ingest_dev.cu) into the stage's
 * E / X / W / Y and trained on — copy, parse and step are all asynchronous, collect() returns the loss.  A line the GPU parser
 * cannot take (a spelling outside its fast path, a short or missing line) drops the WHOLE batch like the reference's swallowed
 * exception does (DataSet.java:96-98): the step applies nothing and collect() reports it as skipped.                       */
void Model::submit_text(const char* text, size_t len, int N, int mode) {
  PS_REQUIRE(in_flight < kStages, PS_ERR_STATE, "model: four steps already in flight; collect first");
  PS_REQUIRE(N > 0 && N <= Bmax && text != nullptr && len > 0, PS_ERR_ARG, "submit_text: bad argument");
  PS_REQUIRE(has_emb, PS_ERR_ARG, "submit_text: the CTR text layout needs a model with embedding fields");
  Stage& S = stage[next_stage];
  if (len > S.text_cap) {
    PS_CUDA(cudaStreamSynchronize(ctx->stream)); PS_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    dfree(S.text); dfree(S.text_ws);
    S.text_cap = len + len / 2 + 4096;
    S.text = dmalloc<char>(S.text_cap);
    S.text_ws = dmalloc<uint32_t>(S.text_cap / 4096 + 4 + (size_t)Bmax);
  }
  if (!S.text_status) { S.text_status = dmalloc<uint8_t>((size_t)Bmax); if (!S.W) S.W = dmalloc<int64_t>((size_t)Bmax * F); }
  cudaStream_t cs = ctx->copy_stream;
  PS_CUDA(cudaMemcpyAsync(S.text, text, len, cudaMemcpyHostToDevice, cs));
  PS_CUDA(cudaEventRecord(S.h2d_done, cs));
  PS_CUDA(cudaStreamWaitEvent(ctx->stream, S.h2d_done, 0));
  libsvm_parse_dev_async(ctx, S.text, len, F, Xn, 100000 /* CTR.wideSize, CTR.java:36 */, N, S.E, S.X, S.W, S.Y, S.text_status, S.text_ws,
                         reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(st_dev) + offsetof(StepStatus, pad)));
  S.N = N;
  pad_dirty = true;
  run_step(S.E, S.X, S.W, S.Y, N, true, S.st_host, mode);
  PS_CUDA(cudaEventRecord(S.step_done, ctx->stream));
  S.busy = true;
  next_stage = (next_stage + 1) % kStages; in_flight++;
  last_N = N; last_train = true;
}

void Model::predict(const HostBatch& b, float* out) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  PS_REQUIRE(in_flight == 0, PS_ERR_STATE, "model: steps in flight; collect first");
  PS_REQUIRE(b.N > 0 && b.N <= Bmax, PS_ERR_ARG, "model: batch size must be in [1, max_batch]");
  Stage& S = stage[0];
  cudaStream_t s = ctx->stream;
  const size_t N = (size_t)b.N;
  if (has_emb) PS_CUDA(cudaMemcpyAsync(S.E, b.E, sizeof(int64_t) * N * F, cudaMemcpyHostToDevice, s));
  if (has_wide) PS_CUDA(cudaMemcpyAsync(S.W, b.W, sizeof(int64_t) * N * F, cudaMemcpyHostToDevice, s));
  PS_CUDA(cudaMemcpyAsync(S.X, b.X, sizeof(float) * N * Xn, cudaMemcpyHostToDevice, s));
  step_device(S.E, S.X, S.W, nullptr, b.N, false, nullptr);
  if (kind == PS_MODEL_FCNN)
    PS_CUDA(cudaMemcpy2DAsync(out, sizeof(float) * width[L], act[L], sizeof(float) * ld[L], sizeof(float) * width[L], N, cudaMemcpyDeviceToHost, s));
  else if (has_wide)
    PS_CUDA(cudaMemcpyAsync(out, P, sizeof(float) * N, cudaMemcpyDeviceToHost, s));
  else
    PS_CUDA(cudaMemcpy2DAsync(out, sizeof(float), act[L], sizeof(float) * ld[L], sizeof(float), N, cudaMemcpyDeviceToHost, s));
  PS_CUDA(cudaStreamSynchronize(s));
  last_N = b.N; last_train = false;
  if (has_emb) emb.check_errors();
  if (has_wide) wide.check_errors();
}

/* ------------------------------------------------------------------ KVStore.get / put by reference key */
/* Java's Float.toString / Double.toString spellings of an integer id ("15757.0", "1.6777216E7"):
 * strtod accepts both (EmbeddingField.java:70-71,88-89; LRLayer.java:78).                      */
int parse_key(const std::string& key, int* field, int64_t* id) {
  if (key.compare(0, 3, "emF") == 0) {
    char* end = nullptr;
    const long f = std::strtol(key.c_str() + 3, &end, 10);
    if (end && *end == '.' && end != key.c_str() + 3) {
      char* end2 = nullptr;
      const double v = std::strtod(end + 1, &end2);
      if (end2 && *end2 == '\0' && end2 != end + 1) { *field = (int)f; *id = (int64_t)v; return 0; }
    }
  }
  static const std::string wp = "wide.weights.";
  if (key.compare(0, wp.size(), wp) == 0) {
    char* end2 = nullptr;
    const double v = std::strtod(key.c_str() + wp.size(), &end2);
    if (end2 && *end2 == '\0' && end2 != key.c_str() + wp.size()) { *field = 0; *id = (int64_t)v; return 1; }
  }
  return 2;
}

static bool fc_key(const std::vector<FcLayer>& fcs, const std::string& key, int* l, bool* is_bias) {
  for (size_t i = 0; i < fcs.size(); ++i) {
    if (key == fcs[i].name + ".weights") { *l = (int)i; *is_bias = false; return true; }
    if (key == fcs[i].name + ".bias") { *l = (int)i; *is_bias = true; return true; }
  }
  return false;
}

static void fetch_fc(Ctx* ctx, const FcLayer& f, bool is_bias, const float* Wsrc, const float* bsrc, std::vector<float>& out) {
  if (is_bias) {
    out.resize(f.out);
    PS_CUDA(cudaMemcpyAsync(out.data(), bsrc, sizeof(float) * f.out, cudaMemcpyDeviceToHost, ctx->stream));
    PS_CUDA(cudaStreamSynchronize(ctx->stream));
    return;
  }
  std::vector<float> tmp((size_t)f.out * f.ldw);
  PS_CUDA(cudaMemcpyAsync(tmp.data(), Wsrc, sizeof(float) * tmp.size(), cudaMemcpyDeviceToHost, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
  out.resize((size_t)f.out * f.in);   /* jblas out x in, column-major: index(o, i) = o + out*i */
  for (int i = 0; i < f.in; ++i)
    for (int o = 0; o < f.out; ++o) out[(size_t)o + (size_t)f.out * i] = tmp[(size_t)o * f.ldw + i];
}

int Model::get(const std::string& key, std::vector<float>& out) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  int field = 0; int64_t id = 0; int l = 0; bool is_bias = false;
  const int k = parse_key(key, &field, &id);
  if (k == 0) {
    if (!has_emb || field < 0 || field >= F) return PS_NOT_FOUND;
    out.resize(D);
    int32_t f32 = field, found = 0;
    emb.get_rows(&f32, &id, 1, out.data(), nullptr, nullptr, &found);
    return found ? PS_OK : PS_NOT_FOUND;
  }
  if (k == 1) {
    if (!has_wide) return PS_NOT_FOUND;
    out.resize(1);
    return wide.get(id, out.data(), nullptr, nullptr) ? PS_OK : PS_NOT_FOUND;
  }
  if (fc_key(fcs, key, &l, &is_bias)) { fetch_fc(ctx, fcs[l], is_bias, fcs[l].W, fcs[l].bias, out); return PS_OK; }
  if (has_wide && key == "wide.bias") {
    out.resize(1);
    PS_CUDA(cudaMemcpyAsync(out.data(), wide_bias, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PS_CUDA(cudaStreamSynchronize(ctx->stream));
    return PS_OK;
  }
  return PS_NOT_FOUND;
}

int Model::get_state(const std::string& key, int which, std::vector<float>& out) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  int field = 0; int64_t id = 0; int l = 0; bool is_bias = false;
  const int k = parse_key(key, &field, &id);
  if (k == 0) {
    if (!has_emb || field < 0 || field >= F) return PS_NOT_FOUND;
    std::vector<float> w(D), a(D), b(D);
    int32_t f32 = field, found = 0;
    emb.get_rows(&f32, &id, 1, w.data(), a.data(), b.data(), &found);
    if (!found) return PS_NOT_FOUND;
    out = which ? b : a;
    return PS_OK;
  }
  if (k == 1) {
    if (!has_wide) return PS_NOT_FOUND;
    float w, a, b;
    if (!wide.get(id, &w, &a, &b)) return PS_NOT_FOUND;
    out.assign(1, which ? b : a);
    return PS_OK;
  }
  if (fc_key(fcs, key, &l, &is_bias)) {
    fetch_fc(ctx, fcs[l], is_bias, which ? fcs[l].sW2 : fcs[l].sW1, which ? fcs[l].sb2 : fcs[l].sb1, out);
    return PS_OK;
  }
  if (has_wide && key == "wide.bias") {
    out.resize(1);
    PS_CUDA(cudaMemcpyAsync(out.data(), wide_bias + 1 + (which ? 1 : 0), sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    PS_CUDA(cudaStreamSynchronize(ctx->stream));
    return PS_OK;
  }
  return PS_NOT_FOUND;
}

void Model::put(const std::string& key, const float* in, int n) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  int field = 0; int64_t id = 0; int l = 0; bool is_bias = false;
  const int k = parse_key(key, &field, &id);
  if (k == 0) {
    PS_REQUIRE(has_emb && field >= 0 && field < F && n == D, PS_ERR_ARG, "put: bad embedding key or length");
    std::vector<float> tmp(in, in + n);
    int32_t f32 = field;
    emb.put_rows(&f32, &id, 1, tmp.data(), 1);
    return;
  }
  if (k == 1) {
    PS_REQUIRE(has_wide && n == 1, PS_ERR_ARG, "put: bad wide key or length");
    wide.put(id, in[0]);
    PS_CUDA(cudaStreamSynchronize(ctx->stream));
    wide.check_errors();
    return;
  }
  if (fc_key(fcs, key, &l, &is_bias)) {
    FcLayer& f = fcs[l];
    if (is_bias) {
      PS_REQUIRE(n == f.out, PS_ERR_ARG, "put: bias length mismatch");
      PS_CUDA(cudaMemcpyAsync(f.bias, in, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
      PS_CUDA(cudaStreamSynchronize(ctx->stream));
      return;
    }
    PS_REQUIRE(n == f.out * f.in, PS_ERR_ARG, "put: weight length mismatch");
    std::vector<float> tmp((size_t)f.out * f.ldw, 0.f), tmpt((size_t)f.in * f.ldwt, 0.f);
    for (int i = 0; i < f.in; ++i)
      for (int o = 0; o < f.out; ++o) {
        const float v = in[(size_t)o + (size_t)f.out * i];
        tmp[(size_t)o * f.ldw + i] = v; tmpt[(size_t)i * f.ldwt + o] = v;
      }
    PS_CUDA(cudaMemcpyAsync(f.W, tmp.data(), sizeof(float) * tmp.size(), cudaMemcpyHostToDevice, ctx->stream));
    PS_CUDA(cudaMemcpyAsync(f.Wt, tmpt.data(), sizeof(float) * tmpt.size(), cudaMemcpyHostToDevice, ctx->stream));
    f.refresh_lo(ctx);
    PS_CUDA(cudaStreamSynchronize(ctx->stream));
    return;
  }
  if (has_wide && key == "wide.bias") {
    PS_REQUIRE(n == 1, PS_ERR_ARG, "put: wide.bias is 1x1");
    PS_CUDA(cudaMemcpyAsync(wide_bias, in, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    PS_CUDA(cudaStreamSynchronize(ctx->stream));
    return;
  }
  PS_REQUIRE(false, PS_ERR_ARG, "put: unknown key");
}

int Model::push(const std::string& key, const float* g, int n, const ps_updater_spec& spec) {
  flush_deferred();
  PS_REQUIRE(in_flight == 0, PS_ERR_STATE, "model: steps in flight; collect first");
  int field = 0; int64_t id = 0; int l = 0; bool is_bias = false;
  const int k = parse_key(key, &field, &id);
  if (k == 0) {
    if (!has_emb || field < 0 || field >= F) return PS_NOT_FOUND;
    PS_REQUIRE(n == D, PS_ERR_ARG, "push: gradient length != embedding dimension");
    int32_t f32 = field, found = 0;
    emb.push_rows(&f32, &id, 1, g, spec, &found);
    return found ? PS_OK : PS_NOT_FOUND;
  }
  if (k == 1) {
    if (!has_wide) return PS_NOT_FOUND;
    PS_REQUIRE(n == 1, PS_ERR_ARG, "push: a wide weight is 1x1");
    return wide.push(id, g[0], spec) ? PS_OK : PS_NOT_FOUND;
  }
  const UpdaterDev u = make_updater_dev(spec);
  auto apply_in_place = [&](float* w, float* s1, float* s2, const std::vector<float>& gh) {
    float* gd = dmalloc<float>(gh.size());
    PS_CUDA(cudaMemcpyAsync(gd, gh.data(), sizeof(float) * gh.size(), cudaMemcpyHostToDevice, ctx->stream));
    updater_apply(ctx, u, w, s1, s2, gd, (int)gh.size());
    PS_CUDA(cudaStreamSynchronize(ctx->stream));
    dfree(gd);
  };
  if (fc_key(fcs, key, &l, &is_bias)) {
    FcLayer& f = fcs[l];
    if (is_bias) {
      PS_REQUIRE(n == f.out, PS_ERR_ARG, "push: bias length mismatch");
      apply_in_place(f.bias, f.sb1, f.sb2, std::vector<float>(g, g + n));
      return PS_OK;
    }
    PS_REQUIRE(n == f.out * f.in, PS_ERR_ARG, "push: weight length mismatch");
    /* the gradient in the layout of W and its states ([out][ldw], padding columns zero: a zero gradient leaves a zero weight with zero state
     * where it is under every updater); element 0 stays element (0, 0), which FtrlUpdater.java:52 looks at */
    std::vector<float> gh((size_t)f.out * f.ldw, 0.f);
    for (int i = 0; i < f.in; ++i)
      for (int o = 0; o < f.out; ++o) gh[(size_t)o * f.ldw + i] = g[(size_t)o + (size_t)f.out * i];
    apply_in_place(f.W, f.sW1, f.sW2, gh);
    std::vector<float> w;                        /* W changed: rebuild the transposed copy and the TF32 residuals the way put() does */
    fetch_fc(ctx, f, false, f.W, nullptr, w);
    put(key, w.data(), (int)w.size());
    return PS_OK;
  }
  if (has_wide && key == "wide.bias") {
    PS_REQUIRE(n == 1, PS_ERR_ARG, "push: wide.bias is 1x1");
    apply_in_place(wide_bias, wide_bias + 1, wide_bias + 2, std::vector<float>(g, g + 1));
    return PS_OK;
  }
  return PS_NOT_FOUND;
}

static void fetch_cols(Ctx* ctx, const float* src, int ldsrc, int cols, int N, std::vector<float>& out) {
  out.resize((size_t)N * cols);
  PS_CUDA(cudaMemcpy2DAsync(out.data(), sizeof(float) * cols, src, sizeof(float) * ldsrc, sizeof(float) * cols, N, cudaMemcpyDeviceToHost, ctx->stream));
  PS_CUDA(cudaStreamSynchronize(ctx->stream));
}

/* Layer.getA() / getDelta() of the last step.  delta semantics follow the reference's in-place
 * updates: fc<l>.delta = W_l^T d_l, later multiplied in place by the activation derivative of
 * fc<l-1> (FcLayer.java:100-102 acting on next.delta).                                         */
int Model::tap(const std::string& layer, int what, std::vector<float>& out) {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  const int N = last_N;
  if (N <= 0) return PS_NOT_FOUND;
  if (what == 1 && !last_train) return PS_NOT_FOUND;
  if (has_emb && layer == "embedding") { fetch_cols(ctx, what ? delta[0] : act[0], ld[0], what ? width[0] : F * D, N, out); return PS_OK; }
  if (has_emb && layer == "concat") { fetch_cols(ctx, what ? delta[0] : act[0], ld[0], width[0], N, out); return PS_OK; }
  for (int l = 0; l < L; ++l)
    if (layer == fcs[l].name) {
      if (what == 0) fetch_cols(ctx, act[l + 1], ld[l + 1], width[l + 1], N, out);
      else fetch_cols(ctx, delta[l], ld[l], width[l], N, out);
      return PS_OK;
    }
  if (has_wide && layer == "wide" && what == 0) { fetch_cols(ctx, wide_z, 1, 1, N, out); return PS_OK; }
  if (has_wide && layer == "addWideDeep") { fetch_cols(ctx, what ? delta[L] : P, what ? ld[L] : 1, 1, N, out); return PS_OK; }
  return PS_NOT_FOUND;
}

int64_t Model::num_keys() {
  flush_deferred();                              /* a sharded step may have left its update to "the next step" */
  int64_t n = 2 * (int64_t)L;
  if (has_emb) n += emb.size();
  if (has_wide) n += wide.size() + 1;
  return n;
}

}  // namespace psb

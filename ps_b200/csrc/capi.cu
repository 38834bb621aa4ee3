/*
 * capi.cu — the extern "C" surface declared in include/ps_b200.h.  Every entry point catches
 * psb::Error, stores the message for ps_last_error() and returns the code; nothing here computes
 * on the CPU — if no CUDA device is present ps_ctx_create fails with PS_ERR_CUDA.
 */
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "fcop.cuh"
#include "ingest.cuh"
#include "model.cuh"

namespace psb {
static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
}  // namespace psb

using namespace psb;

struct ps_ctx { Ctx c; uint32_t* ingest_ws = nullptr; size_t ingest_ws_cap = 0; };
struct ps_model { Model m; };
struct ps_fc { FcOp op; };
struct ps_reader { std::unique_ptr<LibsvmReader> r; };
struct ps_emb {
  Ctx* ctx = nullptr;
  EmbTable t;
  int Ncap = 0, lastN = 0;
  int64_t* dE = nullptr; float* dEf = nullptr; float* dOut = nullptr; float* dDelta = nullptr; size_t delta_cap = 0;
};

#define PS_TRY try {
#define PS_CATCH                                                                  \
  }                                                                               \
  catch (const psb::Error& e) { set_last_error(e.what()); return e.code; }        \
  catch (const std::exception& e) { set_last_error(e.what()); return PS_ERR_ARG; } \
  return PS_OK;

extern "C" {

const char* ps_last_error(void) { return g_last_error.c_str(); }
int ps_abi_version(void) { return 1; }

int ps_ctx_create(int device, uint64_t seed, ps_ctx** out) {
  PS_TRY
  PS_REQUIRE(out != nullptr, PS_ERR_ARG, "ps_ctx_create: out is null");
  int n = 0;
  PS_CUDA(cudaGetDeviceCount(&n));
  PS_REQUIRE(n > 0 && device >= 0 && device < n, PS_ERR_CUDA, "ps_ctx_create: no such CUDA device (there is no CPU fallback)");
  PS_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PS_CUDA(cudaGetDeviceProperties(&prop, device));
  PS_REQUIRE(prop.major >= 10, PS_ERR_CUDA, "ps_ctx_create: libps_b200 is built for sm_100a (Blackwell) only");
  ps_ctx* c = new ps_ctx();
  c->c.device = device; c->c.seed = seed; c->c.num_sms = prop.multiProcessorCount;
  /* the step's critical chain (forward, dgrad, embedding update) runs on the HIGHEST priority: its CTAs are dispatched
   * before those of the side branches (weight gradients, dense update) whenever both are pending (PS_STREAM_PRIO=0: off) */
  int prio_lo = 0, prio_hi = 0;
  PS_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  const char* pe = std::getenv("PS_STREAM_PRIO");
  c->c.prio_main = (pe && pe[0] == '0') ? 0 : prio_hi;
  c->c.prio_side = (pe && pe[0] == '0') ? 0 : prio_lo;
  const char* pd = std::getenv("PS_PDL");
  c->c.pdl = (pd && pd[0] == '0') ? 0 : 1;
  const char* pg = std::getenv("PS_PDL_GEMM");
  c->c.pdl_gemm = (pg && pg[0] == '1') ? 1 : 0;
  const char* pf = std::getenv("PS_P2P_DEFER");
  c->c.p2p_defer = pf ? (pf[0] == '0' ? 0 : pf[0] == '1' ? 1 : 2) : 2;
  const char* px = std::getenv("PS_PDL_EXCHANGE");
  c->c.pdl_exchange = (px && px[0] == '1') ? 1 : 0;
  const char* xe = std::getenv("PS_EXACT_UPDATERS");
  c->c.exact_updaters = (xe && xe[0] == '1') ? 1 : 0;
  const char* ht = std::getenv("PS_HOT_TMA");
  c->c.hot_tma = (ht && ht[0] == '0') ? 0 : 1;
  const char* td = std::getenv("PS_TC_DEEP");
  if (td && td[0] >= '0' && td[0] <= '2') c->c.tc_deep = td[0] - '0';
  const char* tw = std::getenv("PS_TC_DEEP_WGRAD");
  if (tw && tw[0] >= '0' && tw[0] <= '2') c->c.tc_deep_wgrad = tw[0] - '0';
  const char* wr = std::getenv("PS_TC_WIDE_RULE");
  c->c.tc_wide_rule = (wr && wr[0] == '0') ? 0 : 1;
  const char* gn = std::getenv("PS_GEMM_NARROW");
  c->c.gemm_narrow = (gn && gn[0] == '1') ? 1 : 0;
  const char* gw = std::getenv("PS_GROUP_WGRAD");
  c->c.group_wgrad = (gw && gw[0] == '1') ? 1 : 0;
  const char* us = std::getenv("PS_UPDATE_SLAB");
  c->c.update_slab = (us && us[0] == '0') ? 0 : 1;
  const char* ss = std::getenv("PS_SCATTER_SLAB");
  c->c.scatter_slab = (ss && ss[0] == '0') ? 0 : 1;
  const char* hs = std::getenv("PS_HOT_SHARE");
  if (hs && hs[0]) c->c.hot_share = std::max(1, std::atoi(hs));
  const char* hm = std::getenv("PS_HOT_MIN");
  if (hm && hm[0]) { const long v = std::atol(hm); c->c.hot_min = v <= 0 ? 0xFFFFFFFFu : (unsigned)v; }
  PS_CUDA(cudaStreamCreateWithPriority(&c->c.stream, cudaStreamNonBlocking, c->c.prio_main));
  PS_CUDA(cudaStreamCreateWithFlags(&c->c.copy_stream, cudaStreamNonBlocking));
  *out = c;
  PS_CATCH
}
int ps_ctx_destroy(ps_ctx* ctx) {
  PS_TRY
  if (ctx) {
    cudaStreamSynchronize(ctx->c.stream);
    cudaStreamDestroy(ctx->c.stream); cudaStreamDestroy(ctx->c.copy_stream);
    dfree(ctx->ingest_ws);
    delete ctx;
  }
  PS_CATCH
}
int ps_ctx_set_fc_precision(ps_ctx* ctx, int mode) {
  PS_TRY
  PS_REQUIRE(ctx && (mode == PS_FC_FP32 || mode == PS_FC_TF32 || mode == PS_FC_TF32X3), PS_ERR_ARG, "ps_ctx_set_fc_precision: bad mode");
  ctx->c.fc_precision = mode;
  PS_CATCH
}
int ps_ctx_get_fc_precision(ps_ctx* ctx, int* mode) {
  PS_TRY
  PS_REQUIRE(ctx && mode, PS_ERR_ARG, "ps_ctx_get_fc_precision: null argument");
  *mode = ctx->c.fc_precision;
  PS_CATCH
}
int ps_ctx_set_exact_updaters(ps_ctx* ctx, int on) {
  PS_TRY
  PS_REQUIRE(ctx != nullptr, PS_ERR_ARG, "ps_ctx_set_exact_updaters: null context");
  ctx->c.exact_updaters = on ? 1 : 0;
  PS_CATCH
}
int ps_ctx_make_current(ps_ctx* ctx) {
  PS_TRY
  PS_REQUIRE(ctx, PS_ERR_ARG, "null ctx");
  PS_CUDA(cudaSetDevice(ctx->c.device));
  PS_CATCH
}
int ps_ctx_synchronize(ps_ctx* ctx) {
  PS_TRY
  PS_REQUIRE(ctx, PS_ERR_ARG, "null ctx");
  PS_CUDA(cudaStreamSynchronize(ctx->c.copy_stream));
  PS_CUDA(cudaStreamSynchronize(ctx->c.stream));
  PS_CATCH
}
int ps_ctx_launch_count(ps_ctx* ctx, int64_t* out) {
  PS_TRY
  PS_REQUIRE(ctx && out, PS_ERR_ARG, "null argument");
  *out = ctx->c.launches;
  PS_CATCH
}
int ps_ctx_device_info(ps_ctx* ctx, char* name, int cap, int* sms, int* cc_major, int* cc_minor) {
  PS_TRY
  PS_REQUIRE(ctx, PS_ERR_ARG, "null ctx");
  cudaDeviceProp prop;
  PS_CUDA(cudaGetDeviceProperties(&prop, ctx->c.device));
  if (name && cap > 0) { std::strncpy(name, prop.name, cap - 1); name[cap - 1] = 0; }
  if (sms) *sms = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  PS_CATCH
}
int ps_ctx_stream(ps_ctx* ctx, void** stream) {   /* for callers that time with events on the library's stream */
  PS_TRY
  PS_REQUIRE(ctx && stream, PS_ERR_ARG, "null argument");
  *stream = (void*)ctx->c.stream;
  PS_CATCH
}

int ps_host_alloc(size_t bytes, void** out) {
  PS_TRY
  PS_REQUIRE(out, PS_ERR_ARG, "null out");
  PS_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  PS_CATCH
}
int ps_host_free(void* p) {
  PS_TRY
  if (p) PS_CUDA(cudaFreeHost(p));
  PS_CATCH
}

/* ---- update.* ---- */
static bool between(const std::string& s, const std::string& open, float* v) {   /* StringUtils.substringBetween(str, open, "@") */
  const size_t a = s.find(open);
  if (a == std::string::npos) return false;
  const size_t b = s.find('@', a + open.size());
  if (b == std::string::npos) return false;
  const std::string t = s.substr(a + open.size(), b - a - open.size());
  char* end = nullptr;
  *v = std::strtof(t.c_str(), &end);
  return end != t.c_str();
}

int ps_updater_parse(const char* name, ps_updater_spec* out) {
  PS_TRY
  PS_REQUIRE(name && out, PS_ERR_ARG, "null argument");
  const std::string s(name);
  ps_updater_spec r{};
  if (s.find("l1:") != std::string::npos) {       /* FtrlUpdater(String), FtrlUpdater.java:44-49 */
    r.kind = PS_UPD_FTRL;
    PS_REQUIRE(between(s, "alfa:", &r.p[0]) && between(s, "beta:", &r.p[1]) && between(s, "l1:", &r.p[2]) && between(s, "l2:", &r.p[3]),
               PS_ERR_ARG, "ps_updater_parse: malformed ftrl name");
  } else if (s.find("beta1:") != std::string::npos) {   /* AdamUpdater(String), AdamUpdater.java:50-55 */
    r.kind = PS_UPD_ADAM;
    PS_REQUIRE(between(s, "alfa:", &r.p[0]) && between(s, "beta1:", &r.p[1]) && between(s, "beta2:", &r.p[2]) && between(s, "epsilon:", &r.p[3]),
               PS_ERR_ARG, "ps_updater_parse: malformed adam name");
  } else if (s.find("eta:") != std::string::npos) {
    r.kind = PS_UPD_SIMPLE;
    PS_REQUIRE(between(s, "eta:", &r.p[0]), PS_ERR_ARG, "ps_updater_parse: malformed simple name");
  } else {
    PS_REQUIRE(false, PS_ERR_ARG, "ps_updater_parse: unknown updater name");
  }
  *out = r;
  PS_CATCH
}

/* Float.toString: shortest decimal that round-trips; plain notation in [1e-3, 1e7), else d.dddE<n> */
static std::string java_float(float v) {
  if (v == 0.0f) return std::signbit(v) ? "-0.0" : "0.0";
  char buf[64];
  int prec = 1;
  for (; prec <= 9; ++prec) { snprintf(buf, sizeof buf, "%.*e", prec - 1, (double)v); if (std::strtof(buf, nullptr) == v) break; }
  std::string m(buf);
  const size_t epos = m.find('e');
  std::string digits = m.substr(0, epos);
  const int ex = std::atoi(m.c_str() + epos + 1);
  const bool neg = digits[0] == '-';
  if (neg) digits = digits.substr(1);
  std::string d;
  for (char ch : digits) if (ch != '.') d.push_back(ch);
  std::string r;
  const float a = std::fabs(v);
  if (a >= 1e-3f && a < 1e7f) {
    if (ex >= 0) {
      std::string ip = d.substr(0, std::min<size_t>(d.size(), (size_t)ex + 1));
      while ((int)ip.size() < ex + 1) ip.push_back('0');
      std::string fp = (int)d.size() > ex + 1 ? d.substr(ex + 1) : "0";
      r = ip + "." + fp;
    } else {
      r = "0." + std::string((size_t)(-ex - 1), '0') + d;
    }
  } else {
    r = d.substr(0, 1) + "." + (d.size() > 1 ? d.substr(1) : "0") + "E" + std::to_string(ex);
  }
  return neg ? "-" + r : r;
}

int ps_updater_name(const ps_updater_spec* spec, char* buf, int cap) {
  PS_TRY
  PS_REQUIRE(spec && buf && cap > 0, PS_ERR_ARG, "null argument");
  std::string s;
  const float* p = spec->p;
  if (spec->kind == PS_UPD_ADAM) s = "adam@alfa:" + java_float(p[0]) + "@beta1:" + java_float(p[1]) + "@beta2:" + java_float(p[2]) + "@epsilon:" + java_float(p[3]) + "@";
  else if (spec->kind == PS_UPD_FTRL) s = "adam@alfa:" + java_float(p[0]) + "@beta:" + java_float(p[1]) + "@l1:" + java_float(p[2]) + "@l2:" + java_float(p[3]) + "@";
  else s = "simple@eta:" + java_float(p[0]) + "@";
  PS_REQUIRE((int)s.size() + 1 <= cap, PS_ERR_ARG, "ps_updater_name: buffer too small");
  std::memcpy(buf, s.c_str(), s.size() + 1);
  PS_CATCH
}

int ps_updater_apply(ps_ctx* ctx, const ps_updater_spec* spec, float* w, float* s1, float* s2, const float* g, int n) {
  PS_TRY
  PS_REQUIRE(ctx && spec && w && s1 && s2 && g && n > 0, PS_ERR_ARG, "null argument");
  cudaStream_t st = ctx->c.stream;
  float* d = dmalloc<float>((size_t)4 * n);
  PS_CUDA(cudaMemcpyAsync(d, w, sizeof(float) * n, cudaMemcpyHostToDevice, st));
  PS_CUDA(cudaMemcpyAsync(d + n, s1, sizeof(float) * n, cudaMemcpyHostToDevice, st));
  PS_CUDA(cudaMemcpyAsync(d + 2 * n, s2, sizeof(float) * n, cudaMemcpyHostToDevice, st));
  PS_CUDA(cudaMemcpyAsync(d + 3 * n, g, sizeof(float) * n, cudaMemcpyHostToDevice, st));
  updater_apply(&ctx->c, make_updater_dev(*spec), d, d + n, d + 2 * n, d + 3 * n, n);
  PS_CUDA(cudaMemcpyAsync(w, d, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
  PS_CUDA(cudaMemcpyAsync(s1, d + n, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
  PS_CUDA(cudaMemcpyAsync(s2, d + 2 * n, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
  PS_CUDA(cudaStreamSynchronize(st));
  dfree(d);
  PS_CATCH
}

/* ---- layer.EmbeddingLayer ---- */
int ps_emb_create(ps_ctx* ctx, int F, int D, int64_t capacity, const ps_updater_spec* upd, ps_emb** out) {
  PS_TRY
  PS_REQUIRE(ctx && out, PS_ERR_ARG, "null argument");
  ps_updater_spec u;
  if (upd) u = *upd;
  else { u.kind = PS_UPD_ADAM; u.p[0] = (float)0.005; u.p[1] = (float)0.9; u.p[2] = (float)0.999; u.p[3] = (float)std::pow(10.0, -8); }
  ps_emb* e = new ps_emb();
  e->ctx = &ctx->c;
  e->t.create(&ctx->c, F, D, capacity, u, 1);
  PS_CUDA(cudaStreamSynchronize(ctx->c.stream));
  *out = e;
  PS_CATCH
}
int ps_emb_destroy(ps_emb* e) {
  PS_TRY
  if (e) {
    cudaStreamSynchronize(e->ctx->stream);
    e->t.destroy(); dfree(e->dE); dfree(e->dEf); dfree(e->dOut); dfree(e->dDelta);
    delete e;
  }
  PS_CATCH
}
static void emb_reserve(ps_emb* e, int N) {
  if (N <= e->Ncap) return;
  PS_CUDA(cudaStreamSynchronize(e->ctx->stream));
  dfree(e->dE); dfree(e->dEf); dfree(e->dOut);
  e->Ncap = N;
  const size_t L = (size_t)N * e->t.F;
  e->dE = dmalloc<int64_t>(L); e->dEf = dmalloc<float>(L);
  e->dOut = dmalloc<float>(L * e->t.D);
  e->t.reserve((int64_t)L);
}
static int emb_forward_impl(ps_emb* e, const int64_t* E, const float* Ef, int N, float* out) {
  PS_TRY
  PS_REQUIRE(e && (E || Ef) && out && N > 0, PS_ERR_ARG, "null argument");
  if (e->lastN) { e->t.clear_batch(); e->lastN = 0; }   /* a forward that was never followed by backward: drop its counts BEFORE the workspace may be reallocated */
  emb_reserve(e, N);
  cudaStream_t st = e->ctx->stream;
  const size_t L = (size_t)N * e->t.F;
  if (E) PS_CUDA(cudaMemcpyAsync(e->dE, E, sizeof(int64_t) * L, cudaMemcpyHostToDevice, st));
  else PS_CUDA(cudaMemcpyAsync(e->dEf, Ef, sizeof(float) * L, cudaMemcpyHostToDevice, st));
  e->t.lookup(E ? e->dE : nullptr, E ? nullptr : e->dEf, N, e->dOut, e->t.F * e->t.D);
  PS_CUDA(cudaMemcpyAsync(out, e->dOut, sizeof(float) * L * e->t.D, cudaMemcpyDeviceToHost, st));
  e->lastN = N;
  e->t.check_errors();
  PS_CATCH
}
int ps_emb_forward(ps_emb* e, const int64_t* E, int N, float* out) { return emb_forward_impl(e, E, nullptr, N, out); }
int ps_emb_forward_f32ids(ps_emb* e, const float* E, int N, float* out) { return emb_forward_impl(e, nullptr, E, N, out); }

int ps_emb_backward_update(ps_emb* e, const float* delta, int ld, int N, int calls) {
  PS_TRY
  PS_REQUIRE(e && delta && N > 0, PS_ERR_ARG, "null argument");
  PS_REQUIRE(N == e->lastN, PS_ERR_STATE, "ps_emb_backward_update: no matching forward");
  PS_REQUIRE(ld >= e->t.F * e->t.D, PS_ERR_ARG, "ps_emb_backward_update: delta has fewer rows than F*D");
  cudaStream_t st = e->ctx->stream;
  const size_t need = (size_t)N * ld;
  if (need > e->delta_cap) { PS_CUDA(cudaStreamSynchronize(st)); dfree(e->dDelta); e->dDelta = dmalloc<float>(need); e->delta_cap = need; }
  PS_CUDA(cudaMemcpyAsync(e->dDelta, delta, sizeof(float) * need, cudaMemcpyHostToDevice, st));
  e->t.scatter_update(e->dDelta, ld, nullptr, 0, N, calls, nullptr, 0, true);   /* ReLU mask: the bits the forward recorded */
  e->lastN = 0;
  PS_CUDA(cudaStreamSynchronize(st));
  PS_CATCH
}
int ps_emb_get_rows(ps_emb* e, const int32_t* fields, const int64_t* ids, int n, float* w, float* s1, float* s2, int32_t* found) {
  PS_TRY
  PS_REQUIRE(e && fields && ids && w && found && n >= 0, PS_ERR_ARG, "null argument");
  e->t.get_rows(fields, ids, n, w, s1, s2, found);
  PS_CATCH
}
int ps_emb_put_rows(ps_emb* e, const int32_t* fields, const int64_t* ids, int n, float* w, int replace) {
  PS_TRY
  PS_REQUIRE(e && fields && ids && w && n >= 0, PS_ERR_ARG, "null argument");
  e->t.put_rows(fields, ids, n, w, replace);
  PS_CATCH
}
int ps_emb_size(ps_emb* e, int64_t* rows) {
  PS_TRY
  PS_REQUIRE(e && rows, PS_ERR_ARG, "null argument");
  *rows = e->t.size();
  PS_CATCH
}

/* ---- model.* ---- */
int ps_model_create(ps_ctx* ctx, int kind, int F, int D, int Xn, const int32_t* fc_dims, int n_fc, int64_t emb_capacity,
                    const ps_updater_spec* emb_updater, int max_batch, ps_model** out) {
  PS_TRY
  PS_REQUIRE(ctx && fc_dims && out, PS_ERR_ARG, "null argument");
  ps_model* m = new ps_model();
  try { m->m.create(&ctx->c, kind, F, D, Xn, fc_dims, n_fc, emb_capacity, emb_updater, max_batch); }
  catch (...) { delete m; throw; }
  *out = m;
  PS_CATCH
}
int ps_model_destroy(ps_model* m) {
  PS_TRY
  if (m) { m->m.destroy(); delete m; }
  PS_CATCH
}
int ps_model_submit(ps_model* m, const int64_t* E, const float* X, const int64_t* W, const float* Y, int N) {
  PS_TRY
  PS_REQUIRE(m, PS_ERR_ARG, "null model");
  HostBatch b; b.E = E; b.X = X; b.W = W; b.Y = Y; b.N = N;
  m->m.submit(b);
  PS_CATCH
}
int ps_model_collect(ps_model* m, float* loss) {
  PS_TRY
  PS_REQUIRE(m && loss, PS_ERR_ARG, "null argument");
  *loss = m->m.collect();
  PS_CATCH
}
int ps_model_train_step(ps_model* m, const int64_t* E, const float* X, const int64_t* W, const float* Y, int N, float* loss) {
  PS_TRY
  PS_REQUIRE(m && loss, PS_ERR_ARG, "null argument");
  HostBatch b; b.E = E; b.X = X; b.W = W; b.Y = Y; b.N = N;
  m->m.submit(b);
  *loss = m->m.collect();
  PS_CATCH
}
/* ---- DNN.train / WideDeepNN.train call by call (the loss stays in the caller) ---- */
int ps_model_forward(ps_model* m, const int64_t* E, const float* X, const int64_t* W, int N, float* P_out) {
  PS_TRY
  PS_REQUIRE(m && X && P_out, PS_ERR_ARG, "null argument");
  HostBatch b; b.E = E; b.X = X; b.W = W; b.Y = nullptr; b.N = N;
  m->m.forward_host(b, P_out);
  PS_CATCH
}
int ps_model_backward_update(ps_model* m, const float* delta_top, int N, float loss) {
  PS_TRY
  PS_REQUIRE(m && delta_top, PS_ERR_ARG, "null argument");
  m->m.backward_update_host(delta_top, N, loss);
  PS_CATCH
}
int ps_model_submit_text(ps_model* m, const char* text, size_t len, int N) {
  PS_TRY
  PS_REQUIRE(m && text, PS_ERR_ARG, "null argument");
  m->m.submit_text(text, len, N);
  PS_CATCH
}
int ps_model_shape(ps_model* m, int* F, int* D, int* Xn) {
  PS_TRY
  PS_REQUIRE(m != nullptr, PS_ERR_ARG, "null model");
  if (F) *F = m->m.F;
  if (D) *D = m->m.D;
  if (Xn) *Xn = m->m.Xn;
  PS_CATCH
}
int ps_model_step_info(ps_model* m, int* skipped, uint32_t* bad_lines, uint32_t* n_unique) {
  PS_TRY
  PS_REQUIRE(m != nullptr, PS_ERR_ARG, "null model");
  if (skipped) *skipped = m->m.last_status.skip;
  if (bad_lines) *bad_lines = m->m.last_status.pad;
  if (n_unique) *n_unique = m->m.last_status.n_unique;
  PS_CATCH
}
int ps_model_train_step_dev(ps_model* m, const int64_t* E_dev, const float* X_dev, const int64_t* W_dev, const float* Y_dev, int N) {
  PS_TRY
  PS_REQUIRE(m, PS_ERR_ARG, "null model");
  PS_REQUIRE(m->m.in_flight == 0, PS_ERR_STATE, "host steps in flight; collect first");
  m->m.run_step(E_dev, X_dev, W_dev, Y_dev, N, true, nullptr);
  m->m.last_N = N; m->m.last_train = true;
  PS_CATCH
}
int ps_model_read_loss(ps_model* m, float* loss) {
  PS_TRY
  PS_REQUIRE(m && loss, PS_ERR_ARG, "null argument");
  *loss = m->m.read_loss();
  PS_CATCH
}
int ps_model_loss_dev(ps_model* m, const float** loss_dev) {
  PS_TRY
  PS_REQUIRE(m && loss_dev, PS_ERR_ARG, "null argument");
  *loss_dev = reinterpret_cast<const float*>(reinterpret_cast<const char*>(m->m.st_dev) + offsetof(StepStatus, loss));
  PS_CATCH
}
int ps_model_predict(ps_model* m, const int64_t* E, const float* X, const int64_t* W, int N, float* out) {
  PS_TRY
  PS_REQUIRE(m && X && out, PS_ERR_ARG, "null argument");
  HostBatch b; b.E = E; b.X = X; b.W = W; b.Y = nullptr; b.N = N;
  m->m.predict(b, out);
  PS_CATCH
}
static int copy_out(const std::vector<float>& v, float* out, int cap, int* n) {
  if (n) *n = (int)v.size();
  if (out && cap >= (int)v.size()) std::memcpy(out, v.data(), sizeof(float) * v.size());
  return PS_OK;
}
int ps_model_get(ps_model* m, const char* key, float* out, int cap, int* n) {
  PS_TRY
  PS_REQUIRE(m && key, PS_ERR_ARG, "null argument");
  std::vector<float> v;
  const int rc = m->m.get(key, v);
  if (rc != PS_OK) { set_last_error(std::string("key absent: ") + key); return rc; }
  copy_out(v, out, cap, n);
  PS_CATCH
}
int ps_model_put(ps_model* m, const char* key, const float* in, int n) {
  PS_TRY
  PS_REQUIRE(m && key && in, PS_ERR_ARG, "null argument");
  m->m.put(key, in, n);
  PS_CATCH
}
/* PSClient.getList / PServer.getList (PSClient.java:72-97, PServer.java:102-117) and PSClient.updateList / PServer.upsertList
 * (PSClient.java:128-151, PServer.java:144-162) by reference key strings.  Embedding keys ("emF<j>.<id>") — the bulk of any
 * list — go through ONE batched lookup / insert kernel; any other key (dense parameters, wide weights) is served one by one. */
int ps_model_get_list(ps_model* m, const char* const* keys, int n, float* out, int stride, int32_t* found) {
  PS_TRY
  if (m) m->m.flush_deferred();
  PS_REQUIRE(m && keys && out && found && n >= 0 && stride > 0, PS_ERR_ARG, "ps_model_get_list: bad argument");
  Model& M = m->m;
  std::vector<int32_t> fields, idx, fnd;
  std::vector<int64_t> ids;
  for (int i = 0; i < n; ++i) {
    int f = 0; int64_t id = 0;
    found[i] = 0;
    if (parse_key(keys[i], &f, &id) == 0 && M.has_emb && f >= 0 && f < M.F && id >= 0) {
      PS_REQUIRE(stride >= M.D, PS_ERR_ARG, "ps_model_get_list: stride smaller than the embedding dimension");
      fields.push_back(f); ids.push_back(id); idx.push_back(i);
    } else {
      std::vector<float> v;
      if (M.get(keys[i], v) == PS_OK) {
        PS_REQUIRE((int)v.size() <= stride, PS_ERR_ARG, "ps_model_get_list: stride smaller than a listed parameter");
        std::memcpy(out + (size_t)i * stride, v.data(), sizeof(float) * v.size());
        found[i] = (int32_t)v.size();
      }
    }
  }
  if (!idx.empty()) {
    const int k = (int)idx.size();
    std::vector<float> w((size_t)k * M.D);
    fnd.assign(k, 0);
    M.emb.get_rows(fields.data(), ids.data(), k, w.data(), nullptr, nullptr, fnd.data());
    for (int q = 0; q < k; ++q)
      if (fnd[q]) { std::memcpy(out + (size_t)idx[q] * stride, w.data() + (size_t)q * M.D, sizeof(float) * M.D); found[idx[q]] = M.D; }
  }
  PS_CATCH
}
int ps_model_update_list(ps_model* m, const char* const* keys, int n, float* io, int stride, const int32_t* lens, int replace) {
  PS_TRY
  if (m) m->m.flush_deferred();
  PS_REQUIRE(m && keys && io && lens && n >= 0 && stride > 0, PS_ERR_ARG, "ps_model_update_list: bad argument");
  Model& M = m->m;
  std::vector<int32_t> fields, idx;
  std::vector<int64_t> ids;
  for (int i = 0; i < n; ++i) {
    int f = 0; int64_t id = 0;
    if (parse_key(keys[i], &f, &id) == 0 && M.has_emb) {
      PS_REQUIRE(lens[i] == M.D && stride >= M.D, PS_ERR_ARG, "ps_model_update_list: embedding row length mismatch");
      fields.push_back(f); ids.push_back(id); idx.push_back(i);
    } else {
      std::vector<float> cur;
      if (!replace && M.get(keys[i], cur) == PS_OK) {          /* insert-if-absent: the caller receives the winner (PServer.java:150-158) */
        PS_REQUIRE((int)cur.size() <= stride, PS_ERR_ARG, "ps_model_update_list: stride smaller than a listed parameter");
        std::memcpy(io + (size_t)i * stride, cur.data(), sizeof(float) * cur.size());
      } else {
        M.put(keys[i], io + (size_t)i * stride, lens[i]);
      }
    }
  }
  if (!idx.empty()) {
    const int k = (int)idx.size();
    std::vector<float> w((size_t)k * M.D);
    for (int q = 0; q < k; ++q) std::memcpy(w.data() + (size_t)q * M.D, io + (size_t)idx[q] * stride, sizeof(float) * M.D);
    M.emb.put_rows(fields.data(), ids.data(), k, w.data(), replace);
    for (int q = 0; q < k; ++q) std::memcpy(io + (size_t)idx[q] * stride, w.data() + (size_t)q * M.D, sizeof(float) * M.D);
  }
  PS_CATCH
}
int ps_model_push(ps_model* m, const char* key, const float* grad, int n, const ps_updater_spec* spec) {
  PS_TRY
  PS_REQUIRE(m && key && grad && spec && n > 0, PS_ERR_ARG, "bad argument");
  return m->m.push(key, grad, n, *spec);
  PS_CATCH
}
int ps_model_get_state(ps_model* m, const char* key, int which, float* out, int cap, int* n) {
  PS_TRY
  PS_REQUIRE(m && key, PS_ERR_ARG, "null argument");
  std::vector<float> v;
  const int rc = m->m.get_state(key, which, v);
  if (rc != PS_OK) { set_last_error(std::string("key absent: ") + key); return rc; }
  copy_out(v, out, cap, n);
  PS_CATCH
}
int ps_model_tap(ps_model* m, const char* layer, int what, float* out, int cap, int* n) {
  PS_TRY
  PS_REQUIRE(m && layer, PS_ERR_ARG, "null argument");
  std::vector<float> v;
  const int rc = m->m.tap(layer, what, v);
  if (rc != PS_OK) { set_last_error(std::string("no such tap: ") + layer); return rc; }
  copy_out(v, out, cap, n);
  PS_CATCH
}
int ps_model_num_keys(ps_model* m, int64_t* out) {
  PS_TRY
  PS_REQUIRE(m && out, PS_ERR_ARG, "null argument");
  *out = m->m.num_keys();
  PS_CATCH
}
int ps_model_skipped_backward(ps_model* m, int* out) {
  PS_TRY
  PS_REQUIRE(m && out, PS_ERR_ARG, "null argument");
  *out = m->m.last_status.skip;
  PS_CATCH
}
int ps_model_profile(ps_model* m, int enable) {
  PS_TRY
  PS_REQUIRE(m, PS_ERR_ARG, "null model");
  m->m.profile = enable != 0;
  PS_CATCH
}
int ps_model_phase_times(ps_model* m, float* ms, int cap, int* n, char* names, int names_cap) {
  PS_TRY
  PS_REQUIRE(m && n, PS_ERR_ARG, "null argument");
  const auto& v = m->m.phase_ms;
  *n = (int)v.size();
  if (ms && cap >= (int)v.size()) std::memcpy(ms, v.data(), sizeof(float) * v.size());
  if (names && names_cap > 0) {
    std::string s;
    for (size_t i = 1; i < m->m.phase_names.size() && i - 1 < v.size(); ++i) { if (i > 1) s += ";"; s += m->m.phase_names[i]; }
    std::strncpy(names, s.c_str(), names_cap - 1); names[names_cap - 1] = 0;
  }
  PS_CATCH
}

int ps_model_kernel_times(ps_model* m, const int64_t* const* E_dev_ring, int n_ring, int N, int reps, float* us) {
  PS_TRY
  PS_REQUIRE(m && E_dev_ring && us, PS_ERR_ARG, "null argument");
  m->m.kernel_times(E_dev_ring, n_ring, N, reps, us);
  PS_CATCH
}

int ps_model_gemm_times(ps_model* m, int N, int reps, float* us, int cap) {
  PS_TRY
  PS_REQUIRE(m && us && cap >= 3 * m->m.L, PS_ERR_ARG, "bad argument");
  m->m.gemm_times(N, reps, us);
  PS_CATCH
}

/* ---- sharded table ---- */
int ps_key_owner(const char* key, int n_shards, int* owner) {
  PS_TRY
  PS_REQUIRE(key && owner && n_shards > 0, PS_ERR_ARG, "bad argument");
  int field = 0; int64_t id = 0;
  if (psb::parse_key(key, &field, &id) != 0) { *owner = -1; return PS_OK; }      /* wide / dense: replicated */
  PS_REQUIRE(field >= 0 && id >= 0 && id <= (int64_t)PS_KEY_ID_MASK, PS_ERR_ARG, "embedding key outside the (field, id) domain");
  *owner = (int)ps_owner_of(ps_pack_key((uint32_t)field, (uint64_t)id), (uint32_t)n_shards);
  PS_CATCH
}
int ps_shard_route_dev(ps_ctx* ctx, const int64_t* E_dev, int N, int F, int R, uint64_t* send_keys_dev, int32_t* send_pos_dev,
                       int32_t* counts_dev, int32_t* cursor_dev) {
  PS_TRY
  PS_REQUIRE(ctx && E_dev && send_keys_dev && send_pos_dev && counts_dev && cursor_dev && N > 0 && F > 0 && R > 0, PS_ERR_ARG, "bad argument");
  PS_CUDA(cudaMemsetAsync(counts_dev, 0, sizeof(int32_t) * R, ctx->c.stream));
  PS_CUDA(cudaMemsetAsync(cursor_dev, 0, sizeof(int32_t) * R, ctx->c.stream));
  shard_count(&ctx->c, E_dev, N, F, R, counts_dev);
  shard_place(&ctx->c, E_dev, N, F, R, counts_dev, cursor_dev, send_keys_dev, send_pos_dev);
  PS_CATCH
}
int ps_shard_route_padded_dev(ps_ctx* ctx, const int64_t* E_dev, int N, int F, int R, int cap, uint64_t* send_keys_dev, int32_t* send_pos_dev,
                              int32_t* cursor_dev, int32_t* overflow_dev) {
  PS_TRY
  PS_REQUIRE(ctx && E_dev && send_keys_dev && send_pos_dev && cursor_dev && overflow_dev && N > 0 && F > 0 && R > 0 && cap > 0, PS_ERR_ARG, "bad argument");
  shard_place_padded(&ctx->c, E_dev, N, F, R, cap, cursor_dev, send_keys_dev, send_pos_dev, overflow_dev);
  PS_CATCH
}
int ps_model_shard_lookup_dev(ps_model* m, const uint64_t* keys_dev, int n, float* rows_out_dev) {
  PS_TRY
  PS_REQUIRE(m && n >= 0, PS_ERR_ARG, "bad argument");
  m->m.shard_emb_lookup(keys_dev, n, rows_out_dev);
  PS_CATCH
}
int ps_model_shard_row_stride(ps_model* m, int* Dp) {
  PS_TRY
  PS_REQUIRE(m && Dp && m->m.has_emb, PS_ERR_ARG, "bad argument");
  *Dp = m->m.emb.Dp;
  PS_CATCH
}
int ps_model_shard_unpack_dev(ps_model* m, const float* rows_dev, const int32_t* send_pos_dev, int N) {
  PS_TRY
  PS_REQUIRE(m && rows_dev && send_pos_dev, PS_ERR_ARG, "bad argument");
  m->m.shard_unpack_rows(rows_dev, send_pos_dev, N);
  PS_CATCH
}
int ps_model_shard_dense_step_dev(ps_model* m, const float* X_dev, const int64_t* W_local_dev, const int64_t* W_all_dev, int n_all,
                                  const float* Y_dev, int N) {
  PS_TRY
  PS_REQUIRE(m && X_dev && Y_dev, PS_ERR_ARG, "bad argument");
  m->m.shard_dense_step(X_dev, W_local_dev, W_all_dev, n_all, Y_dev, N);
  PS_CATCH
}
int ps_model_shard_grad_buffer(ps_model* m, float** buf_dev, int64_t* count) {
  PS_TRY
  PS_REQUIRE(m && buf_dev && count, PS_ERR_ARG, "bad argument");
  if (!m->m.gsum) {
    const DenseUpdateArgs u = m->m.dense_args(1);
    m->m.gsum_len = u.total + 3;
    m->m.gsum = dmalloc_zero<float>((size_t)m->m.gsum_len, m->m.ctx->stream);
    PS_CUDA(cudaStreamSynchronize(m->m.ctx->stream));
  }
  *buf_dev = m->m.gsum; *count = m->m.gsum_len;
  PS_CATCH
}
int ps_model_shard_pack_grads_dev(ps_model* m, const int32_t* send_pos_dev, int N, float* grads_send_dev) {
  PS_TRY
  PS_REQUIRE(m && send_pos_dev && grads_send_dev, PS_ERR_ARG, "bad argument");
  m->m.shard_pack(send_pos_dev, N, grads_send_dev);
  PS_CATCH
}
int ps_model_shard_finish_dev(ps_model* m, int N_global, int R) {
  PS_TRY
  PS_REQUIRE(m && N_global > 0 && R > 0, PS_ERR_ARG, "bad argument");
  m->m.shard_finish(N_global, R);
  PS_CATCH
}
int ps_model_shard_apply_dev(ps_model* m, const float* grads_recv_dev, int n) {
  PS_TRY
  PS_REQUIRE(m && n >= 0, PS_ERR_ARG, "bad argument");
  m->m.shard_emb_apply(grads_recv_dev, n);
  PS_CATCH
}

int ps_model_p2p_init(ps_model* m, int R, int rank, int cap, void* ipc_handle_out64) {
  PS_TRY
  PS_REQUIRE(m && ipc_handle_out64, PS_ERR_ARG, "null argument");
  m->m.p2p_init(R, rank, cap, ipc_handle_out64);
  PS_CATCH
}
int ps_model_p2p_connect(ps_model* m, const void* all_handles) {
  PS_TRY
  PS_REQUIRE(m && all_handles, PS_ERR_ARG, "null argument");
  m->m.p2p_connect(all_handles);
  PS_CATCH
}
int ps_model_p2p_step_dev(ps_model* m, const int64_t* E_dev, const float* X_dev, const int64_t* W_dev, const float* Y_dev, int N) {
  PS_TRY
  PS_REQUIRE(m && X_dev && Y_dev, PS_ERR_ARG, "null argument");
  PS_REQUIRE(m->m.in_flight == 0, PS_ERR_STATE, "host steps in flight; collect first");
  m->m.run_step(E_dev, X_dev, W_dev, Y_dev, N, true, nullptr, 1);
  PS_CATCH
}
int ps_model_p2p_submit(ps_model* m, const int64_t* E, const float* X, const int64_t* W, const float* Y, int N) {
  PS_TRY
  PS_REQUIRE(m != nullptr, PS_ERR_ARG, "null model");
  HostBatch b; b.E = E; b.X = X; b.W = W; b.Y = Y; b.N = N;
  m->m.submit(b, 1);
  PS_CATCH
}
int ps_model_p2p_overflowed(ps_model* m, int* out) {
  PS_TRY
  PS_REQUIRE(m && out, PS_ERR_ARG, "null argument");
  *out = m->m.p2p.slab ? (m->m.p2p.overflowed() ? 1 : 0) : 0;
  PS_CATCH
}

/* ---- test hook ---- */
/* ---- store dump / load ---- */
int ps_model_save(ps_model* m, const char* path) {
  PS_TRY
  PS_REQUIRE(m && path, PS_ERR_ARG, "ps_model_save: null argument");
  m->m.save(path);
  PS_CATCH
}
int ps_model_load(ps_model* m, const char* path) {
  PS_TRY
  PS_REQUIRE(m && path, PS_ERR_ARG, "ps_model_load: null argument");
  m->m.load(path);
  PS_CATCH
}

/* ---- layer.FcLayer standalone ---- */
int ps_fc_create(ps_ctx* ctx, const char* name, int in, int out, int act, const ps_updater_spec* upd, int max_batch, ps_fc** out_fc) {
  PS_TRY
  PS_REQUIRE(ctx && name && out_fc, PS_ERR_ARG, "ps_fc_create: null argument");
  ps_updater_spec u;
  if (upd) u = *upd;
  else { u.kind = PS_UPD_ADAM; u.p[0] = (float)0.005; u.p[1] = (float)0.9; u.p[2] = (float)0.999; u.p[3] = (float)std::pow(10.0, -8); }
  std::unique_ptr<ps_fc> h(new ps_fc);
  h->op.create(&ctx->c, name, in, out, act, u, max_batch);
  *out_fc = h.release();
  PS_CATCH
}
int ps_fc_destroy(ps_fc* fc) {
  PS_TRY
  if (fc) { fc->op.destroy(); delete fc; }
  PS_CATCH
}
int ps_fc_forward(ps_fc* fc, const float* A_prev, int N, float* A) {
  PS_TRY
  PS_REQUIRE(fc != nullptr, PS_ERR_ARG, "ps_fc_forward: null layer");
  fc->op.forward(A_prev, N, A);
  PS_CATCH
}
int ps_fc_backward(ps_fc* fc, const float* delta, int N, float* delta_prev) {
  PS_TRY
  PS_REQUIRE(fc != nullptr, PS_ERR_ARG, "ps_fc_backward: null layer");
  fc->op.backward(delta, N, delta_prev);
  PS_CATCH
}
int ps_fc_gradients(ps_fc* fc, float* dW, float* db) {
  PS_TRY
  PS_REQUIRE(fc != nullptr, PS_ERR_ARG, "ps_fc_gradients: null layer");
  fc->op.gradients(dW, db);
  PS_CATCH
}
int ps_fc_update(ps_fc* fc) {
  PS_TRY
  PS_REQUIRE(fc != nullptr, PS_ERR_ARG, "ps_fc_update: null layer");
  fc->op.update();
  PS_CATCH
}
int ps_fc_get(ps_fc* fc, int which, float* out, int cap, int* n) {
  PS_TRY
  PS_REQUIRE(fc != nullptr && (which == 0 || which == 1), PS_ERR_ARG, "ps_fc_get: bad argument");
  std::vector<float> v;
  fc->op.get(which, v);
  if (n) *n = (int)v.size();
  if (out && cap >= (int)v.size()) std::memcpy(out, v.data(), sizeof(float) * v.size());
  PS_CATCH
}
int ps_fc_put(ps_fc* fc, int which, const float* in, int n) {
  PS_TRY
  PS_REQUIRE(fc != nullptr && in != nullptr && (which == 0 || which == 1), PS_ERR_ARG, "ps_fc_put: bad argument");
  fc->op.put(which, in, n);
  PS_CATCH
}

/* ---- libsvm ingest ---- */
int ps_libsvm_parse_line(const char* line, size_t len, int F, int Xn, int64_t wide_size, int64_t* E, float* X, int64_t* W, float* Y, int* status) {
  PS_TRY
  PS_REQUIRE(line != nullptr && status != nullptr && F >= 0 && Xn >= 0 && wide_size > 0, PS_ERR_ARG, "ps_libsvm_parse_line: bad argument");
  const char* e = line + len;
  if (e > line && e[-1] == '\n') --e;
  if (e > line && e[-1] == '\r') --e;
  *status = parse_ctr_line(line, e, F, Xn, wide_size, E, X, W, Y);
  PS_CATCH
}
int ps_libsvm_parse_dev(ps_ctx* ctx, const char* text_dev, size_t len, int F, int Xn, int64_t wide_size, int max_rows, int64_t* E_dev, float* X_dev,
                        int64_t* W_dev, float* Y_dev, uint8_t* status_dev, int* rows) {
  PS_TRY
  PS_REQUIRE(ctx != nullptr && rows != nullptr && max_rows > 0, PS_ERR_ARG, "ps_libsvm_parse_dev: bad argument");
  const size_t need = len / 4096 + 2 + (size_t)max_rows;
  if (need > ctx->ingest_ws_cap) {
    PS_CUDA(cudaStreamSynchronize(ctx->c.stream));
    dfree(ctx->ingest_ws);
    ctx->ingest_ws = nullptr; ctx->ingest_ws_cap = 0;
    ctx->ingest_ws = dmalloc<uint32_t>(need);
    ctx->ingest_ws_cap = need;
  }
  *rows = libsvm_parse_dev(&ctx->c, text_dev, len, F, Xn, wide_size, max_rows, E_dev, X_dev, W_dev, Y_dev, status_dev, ctx->ingest_ws);
  PS_CATCH
}
int ps_reader_open(const char* path, int F, int Xn, int64_t wide_size, int batch, int offset, int step, int threads, ps_reader** out) {
  PS_TRY
  PS_REQUIRE(path != nullptr && out != nullptr, PS_ERR_ARG, "ps_reader_open: null argument");
  std::unique_ptr<ps_reader> h(new ps_reader);
  h->r.reset(new LibsvmReader(path, F, Xn, wide_size, batch, offset, step, threads, 2));
  *out = h.release();
  PS_CATCH
}
int ps_reader_next(ps_reader* r, int64_t* E, float* X, int64_t* W, float* Y, int* rows) {
  PS_TRY
  PS_REQUIRE(r != nullptr && rows != nullptr, PS_ERR_ARG, "ps_reader_next: null argument");
  *rows = r->r->next(E, X, W, Y);
  PS_CATCH
}
int ps_reader_shape(ps_reader* r, int* batch, int* F, int* Xn) {
  PS_TRY
  PS_REQUIRE(r != nullptr, PS_ERR_ARG, "ps_reader_shape: null reader");
  if (batch) *batch = r->r->batch;
  if (F) *F = r->r->F;
  if (Xn) *Xn = r->r->Xn;
  PS_CATCH
}
int ps_reader_reset(ps_reader* r) {
  PS_TRY
  PS_REQUIRE(r != nullptr, PS_ERR_ARG, "ps_reader_reset: null reader");
  r->r->reset();
  PS_CATCH
}
int ps_reader_stats(ps_reader* r, int64_t* lines, int64_t* batches, int64_t* dropped_batches) {
  PS_TRY
  PS_REQUIRE(r != nullptr, PS_ERR_ARG, "ps_reader_stats: null reader");
  if (lines) *lines = r->r->lines_read.load();
  if (batches) *batches = r->r->batches.load();
  if (dropped_batches) *dropped_batches = r->r->dropped.load();
  PS_CATCH
}
int ps_reader_close(ps_reader* r) {
  PS_TRY
  delete r;
  PS_CATCH
}

int ps_test_gemm_nt(ps_ctx* ctx, int mode, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc) {
  PS_TRY
  PS_REQUIRE(ctx && A && B && C && M > 0 && N > 0 && K > 0, PS_ERR_ARG, "bad argument");
  cudaStream_t st = ctx->c.stream;
  float* dA = dmalloc<float>((size_t)M * lda); float* dB = dmalloc<float>((size_t)N * ldb); float* dC = dmalloc_zero<float>((size_t)M * ldc, st);
  float* dbias = dmalloc_zero<float>(N, st);
  PS_CUDA(cudaMemcpyAsync(dA, A, sizeof(float) * M * lda, cudaMemcpyHostToDevice, st));
  PS_CUDA(cudaMemcpyAsync(dB, B, sizeof(float) * N * ldb, cudaMemcpyHostToDevice, st));
  FcFwdArgs a{};
  a.B = M; a.in = K; a.out = N; a.A = dA; a.lda = lda; a.W = dB; a.ldw = ldb; a.bias = dbias; a.act = PS_ACT_NONE; a.Z = dC; a.ldz = ldc;
  const int saved = ctx->c.fc_precision;
  ctx->c.fc_precision = mode;
  try { if (mode == PS_FC_FP32) fc_forward_fp32(&ctx->c, a); else fc_forward_tf32(&ctx->c, a); } catch (...) { ctx->c.fc_precision = saved; throw; }
  ctx->c.fc_precision = saved;
  PS_CUDA(cudaMemcpyAsync(C, dC, sizeof(float) * M * ldc, cudaMemcpyDeviceToHost, st));
  PS_CUDA(cudaStreamSynchronize(st));
  dfree(dA); dfree(dB); dfree(dC); dfree(dbias);
  PS_CATCH
}

}  // extern "C"

/*
 * ps_oracle.cpp — CPU restatement of the wudikua/ps standalone training path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ps_b200/ (the product) may include, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference (pure Java + jblas 1.2.4 + grpc) cannot be built or run
 * in this image (no JDK, no jars) and its own tests assert nothing, so this
 * oracle is pinned only against (i) derived known answers — Java String.hashCode values,
 * the TestAuc vector, the updater-name grammar — and (ii) closed forms that follow from the
 * cited Java lines (g_eff = S(n+1)/(2n^2), Adam first-step identity).  Every function cites
 * the reference file:line it restates (paths relative to /root/reference/src/main/java/).
 *
 * Restated bug-for-bug (SURVEY.md §8a "Quirks"): string keys, EmbeddingLayer.backward being
 * called twice per step with the gradient object aliased inside KVStore.sum, constant-bias
 * Adam, mis-parenthesised FTRL, clipped sigmoid, LRLayer's ever-growing key set.
 * The ONLY deliberate difference: the reference's unseeded initialiser
 * (util/MatrixUtil.java:62-74) is replaced by the counter-based ps_init_value() of
 * include/ps_spec.h (same distribution) so the CUDA path can start from identical tables.
 *
 * jblas semantics honoured: column-major storage, in-place *i ops
 * returning the receiver, rowMeans = rowSums / columns, getRange half-open.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../include/ps_spec.h"

namespace {

/* ------------------------------------------------------------------ FloatMatrix */
struct FM {
  int rows = 0, cols = 0;
  std::vector<float> d;
  FM() {}
  FM(int r, int c) : rows(r), cols(c), d((size_t)r * c, 0.0f) {}
  int length() const { return rows * cols; }
  float& at(int i, int j) { return d[(size_t)i + (size_t)rows * j]; }
  float at(int i, int j) const { return d[(size_t)i + (size_t)rows * j]; }
};
using MP = std::shared_ptr<FM>;
static MP mk(int r, int c) { return std::make_shared<FM>(r, c); }
static MP dup(const MP& a) { return std::make_shared<FM>(*a); }

/* sgemm back-ends.  kind 0: plain ordered loops (deterministic, used by parity tests);
 * kind 1: OpenMP axpy-form loops; kind 2: OpenBLAS cblas_sgemm found by dlopen (what jblas'
 * NativeBlas.sgemm amounts to).  All compute C(m x n) = A(m x k) * B(k x n), column-major. */
typedef void (*cblas_sgemm_t)(int, int, int, int64_t, int64_t, int64_t, float, const float*, int64_t,
                              const float*, int64_t, float, float*, int64_t);
typedef void (*blas_set_threads_t)(int);
typedef int (*blas_get_threads_t)(void);
static cblas_sgemm_t g_cblas = nullptr;
static blas_set_threads_t g_blas_set_threads = nullptr;
static blas_get_threads_t g_blas_get_threads = nullptr;
static int g_gemm_kind = 0;

__attribute__((target_clones("avx512f", "avx2", "default")))
static void sgemm_axpy(int m, int n, int k, const float* A, const float* B, float* C) {
#pragma omp parallel for schedule(static)
  for (int j = 0; j < n; ++j) {
    float* c = C + (size_t)j * m;
    for (int i = 0; i < m; ++i) c[i] = 0.0f;
    for (int p = 0; p < k; ++p) {
      const float b = B[(size_t)p + (size_t)k * j];
      const float* a = A + (size_t)p * m;
      for (int i = 0; i < m; ++i) c[i] += a[i] * b;
    }
  }
}

static MP mmul(const MP& A, const MP& B) { /* FloatMatrix.mmul (FcLayer.java:76,105,108) */
  const int m = A->rows, k = A->cols, n = B->cols;
  MP C = mk(m, n);
  if (g_gemm_kind == 2 && g_cblas) {
    g_cblas(102 /*ColMajor*/, 111, 111, m, n, k, 1.0f, A->d.data(), m, B->d.data(), k, 0.0f, C->d.data(), m);
  } else if (g_gemm_kind == 1) {
    sgemm_axpy(m, n, k, A->d.data(), B->d.data(), C->d.data());
  } else {
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < m; ++i) {
        float s = 0.0f;
        for (int p = 0; p < k; ++p) s += A->d[(size_t)i + (size_t)m * p] * B->d[(size_t)p + (size_t)k * j];
        C->d[(size_t)i + (size_t)m * j] = s;
      }
  }
  return C;
}
static MP transpose(const MP& A) {
  MP T = mk(A->cols, A->rows);
  for (int j = 0; j < A->cols; ++j)
    for (int i = 0; i < A->rows; ++i) T->at(j, i) = A->at(i, j);
  return T;
}
static MP rowMeans(const MP& A) {
  MP r = mk(A->rows, 1);
  for (int j = 0; j < A->cols; ++j)
    for (int i = 0; i < A->rows; ++i) r->d[i] += A->at(i, j);
  for (int i = 0; i < A->rows; ++i) r->d[i] /= (float)A->cols;
  return r;
}
static void addiColumnVector(FM& A, const FM& v) {
  for (int j = 0; j < A.cols; ++j)
    for (int i = 0; i < A.rows; ++i) A.at(i, j) += v.d[i];
}
static void addi(FM& a, const FM& b) { /* a.addi(a) doubles a — aliasing is intended */
  const size_t n = a.d.size();
  for (size_t i = 0; i < n; ++i) a.d[i] += b.d[i];
}
static void divi(FM& a, float s) { for (auto& x : a.d) x /= s; }

/* ------------------------------------------------------------------ Java strings */
/* String.valueOf(double) / String.valueOf(float) for non-negative INTEGER values: below 1e7
 * "<int>.0", otherwise computerised scientific "d.dddE<n>" with the integer's digits minus
 * trailing zeros (every integer < 2^53 is a distinct double, so that digit string is the
 * shortest one that round-trips).  EmbeddingField.java:70 (float, forward) and :88
 * (double, backward) agree on this domain below 2^24.                  */
static std::string java_num(int64_t id) {
  char buf[48];
  if (id < 10000000LL) { snprintf(buf, sizeof buf, "%lld.0", (long long)id); return buf; }
  snprintf(buf, sizeof buf, "%lld", (long long)id);
  std::string s(buf);
  const int e = (int)s.size() - 1;
  while (s.size() > 2 && s.back() == '0') s.pop_back();
  std::string r = s.substr(0, 1) + "." + (s.size() > 1 ? s.substr(1) : std::string("0"));
  snprintf(buf, sizeof buf, "E%d", e);
  return r + buf;
}
static std::string emb_key(int field, int64_t id) { return "emF" + std::to_string(field) + "." + java_num(id); }
static std::string wide_key(int64_t id) { return "wide.weights." + java_num(id); }

static float xavier(int in, int out) { /* EmbeddingField.java:40, FcLayer.java:39,46 */
  return (float)(4 * (std::sqrt(6.0) / std::sqrt((double)(in + out))));
}

/* deterministic MatrixUtil.rand(row, col, max) (MatrixUtil.java:62-74), see header */
static MP init_rand(uint64_t seed, uint64_t key64, int rows, int cols, float maxv) {
  MP m = mk(rows, cols);
  for (int j = 0; j < rows * cols; ++j) m->d[j] = ps_init_value(seed, key64, (uint32_t)j, maxv);
  return m;
}

/* ------------------------------------------------------------------ update.* */
struct Updater {
  virtual ~Updater() {}
  virtual void update(const std::string& key, const MP& w, const MP& dw) = 0;
  virtual std::string name() const = 0;
};

static std::string jf(float v) { /* Float.toString for the few hyper-parameter values; informative only */
  char b[32]; snprintf(b, sizeof b, "%g", v); std::string s(b);
  if (s.find('.') == std::string::npos && s.find('e') == std::string::npos) s += ".0";
  return s;
}

struct AdamUpdater : Updater { /* update/AdamUpdater.java:57-84 */
  float alfa, beta1, beta2, epsilon;
  std::unordered_map<std::string, MP> M, V;
  AdamUpdater(double a, double b1, double b2, double e) : alfa((float)a), beta1((float)b1), beta2((float)b2), epsilon((float)e) {}
  void update(const std::string& key, const MP& w, const MP& dw) override {
    if (!M.count(key)) { MP t = mk(dw->rows, dw->cols); M[key] = t; V[key] = t; } /* :76-84 shared zero matrix */
    const float omb1 = 1 - beta1, omb2 = 1 - beta2;
    const int n = dw->length();
    MP m = mk(dw->rows, dw->cols), v = mk(dw->rows, dw->cols);
    FM& m0 = *M[key];
    for (int i = 0; i < n; ++i) m0.d[i] *= beta1;                       /* M.get(key).muli(beta1) */
    for (int i = 0; i < n; ++i) m->d[i] = dw->d[i] * omb1 + m0.d[i];    /* dw.mul(1-beta1).addi(..) :61 */
    M[key] = m;
    FM& v0 = *V[key];
    for (int i = 0; i < n; ++i) v0.d[i] *= beta2;
    for (int i = 0; i < n; ++i) { float t = dw->d[i] * dw->d[i]; t *= omb2; v->d[i] = t + v0.d[i]; } /* :62 */
    V[key] = v;
    for (int i = 0; i < n; ++i) {
      const float Mm = m->d[i] / omb1;                                  /* :63 */
      const float Vv = v->d[i] / omb2;                                  /* :64 */
      float den = (float)std::sqrt((double)Vv); den += epsilon;         /* :69 */
      float stp = Mm / den; stp *= (-1 * alfa);
      w->d[i] += stp;
    }
  }
  std::string name() const override { /* :72-74 */
    return "adam@alfa:" + jf(alfa) + "@beta1:" + jf(beta1) + "@beta2:" + jf(beta2) + "@epsilon:" + jf(epsilon) + "@";
  }
};

struct FtrlUpdater : Updater { /* update/FtrlUpdater.java:51-76 */
  float alfa, beta, l1, l2;
  std::unordered_map<std::string, MP> Z, N;
  FtrlUpdater(float a, float b, float l1_, float l2_) : alfa(a), beta(b), l1(l1_), l2(l2_) {}
  void update(const std::string& key, const MP& w, const MP& dw) override {
    if (dw->d[0] == 0) return;                                          /* :52 */
    if (!N.count(key)) N[key] = mk(w->length(), 1);
    if (!Z.count(key)) Z[key] = mk(w->length(), 1);
    FM& zi = *Z[key]; FM& ni = *N[key];
    const int n = w->length();
    for (int i = 0; i < n; ++i) {                                       /* :64-71 */
      if (std::fabs(zi.d[i]) <= l1) w->d[i] = 0;
      else {
        const float sign = zi.d[i] >= 0 ? 1.0f : -1.0f;
        w->d[i] = -(zi.d[i] - sign * l1) / ((l2 + (beta + (float)std::sqrt((double)ni.d[i]))) / alfa);
      }
    }
    for (int i = 0; i < n; ++i) {                                       /* :72-74 */
      const float g = dw->d[i];
      const float g2 = (float)((double)g * (double)g);                  /* MatrixFunctions.pow(dw,2) */
      float s = (float)std::sqrt((double)(ni.d[i] + g2));
      s -= (float)std::sqrt((double)(ni.d[i] / alfa));
      zi.d[i] += g - s * w->d[i];
      ni.d[i] += g2;
    }
  }
  std::string name() const override { /* :78-80 — says "adam@", sic */
    return "adam@alfa:" + jf(alfa) + "@beta:" + jf(beta) + "@l1:" + jf(l1) + "@l2:" + jf(l2) + "@";
  }
};

struct SimpleUpdater : Updater { /* update/SimpleUpdater.java:20-22 */
  float eta;
  explicit SimpleUpdater(float e) : eta(e) {}
  void update(const std::string&, const MP& w, const MP& dw) override {
    for (auto& x : dw->d) x *= -eta;
    addi(*w, *dw);
  }
  std::string name() const override { return "simple@eta:" + jf(eta) + "@"; }
};

/* ------------------------------------------------------------------ store.KVStore (standalone mode) */
struct KVStore {
  uint64_t seed = 0;
  std::unordered_map<std::string, MP> store, storeInit, sum;
  std::unordered_map<std::string, long> sumCnt;
  template <class Init> MP get(const std::string& key, Init init) { /* KVStore.java:136-159,168-190 */
    auto it = store.find(key);
    if (it != store.end()) return it->second;
    MP m = init();
    store[key] = m;
    storeInit[key] = dup(m);
    return m;
  }
  MP get(const std::string& key) { auto it = store.find(key); return it == store.end() ? nullptr : it->second; } /* :129-134 */
  void sumf(const std::string& key, const MP& val) { /* :192-200 — first caller's object kept BY REFERENCE */
    auto it = sum.find(key);
    if (it == sum.end()) { sum[key] = val; sumCnt[key] = 1; }
    else { addi(*it->second, *val); sumCnt[key]++; }
  }
  void update(const std::vector<std::pair<std::string, Updater*>>& updaters) { /* :240-268 */
    for (auto& kv : sum) {
      const std::string& key = kv.first;
      Updater* u = nullptr;
      for (auto& p : updaters) if (p.first == key) u = p.second;
      if (!u) for (auto& p : updaters) if (key.compare(0, p.first.size(), p.first) == 0) u = p.second; /* last startsWith hit */
      if (!u) for (auto& p : updaters) if (p.first == "default") u = p.second;
      divi(*kv.second, (float)sumCnt[key]);                             /* :253 */
      u->update(key, store[key], kv.second);
    }
  }
  void clear() { sum.clear(); sumCnt.clear(); }                         /* :270-277 (standalone) */
};

/* ------------------------------------------------------------------ activations / loss */
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_SOFTMAX = 3 };

static void relu_fwd(FM& x) { for (auto& v : x.d) v = std::max(0.0f, v); }                 /* Relu.java:7-12 */
static void relu_bwd(FM& dy, const FM& y) { for (size_t i = 0; i < y.d.size(); ++i) dy.d[i] *= y.d[i] > 0 ? 1 : 0; } /* :14-19 */
static void sigmoid_fwd(FM& x) {                                                           /* Sigmoid.java:9-14 */
  for (auto& v : x.d) v = (float)((double)0.001f + (double)(.999f - 0.001f) / ((double)1.0f + std::exp((double)(-v))));
}
static void sigmoid_bwd(FM& dy, const FM& y) { for (size_t i = 0; i < y.d.size(); ++i) dy.d[i] *= y.d[i] * (1 - y.d[i]); } /* :16-21 */
static void softmax_fwd(FM& x, int scale) {                                                /* Softmax.java:21-45 */
  for (auto& v : x.d) v /= (float)scale;
  for (int j = 0; j < x.cols; ++j) {
    float mx = x.at(0, j);
    for (int i = 1; i < x.rows; ++i) mx = std::max(mx, x.at(i, j));
    for (int i = 0; i < x.rows; ++i) x.at(i, j) = (float)std::exp((double)(x.at(i, j) - mx));
    float s = 0; for (int i = 0; i < x.rows; ++i) s += x.at(i, j);
    for (int i = 0; i < x.rows; ++i) {
      float& v = x.at(i, j); v = v / s;
      if (v == 0) v = 0.001f; else if (v == 1) v = 0.999f;
    }
  }
}
static MP softmax_bwd(const FM& dy, const FM& y) {                                         /* Softmax.java:47-67 */
  MP delta = mk(y.rows, y.cols);
  for (int i = 0; i < y.cols; ++i)
    for (int j = 0; j < y.rows; ++j) {
      if (dy.at(j, i) == 0) continue;
      const float d = dy.at(j, i);
      for (int k = 0; k < y.rows; ++k) {
        if (j == k) delta->at(k, i) += y.at(k, i) * (1 - y.at(k, i));
        else delta->at(k, i) += -y.at(j, i) * y.at(k, i);
        delta->at(k, i) *= d;
      }
    }
  return delta;
}
/* activation.backward(dy, Z, A): in place for Relu/Sigmoid, fresh matrix for Softmax */
static MP act_bwd(int act, const MP& dy, const MP& y) {
  if (act == ACT_RELU) { relu_bwd(*dy, *y); return dy; }
  if (act == ACT_SIGMOID) { sigmoid_bwd(*dy, *y); return dy; }
  if (act == ACT_SOFTMAX) return softmax_bwd(*dy, *y);
  return dy;
}
static void act_fwd(int act, FM& z) {
  if (act == ACT_RELU) relu_fwd(z); else if (act == ACT_SIGMOID) sigmoid_fwd(z); else if (act == ACT_SOFTMAX) softmax_fwd(z, 10000);
}

static float ce_forward(const FM& p, const FM& l) {                                        /* CrossEntropy.java:10-18 */
  float sum = 0;
  for (int i = 0; i < p.cols; ++i) {
    const float pi = p.at(0, i), li = l.at(0, i);
    sum += (float)((double)(-li) * std::log((double)pi) - ((double)(1 - li) * std::log((double)(1 - pi))));
  }
  return sum / p.cols;
}
static MP ce_backward(const FM& p, const FM& l) {                                          /* :20-28 */
  MP d = std::make_shared<FM>(p);
  for (int i = 0; i < p.cols; ++i) { const float pi = p.at(0, i), li = l.at(0, i); d->at(0, i) = (pi - li) / (pi * (1 - pi)); }
  return d;
}
static float sml_forward(const FM& p, const FM& l) {                                       /* SoftmaxLoss.java:9-17 */
  float sum = 0;
  /* `sum += -FastMath.log(p)` is a float/double compound assignment: the add happens in double, then narrows */
  for (int i = 0; i < p.cols; ++i) { const int hot = (int)l.at(0, i); sum = (float)((double)sum + (-std::log((double)p.at(hot, i)))); }
  return sum / p.cols;
}
static MP sml_backward(const FM& p, const FM& l) {                                         /* :20-28 */
  MP d = mk(p.rows, p.cols);
  for (int i = 0; i < p.cols; ++i) { const int hot = (int)l.at(0, i); d->at(hot, i) = -1 / p.at(hot, i); }
  return d;
}

/* ------------------------------------------------------------------ layer.* */
struct Layer {
  std::string name; int inputDims = 0, outputDims = 0;
  MP A, delta; Layer* next = nullptr; Layer* pre = nullptr; KVStore* kv = nullptr;
  virtual ~Layer() {}
  virtual void forward() = 0; virtual void backward() = 0; virtual void pullWeights() = 0;
  void setNext(Layer* l) { next = l; l->pre = this; }                   /* Layer.java:52-56 */
};
struct InputLayer : Layer {                                             /* InputLayer.java */
  void forward() override {} void backward() override {} void pullWeights() override {}
};

struct EmbeddingField {                                                 /* layer/EmbeddingField.java */
  int field; int D; KVStore* kv; float xav;
  std::unordered_map<std::string, MP> weights, wg; std::unordered_map<std::string, int> wgN;
  std::vector<std::string> order;                                       /* keySet() order is irrelevant numerically */
  std::vector<int64_t> nSample; MP Z, A;
  EmbeddingField(int f, int d, KVStore* k) : field(f), D(d), kv(k), xav(xavier(1, d)) {}
  MP forward(const std::vector<int64_t>& ids) {                         /* :66-78 */
    nSample = ids;
    const int n = (int)ids.size();
    MP WX = mk(D, n);
    for (int i = 0; i < n; ++i) {
      const std::string key = emb_key(field, ids[i]);
      auto it = weights.find(key);
      if (it == weights.end()) {                                        /* checkExists :49-54 */
        const int f = field; const int64_t id = ids[i]; const int dd = D; const float xv = xav; const uint64_t sd = kv->seed;
        MP w = kv->get(key, [=] { return init_rand(sd, ps_pack_key((uint32_t)f, (uint64_t)id), dd, 1, xv); });
        it = weights.emplace(key, w).first;
      }
      std::memcpy(&WX->d[(size_t)i * D], it->second->d.data(), sizeof(float) * D);  /* JavaBlas.rcopy :73 */
    }
    Z = WX; relu_fwd(*Z); A = Z;                                        /* :75-76, Z aliases A */
    return A;
  }
  void clear() { weights.clear(); wg.clear(); wgN.clear(); order.clear(); }      /* :80-84 */
  void backward(int offset, const FM& delta) {                          /* :86-104 */
    const int n = (int)nSample.size();
    for (int k = 0; k < n; ++k) {
      const std::string key = emb_key(field, nSample[k]);
      MP g = mk(D, 1);
      for (int i = 0; i < D; ++i) g->d[i] = delta.at(offset + i, k);    /* getRange */
      for (int i = 0; i < D; ++i) g->d[i] *= A->at(i, k) > 0 ? 1 : 0;   /* Relu.backward with y = A.getColumn(k) */
      auto it = wg.find(key);
      if (it == wg.end()) { wg[key] = g; wgN[key] = 1; order.push_back(key); }
      else { addi(*it->second, *g); wgN[key] += 1; }
    }
    for (const std::string& key : order) {                              /* :99-102 */
      MP G = wg[key];
      divi(*G, (float)wgN[key]);
      kv->sumf(key, G);
    }
  }
};

struct EmbeddingLayer : Layer {                                         /* layer/EmbeddingLayer.java */
  std::vector<EmbeddingField> fields; int D = 0;
  void build(int F, int d) { D = d; for (int j = 0; j < F; ++j) fields.emplace_back(j, d, kv); }  /* :50-57 */
  std::vector<int64_t> ids; int N = 0;                                  /* F x N column-major int64 ids (the "E" input) */
  void forward() override { forward_ids(); }
  void forward_ids() {                                                  /* :25-48 */
    const int F = (int)fields.size();
    MP out = mk(F * D, N);
    std::vector<int64_t> row(N);
    for (int i = 0; i < F; ++i) {
      for (int n = 0; n < N; ++n) row[n] = ids[(size_t)i + (size_t)F * n];      /* E.getRow(i).toArray() */
      MP emb = fields[i].forward(row);
      for (int r = 0; r < D; ++r)                                       /* MatrixUtil.appendRows :76-82 + new FloatMatrix(EX) */
        for (int n = 0; n < N; ++n) out->at(i * D + r, n) = emb->at(r, n);
    }
    A = out;
  }
  void backward() override {                                            /* :59-69 */
    delta = next->delta;
    int offset = 0;
    for (auto& f : fields) { f.backward(offset, *delta); offset += f.D; }
  }
  void pullWeights() override { for (auto& f : fields) f.clear(); }     /* :71-75 */
};

struct ConcatLayer : Layer {                                            /* layer/ConcatLayer.java */
  std::vector<Layer*> inputs;
  void forward() override {                                             /* :30-37 */
    A = inputs[0]->A;
    for (size_t i = 1; i < inputs.size(); ++i) {
      const FM& a = *A; const FM& b = *inputs[i]->A;
      MP c = mk(a.rows + b.rows, a.cols);
      for (int j = 0; j < a.cols; ++j) {
        std::memcpy(&c->at(0, j), &a.d[(size_t)a.rows * j], sizeof(float) * a.rows);
        std::memcpy(&c->at(a.rows, j), &b.d[(size_t)b.rows * j], sizeof(float) * b.rows);
      }
      A = c;
    }
  }
  void backward() override { delta = next->delta; for (Layer* l : inputs) l->backward(); }   /* :39-48 — re-invokes inputs */
  void pullWeights() override {}
};

struct FcLayer : Layer {                                                /* layer/FcLayer.java */
  MP weights, bias, Z; int act = ACT_NONE;
  void forward() override {                                             /* :74-91 */
    MP WX = mmul(weights, pre->A);
    addiColumnVector(*WX, *bias);
    Z = WX; act_fwd(act, *Z); A = Z;
  }
  void backward() override {                                            /* :93-110 */
    MP d = next == nullptr ? delta : next->delta;
    if (act != ACT_NONE) d = act_bwd(act, d, A);
    MP db = rowMeans(d);
    kv->sumf(name + ".bias", db);
    MP dW = mmul(d, transpose(pre->A));
    divi(*dW, (float)d->cols);
    kv->sumf(name + ".weights", dW);
    delta = mmul(transpose(weights), d);
  }
  void pullWeights() override {                                         /* :112-115 */
    const uint64_t sd = kv->seed; const int in = inputDims, out = outputDims;
    const std::string wk = name + ".weights", bk = name + ".bias";
    weights = kv->get(wk, [=] { return init_rand(sd, ps_name_key(wk.c_str()), out, in, xavier(in, out)); });
    bias = kv->get(bk, [=] { return init_rand(sd, ps_name_key(bk.c_str()), out, 1, xavier(in, 1)); });
  }
};

struct LRLayer : Layer {                                                /* layer/LRLayer.java */
  std::map<std::string, MP> weights; MP bias, Z; int act = ACT_NONE;    /* `weights` is never cleared (quirk 7) */
  std::vector<int64_t> ids; int F = 0, N = 0;                           /* the "W" input, F x N column-major */
  void init() { bias = kv->get(name + ".bias", [] { return mk(1, 1); }); }       /* ctor :52 */
  void forward() override {                                             /* :62-98 */
    MP WX = mk(1, N);
    for (int i = 0; i < N; ++i) {
      float sumW = 0.0f;
      for (int j = 0; j < F; ++j) {
        const std::string key = wide_key(ids[(size_t)j + (size_t)F * i]);
        MP wi = kv->get(key, [] { return mk(1, 1); });
        weights[key] = wi;
        sumW += wi->d[0];
      }
      WX->d[i] = sumW;
    }
    addiColumnVector(*WX, *bias);
    Z = WX; act_fwd(act, *Z); A = Z;
  }
  void backward() override {                                            /* :100-120 */
    MP d = next == nullptr ? delta : next->delta;
    if (act != ACT_NONE) d = act_bwd(act, d, A);
    d = rowMeans(d);
    kv->sumf(name + ".bias", d);
    for (auto& kvp : weights) kv->sumf(kvp.first, d);                   /* the SAME 1x1 object for every key ever seen */
  }
  void pullWeights() override { bias = kv->get(name + ".bias", [] { return mk(1, 1); }); }
};

struct AddLayer : Layer {                                               /* layer/AddLayer.java */
  Layer *left = nullptr, *right = nullptr; MP Z; int act = ACT_NONE;
  void forward() override {                                             /* :33-47 */
    MP z = dup(left->A); addi(*z, *right->A);
    Z = z; act_fwd(act, *Z); A = Z;
  }
  void backward() override {                                            /* :49-61 */
    MP d = next == nullptr ? delta : next->delta;
    if (act != ACT_NONE) d = act_bwd(act, d, A);
    delta = d;
  }
  void pullWeights() override {}
};

/* ------------------------------------------------------------------ model.* + train.Trainer (thread = 1) */
enum Kind { KIND_DNN = 0, KIND_WIDEDEEP = 1, KIND_FCNN = 2 };

struct Model {
  int kind; KVStore kv;
  std::vector<std::unique_ptr<Layer>> owned; std::vector<Layer*> layers;
  InputLayer *category = nullptr, *number = nullptr, *wideIn = nullptr;
  EmbeddingLayer* emb = nullptr; LRLayer* wide = nullptr;
  std::vector<std::pair<std::string, Updater*>> updaters; std::vector<std::unique_ptr<Updater>> ownedUpd;
  int F = 0, D = 0, Xn = 0; bool softmaxLoss = false; bool skipped_backward = false;

  template <class T> T* add(const std::string& name, int in, int out) {
    T* l = new T(); l->name = name; l->inputDims = in; l->outputDims = out; l->kv = &kv; owned.emplace_back(l); return l;
  }
  std::vector<FcLayer*> buildFc(int inputSize, const std::vector<int>& dims) { /* FcLayer.build :53-70 */
    std::vector<FcLayer*> r;
    for (size_t i = 0; i < dims.size(); ++i) {
      FcLayer* fc = add<FcLayer>("fc" + std::to_string(i), inputSize, dims[i]);
      fc->act = (i + 1 == dims.size()) ? ACT_SIGMOID : ACT_RELU;
      if (i) r[i - 1]->setNext(fc);
      r.push_back(fc); inputSize = dims[i];
    }
    return r;
  }
  Model(int kind_, int F_, int D_, int Xn_, const std::vector<int>& fc, uint64_t seed, int emb_opt) : kind(kind_), F(F_), D(D_), Xn(Xn_) {
    kv.seed = seed;
    Updater* adam = new AdamUpdater(0.005, 0.9, 0.999, std::pow(10.0, -8)); ownedUpd.emplace_back(adam);
    if (kind == KIND_FCNN) {                                            /* FullConnectedNN.buildModel :86-110 */
      number = add<InputLayer>("number", 0, Xn);
      auto fcs = buildFc(Xn, fc);
      fcs.back()->act = ACT_SOFTMAX; softmaxLoss = true;
      number->setNext(fcs[0]);
      for (auto* l : fcs) layers.push_back(l);
      updaters.emplace_back("default", adam);
      return;
    }
    /* DNN.buildModel :92-128 / WideDeepNN.buildModel :105-161 */
    category = add<InputLayer>("category", 0, F * D);
    number = add<InputLayer>("number", 0, Xn);
    emb = add<EmbeddingLayer>("embedding", F, F * D); emb->build(F, D);
    ConcatLayer* concat = add<ConcatLayer>("concat", F * D + Xn, F * D + Xn);
    concat->inputs = {emb, number};
    auto fcs = buildFc(F * D + Xn, fc);
    category->setNext(emb); number->setNext(concat); emb->setNext(concat); concat->setNext(fcs[0]);
    layers.push_back(emb); layers.push_back(concat);
    for (auto* l : fcs) layers.push_back(l);
    if (kind == KIND_WIDEDEEP) {
      Updater* ftrl = new FtrlUpdater(0.005f, 1.0f, 0.001f, 0.001f); ownedUpd.emplace_back(ftrl);
      updaters.emplace_back("wide.weights", ftrl); updaters.emplace_back("wide.bias", ftrl);
      fcs.back()->act = ACT_NONE;                                       /* :128 */
      wideIn = add<InputLayer>("wideCategory", 0, 100000);
      wide = add<LRLayer>("wide", 100000, 1); wide->init(); wide->act = ACT_NONE;
      AddLayer* addl = add<AddLayer>("addWideDeep", 0, 0); addl->left = fcs.back(); addl->right = wide; addl->act = ACT_SIGMOID;
      fcs.back()->setNext(addl); wideIn->setNext(wide); wide->setNext(addl);
      layers.push_back(wide); layers.push_back(addl);
    }
    if (emb_opt == 1) {                                                 /* put("emF", ftrl) through the prefix rule */
      Updater* eftrl = new FtrlUpdater(0.005f, 1.0f, 0.001f, 0.001f); ownedUpd.emplace_back(eftrl);
      updaters.emplace_back("emF", eftrl);
    }
    updaters.emplace_back("default", adam);
  }
  void setInputs(const int64_t* E, const float* X, const int64_t* W, int N) {
    if (emb) { emb->ids.assign(E, E + (size_t)F * N); emb->N = N; }
    if (number) { MP x = mk(Xn, N); std::memcpy(x->d.data(), X, sizeof(float) * (size_t)Xn * N); number->A = x; }
    if (wide) { wide->ids.assign(W, W + (size_t)F * N); wide->F = F; wide->N = N; }
  }
  void forwardAll() { for (Layer* l : layers) l->forward(); }
  float train(const int64_t* E, const float* X, const int64_t* W, const float* Y, int N) { /* DNN.java:35-70 etc. */
    setInputs(E, X, W, N);
    forwardAll();
    MP y = mk(1, N); std::memcpy(y->d.data(), Y, sizeof(float) * N);
    const FM& P = *layers.back()->A;
    const float loss = softmaxLoss ? sml_forward(P, *y) : ce_forward(P, *y);
    MP delta = softmaxLoss ? sml_backward(P, *y) : ce_backward(P, *y);
    skipped_backward = false;
    if (loss <= (float)std::pow(10.0, -2) || std::isnan(loss)) { skipped_backward = true; return loss; }  /* :58-63 */
    layers.back()->delta = delta;
    for (int i = (int)layers.size() - 1; i >= 0; --i) layers[i]->backward();
    return loss;
  }
  float trainerStep(const int64_t* E, const float* X, const int64_t* W, const float* Y, int N) {
    for (Layer* l : layers) l->pullWeights();                           /* TrainerThread.java:33 */
    const float loss = train(E, X, W, Y, N);                            /* :34 */
    kv.update(updaters);                                                /* Trainer.java:93 */
    kv.clear();                                                         /* :95 */
    return loss;
  }
  void predict(const int64_t* E, const float* X, const int64_t* W, int N, float* out) { /* PredictThread + Trainer.predict */
    for (Layer* l : layers) l->pullWeights();
    setInputs(E, X, W, N);
    forwardAll();
    const FM& P = *layers.back()->A;
    std::memcpy(out, P.d.data(), sizeof(float) * P.d.size());
    kv.clear();
  }
  Layer* find(const std::string& n) { for (auto& l : owned) if (l->name == n) return l.get(); return nullptr; }
};

}  // namespace

/* ====================================================================== C API (ctypes) */
extern "C" {

int32_t pso_java_hash(const char* s) { return ps_java_string_hash(s); }           /* net/Mod.java:14 */
int32_t pso_router_mod(const char* s, int32_t n) { return ps_router_mod_java(s, n); }
int32_t pso_router_floormod(const char* s, int32_t n) { return ps_router_floormod_java(s, n); }
int pso_key_string(int kind, int field, int64_t id, char* buf, int cap) {
  const std::string s = kind == 0 ? emb_key(field, id) : wide_key(id);
  if ((int)s.size() + 1 > cap) return -1;
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return (int)s.size();
}
uint64_t pso_pack_key(uint32_t ns, uint64_t id) { return ps_pack_key(ns, id); }
uint64_t pso_name_key(const char* s) { return ps_name_key(s); }
float pso_init_value(uint64_t seed, uint64_t key, uint32_t j, float maxv) { return ps_init_value(seed, key, j, maxv); }
uint32_t pso_owner_of(uint64_t key, uint32_t n) { return ps_owner_of(key, n); }
float pso_xavier(int in, int out) { return xavier(in, out); }

double pso_auc(const float* p, const float* y, int n) {                           /* evaluate/AUC.java:32-82 */
  std::vector<std::pair<double, double>> v(n);
  for (int i = 0; i < n; ++i) v[i] = {(double)p[i], (double)y[i]};
  std::stable_sort(v.begin(), v.end(), [](const std::pair<double, double>& a, const std::pair<double, double>& b) { return a.first < b.first; });
  double pos = 0, neg = 0;
  for (auto& q : v) { if (q.second > 0.0) pos += 1; else neg += 1; }
  double tp = 0, fp = 0, prev = 0, auc = 0;
  for (int i = n - 1; i >= 0; --i) {
    if (v[i].second > 0.0) fp += 1; else tp += 1;
    const double x = tp / pos, yy = fp / neg;
    if (x != prev) { auc += (x - prev) * yy; prev = x; }
  }
  return auc;
}

/* stateless forms of the updaters for property tests: state arrays are caller-held */
void pso_adam_update(float* w, float* m, float* v, const float* g, int n, float alfa, float b1, float b2, float eps) {
  AdamUpdater u(0, 0, 0, 0); u.alfa = alfa; u.beta1 = b1; u.beta2 = b2; u.epsilon = eps;
  MP W = mk(n, 1), G = mk(n, 1), M0 = mk(n, 1), V0 = mk(n, 1);
  std::memcpy(W->d.data(), w, 4 * n); std::memcpy(G->d.data(), g, 4 * n); std::memcpy(M0->d.data(), m, 4 * n); std::memcpy(V0->d.data(), v, 4 * n);
  u.M["k"] = M0; u.V["k"] = V0; u.update("k", W, G);
  std::memcpy(w, W->d.data(), 4 * n); std::memcpy(m, u.M["k"]->d.data(), 4 * n); std::memcpy(v, u.V["k"]->d.data(), 4 * n);
}
void pso_ftrl_update(float* w, float* z, float* nn, const float* g, int n, float alfa, float beta, float l1, float l2) {
  FtrlUpdater u(alfa, beta, l1, l2);
  MP W = mk(n, 1), G = mk(n, 1), Z0 = mk(n, 1), N0 = mk(n, 1);
  std::memcpy(W->d.data(), w, 4 * n); std::memcpy(G->d.data(), g, 4 * n); std::memcpy(Z0->d.data(), z, 4 * n); std::memcpy(N0->d.data(), nn, 4 * n);
  u.Z["k"] = Z0; u.N["k"] = N0; u.update("k", W, G);
  std::memcpy(w, W->d.data(), 4 * n); std::memcpy(z, Z0->d.data(), 4 * n); std::memcpy(nn, N0->d.data(), 4 * n);
}
int pso_updater_name(int kind, float a, float b, float c, float d, char* buf, int cap) { /* T/TestPs.java:26-27 grammar */
  std::string s;
  if (kind == 0) { AdamUpdater u(a, b, c, d); s = u.name(); } else if (kind == 1) { FtrlUpdater u(a, b, c, d); s = u.name(); } else { SimpleUpdater u(a); s = u.name(); }
  if ((int)s.size() + 1 > cap) return -1;
  std::memcpy(buf, s.c_str(), s.size() + 1); return (int)s.size();
}

/* sgemm back-end: 0 ordered loops, 1 OpenMP loops, 2 OpenBLAS (path to the .so, ILP64 scipy build) */
int pso_set_gemm(int kind, const char* openblas_path) {
  if (kind == 2) {
    if (!g_cblas) {
      void* h = dlopen(openblas_path, RTLD_NOW | RTLD_LOCAL);
      if (!h) return -1;
      g_cblas = (cblas_sgemm_t)dlsym(h, "scipy_cblas_sgemm64_");
      if (!g_cblas) return -2;
      g_blas_set_threads = (blas_set_threads_t)dlsym(h, "scipy_openblas_set_num_threads64_");
      g_blas_get_threads = (blas_get_threads_t)dlsym(h, "scipy_openblas_get_num_threads64_");
    }
  }
  g_gemm_kind = kind; return 0;
}
/* Threads the dlopen'ed OpenBLAS uses per sgemm call.  numpy initialises the same library with one thread per core long
 * before bench.py can set OPENBLAS_NUM_THREADS, so the environment variable alone does NOT pin it: R replicas x C BLAS
 * threads oversubscribed the host in round 1.  Returns the value in force afterwards (-1: no OpenBLAS loaded). */
int pso_set_blas_threads(int n) {
  if (!g_blas_set_threads) return -1;
  g_blas_set_threads(n);
  return g_blas_get_threads ? g_blas_get_threads() : n;
}

void* pso_model_create(int kind, int F, int D, int Xn, const int* fc, int n_fc, uint64_t seed, int emb_opt) {
  return new Model(kind, F, D, Xn, std::vector<int>(fc, fc + n_fc), seed, emb_opt);
}
void pso_model_destroy(void* m) { delete (Model*)m; }
float pso_train_step(void* m, const int64_t* E, const float* X, const int64_t* W, const float* Y, int N) {
  return ((Model*)m)->trainerStep(E, X, W, Y, N);
}
void pso_predict(void* m, const int64_t* E, const float* X, const int64_t* W, int N, float* out) { ((Model*)m)->predict(E, X, W, N, out); }
int pso_skipped_backward(void* m) { return ((Model*)m)->skipped_backward ? 1 : 0; }
int64_t pso_num_keys(void* m) { return (int64_t)((Model*)m)->kv.store.size(); }

static int copy_out(const MP& p, float* out, int cap) {
  if (!p) return -1;
  const int n = p->length();
  if (out && cap >= n) std::memcpy(out, p->d.data(), sizeof(float) * n);
  return n;
}
/* KVStore.get(String) (KVStore.java:129-134): host copy of the live weight, -1 when absent */
int pso_get(void* m, const char* key, float* out, int cap) { return copy_out(((Model*)m)->kv.get(key), out, cap); }
int pso_get_init(void* m, const char* key, float* out, int cap) {
  auto& s = ((Model*)m)->kv.storeInit; auto it = s.find(key); return it == s.end() ? -1 : copy_out(it->second, out, cap);
}
/* updater state: which 0 = Adam M / Ftrl Z, 1 = Adam V / Ftrl N; looked up in the updater that owns `key` */
int pso_get_state(void* mm, const char* key, int which, float* out, int cap) {
  Model* m = (Model*)mm;
  for (auto& u : m->ownedUpd) {
    if (auto* a = dynamic_cast<AdamUpdater*>(u.get())) { auto& mp = which ? a->V : a->M; auto it = mp.find(key); if (it != mp.end()) return copy_out(it->second, out, cap); }
    if (auto* f = dynamic_cast<FtrlUpdater*>(u.get())) { auto& mp = which ? f->N : f->Z; auto it = mp.find(key); if (it != mp.end()) return copy_out(it->second, out, cap); }
  }
  return -1;
}
/* layer taps after the last train/predict: what 0 = A, 1 = delta */
int pso_layer_tap(void* mm, const char* layer, int what, float* out, int cap) {
  Layer* l = ((Model*)mm)->find(layer);
  if (!l) return -1;
  return copy_out(what ? l->delta : l->A, out, cap);
}

/* EmbeddingLayer in isolation (A1-A6 without the dense stack): forward, the two backward calls, update, clear */
void* pso_emb_create(int F, int D, uint64_t seed, int opt) { return new Model(KIND_DNN, F, D, 1, std::vector<int>{1}, seed, opt); }
void pso_emb_forward(void* mm, const int64_t* E, int N, float* out) {
  Model* m = (Model*)mm;
  m->emb->pullWeights(); m->emb->ids.assign(E, E + (size_t)m->F * N); m->emb->N = N; m->emb->forward_ids();
  std::memcpy(out, m->emb->A->d.data(), sizeof(float) * m->emb->A->d.size());
}
void pso_emb_backward_update(void* mm, const float* delta, int ld, int N, int calls) {
  Model* m = (Model*)mm;
  MP d = mk(ld, N); std::memcpy(d->d.data(), delta, sizeof(float) * (size_t)ld * N);
  struct Dummy : Layer { void forward() override {} void backward() override {} void pullWeights() override {} } nx;
  nx.delta = d; Layer* saved = m->emb->next; m->emb->next = &nx;
  for (int c = 0; c < calls; ++c) m->emb->backward();                   /* the reference makes 2 calls per step */
  m->emb->next = saved;
  m->kv.update(m->updaters); m->kv.clear();
}

}  // extern "C"

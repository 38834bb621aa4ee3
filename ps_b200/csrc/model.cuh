/*
 * model.cuh — host-side mirror of the reference's class surface for the hot path, in C++
 * because the reference is compiled (Java) code and no JDK exists in this image
 * (SURVEY.md §8b).  Names follow the reference: layer.FcLayer, layer.EmbeddingLayer,
 * layer.LRLayer, layer.AddLayer, model.DNN / WideDeepNN / FullConnectedNN, train.Trainer.
 * Every object keeps its parameters in the GPU-resident store (store.KVStore's role); the
 * Java classes of integration/java/ forward to these through the C ABI.
 */
#pragma once
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "dense.cuh"
#include "gemm.cuh"
#include "ingest.cuh"
#include "p2p.cuh"
#include "shard.cuh"
#include "table.cuh"

namespace psb {

/* layer/FcLayer.java — weights "fc<i>.weights" (out x in), bias "fc<i>.bias" */
struct FcLayer {
  std::string name;
  int in = 0, out = 0, act = PS_ACT_NONE;
  int ldw = 0, ldwt = 0;
  float *W = nullptr, *Wt = nullptr, *bias = nullptr;
  float *Wlo = nullptr, *Wtlo = nullptr;   /* residuals W - tf32(W), Wt - tf32(Wt) for the 3xTF32 GEMMs (TMA operands) */
  void refresh_lo(Ctx* ctx);               /* after W / Wt were written from the host */
  float *sW1 = nullptr, *sW2 = nullptr, *sb1 = nullptr, *sb2 = nullptr;
  float* G = nullptr;            /* wgrad partial slabs [nsplit][out][ldw] */
  int nsplit = 1;
  ps_updater_spec updW, updB;
  void create(Ctx* ctx, const std::string& nm, int in_, int out_, int act_, const ps_updater_spec& u, int nsplit_);
  void destroy();
};

struct HostBatch {               /* host pointers of one submitted step (CTR.parseFeature's map, CTR.java:47-68) */
  const int64_t* E = nullptr; const float* X = nullptr; const int64_t* W = nullptr; const float* Y = nullptr; int N = 0;
};

struct Model {
  Ctx* ctx = nullptr;
  int kind = 0, F = 0, D = 0, Xn = 0, Bmax = 0, L = 0;
  bool has_emb = false, has_wide = false;
  EmbTable emb;                  /* layer.EmbeddingLayer "embedding" (EmbeddingLayer.java) */
  WideTable wide;                /* layer.LRLayer "wide" weights (LRLayer.java) */
  float* wide_bias = nullptr;    /* {w, s1, s2} of "wide.bias" */
  ps_updater_spec upd_default, upd_wide;
  std::vector<FcLayer> fcs;
  std::vector<float*> act;       /* act[0] = concat output, act[l+1] = fc<l>.A ; [Bmax][ld[l]] */
  std::vector<float*> delta;     /* delta[l] = fc<l>.delta (input side, shape of act[l]); delta[L] = top delta */
  std::vector<float*> act_t;     /* act_t[l] = [act[l] | 1]^T, [width[l]+1][ldt]: K-major operand of the TF32 wgrad */
  std::vector<float*> delta_t;   /* delta_t[l] = delta[l]^T, [width[l]][ldt] */
  int ldt = 0;
  std::vector<int> width, ld;
  float *wide_z = nullptr, *P = nullptr, *tail_ws = nullptr;
  StepStatus* st_dev = nullptr;
  /* side streams: independent kernels of a step (wide branch, weight gradients, dense update) run
   * beside the critical chain probe → gather → fc forward → tail → dgrad chain → scatter/update;
   * inside a capture they become parallel branches of the step graph */
  cudaStream_t aux[2] = {nullptr, nullptr};
  cudaEvent_t sync_ev[24] = {};
  int sync_n = 0;
  void fork(cudaStream_t from, cudaStream_t to);
  /* one CUDA graph per (input buffers, batch size, mode): a step is ~25 small launches, replayed as one */
  bool use_graph = true;
  int emb_generation = 0;        /* EmbTable::generation the cached graphs were captured with */
  struct GraphEntry { cudaGraphExec_t exec = nullptr; long kernels = 0; std::vector<std::string> names; };
  std::map<std::tuple<const void*, const void*, const void*, const void*, int, int, const void*, int>, GraphEntry> graphs;
  /* kStages staging sets: the H2D of step i+1 overlaps the kernels of step i, and the host may run up to kStages steps ahead of the
   * device — with several ranks stepping in lockstep, one rank's late host thread otherwise stalls every GPU (8 GPUs: e2e 20 % under the
   * device-resident rate with two sets) */
  static constexpr int kStages = 4;
  struct Stage {
    int64_t *E = nullptr, *W = nullptr; float *X = nullptr, *Y = nullptr;
    char* text = nullptr; size_t text_cap = 0; uint32_t* text_ws = nullptr; uint8_t* text_status = nullptr;   /* submit_text: raw libsvm text, parsed on the device */
    StepStatus* st_host = nullptr;       /* mapped pinned */
    cudaEvent_t h2d_done = nullptr, step_done = nullptr;
    int N = 0; bool busy = false;
  } stage[kStages];
  int next_stage = 0, oldest_stage = 0, in_flight = 0;
  uint32_t seq = 0;
  int last_N = 0; bool last_train = false;
  StepStatus last_status{};
  /* per-phase device timing */
  /* per-phase device timing: event-record nodes INSIDE the step's CUDA graph (cudaEventRecordExternal),
   * so the times are those of the graph-replayed kernels, not of host launch cadence */
  bool profile = false, capturing = false;
  std::vector<cudaEvent_t> ev_pool;
  int ev_n = 0;
  std::vector<std::string> ev_names, phase_names;
  std::vector<float> phase_ms;

  void create(Ctx* c, int kind_, int F_, int D_, int Xn_, const int32_t* fc_dims, int n_fc, int64_t emb_capacity,
              const ps_updater_spec* emb_updater, int max_batch);
  void destroy();
  /* the whole step on device-resident inputs, enqueued on ctx->stream; status lands in st_dev */
  /* per-layer pieces shared by the single-GPU step, the sharded step and the timing hooks */
  bool top_is_unit() const { return kind != PS_MODEL_FCNN && fcs[L - 1].out == 1; }   /* the fused GEMV + tail path applies */
  void fwd_layer(int l, int N);
  void run_tail(const float* Y, int N, bool train);
  void wgrad_layer(int l, int N);
  void dgrad_layer(int l, int N);
  /* everything between "act[0] is filled" and "delta[0] is ready": wide branch (side stream 1), FcLayer
   * forward, tail, dgrad chain on the main stream with each wgrad beside it on side stream 1.  On return
   * the main stream holds delta[0]; side stream 1 holds the wgrads; side stream 2 the transposes / wide update. */
  void forward_backward(const int64_t* W, const int64_t* W_all, int n_all, const float* Y, int N, bool train, bool wide_update_now);
  void forward_layers(int N, bool train);
  void backward_layers(int N, bool wide_update_now);
  /* DNN.train call by call: forward loop | (loss in the caller) | reverse loop + KVStore.update (see model.cu) */
  bool p2p_scalars_now = false;        /* forward_backward exchanges the global scalars right after the tail (set by p2p_step) */
  int pending_forward_N = 0;
  float* dtop_stage = nullptr;         /* FullConnectedNN: the caller's C x N delta on the device */
  bool pad_dirty = false;
  void forward_host(const HostBatch& b, float* P_out);
  void backward_update_host(const float* delta_top, int N, float loss);
  void submit_text(const char* text, size_t len, int N, int mode = 0);
  void step_device(const int64_t* E, const float* X, const int64_t* W, const float* Y, int N, bool train, StepStatus* publish_to);
  /* the same through the graph cache (falls back to direct launches while profiling) */
  void run_step(const int64_t* E, const float* X, const int64_t* W, const float* Y, int N, bool train, StepStatus* publish_to, int mode = 0);
  /* ---- key-hash sharded (multi-GPU) step, split at the exchanges ---- */
  float* gsum = nullptr; long gsum_len = 0;          /* flat [dense gradient sums | loss | gbar]: the all-reduce buffer */
  DenseUpdateArgs dense_args(int N);
  void shard_emb_lookup(const uint64_t* keys, int n, float* rows_out);                    /* owner: find-or-insert + gather */
  void shard_unpack_rows(const float* rows, const int32_t* send_pos, int N);              /* requester: rows → act[0] */
  void shard_dense_step(const float* X, const int64_t* W_local, const int64_t* W_all, int n_all, const float* Y, int N);
  void shard_pack(const int32_t* send_pos, int N, float* grads_send);
  void shard_finish(int N_global, int R);                                                 /* after the all-reduce of gsum */
  void shard_emb_apply(const float* grads_recv, int n);                                   /* owner: scatter + update */
  /* average device time (us) of the embedding kernels over `reps` graph-replayed repetitions on a ring
   * of device batches: out = {probe, gather, scatter_update, clear_batch} (differences of four loops) */
  void kernel_times(const int64_t* const* E_ring, int n_ring, int N, int reps, float* out);
  /* the same for the FcLayer GEMMs on the buffers of the last step: out[3*l + {0,1,2}] = {forward, dgrad, wgrad} of layer l */
  void gemm_times(int N, int reps, float* out);
  /* ---- the same sharded step over NVLink peer memory (p2p.cuh): no collective calls, one CUDA graph per rank ---- */
  P2P p2p;
  int32_t* send_pos = nullptr;
  void p2p_init(int R, int rank, int cap, void* handle_out64);
  void p2p_connect(const void* all_handles);
  /* consume_N > 0: the previous sharded step (of consume_N samples per rank) left its owner-side embedding update and its dense update
   * to this one; defer: this step leaves its own to the next */
  void p2p_step(const int64_t* E, const float* X, const int64_t* W, const float* Y, int N, int consume_N, bool defer);
  bool deferred_pending = false; int deferred_N = 0;
  void flush_deferred();             /* runs a pending deferred update now (every entry point that is not a sharded step calls it first) */
  void submit(const HostBatch& b, int mode = 0);     /* mode 1: the peer-memory sharded step (this rank's slice of the global batch) */
  float collect();
  float read_loss();
  void predict(const HostBatch& b, float* out);
  int get(const std::string& key, std::vector<float>& out);
  void put(const std::string& key, const float* in, int n);
  int get_state(const std::string& key, int which, std::vector<float>& out);
  /* PServer.push (PServer.java:164-184 → KVStore.update(updater, key), KVStore.java:202-208): one step of the updater `spec` names on an EXISTING
   * key with the gradient a legacy worker pushed (reference layout: rows x cols column-major, like put).  PS_NOT_FOUND for a missing key. */
  int push(const std::string& key, const float* g, int n, const ps_updater_spec& spec);
  int tap(const std::string& layer, int what, std::vector<float>& out);
  int64_t num_keys();
  /* bulk dump / load of every key of the store with its updater state (checkpoint.cu) */
  void save(const std::string& path);
  void load(const std::string& path);
  void mark(const char* phase);
  void finish_profile();
};

/* "emF3.15757.0" → (kind 0, field 3, id 15757); "wide.weights.77.0" → (kind 1, 0, 77); else kind 2 (dense name) */
int parse_key(const std::string& key, int* field, int64_t* id);

}  // namespace psb

/*
 * dense.cuh — dense-parameter side of store.KVStore (FcLayer weights / biases, wide.bias):
 * initialisation, the fused gradient-mean + updater step, and the network tail
 * (AddLayer + Sigmoid/Softmax + loss + loss gradient + last activation derivative).
 */
#pragma once
#include "common.cuh"
#include "updaters.cuh"

namespace psb {

struct P2PState;

/* device-resident step status; published to mapped host memory by the last kernel of a step */
struct StepStatus {
  float loss;       /* Model.train return value (model/DNN.java:47) */
  int skip;         /* 1 when loss <= CrossEntropy.slim or NaN: backward + update skipped (DNN.java:58-63) */
  float gbar;       /* rowMeans of the top delta: LRLayer's gradient (layer/LRLayer.java:110) */
  uint32_t emb_err; /* embedding table full */
  uint32_t wide_err;
  uint32_t n_unique;/* unique embedding keys of the batch */
  uint32_t seq;     /* step sequence number */
  uint32_t pad;     /* submit_text: lines of the batch the device parser could not take (non-zero = the batch is dropped) */
};

/* one FcLayer's parameters as the fused update kernel sees them */
struct DenseLayerDesc {
  float *W, *Wt, *bias;           /* W [out][ldw]; Wt [in][ldwt] transposed copy or null; bias [out] */
  float *Wlo, *Wtlo;              /* W - tf32(W), Wt - tf32(Wt): the residual operands of the 3xTF32 GEMMs, kept current by whoever writes W (or null) */
  float *sW1, *sW2, *sb1, *sb2;   /* updater state of "fc<i>.weights" / "fc<i>.bias" */
  const float* G;                 /* [nsplit][out][ldg]; column `in` is the bias-gradient sum */
  size_t slab;
  int nsplit, out, in, ldw, ldwt, ldg;
  UpdaterDev updW, updB;
  long first;                     /* index of this layer's first work item */
};
constexpr int kMaxDenseLayers = 16;
struct DenseUpdateArgs {
  DenseLayerDesc l[kMaxDenseLayers];
  int n_layers;
  long total;
  int N;                          /* batch size: gradients are batch means (FcLayer.java:103,105) */
};

void dense_init(Ctx* ctx, float* W, int out, int in, int ldw, float* Wt, int ldwt, uint64_t key, float maxv);
void fill_column(Ctx* ctx, float* buf, int ld, int col, int rows, float value);
/* lo[i] = x[i] - tf32_trunc(x[i]) over n floats: refreshes a weight matrix's residual copy after it was written from the host */
void split_lo(Ctx* ctx, const float* x, float* lo, size_t n);
/* also publishes the step status to mapped host memory when host_mapped is non-null */
void dense_update(Ctx* ctx, const DenseUpdateArgs& a, StepStatus* st, const uint32_t* emb_counters, const uint32_t* wide_counters,
                  StepStatus* host_mapped, const P2PState* p2p = nullptr /* gradients come from this step's gsum_in mailbox */);
/* peer-memory forms of dense_reduce / shard_finish_scalars (p2p.cuh): sums stored straight into every rank's mailbox */
void dense_reduce_send(Ctx* ctx, const DenseUpdateArgs& a, const StepStatus* st, P2PState* p2p, const uint32_t* emb_counters = nullptr);   /* publishes CH_GSUM */
void shard_finish_scalars_p2p(Ctx* ctx, StepStatus* st, const P2PState* p2p, long total);
void scalars_send(Ctx* ctx, const StepStatus* st, P2PState* p2p, long total, const uint32_t* emb_counters);   /* publishes CH_SCAL */
constexpr int kTailWorkspaceFloats = 2 * 1024 + 4;

/* binary tail: z = deep (+ wide); p = clipped sigmoid; CrossEntropy forward/backward; sigmoid
 * derivative.  Writes p to p_out (stride ldp), the post-derivative delta to d_out (stride ldd). */
void tail_binary(Ctx* ctx, int N, const float* zdeep, int ldz, const float* zwide, const float* Y, float* p_out, int ldp,
                 float* d_out, int ldd, float* dt_out, int train, StepStatus* st, float* ws, const uint32_t* emb_counters = nullptr);
/* the reverse loop's first step when the caller computed the loss itself: d = delta_top * p * (1 - p), gbar, skip; `loss` is recorded as given */
void tail_binary_from_delta(Ctx* ctx, int N, const float* p, int ldp, const float* dtop, float* d_out, int ldd, float* dt_out, float loss,
                            StepStatus* st, float* ws, const uint32_t* emb_counters);
void tail_softmax_from_delta(Ctx* ctx, int N, int C, const float* P, int ldp, const float* dtop, float* d_out, int ldd, float* dt_out, int ldt, float loss,
                             StepStatus* st, float* ws);
/* multi-class tail (FullConnectedNN): Softmax(10000) in place on Z, SoftmaxLoss, Softmax.backward */
void tail_softmax(Ctx* ctx, int N, int C, float* Z, int ldz, const float* Y, float* d_out, int ldd, float* dt_out, int ldt, int train,
                  StepStatus* st, float* ws);
/* the 1-unit top FcLayer on CUDA cores (see dense.cu): forward GEMV fused with the binary tail, dgrad, wgrad */
void fc1_forward_tail(Ctx* ctx, int N, int in, const float* A, int lda, const float* w, const float* bias, const float* zwide, const float* Y,
                      float* z_out, int ldz, float* p_out, int ldp, float* d_out, int ldd, float* dt_out, int train, StepStatus* st, float* ws,
                      const uint32_t* emb_counters = nullptr /* EmbTable::counters: a full table turns into the step's skip flag */);
void fc1_dgrad(Ctx* ctx, int N, int in, const float* d, int ldd, const float* dT /* contiguous copy of d or null */, const float* w, int act_below, const float* Y, int ldy, float* dX, int ldx,
               const float* Yt, int ldyt, float* dXt, int ldxt);
void fc1_wgrad(Ctx* ctx, int N, int in, const float* d, int ldd, const float* A, int lda, float* G, size_t slab, int nsplit);

/* copies the status (plus table error flags) to mapped host memory */
void publish_status(Ctx* ctx, StepStatus* st, const uint32_t* emb_counters, const uint32_t* wide_counters, StepStatus* host_mapped);
/* sharded (multi-GPU) step: gsum[compact (o, c) index] = sum of the wgrad slabs, then [loss, gbar] at
 * gsum[total], gsum[total+1] (+ the table-full flag at gsum[total+2]) — ONE flat buffer the host all-reduces across ranks               */
void dense_reduce(Ctx* ctx, const DenseUpdateArgs& a, const StepStatus* st, float* gsum, const uint32_t* emb_counters = nullptr);
/* after the all-reduce (sum over R ranks of per-rank means): global loss / gbar / early-exit flag */
void shard_finish_scalars(Ctx* ctx, StepStatus* st, const float* gsum_tail, int R);

/* Updater.update on caller-provided device arrays (test hook for ps_updater_apply) */
void updater_apply(Ctx* ctx, const UpdaterDev& u, float* w, float* s1, float* s2, const float* g, int n);
/* out[c*ldo + r] = in[r*ldi + c]  (layout conversion at the get/put boundary) */
void transpose_copy(Ctx* ctx, const float* in, int ldi, float* out, int ldo, int rows, int cols);

}  // namespace psb

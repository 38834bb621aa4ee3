/*
 * ps_b200.h — the C ABI of libps_b200.so: the drop-in boundary for the parameter-server
 * training hot path of wudikua/ps.  The reference has no FFI layer; its
 * boundary is a Java class surface.  Each entry point below names the Java method(s) it
 * stands behind (paths relative to /root/reference/src/main/java/); INTEGRATION.md shows
 * the JNI stubs a maintainer adds to call them from store.KVStore, layer.* and train.*.
 *
 * Conventions
 *  - every function returns PS_OK (0) or a PS_ERR_* code; ps_last_error() gives the text.
 *    (The reference swallows exceptions and returns null, KVStore.java:173-176,
 *    PSClient.java:64-69, or System.exit(0)s, AdamUpdater.java:65-68; here misuse is an
 *    error code and "key absent" is PS_NOT_FOUND, the 204 of net/PServer.java:84.)
 *  - all pointers are HOST pointers owned by the caller unless the name ends in _dev;
 *    device memory is owned by the library.  Matrices use the reference's layout: jblas
 *    FloatMatrix, column-major rows x columns, i.e. a features x N activation is N
 *    consecutive feature vectors.  "E" and "W" ids are int64 (the reference carries them in
 *    a float matrix, exact below 2^24 only — SURVEY quirk 2; a float variant is provided).
 *  - there is no CPU fallback: without a CUDA device every call fails with PS_ERR_CUDA.
 */
#ifndef PS_B200_H_
#define PS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PS_OK 0
#define PS_NOT_FOUND 204      /* net/PServer.java:28,84 — key absent */
#define PS_ERR_ARG 400        /* bad argument / misuse (RuntimeException in the reference) */
#define PS_ERR_CUDA 500       /* CUDA failure or no device (net/PServer.java:50-52 "500") */
#define PS_ERR_CAPACITY 507   /* embedding table full */
#define PS_ERR_STATE 409      /* call sequence violated (e.g. backward before forward) */

/* FcLayer arithmetic: fp32 FFMA (exact mode), TF32 tcgen05 tensor cores, or 3xTF32 tcgen05 (fp32-grade) */
#define PS_FC_FP32 0
#define PS_FC_TF32 1
#define PS_FC_TF32X3 2     /* tcgen05 with error-compensated operand split (3 MMAs per product): fp32-grade results */

/* activations.* */
#define PS_ACT_NONE 0
#define PS_ACT_RELU 1      /* activations/Relu.java:7-19 */
#define PS_ACT_SIGMOID 2   /* activations/Sigmoid.java:9-21 (clipped: 0.001 + 0.998*sigma) */
#define PS_ACT_SOFTMAX 3   /* activations/Softmax.java:21-67 (inputs pre-scaled by 1/10000) */

/* update.* */
#define PS_UPD_ADAM 0      /* update/AdamUpdater.java:57-70  p = {alfa, beta1, beta2, epsilon} */
#define PS_UPD_FTRL 1      /* update/FtrlUpdater.java:51-76  p = {alfa, beta, l1, l2}          */
#define PS_UPD_SIMPLE 2    /* update/SimpleUpdater.java:20-22 p = {eta}                        */
typedef struct ps_updater_spec {
  int32_t kind;
  float p[4];
} ps_updater_spec;

/* model.* */
#define PS_MODEL_DNN 0        /* model/DNN.java:92-128 */
#define PS_MODEL_WIDEDEEP 1   /* model/WideDeepNN.java:105-161 */
#define PS_MODEL_FCNN 2       /* model/FullConnectedNN.java:86-110 */

typedef struct ps_ctx ps_ctx;       /* the process-wide store.KVStore + device (KVStore.java:36,70) */
typedef struct ps_emb ps_emb;       /* layer.EmbeddingLayer with its F EmbeddingFields */
typedef struct ps_model ps_model;   /* a model.* network driven by train.Trainer, thread = 1 */

const char* ps_last_error(void);
int ps_abi_version(void);

/* ---- context ------------------------------------------------------------------- */
/* KVStore.ins() (KVStore.java:70) + Context.init() (context/Context.java:60-88).
 * `seed` keys the deterministic replacement of the unseeded MatrixUtil.rand (ps_spec.h). */
int ps_ctx_create(int device, uint64_t seed, ps_ctx** out);
int ps_ctx_destroy(ps_ctx* ctx);
/* FcLayer arithmetic (FcLayer.java:76,105,108).  Default PS_FC_TF32X3: tcgen05 tensor cores with the error-compensated operand
 * split (fp32-grade results: the reference computes in fp32).  PS_FC_TF32: plain TF32 tensor cores; PS_FC_FP32: FFMA exact mode. */
int ps_ctx_set_fc_precision(ps_ctx* ctx, int mode);   /* PS_FC_FP32 | PS_FC_TF32 | PS_FC_TF32X3 */
int ps_ctx_get_fc_precision(ps_ctx* ctx, int* mode);
/* the sparse (embedding-row) update: 0 = fast forms (reciprocal multiplications, approximate quotient / root; post-update
 * rows within 1e-6 relative of the exact forms), 1 = the IEEE operation sequence of AdamUpdater.java:57-70 /
 * FtrlUpdater.java:51-76 (bit-exact given the gradient).  Dense parameters always use the exact forms.  Default 0
 * (PS_EXACT_UPDATERS=1 in the environment flips it).  Takes effect for graphs captured afterwards: call before the first step. */
int ps_ctx_set_exact_updaters(ps_ctx* ctx, int on);
int ps_ctx_synchronize(ps_ctx* ctx);
/* A context belongs to one device; its creating thread has that device current.  Any OTHER host thread that is going to call into the
 * context (a JVM worker pool, the gRPC threads of ps_b200/wire.py) calls this first.  Calls on one context must not overlap in time. */
int ps_ctx_make_current(ps_ctx* ctx);
int ps_ctx_launch_count(ps_ctx* ctx, int64_t* out);   /* kernels this library launched so far */
int ps_ctx_device_info(ps_ctx* ctx, char* name, int cap, int* sms, int* cc_major, int* cc_minor);
int ps_ctx_stream(ps_ctx* ctx, void** stream);        /* the cudaStream_t every step is ordered on (for event timing) */

/* page-locked host memory for batch buffers (what a JNI caller wraps in a direct ByteBuffer):
 * submit()/train_step() copy asynchronously only from such memory.                          */
int ps_host_alloc(size_t bytes, void** out);
int ps_host_free(void* p);

/* ---- update.* -------------------------------------------------------------------
 * AdamUpdater(String) / FtrlUpdater(String) (AdamUpdater.java:50-55, FtrlUpdater.java:44-49):
 * parses "adam@alfa:0.005@beta1:0.9@beta2:0.999@epsilon:1.0E-8@" or the Ftrl spelling
 * "adam@alfa:..@beta:..@l1:..@l2:..@" (FtrlUpdater.getName says "adam@", sic, :78-80).   */
int ps_updater_parse(const char* name, ps_updater_spec* out);
/* getName() (AdamUpdater.java:72-74, FtrlUpdater.java:78-80) with Java float spelling    */
int ps_updater_name(const ps_updater_spec* spec, char* buf, int cap);
/* Updater.update(key, w, dw) (update/Updater.java:8) on caller-held state, n elements,
 * executed by the same device code the fused kernels use (test hook).                     */
int ps_updater_apply(ps_ctx* ctx, const ps_updater_spec* spec, float* w, float* s1, float* s2, const float* g, int n);

/* ---- layer.EmbeddingLayer / layer.EmbeddingField ---------------------------------
 * EmbeddingLayer(name, F, F*D).build(F, D) (EmbeddingLayer.java:21,50-57).  Rows live in a
 * GPU-resident open-addressing table of `capacity` slots keyed (field, id) — the reference's
 * "emF<field>.<id>" strings (EmbeddingField.java:70) — created on first touch with the
 * Xavier bound of EmbeddingField.java:40 (KVStore.java:136-159,168-190).                  */
int ps_emb_create(ps_ctx* ctx, int F, int D, int64_t capacity, const ps_updater_spec* upd, ps_emb** out);
int ps_emb_destroy(ps_emb* emb);
/* EmbeddingLayer.forward() (EmbeddingLayer.java:25-48) → EmbeddingField.forward ×F
 * (EmbeddingField.java:66-78): E is F x N (ids), out is (F*D) x N, ReLU applied.          */
int ps_emb_forward(ps_emb* emb, const int64_t* E, int N, float* out);
int ps_emb_forward_f32ids(ps_emb* emb, const float* E, int N, float* out);   /* the FloatMatrix "E" as the reference carries it */
/* EmbeddingLayer.backward() called `calls` times (the reference makes 2 per step:
 * ConcatLayer.java:44 and DNN.java:66-68) on delta (ld x N, rows >= F*D ignored), then
 * KVStore.update(updaters) + KVStore.clear() (Trainer.java:93,95) for the touched rows:
 * a pre-summed scatter-add kernel, then (as its programmatic dependent) one kernel for the
 * occurrence normalisation, the Adam/Ftrl step and the per-batch reset.                   */
int ps_emb_backward_update(ps_emb* emb, const float* delta, int ld, int N, int calls);
/* KVStore.get(String) / PSClient.getList (KVStore.java:129-134, PSClient.java:72-97):
 * host snapshot of rows (and optimiser state when s1/s2 non-null); found[i] = 0 if absent. */
int ps_emb_get_rows(ps_emb* emb, const int32_t* fields, const int64_t* ids, int n, float* w, float* s1, float* s2, int32_t* found);
/* KVStore.put / PSClient.updateList (KVStore.java:161-166, PSClient.java:128-151):
 * replace != 0 overwrites, replace == 0 is insert-if-absent and returns the winner in w.   */
int ps_emb_put_rows(ps_emb* emb, const int32_t* fields, const int64_t* ids, int n, float* w, int replace);
int ps_emb_size(ps_emb* emb, int64_t* rows);

/* ---- layer.FcLayer as a standalone operator ---------------------------------------
 * For callers that walk the reference's layer list themselves (Model.train's forward / backward loops, DNN.java:44-68)
 * instead of handing the whole step to ps_model_train_step.  Matrices are jblas column-major: A_prev is in x N, A and
 * delta are out x N, delta_prev is in x N.  Weights "<name>.weights" / bias "<name>.bias" are created on first use with
 * the Xavier bounds of FcLayer.java:39,46 (FcLayer.pullWeights, :112-115); `act` is PS_ACT_NONE | RELU | SIGMOID
 * (FcLayer.setActivation incl. null, WideDeepNN.java:128); upd NULL = the "default" Adam of DNN.java:95.
 *   ps_fc_forward   FcLayer.forward  (FcLayer.java:74-91):  A = act(W * A_prev + b 1^T)
 *   ps_fc_backward  FcLayer.backward (FcLayer.java:93-110): delta <- act'(delta); db = rowMeans(delta), dW = delta * A_prev^T / N
 *                   are kept on the device as the pending KVStore.sum; delta_prev = W^T * delta is returned (may be NULL)
 *   ps_fc_gradients the pending dW (out x in) and db (out) as KVStore.sum received them
 *   ps_fc_update    KVStore.update(updaters) + clear() for this layer's two keys (KVStore.java:240-277)
 *   ps_fc_get / ps_fc_put  KVStore.get / put of "<name>.weights" (which = 0, out x in) or "<name>.bias" (which = 1)      */
typedef struct ps_fc ps_fc;
int ps_fc_create(ps_ctx* ctx, const char* name, int in, int out, int act, const ps_updater_spec* upd, int max_batch, ps_fc** out_fc);
int ps_fc_destroy(ps_fc* fc);
int ps_fc_forward(ps_fc* fc, const float* A_prev, int N, float* A);
int ps_fc_backward(ps_fc* fc, const float* delta, int N, float* delta_prev);
int ps_fc_gradients(ps_fc* fc, float* dW, float* db);
int ps_fc_update(ps_fc* fc);
int ps_fc_get(ps_fc* fc, int which, float* out, int cap, int* n);
int ps_fc_put(ps_fc* fc, int which, const float* in, int n);

/* ---- model.* driven by train.Trainer (thread = 1) --------------------------------
 * DNN.buildModel / WideDeepNN.buildModel / FullConnectedNN.buildModel with the reference's
 * default updaters (DNN.java:95, WideDeepNN.java:109-113).  emb_updater may be NULL (the
 * "default" Adam) or e.g. Ftrl to express updaters.put("emF", ftrl) (KVStore.java:244-248). */
int ps_model_create(ps_ctx* ctx, int kind, int F, int D, int Xn, const int32_t* fc_dims, int n_fc,
                    int64_t emb_capacity, const ps_updater_spec* emb_updater, int max_batch, ps_model** out);
int ps_model_destroy(ps_model* m);
/* TrainerThread.call (TrainerThread.java:29-39: pullWeights + Model.train) followed by
 * Trainer.train's KVStore.update(updaters) + clear() (Trainer.java:93,95).  Inputs as
 * CTR.parseFeature builds them (CTR.java:47-68): E, W are F x N, X is Xn x N, Y is 1 x N.
 * Returns the loss Model.train returns (Model.java:11).                                    */
int ps_model_train_step(ps_model* m, const int64_t* E, const float* X, const int64_t* W, const float* Y, int N, float* loss);
/* the same step split in two so the host copy of batch i+1 overlaps the kernels of batch i
 * (what the reference's DataSet reader thread does for parsing, data/DataSet.java:77-100):
 * submit() enqueues H2D + the whole step, collect() waits for the oldest and returns its loss.
 * At most 4 steps may be in flight; host buffers must stay valid until the matching collect. */
int ps_model_submit(ps_model* m, const int64_t* E, const float* X, const int64_t* W, const float* Y, int N);
int ps_model_collect(ps_model* m, float* loss);
/* Model.train CALL BY CALL, for an unchanged model.DNN / model.WideDeepNN whose loss stays in Java:
 *   ps_model_forward          the forward loop `for (Layer layer : layers) layer.forward()` (DNN.java:44-46, WideDeepNN.java:52-58):
 *                             P_out[N] = layers.get(last).getA(); the batch stays pending on the device (occurrence counts, ReLU
 *                             masks, activations).  No label is passed: loss.forward / loss.backward run in the caller (DNN.java:47-49).
 *   ps_model_backward_update  delta_top[N] = loss.backward(P, Y) exactly as setDelta() receives it (DNN.java:64: dLoss/dP, BEFORE the
 *                             output Sigmoid's derivative); runs the reverse loop (DNN.java:65-68: FcLayer.backward incl. its
 *                             activation.backward, LRLayer.backward, EmbeddingLayer.backward x2) and then KVStore.update + clear
 *                             (Trainer.java:93,95).  `loss` = the caller's loss.forward value, recorded as the step's loss; the early
 *                             exit of DNN.java:58-63 is the caller's to take (it simply does not call this; the next forward forgets
 *                             the pending batch).  FullConnectedNN (Softmax + SoftmaxLoss, FullConnectedNN.java:37-70): P_out and
 *                             delta_top are C x N (N rows of C floats); Softmax.backward (Softmax.java:45-67) runs in the library.   */
int ps_model_forward(ps_model* m, const int64_t* E, const float* X, const int64_t* W, int N, float* P_out);
int ps_model_backward_update(ps_model* m, const float* delta_top, int N, float loss);
/* DataSet.next + Trainer.train in one submission (DataSet.java:77-100, CTR.parseFeature CTR.java:47-68, CTR.wideSize CTR.java:36):
 * `len` bytes of libsvm text (host memory, ideally pinned) holding exactly N complete '\n'-terminated lines; the text is copied
 * to the device as it is, parsed THERE into E / X / W / Y and trained on, all asynchronously; ps_model_collect returns the loss.
 * A line the device parser cannot take (short, blank, missing, or a spelling outside its fast path — see ps_libsvm_parse_dev)
 * drops the whole batch, as the reference's swallowed exception does (DataSet.java:96-98): nothing is applied, ps_model_step_info
 * reports skipped = 1 and the number of offending lines.                                                                    */
int ps_model_submit_text(ps_model* m, const char* text, size_t len, int N);
/* status of the step last collected: skipped (early exit / dropped batch / full table), bad text lines, unique embedding keys */
int ps_model_step_info(ps_model* m, int* skipped, uint32_t* bad_lines, uint32_t* n_unique);
/* the geometry the model was created with (a binding checks its callers' array lengths against it before any native read) */
int ps_model_shape(ps_model* m, int* F, int* D, int* Xn);
/* the step on inputs ALREADY resident in device memory (bench `value` leg); loss stays on the
 * device until ps_model_read_loss.                                                         */
int ps_model_train_step_dev(ps_model* m, const int64_t* E_dev, const float* X_dev, const int64_t* W_dev, const float* Y_dev, int N);
int ps_model_read_loss(ps_model* m, float* loss);
/* device address of the step's result (a float: the loss Model.train returns, Model.java:11), valid after any *_dev step:
 * lets a pipelined caller copy it out asynchronously on ps_ctx_stream instead of synchronising in ps_model_read_loss */
int ps_model_loss_dev(ps_model* m, const float** loss_dev);
/* PredictThread.call + Trainer.predict (train/PredictThread.java, Trainer.java:44-68)      */
int ps_model_predict(ps_model* m, const int64_t* E, const float* X, const int64_t* W, int N, float* out);
/* KVStore.get(String) on any key the reference would hold: "fc0.weights" (out x in, column
 * major), "fc0.bias", "wide.bias", "emF3.15757.0", "wide.weights.77.0".  *n = element count. */
int ps_model_get(ps_model* m, const char* key, float* out, int cap, int* n);
int ps_model_put(ps_model* m, const char* key, const float* in, int n);      /* KVStore.put */
/* updater state of a key: which = 0 Adam M / Ftrl Z, 1 Adam V / Ftrl N                       */
/* PSClient.getList / updateList (net/PSClient.java:72-97,128-151; server side net/PServer.java:102-117,144-162) by reference key
 * strings, batched: embedding keys ("emF<j>.<id>", the bulk of any list) are served by ONE lookup / insert kernel, other keys one
 * by one.  Row i of out / io starts at i*stride floats.  get: found[i] = number of floats written (0 = absent, Resp.ec 204).
 * update: lens[i] floats of io row i are offered; replace = 0 is the reference's insert-if-absent ("重复key不替换"), io then
 * receives the winning value of every key.                                                                                   */
int ps_model_get_list(ps_model* m, const char* const* keys, int n, float* out, int stride, int32_t* found);
int ps_model_update_list(ps_model* m, const char* const* keys, int n, float* io, int stride, const int32_t* lens, int replace);
int ps_model_get_state(ps_model* m, const char* key, int which, float* out, int cap, int* n);
/* PServer.push (net/PServer.java:164-184) → KVStore.update(updater, key) (store/KVStore.java:202-208): ONE step of the updater `spec` names
 * (the request's updaterKey, parsed with ps_updater_parse) on an existing key, with the gradient a worker pushed — in the reference's layout
 * (rows x cols column-major, as ps_model_put takes values).  For workers that still compute gradients themselves and speak the gRPC protocol
 * (ps_b200/wire.py serves it); the native step never needs it.  PS_NOT_FOUND: no such key (the reference's updaters exit on a null weight). */
int ps_model_push(ps_model* m, const char* key, const float* grad, int n, const ps_updater_spec* spec);
/* Layer.getA() / getDelta() after the last step (layer/Layer.java:16-45): what = 0 A, 1 delta;
 * layers: "embedding", "concat", "fc<i>", "wide", "addWideDeep"                               */
int ps_model_tap(ps_model* m, const char* layer, int what, float* out, int cap, int* n);
int ps_model_num_keys(ps_model* m, int64_t* out);            /* KVStore.store.size() */
int ps_model_skipped_backward(ps_model* m, int* out);        /* DNN.java:58-63 early exit taken */
/* Bulk dump / load of the store: every key store.KVStore would hold (KVStore.java:38-44; PServer.getList /
 * upsertList, PServer.java:102-162, are the reference's only bulk accessors — it has no checkpoint) with its updater state,
 * to / from one file.  Load needs a freshly created model of the same shape; embedding rows are re-inserted by key, so the
 * table capacity may differ from the one that wrote the file.  PS_NOT_FOUND when the file cannot be opened.               */
int ps_model_save(ps_model* m, const char* path);
int ps_model_load(ps_model* m, const char* path);
/* step-level timing of the last ps_model_train_step_dev: device milliseconds per phase
 * (CUDA events on the library's stream); names returned as a ';'-separated list.             */
int ps_model_profile(ps_model* m, int enable);
int ps_model_phase_times(ps_model* m, float* ms, int cap, int* n, char* names, int names_cap);

/* measurement hook: average device microseconds of the embedding kernels of this model, each
 * replayed `reps` times inside a CUDA graph over a ring of device-resident E batches and timed with
 * CUDA events on the library's stream: us[4] = {probe, gather, scatter_update, clear_batch}.  The
 * scatter_update repetitions apply real (meaningless) updates: call it after the timed training. */
int ps_model_kernel_times(ps_model* m, const int64_t* const* E_dev_ring, int n_ring, int N, int reps, float* us);
/* the same for the FcLayer GEMMs (idempotent on the buffers of the last step): us[3*l + {0,1,2}] = {forward, dgrad, wgrad} of fc<l> */
int ps_model_gemm_times(ps_model* m, int N, int reps, float* us, int cap);

/* ---- key-hash sharded table across the GPUs of one box (net/PSRouterClient.java:60-151) ----------
 * One process per GPU.  PSRouterClient buckets keys by router.shard(key), sends one batched
 * getList / updateList per shard and merges (PSRouterClient.java:60-85, 93-122); here the buckets
 * travel in ONE all-to-all over NVLink per direction, issued by the host between these calls
 * (ps_b200/sharded.py with torch.distributed; INTEGRATION.md shows the sequence).  All pointers
 * in this group are DEVICE pointers; calls are asynchronous on the context's stream.
 *   requester: ps_shard_route_dev → [all-to-all keys] → owner: ps_model_shard_lookup_dev →
 *   [all-to-all rows] → requester: ps_model_shard_unpack_dev, ps_model_shard_dense_step_dev,
 *   ps_model_shard_pack_grads_dev → [all-reduce grad buffer] [all-to-all row gradients] →
 *   ps_model_shard_finish_dev, owner: ps_model_shard_apply_dev.
 * Semantics: the R ranks together perform ONE Trainer step on the concatenated batch
 * (thread = 1): an N-GPU step equals the 1-GPU step on the same global batch.   */
/* owner(key) = ps_owner_of(pack(field, id), R) (include/ps_spec.h): a net/Router.java:5.  counts_dev[R]
 * and cursor_dev[R] are int32 scratch (zeroed by the call); send_keys_dev[N*F] is grouped by owner,
 * send_pos_dev[N*F] maps each lookup to its place in that order.                                      */
/* net/Router.java:5 `int shard(String key)` for the reference's key strings, host side, no device: *owner = the shard of an embedding key
 * ("emF<field>.<id>.0" → ps_owner_of(pack(field, id), n_shards)), or -1 for keys every shard holds (wide and dense parameters are replicated).
 * What PSRouterClient.java:55-57 asks its Router before it picks a client; the device-side route kernels use the same function. */
int ps_key_owner(const char* key, int n_shards, int* owner);
int ps_shard_route_dev(ps_ctx* ctx, const int64_t* E_dev, int N, int F, int R, uint64_t* send_keys_dev, int32_t* send_pos_dev,
                       int32_t* counts_dev, int32_t* cursor_dev);
/* the same with a FIXED bucket capacity per owner (send_keys_dev[R*cap], unused entries = 0 = EMPTY, which
 * the owner skips): no count has to reach the host, so route → all-to-all → lookup → … → apply can be
 * captured in ONE CUDA graph per rank.  *overflow_dev is set to 1 if a bucket did not fit.              */
int ps_shard_route_padded_dev(ps_ctx* ctx, const int64_t* E_dev, int N, int F, int R, int cap, uint64_t* send_keys_dev, int32_t* send_pos_dev,
                              int32_t* cursor_dev, int32_t* overflow_dev);
/* PServer.getList (net/PServer.java:102-117) for the keys this rank owns: rows_out_dev[n][Dp], ReLU applied */
int ps_model_shard_lookup_dev(ps_model* m, const uint64_t* keys_dev, int n, float* rows_out_dev);
int ps_model_shard_row_stride(ps_model* m, int* Dp);       /* floats per exchanged row (D rounded up to 4) */
int ps_model_shard_unpack_dev(ps_model* m, const float* rows_dev, const int32_t* send_pos_dev, int N);
/* concat + wide + FcLayer forward/backward on this rank's N samples; W_all_dev holds the wide ids of
 * EVERY rank (the replicated wide table must learn them all, LRLayer.java:79)                          */
int ps_model_shard_dense_step_dev(ps_model* m, const float* X_dev, const int64_t* W_local_dev, const int64_t* W_all_dev, int n_all,
                                  const float* Y_dev, int N);
/* the flat fp32 buffer [dense gradient sums | loss | gbar] to all-reduce (SUM) across ranks           */
int ps_model_shard_grad_buffer(ps_model* m, float** buf_dev, int64_t* count);
int ps_model_shard_pack_grads_dev(ps_model* m, const int32_t* send_pos_dev, int N, float* grads_send_dev);
/* KVStore.update for dense keys and the wide branch with the GLOBAL batch mean (Trainer.java:93)      */
int ps_model_shard_finish_dev(ps_model* m, int N_global, int R);
/* PServer.push + psUpdate (net/PServer.java:164-214) for the rows this rank owns: fused scatter + update */
int ps_model_shard_apply_dev(ps_model* m, const float* grads_recv_dev, int n);

/* ---- the same sharded step over NVLink PEER MEMORY (no collective library on the data path) --------
 * Every rank maps every peer's mailbox slab (CUDA IPC); the kernel that produces a bucket (routed keys,
 * gathered rows, row gradients, dense gradient sums, wide ids) stores it straight into the consumer's HBM
 * through NVSwitch and, when its last block ends, flags the consumers, whose kernels wait for the flags in their prologue — there
 * are no flag kernels and no collective calls.  A whole step is ONE CUDA
 * graph per rank: ps_model_p2p_step_dev enqueues it (asynchronous), ps_model_read_loss reads the result.
 *   ps_model_p2p_init    → allocates the slab, returns its 64-byte cudaIpcMemHandle_t
 *   [host: all-gather the R handles, e.g. torch.distributed.all_gather]
 *   ps_model_p2p_connect → opens the peers' slabs (rank order); all ranks must then barrier once
 * cap = capacity of one per-owner bucket (>= the most lookups one rank sends to one owner in a step). */
int ps_model_p2p_init(ps_model* m, int R, int rank, int cap, void* ipc_handle_out64);
int ps_model_p2p_connect(ps_model* m, const void* all_handles /* R x 64 bytes */);
int ps_model_p2p_step_dev(ps_model* m, const int64_t* E_dev, const float* X_dev, const int64_t* W_dev, const float* Y_dev, int N);
/* the pipelined host-facing form (ps_model_submit's twin): this rank's slice from (pinned) HOST memory, at most 2 steps in
 * flight, ps_model_collect returns the GLOBAL loss of the oldest                                                           */
int ps_model_p2p_submit(ps_model* m, const int64_t* E, const float* X, const int64_t* W, const float* Y, int N);
int ps_model_p2p_overflowed(ps_model* m, int* out);   /* a bucket exceeded cap at some step: results invalid, raise cap */

/* ---- libsvm ingest feeding the step (host code: callable without a GPU) ---------------------------
 * ps_libsvm_parse_line: data.LibsvmParser.parse (LibsvmParser.java:13-25) + one column of CTR.parseFeature (CTR.java:47-68):
 *   "label idx:val idx:val ..." -> Y = label, E[j] = (float) idx of columns 1..F, X[x] = val of columns F+1..F+Xn,
 *   W[j] = E[j] % wide_size in float arithmetic (MatrixUtil.hash, MatrixUtil.java:27-33; CTR.wideSize = 100000).
 *   *status: 0 ok, 1 blank/short line (IndexOutOfBounds in parseFeature), 2 unparsable (exception in parser.parse).
 * ps_reader_*: data.DataSet over data.FileSource (DataSet.java:37-100, DataSource.java:25-46): this reader sees file lines
 *   offset, offset+step, ... in batches of `batch` (a short last batch is delivered); batches the reference loses to its
 *   swallowed exceptions are lost here too (counted in ps_reader_stats).  A background thread keeps parsed batches ahead
 *   of the consumer; ps_reader_next copies one into the caller's buffers — E, W: [rows][F] int64, X: [rows][Xn], Y: [rows],
 *   the layout ps_model_train_step / ps_model_submit take — and returns *rows = 0 at end of data (DataSet.next() == null). */
/* The same parse ON the GPU, for training straight from text at step speed (raw text is ~700 B per line: PCIe carries it; the host
 * parser does not keep up).  text_dev: `len` bytes (< 4 GiB) of complete lines in DEVICE memory, each terminated by '\n'.  At most
 * max_rows lines are parsed into E_dev, W_dev [rows][F], X_dev [rows][Xn], Y_dev [rows]; status_dev[r]: 0 ok, 1 blank / short line,
 * 2 = a spelling outside the fast path (exponent, suffix, hex float, NaN / Infinity, signed index, empty token ...): nothing was
 * guessed — re-parse that line with ps_libsvm_parse_line.  Rows with status 0 are bit-identical to the host parser's.
 * *rows = number of lines parsed (the first max_rows of those present; a last line without '\n' is not a line yet).  Kernels
 * run on ps_ctx_stream; the call returns once the line count is known.                                                      */
int ps_libsvm_parse_dev(ps_ctx* ctx, const char* text_dev, size_t len, int F, int Xn, int64_t wide_size, int max_rows,
                        int64_t* E_dev, float* X_dev, int64_t* W_dev, float* Y_dev, uint8_t* status_dev, int* rows);
typedef struct ps_reader ps_reader;
int ps_libsvm_parse_line(const char* line, size_t len, int F, int Xn, int64_t wide_size, int64_t* E, float* X, int64_t* W, float* Y, int* status);
int ps_reader_open(const char* path, int F, int Xn, int64_t wide_size, int batch, int offset, int step, int threads, ps_reader** out);
int ps_reader_next(ps_reader* r, int64_t* E, float* X, int64_t* W, float* Y, int* rows);
int ps_reader_shape(ps_reader* r, int* batch, int* F, int* Xn);   /* ps_reader_next writes up to batch rows into every output */
int ps_reader_reset(ps_reader* r);                     /* DataSet.reset (DataSet.java:61-67) */
int ps_reader_stats(ps_reader* r, int64_t* lines, int64_t* batches, int64_t* dropped_batches);
int ps_reader_close(ps_reader* r);

/* ---- test hooks -------------------------------------------------------------------- */
/* C (M x N, row-major, ldc) = A (M x K, row-major, lda) * B^T (B is N x K, row-major, ldb),
 * through the FcLayer GEMM of the given precision mode.                                      */
int ps_test_gemm_nt(ps_ctx* ctx, int mode, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc);

#ifdef __cplusplus
}
#endif
#endif /* PS_B200_H_ */

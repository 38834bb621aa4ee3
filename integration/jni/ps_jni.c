/*
 * ps_jni.c — the JNI shim between integration/java/nativeps/PsNative.java and the C ABI of
 * include/ps_b200.h.  Deliberately thin: pin the Java arrays, widen float ids to int64, call.
 * NOT built in this repository's image (no JDK here; tests/test_capi.py type-checks it against include/ps_b200.h with the stub
 * header integration/jni/stub/jni.h); build on a host with a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include \
 *       ps_jni.c -L../../ps_b200/lib -lps_b200 -o libps_b200_jni.so
 */
#include <jni.h>
#include <stdlib.h>
#include <string.h>

#include "ps_b200.h"

static void throw_ps(JNIEnv* env, int rc) {
  if (rc == PS_OK) return;
  jclass cls = (*env)->FindClass(env, "java/lang/RuntimeException");   /* the reference throws RuntimeException on misuse (KVStore.java:163) */
  (*env)->ThrowNew(env, cls, ps_last_error());
}

/* every Java array is checked against the batch geometry BEFORE the native call reads it: a short array would otherwise be read
 * (or, for the reader's outputs, written) past its end */
static int check_len(JNIEnv* env, jarray a, jlong need, const char* what) {
  if (a == NULL) return 1;
  if ((jlong)(*env)->GetArrayLength(env, a) >= need) return 1;
  jclass cls = (*env)->FindClass(env, "java/lang/IllegalArgumentException");
  if (cls) (*env)->ThrowNew(env, cls, what);
  return 0;
}

static int64_t* widen_ids(JNIEnv* env, jfloatArray a, jsize* n) {   /* float-carried ids (exact < 2^24) */
  if (!a) { *n = 0; return NULL; }
  *n = (*env)->GetArrayLength(env, a);
  jfloat* f = (*env)->GetPrimitiveArrayCritical(env, a, NULL);
  int64_t* out = (int64_t*)malloc(sizeof(int64_t) * (size_t)(*n > 0 ? *n : 1));
  for (jsize i = 0; out != NULL && i < *n; ++i) out[i] = (int64_t)f[i];
  (*env)->ReleasePrimitiveArrayCritical(env, a, f, JNI_ABORT);
  return out;
}

JNIEXPORT jlong JNICALL Java_nativeps_PsNative_ctxCreate(JNIEnv* env, jclass c, jint device, jlong seed) {
  ps_ctx* ctx = NULL;
  throw_ps(env, ps_ctx_create(device, (uint64_t)seed, &ctx));
  return (jlong)(intptr_t)ctx;
}
JNIEXPORT void JNICALL Java_nativeps_PsNative_ctxDestroy(JNIEnv* env, jclass c, jlong ctx) { ps_ctx_destroy((ps_ctx*)(intptr_t)ctx); }
JNIEXPORT void JNICALL Java_nativeps_PsNative_ctxSetFcPrecision(JNIEnv* env, jclass c, jlong ctx, jint mode) {
  throw_ps(env, ps_ctx_set_fc_precision((ps_ctx*)(intptr_t)ctx, mode));
}
JNIEXPORT jstring JNICALL Java_nativeps_PsNative_lastError(JNIEnv* env, jclass c) { return (*env)->NewStringUTF(env, ps_last_error()); }

JNIEXPORT jlong JNICALL Java_nativeps_PsNative_modelCreate(JNIEnv* env, jclass c, jlong ctx, jint kind, jint F, jint D, jint Xn, jintArray fc,
                                                           jlong cap, jfloatArray upd, jint maxBatch) {
  jsize nfc = (*env)->GetArrayLength(env, fc);
  jint* dims = (*env)->GetIntArrayElements(env, fc, NULL);
  ps_updater_spec spec, *sp = NULL;
  if (upd) {
    jfloat* u = (*env)->GetFloatArrayElements(env, upd, NULL);
    spec.kind = (int32_t)u[0]; for (int i = 0; i < 4; ++i) spec.p[i] = u[1 + i];
    (*env)->ReleaseFloatArrayElements(env, upd, u, JNI_ABORT);
    sp = &spec;
  }
  ps_model* m = NULL;
  throw_ps(env, ps_model_create((ps_ctx*)(intptr_t)ctx, kind, F, D, Xn, (const int32_t*)dims, nfc, cap, sp, maxBatch, &m));
  (*env)->ReleaseIntArrayElements(env, fc, dims, JNI_ABORT);
  return (jlong)(intptr_t)m;
}
JNIEXPORT void JNICALL Java_nativeps_PsNative_modelDestroy(JNIEnv* env, jclass c, jlong m) { ps_model_destroy((ps_model*)(intptr_t)m); }

JNIEXPORT jfloat JNICALL Java_nativeps_PsNative_modelTrainStep(JNIEnv* env, jclass c, jlong m, jfloatArray E, jfloatArray X, jfloatArray W,
                                                               jfloatArray Y, jint N) {
  jsize ne, nw;
  int F = 0, Xn = 0;
  throw_ps(env, ps_model_shape((ps_model*)(intptr_t)m, &F, NULL, &Xn));
  if (N <= 0 || !check_len(env, E, (jlong)N * F, "E shorter than F x N") || !check_len(env, W, (jlong)N * F, "W shorter than F x N") ||
      !check_len(env, X, (jlong)N * Xn, "X shorter than Xn x N") || !check_len(env, Y, N, "Y shorter than N")) return 0.f;
  int64_t* e = widen_ids(env, E, &ne);
  int64_t* w = widen_ids(env, W, &nw);
  jfloat* x = (*env)->GetFloatArrayElements(env, X, NULL);
  jfloat* y = (*env)->GetFloatArrayElements(env, Y, NULL);
  float loss = 0.f;
  int rc = ps_model_train_step((ps_model*)(intptr_t)m, e, x, w, y, N, &loss);   /* Model.train's return value (Model.java:11) */
  (*env)->ReleaseFloatArrayElements(env, X, x, JNI_ABORT);
  (*env)->ReleaseFloatArrayElements(env, Y, y, JNI_ABORT);
  free(e); free(w);
  throw_ps(env, rc);
  return loss;
}

JNIEXPORT jfloatArray JNICALL Java_nativeps_PsNative_modelGet(JNIEnv* env, jclass c, jlong m, jstring key) {
  const char* k = (*env)->GetStringUTFChars(env, key, NULL);
  int n = 0;
  int rc = ps_model_get((ps_model*)(intptr_t)m, k, NULL, 0, &n);
  jfloatArray out = NULL;
  if (rc == PS_OK) {
    out = (*env)->NewFloatArray(env, n);
    jfloat* p = (*env)->GetFloatArrayElements(env, out, NULL);
    rc = ps_model_get((ps_model*)(intptr_t)m, k, p, n, &n);
    (*env)->ReleaseFloatArrayElements(env, out, p, 0);
  }
  (*env)->ReleaseStringUTFChars(env, key, k);
  if (rc != PS_OK && rc != PS_NOT_FOUND) throw_ps(env, rc);
  return rc == PS_OK ? out : NULL;                                              /* KVStore.get(String) returns null when absent */
}
JNIEXPORT void JNICALL Java_nativeps_PsNative_modelPut(JNIEnv* env, jclass c, jlong m, jstring key, jfloatArray v) {
  const char* k = (*env)->GetStringUTFChars(env, key, NULL);
  jsize n = (*env)->GetArrayLength(env, v);
  jfloat* p = (*env)->GetFloatArrayElements(env, v, NULL);
  int rc = ps_model_put((ps_model*)(intptr_t)m, k, p, n);
  (*env)->ReleaseFloatArrayElements(env, v, p, JNI_ABORT);
  (*env)->ReleaseStringUTFChars(env, key, k);
  throw_ps(env, rc);
}
/* PServer.push → KVStore.update(updater, key): false when the key does not exist (the reference's updaters exit the JVM on a null weight) */
JNIEXPORT jboolean JNICALL Java_nativeps_PsNative_modelPush(JNIEnv* env, jclass c, jlong m, jstring key, jfloatArray grad, jstring updaterKey) {
  const char* k = (*env)->GetStringUTFChars(env, key, NULL);
  const char* u = (*env)->GetStringUTFChars(env, updaterKey, NULL);
  ps_updater_spec spec;
  int rc = ps_updater_parse(u, &spec);
  if (rc == PS_OK) {
    jsize n = (*env)->GetArrayLength(env, grad);
    jfloat* p = (*env)->GetFloatArrayElements(env, grad, NULL);
    rc = ps_model_push((ps_model*)(intptr_t)m, k, p, n, &spec);
    (*env)->ReleaseFloatArrayElements(env, grad, p, JNI_ABORT);
  }
  (*env)->ReleaseStringUTFChars(env, updaterKey, u);
  (*env)->ReleaseStringUTFChars(env, key, k);
  if (rc != PS_OK && rc != PS_NOT_FOUND) throw_ps(env, rc);
  return rc == PS_OK ? JNI_TRUE : JNI_FALSE;
}
/* for JVM threads other than the one that created the context (a gRPC executor, Trainer's pool) */
JNIEXPORT void JNICALL Java_nativeps_PsNative_ctxMakeCurrent(JNIEnv* env, jclass c, jlong ctx) {
  throw_ps(env, ps_ctx_make_current((ps_ctx*)(intptr_t)ctx));
}
JNIEXPORT jfloatArray JNICALL Java_nativeps_PsNative_modelTap(JNIEnv* env, jclass c, jlong m, jstring layer, jint what) {
  const char* k = (*env)->GetStringUTFChars(env, layer, NULL);
  int n = 0;
  int rc = ps_model_tap((ps_model*)(intptr_t)m, k, what, NULL, 0, &n);
  jfloatArray out = NULL;
  if (rc == PS_OK) {
    out = (*env)->NewFloatArray(env, n);
    jfloat* p = (*env)->GetFloatArrayElements(env, out, NULL);
    rc = ps_model_tap((ps_model*)(intptr_t)m, k, what, p, n, &n);
    (*env)->ReleaseFloatArrayElements(env, out, p, 0);
  }
  (*env)->ReleaseStringUTFChars(env, layer, k);
  return rc == PS_OK ? out : NULL;
}
JNIEXPORT jboolean JNICALL Java_nativeps_PsNative_modelSkippedBackward(JNIEnv* env, jclass c, jlong m) {
  int v = 0;
  ps_model_skipped_backward((ps_model*)(intptr_t)m, &v);
  return v ? JNI_TRUE : JNI_FALSE;
}
/* net/Router.java:5 for the native store: the shard of a key, -1 when every shard holds it */
JNIEXPORT jint JNICALL Java_nativeps_PsNative_keyOwner(JNIEnv* env, jclass c, jstring key, jint nShards) {
  const char* k = (*env)->GetStringUTFChars(env, key, NULL);
  int owner = -1;
  const int rc = ps_key_owner(k, nShards, &owner);
  (*env)->ReleaseStringUTFChars(env, key, k);
  throw_ps(env, rc);
  return owner;
}
JNIEXPORT jfloatArray JNICALL Java_nativeps_PsNative_updaterParse(JNIEnv* env, jclass c, jstring name) {
  const char* k = (*env)->GetStringUTFChars(env, name, NULL);
  ps_updater_spec spec;
  int rc = ps_updater_parse(k, &spec);                                          /* "adam@alfa:..@beta1:..@" (T/TestPs.java:26-27) */
  (*env)->ReleaseStringUTFChars(env, name, k);
  if (rc != PS_OK) { throw_ps(env, rc); return NULL; }
  jfloatArray out = (*env)->NewFloatArray(env, 5);
  jfloat* p = (*env)->GetFloatArrayElements(env, out, NULL);
  p[0] = (jfloat)spec.kind;
  for (int i = 0; i < 4; ++i) p[1 + i] = spec.p[i];
  (*env)->ReleaseFloatArrayElements(env, out, p, 0);
  return out;
}
JNIEXPORT jfloatArray JNICALL Java_nativeps_PsNative_modelPredict(JNIEnv* env, jclass c, jlong m, jfloatArray E, jfloatArray X, jfloatArray W,
                                                                  jint N, jint outRows) {
  jsize ne, nw;
  int F = 0, Xn = 0;
  throw_ps(env, ps_model_shape((ps_model*)(intptr_t)m, &F, NULL, &Xn));
  if (N <= 0 || outRows <= 0 || !check_len(env, E, (jlong)N * F, "E shorter than F x N") || !check_len(env, W, (jlong)N * F, "W shorter than F x N") ||
      !check_len(env, X, (jlong)N * Xn, "X shorter than Xn x N")) return NULL;
  int64_t* e = widen_ids(env, E, &ne);
  int64_t* w = widen_ids(env, W, &nw);
  jfloat* x = (*env)->GetFloatArrayElements(env, X, NULL);
  jfloatArray out = (*env)->NewFloatArray(env, N * outRows);
  jfloat* p = (*env)->GetFloatArrayElements(env, out, NULL);
  int rc = ps_model_predict((ps_model*)(intptr_t)m, e, x, w, N, p);             /* PredictThread.call + Trainer.predict (Trainer.java:44-68) */
  (*env)->ReleaseFloatArrayElements(env, out, p, 0);
  (*env)->ReleaseFloatArrayElements(env, X, x, JNI_ABORT);
  free(e); free(w);
  throw_ps(env, rc);
  return out;
}

/* ---- layer.FcLayer standalone (ps_fc_*): jblas column-major float[] is exactly the C ABI's layout, no conversion ---- */
JNIEXPORT void JNICALL Java_nativeps_PsNative_fcForward(JNIEnv* env, jclass c, jlong fc, jfloatArray aPrev, jint N, jfloatArray aOut) {
  jfloat* in = (*env)->GetPrimitiveArrayCritical(env, aPrev, NULL);
  jfloat* out = (*env)->GetPrimitiveArrayCritical(env, aOut, NULL);
  int rc = ps_fc_forward((ps_fc*)(intptr_t)fc, in, N, out);                     /* FcLayer.forward (FcLayer.java:74-91) */
  (*env)->ReleasePrimitiveArrayCritical(env, aOut, out, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, aPrev, in, JNI_ABORT);
  throw_ps(env, rc);
}
JNIEXPORT void JNICALL Java_nativeps_PsNative_fcBackward(JNIEnv* env, jclass c, jlong fc, jfloatArray delta, jint N, jfloatArray deltaPrev) {
  jfloat* d = (*env)->GetPrimitiveArrayCritical(env, delta, NULL);
  jfloat* p = (*env)->GetPrimitiveArrayCritical(env, deltaPrev, NULL);
  int rc = ps_fc_backward((ps_fc*)(intptr_t)fc, d, N, p);                       /* FcLayer.backward (FcLayer.java:93-110) */
  (*env)->ReleasePrimitiveArrayCritical(env, deltaPrev, p, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, delta, d, JNI_ABORT);
  throw_ps(env, rc);
}
JNIEXPORT void JNICALL Java_nativeps_PsNative_fcUpdate(JNIEnv* env, jclass c, jlong fc) { throw_ps(env, ps_fc_update((ps_fc*)(intptr_t)fc)); }
JNIEXPORT jlong JNICALL Java_nativeps_PsNative_fcCreate(JNIEnv* env, jclass c, jlong ctx, jstring name, jint in, jint out, jint act, jfloatArray upd,
                                                        jint maxBatch) {
  const char* k = (*env)->GetStringUTFChars(env, name, NULL);
  ps_updater_spec spec, *sp = NULL;
  if (upd) {
    jfloat* u = (*env)->GetFloatArrayElements(env, upd, NULL);
    spec.kind = (int32_t)u[0]; for (int i = 0; i < 4; ++i) spec.p[i] = u[1 + i];
    (*env)->ReleaseFloatArrayElements(env, upd, u, JNI_ABORT);
    sp = &spec;
  }
  ps_fc* fc = NULL;
  int rc = ps_fc_create((ps_ctx*)(intptr_t)ctx, k, in, out, act, sp, maxBatch, &fc);   /* FcLayer ctor + pullWeights (FcLayer.java:34-51,112-115) */
  (*env)->ReleaseStringUTFChars(env, name, k);
  throw_ps(env, rc);
  return (jlong)(intptr_t)fc;
}
JNIEXPORT void JNICALL Java_nativeps_PsNative_fcDestroy(JNIEnv* env, jclass c, jlong fc) { ps_fc_destroy((ps_fc*)(intptr_t)fc); }
JNIEXPORT jfloatArray JNICALL Java_nativeps_PsNative_fcGet(JNIEnv* env, jclass c, jlong fc, jint which) {
  int n = 0;
  int rc = ps_fc_get((ps_fc*)(intptr_t)fc, which, NULL, 0, &n);
  if (rc != PS_OK) { throw_ps(env, rc); return NULL; }
  jfloatArray out = (*env)->NewFloatArray(env, n);
  jfloat* p = (*env)->GetFloatArrayElements(env, out, NULL);
  rc = ps_fc_get((ps_fc*)(intptr_t)fc, which, p, n, &n);                        /* KVStore.get("<name>.weights" | ".bias") */
  (*env)->ReleaseFloatArrayElements(env, out, p, 0);
  throw_ps(env, rc);
  return out;
}
JNIEXPORT void JNICALL Java_nativeps_PsNative_fcPut(JNIEnv* env, jclass c, jlong fc, jint which, jfloatArray v) {
  jsize n = (*env)->GetArrayLength(env, v);
  jfloat* p = (*env)->GetFloatArrayElements(env, v, NULL);
  int rc = ps_fc_put((ps_fc*)(intptr_t)fc, which, p, n);
  (*env)->ReleaseFloatArrayElements(env, v, p, JNI_ABORT);
  throw_ps(env, rc);
}

/* ---- data.DataSet (ps_reader_*): ids are widened back to floats because CTR.parseFeature's "E"/"W" are FloatMatrix ---- */
JNIEXPORT jint JNICALL Java_nativeps_PsNative_readerNext(JNIEnv* env, jclass c, jlong r, jfloatArray E, jfloatArray X, jfloatArray W, jfloatArray Y) {
  int batch = 0, F = 0, Xn = 0;
  throw_ps(env, ps_reader_shape((ps_reader*)(intptr_t)r, &batch, &F, &Xn));
  /* ps_reader_next writes up to batch rows: every output array must hold a whole batch */
  if (!check_len(env, E, (jlong)batch * F, "E shorter than F x batch") || !check_len(env, W, (jlong)batch * F, "W shorter than F x batch") ||
      !check_len(env, X, (jlong)batch * Xn, "X shorter than Xn x batch") || !check_len(env, Y, batch, "Y shorter than batch")) return 0;
  jsize ne = (jsize)batch * F;
  int64_t* e = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ne > 0 ? ne : 1));
  int64_t* w = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ne > 0 ? ne : 1));
  if (e == NULL || w == NULL) { free(e); free(w); throw_ps(env, PS_ERR_ARG); return 0; }
  jfloat* x = (*env)->GetFloatArrayElements(env, X, NULL);
  jfloat* y = (*env)->GetFloatArrayElements(env, Y, NULL);
  int rows = 0;
  int rc = ps_reader_next((ps_reader*)(intptr_t)r, e, x, w, y, &rows);
  (*env)->ReleaseFloatArrayElements(env, X, x, 0);
  (*env)->ReleaseFloatArrayElements(env, Y, y, 0);
  if (rc == PS_OK && rows > 0) {
    jfloat* ef = (*env)->GetFloatArrayElements(env, E, NULL);
    jfloat* wf = (*env)->GetFloatArrayElements(env, W, NULL);
    for (jsize i = 0; i < (jsize)rows * F; ++i) { ef[i] = (jfloat)e[i]; wf[i] = (jfloat)w[i]; }   /* only the rows delivered; exact: the reader already applied (float) idx */
    (*env)->ReleaseFloatArrayElements(env, E, ef, 0);
    (*env)->ReleaseFloatArrayElements(env, W, wf, 0);
  }
  free(e); free(w);
  throw_ps(env, rc);
  return rows;
}
JNIEXPORT jlong JNICALL Java_nativeps_PsNative_readerOpen(JNIEnv* env, jclass c, jstring path, jint F, jint Xn, jlong wideSize, jint batch, jint offset,
                                                          jint step, jint threads) {
  const char* k = (*env)->GetStringUTFChars(env, path, NULL);
  ps_reader* r = NULL;
  int rc = ps_reader_open(k, F, Xn, wideSize, batch, offset, step, threads, &r);   /* new CTR(new LibsvmParser(), new FileSource(file), batch, thread) */
  (*env)->ReleaseStringUTFChars(env, path, k);
  throw_ps(env, rc);
  return (jlong)(intptr_t)r;
}
JNIEXPORT void JNICALL Java_nativeps_PsNative_readerReset(JNIEnv* env, jclass c, jlong r) { throw_ps(env, ps_reader_reset((ps_reader*)(intptr_t)r)); }
JNIEXPORT void JNICALL Java_nativeps_PsNative_readerClose(JNIEnv* env, jclass c, jlong r) { ps_reader_close((ps_reader*)(intptr_t)r); }

/* ---- Model.train call by call: the loss stays in Java (DNN.java:44-68) ---- */
JNIEXPORT jfloatArray JNICALL Java_nativeps_PsNative_modelForward(JNIEnv* env, jclass c, jlong m, jfloatArray E, jfloatArray X, jfloatArray W, jint N) {
  jsize ne, nw;
  int F = 0, Xn = 0;
  throw_ps(env, ps_model_shape((ps_model*)(intptr_t)m, &F, NULL, &Xn));
  if (N <= 0 || !check_len(env, E, (jlong)N * F, "E shorter than F x N") || !check_len(env, W, (jlong)N * F, "W shorter than F x N") ||
      !check_len(env, X, (jlong)N * Xn, "X shorter than Xn x N")) return NULL;
  int64_t* e = widen_ids(env, E, &ne);
  int64_t* w = widen_ids(env, W, &nw);
  jfloat* x = (*env)->GetFloatArrayElements(env, X, NULL);
  jfloatArray out = (*env)->NewFloatArray(env, N);
  jfloat* p = (*env)->GetFloatArrayElements(env, out, NULL);
  int rc = ps_model_forward((ps_model*)(intptr_t)m, e, x, w, N, p);             /* the forward loop; P = layers.get(last).getA() */
  (*env)->ReleaseFloatArrayElements(env, out, p, 0);
  (*env)->ReleaseFloatArrayElements(env, X, x, JNI_ABORT);
  free(e); free(w);
  throw_ps(env, rc);
  return out;
}
JNIEXPORT void JNICALL Java_nativeps_PsNative_modelBackwardUpdate(JNIEnv* env, jclass c, jlong m, jfloatArray deltaTop, jint N, jfloat loss) {
  if (N <= 0 || !check_len(env, deltaTop, N, "delta shorter than N")) return;
  jfloat* d = (*env)->GetFloatArrayElements(env, deltaTop, NULL);
  int rc = ps_model_backward_update((ps_model*)(intptr_t)m, d, N, loss);        /* the reverse loop + KVStore.update + clear */
  (*env)->ReleaseFloatArrayElements(env, deltaTop, d, JNI_ABORT);
  throw_ps(env, rc);
}

/* ---- PSClient.getList / updateList: one batched native call per list ---- */
#define PS_JNI_LIST_STRIDE 65536
JNIEXPORT jobjectArray JNICALL Java_nativeps_PsNative_modelGetList(JNIEnv* env, jclass c, jlong m, jobjectArray keys) {
  const jsize n = (*env)->GetArrayLength(env, keys);
  int D = 0;
  throw_ps(env, ps_model_shape((ps_model*)(intptr_t)m, NULL, &D, NULL));
  const char** ks = (const char**)malloc(sizeof(char*) * (size_t)(n > 0 ? n : 1));
  int32_t* found = (int32_t*)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
  int stride = D > 0 ? D : 1, all_emb = 1;
  for (jsize i = 0; i < n; ++i) {
    ks[i] = (*env)->GetStringUTFChars(env, (jstring)(*env)->GetObjectArrayElement(env, keys, i), NULL);
    if (strncmp(ks[i], "emF", 3) != 0) all_emb = 0;
  }
  if (!all_emb) stride = PS_JNI_LIST_STRIDE;        /* dense parameters in the list: rows as wide as the widest FcLayer weight may get */
  float* out = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1) * (size_t)stride);
  int rc = ps_model_get_list((ps_model*)(intptr_t)m, ks, n, out, stride, found);
  jobjectArray res = (*env)->NewObjectArray(env, n, (*env)->FindClass(env, "[F"), NULL);
  for (jsize i = 0; i < n; ++i) {
    if (rc == PS_OK && found[i] > 0) {
      jfloatArray row = (*env)->NewFloatArray(env, found[i]);
      (*env)->SetFloatArrayRegion(env, row, 0, found[i], out + (size_t)i * stride);
      (*env)->SetObjectArrayElement(env, res, i, row);
      (*env)->DeleteLocalRef(env, row);
    }
    (*env)->ReleaseStringUTFChars(env, (jstring)(*env)->GetObjectArrayElement(env, keys, i), ks[i]);
  }
  free(out); free(found); free((void*)ks);
  throw_ps(env, rc);
  return res;
}
JNIEXPORT jobjectArray JNICALL Java_nativeps_PsNative_modelUpdateList(JNIEnv* env, jclass c, jlong m, jobjectArray keys, jobjectArray values, jboolean replace) {
  const jsize n = (*env)->GetArrayLength(env, keys);
  if (!check_len(env, values, n, "fewer values than keys")) return NULL;
  const char** ks = (const char**)malloc(sizeof(char*) * (size_t)(n > 0 ? n : 1));
  int32_t* lens = (int32_t*)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
  int stride = 1;
  for (jsize i = 0; i < n; ++i) {
    lens[i] = (*env)->GetArrayLength(env, (jarray)(*env)->GetObjectArrayElement(env, values, i));
    if (lens[i] > stride) stride = lens[i];
  }
  float* io = (float*)calloc((size_t)(n > 0 ? n : 1) * (size_t)stride, sizeof(float));
  for (jsize i = 0; i < n; ++i) {
    ks[i] = (*env)->GetStringUTFChars(env, (jstring)(*env)->GetObjectArrayElement(env, keys, i), NULL);
    jfloatArray v = (jfloatArray)(*env)->GetObjectArrayElement(env, values, i);
    jfloat* p = (*env)->GetFloatArrayElements(env, v, NULL);
    memcpy(io + (size_t)i * stride, p, sizeof(float) * (size_t)lens[i]);
    (*env)->ReleaseFloatArrayElements(env, v, p, JNI_ABORT);
  }
  int rc = ps_model_update_list((ps_model*)(intptr_t)m, ks, n, io, stride, lens, replace ? 1 : 0);   /* PServer.upsertList: insert-if-absent unless replace */
  jobjectArray res = (*env)->NewObjectArray(env, n, (*env)->FindClass(env, "[F"), NULL);
  for (jsize i = 0; i < n; ++i) {
    jfloatArray row = (*env)->NewFloatArray(env, lens[i]);
    (*env)->SetFloatArrayRegion(env, row, 0, lens[i], io + (size_t)i * stride);
    (*env)->SetObjectArrayElement(env, res, i, row);
    (*env)->DeleteLocalRef(env, row);
    (*env)->ReleaseStringUTFChars(env, (jstring)(*env)->GetObjectArrayElement(env, keys, i), ks[i]);
  }
  free(io); free(lens); free((void*)ks);
  throw_ps(env, rc);
  return res;
}

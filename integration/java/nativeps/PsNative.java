package nativeps;

/**
 * Thin JNI binding of libps_b200.so (include/ps_b200.h).  One static native method per C entry
 * point; handles are opaque longs.  SOURCE ONLY in this repository: no JDK exists in the build
 * image, so these files are not compiled or run here (INTEGRATION.md, DESIGN.md §1).
 */
public final class PsNative {
	static { System.loadLibrary("ps_b200_jni"); }   // integration/jni/ps_jni.c, links libps_b200.so
	private PsNative() {}

	public static native String lastError();
	public static native long ctxCreate(int device, long seed);                       // ps_ctx_create
	public static native void ctxDestroy(long ctx);
	public static native void ctxSetFcPrecision(long ctx, int mode);                   // 0 fp32 | 1 tf32 tcgen05
	public static native float[] updaterParse(String name);                            // {kind, p0..p3}
	public static native int keyOwner(String key, int nShards);                        // ps_key_owner: net/Router.java:5 for the native store; -1 = replicated
	public static native long modelCreate(long ctx, int kind, int F, int D, int Xn, int[] fc, long embCapacity, float[] embUpdater, int maxBatch);
	public static native void modelDestroy(long model);
	/** E, W: F x N ids carried as floats exactly as CTR.parseFeature builds them (CTR.java:47-68). */
	public static native float modelTrainStep(long model, float[] E, float[] X, float[] W, float[] Y, int N);
	public static native float[] modelPredict(long model, float[] E, float[] X, float[] W, int N, int outRows);
	/** the forward loop of DNN.train / WideDeepNN.train (DNN.java:44-46): returns P (N floats); the batch stays pending on the device */
	public static native float[] modelForward(long model, float[] E, float[] X, float[] W, int N);
	/** the reverse loop + KVStore.update + clear given deltaTop = loss.backward(P, Y) (DNN.java:49,64-68; Trainer.java:93,95) */
	public static native void modelBackwardUpdate(long model, float[] deltaTop, int N, float loss);
	/** PSClient.getList / updateList: one batched call; row i is null when key i is absent / the winning value of key i */
	public static native float[][] modelGetList(long model, String[] keys);
	public static native float[][] modelUpdateList(long model, String[] keys, float[][] values, boolean replace);
	public static native float[] modelGet(long model, String key);                     // null when absent (KVStore.get)
	public static native void modelPut(long model, String key, float[] value);
	public static native boolean modelPush(long model, String key, float[] gradient, String updaterKey);   // ps_model_push: PServer.push; false = no such key
	public static native void ctxMakeCurrent(long ctx);                                 // ps_ctx_make_current: first call of any other thread
	public static native float[] modelTap(long model, String layer, int what);        // 0 = A, 1 = delta
	public static native boolean modelSkippedBackward(long model);

	// layer.FcLayer as a standalone operator (ps_fc_*): for models that walk their own layer list (layer/StandaloneFcLayer.java)
	public static native long fcCreate(long ctx, String name, int in, int out, int act, float[] updater, int maxBatch);
	public static native void fcDestroy(long fc);
	public static native void fcForward(long fc, float[] aPrev, int N, float[] aOut);         // FcLayer.forward  (FcLayer.java:74-91)
	public static native void fcBackward(long fc, float[] delta, int N, float[] deltaPrev);   // FcLayer.backward (FcLayer.java:93-110)
	public static native void fcUpdate(long fc);                                              // KVStore.update + clear for its two keys
	public static native float[] fcGet(long fc, int which);                                   // 0 weights (out x in), 1 bias
	public static native void fcPut(long fc, int which, float[] value);

	// data.DataSet over data.FileSource with LibsvmParser + CTR.parseFeature (ps_reader_*): data/NativeCtrDataSet.java
	public static native long readerOpen(String path, int F, int Xn, long wideSize, int batch, int offset, int step, int threads);
	/** fills E, W (F x rows ids as floats, like CTR.parseFeature), X (Xn x rows), Y (rows); returns rows, 0 at end of data */
	public static native int readerNext(long reader, float[] E, float[] X, float[] W, float[] Y);
	public static native void readerReset(long reader);
	public static native void readerClose(long reader);
}

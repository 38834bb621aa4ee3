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
	public static native long modelCreate(long ctx, int kind, int F, int D, int Xn, int[] fc, long embCapacity, float[] embUpdater, int maxBatch);
	public static native void modelDestroy(long model);
	/** E, W: F x N ids carried as floats exactly as CTR.parseFeature builds them (CTR.java:47-68). */
	public static native float modelTrainStep(long model, float[] E, float[] X, float[] W, float[] Y, int N);
	public static native float[] modelPredict(long model, float[] E, float[] X, float[] W, int N, int outRows);
	public static native float[] modelGet(long model, String key);                     // null when absent (KVStore.get)
	public static native void modelPut(long model, String key, float[] value);
	public static native float[] modelTap(long model, String layer, int what);        // 0 = A, 1 = delta
	public static native boolean modelSkippedBackward(long model);
}

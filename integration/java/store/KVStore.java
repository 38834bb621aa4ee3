package store;

import nativeps.PsNative;
import org.jblas.FloatMatrix;
import update.Updater;

import java.util.Map;
import java.util.concurrent.Callable;

/**
 * Drop-in for store/KVStore.java of the reference: same package, class and public methods
 * (ins :70, get :129/:136, put :161, sum :192, update :202/:220/:240, clear :270, asyncGet :109,
 * asyncWait :113), backed by the GPU-resident store of libps_b200.so instead of Java maps.
 *
 * Differences a caller can observe (DESIGN.md §5.7): get() returns a host SNAPSHOT, not the live
 * matrix; lazily created weights come from the seeded initialiser of ps_spec.h; sum() / update() / clear() are no-ops
 * because ps_model_backward_update — started by the last layer's backward() — applies KVStore.sum + update + clear itself,
 * fused into the backward kernels.  FC weights come back as out x in matrices (column-major, like the reference's).
 */
public class KVStore {
	private static final KVStore ins = new KVStore();
	public static KVStore ins() { return ins; }

	private long ctx, model;                       // opaque native handles, bound by GpuStep.bind()
	public void bind(long ctx, long model) { this.ctx = ctx; this.model = model; }
	public long model() { return model; }
	public long nativeCtx() { return ctx; }

	// layer-at-a-time mode (layer/StandaloneFcLayer.java): dense layers whose pending KVStore.sum lives on the device
	private final java.util.Map<String, Long> dense = new java.util.LinkedHashMap<String, Long>();
	public synchronized void registerDense(String name, long fcHandle) { dense.put(name, fcHandle); }

	public FloatMatrix get(String key) {           // KVStore.java:129-134
		float[] v = PsNative.modelGet(model, key);
		return v == null ? null : shaped(key, v);
	}
	public synchronized FloatMatrix get(String key, Callable<FloatMatrix> init) {   // KVStore.java:136-159
		FloatMatrix m = get(key);
		if (m != null) return m;
		try { m = init.call(); } catch (Exception e) { e.printStackTrace(); return null; }
		put(key, m);
		return m;
	}
	public void put(String key, FloatMatrix m) { PsNative.modelPut(model, key, m.data); }   // KVStore.java:161-166
	public synchronized void sum(String key, FloatMatrix val) { /* gradients are accumulated on the device by the native step */ }
	public void update(Updater updater, String key) {}
	public void update(Updater updater) {}
	public void update(Map<String, Updater> updaters) {                                    // KVStore.java:240-268
		for (long fc : dense.values()) PsNative.fcUpdate(fc);                                // whole-step mode: nothing registered, applied inside modelTrainStep
	}
	public void clear() {}
	public void asyncGet(String key, Callable<FloatMatrix> init) {}                        // the probe kernel is the batched prefetch
	public void asyncWait() {}

	private final java.util.Map<String, int[]> shapes = new java.util.HashMap<String, int[]>();
	/** layer.FcLayer registers out x in for "fc<i>.weights" so that get() hands back the reference's shape */
	public void shape(String key, int rows, int cols) { shapes.put(key, new int[]{rows, cols}); }
	public int[] shapeOf(String key) { return shapes.get(key); }                           // net/PServer.java: rows x cols on the wire
	public void rememberShape(String key, int rows, int cols) { shape(key, rows, cols); }
	private FloatMatrix shaped(String key, float[] v) {
		int[] s = shapes.get(key);
		return s != null && s[0] * s[1] == v.length ? new FloatMatrix(s[0], s[1], v) : new FloatMatrix(v.length, 1, v);
	}
}

package layer;

import nativeps.PsNative;
import org.jblas.FloatMatrix;
import store.KVStore;

import java.util.Map;

/**
 * Runs ONE native Trainer step the first time a layer of the model is asked to go forward, then
 * serves every later forward()/backward() of that step from the taps.  model.DNN / WideDeepNN /
 * FullConnectedNN and train.Trainer stay UNCHANGED: they still walk their layer list, compute the
 * loss on P with the reference's own loss class and call KVStore.update/clear (now no-ops).
 */
public final class GpuStep {
	private static final ThreadLocal<GpuStep> cur = ThreadLocal.withInitial(GpuStep::new);
	public static GpuStep current() { return cur.get(); }

	private boolean ran;
	private float loss;
	public void begin() { ran = false; }           // called by InputLayer.setA / Model.pullWeights

	public void ensureRan(FloatMatrix E, FloatMatrix X, FloatMatrix W, FloatMatrix Y) {
		if (ran) return;
		loss = PsNative.modelTrainStep(KVStore.ins().model(), E == null ? null : E.data, X.data, W == null ? null : W.data, Y.data, X.columns);
		ran = true;
	}
	public float loss() { return loss; }
	public FloatMatrix A(String layer, int rows, int cols) { return new FloatMatrix(rows, cols, PsNative.modelTap(KVStore.ins().model(), layer, 0)); }
	public FloatMatrix delta(String layer, int rows, int cols) { return new FloatMatrix(rows, cols, PsNative.modelTap(KVStore.ins().model(), layer, 1)); }
}

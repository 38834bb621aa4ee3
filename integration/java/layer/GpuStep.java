package layer;

import nativeps.PsNative;
import org.jblas.FloatMatrix;
import store.KVStore;

/**
 * The glue that lets model.DNN / model.WideDeepNN and train.Trainer stay UNCHANGED (DNN.java:35-70, WideDeepNN.java:39-83):
 * they still walk their layer list forwards, compute the loss on P with the reference's own loss class, call setDelta on
 * the last layer, walk the list backwards and call KVStore.update / clear.  The drop-in layers turn that walk into TWO
 * native calls: the first forward() of a step runs ps_model_forward (no label crosses the boundary), the first backward()
 * runs ps_model_backward_update with the delta Java's loss.backward produced.  Everything in between is a read.
 * SOURCE ONLY: no JDK in the build image.
 */
public final class GpuStep {
	private static final ThreadLocal<GpuStep> cur = ThreadLocal.withInitial(GpuStep::new);
	public static GpuStep current() { return cur.get(); }

	private boolean forwardRan, backwardRan;
	private float[] P;
	private int N;
	public void begin() { forwardRan = false; backwardRan = false; }        // Model.pullWeights: first call of TrainerThread.call

	/** the forward loop: E (F x N float-carried ids), X (Xn x N), W (F x N or null) exactly as CTR.parseFeature builds them */
	public void ensureForward(FloatMatrix E, FloatMatrix X, FloatMatrix W) {
		if (forwardRan) return;
		N = X.columns;
		P = PsNative.modelForward(KVStore.ins().model(), E.data, X.data, W == null ? null : W.data, N);
		forwardRan = true; backwardRan = false;
	}
	/** the reverse loop + KVStore.update: delta = loss.backward(P, Y) as the model handed it to setDelta (DNN.java:64) */
	public void ensureBackward(FloatMatrix deltaTop) {
		if (backwardRan || !forwardRan) return;
		PsNative.modelBackwardUpdate(KVStore.ins().model(), deltaTop.data, N, Float.NaN);
		backwardRan = true; forwardRan = false;
	}
	public FloatMatrix P() { return new FloatMatrix(1, N, P); }
	public FloatMatrix A(String layer, int rows, int cols) { return new FloatMatrix(rows, cols, PsNative.modelTap(KVStore.ins().model(), layer, 0)); }
	public FloatMatrix delta(String layer, int rows, int cols) { return new FloatMatrix(rows, cols, PsNative.modelTap(KVStore.ins().model(), layer, 1)); }
}

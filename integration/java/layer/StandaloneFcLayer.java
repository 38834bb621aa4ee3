package layer;

import activations.Activation;
import activations.Relu;
import activations.Sigmoid;
import nativeps.PsNative;
import org.jblas.FloatMatrix;
import store.KVStore;

/**
 * layer/FcLayer.java with every matrix product on the GPU but the layer list still walked by the reference's own
 * Model.train loops (DNN.java:44-68): forward / backward go through ps_fc_forward / ps_fc_backward one layer at a time,
 * and Trainer.train's KVStore.update(updaters) reaches ps_fc_update through KVStore.registerDense.  Use this instead of
 * the whole-step proxy (layer/FcLayer.java + GpuStep) when a model mixes native and Java layers.
 * SOURCE ONLY: no JDK in the build image.
 */
public class StandaloneFcLayer extends Layer {
	protected Activation activation;
	private long handle;
	private final int maxBatch;

	public StandaloneFcLayer(String name, int inputDims, int outputDims, int maxBatch) {
		super(name, inputDims, outputDims);
		this.maxBatch = maxBatch;
	}
	public void setActivation(Activation a) { this.activation = a; }

	private long handle() {
		if (handle == 0) {
			int act = activation == null ? 0 : activation instanceof Relu ? 1 : activation instanceof Sigmoid ? 2 : -1;
			if (act < 0) throw new RuntimeException("unsupported activation for the native FcLayer");
			handle = PsNative.fcCreate(KVStore.ins().nativeCtx(), name, inputDims, outputDims, act, null, maxBatch);
			KVStore.ins().registerDense(name, handle);       // KVStore.update(...) then calls PsNative.fcUpdate(handle)
		}
		return handle;
	}

	public FloatMatrix forward() {                 // FcLayer.java:74-91
		FloatMatrix in = pre.getA();
		FloatMatrix out = new FloatMatrix(outputDims, in.columns);
		PsNative.fcForward(handle(), in.data, in.columns, out.data);   // jblas column-major == the C ABI's layout
		this.A = out;
		return out;
	}
	public FloatMatrix backward() {                // FcLayer.java:93-110
		FloatMatrix d = next == null ? this.delta : next.getDelta();
		FloatMatrix prev = new FloatMatrix(inputDims, d.columns);
		PsNative.fcBackward(handle(), d.data, d.columns, prev.data);
		this.delta = prev;
		return prev;
	}
	public void pullWeights() { handle(); }        // FcLayer.java:112-115: weights are created on first use inside the native store
}

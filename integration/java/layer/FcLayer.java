package layer;

import activations.Activation;
import activations.Relu;
import activations.Sigmoid;
import com.google.common.collect.Lists;
import org.jblas.FloatMatrix;

import java.util.List;

/**
 * Drop-in for layer/FcLayer.java (ctor :34, build :53, forward :74, backward :93, pullWeights :112): same signatures incl. the
 * static build(int, int[]) that DNN.buildModel / WideDeepNN.buildModel call (DNN.java:108, WideDeepNN.java:125) and
 * setActivation(null) (WideDeepNN.java:128).  The arithmetic (W·A + b, activation, dW = δ·Aᵀ/N, db = rowMeans(δ),
 * δ_prev = Wᵀ·δ) runs in gemm_tf32_kernel (tcgen05) inside the native forward / reverse loops.
 */
public class FcLayer extends Layer {
	protected Activation activation;
	public FcLayer(String name, int inputDims, int outputDims) {
		super(name, inputDims, outputDims);
		store.KVStore.ins().shape(name + ".weights", outputDims, inputDims);
	}
	public void setActivation(Activation a) { this.activation = a; }

	public static List<Layer> build(int inputSize, int[] fcLayers) {       // FcLayer.java:53-70, verbatim wiring
		List<Layer> result = Lists.newArrayList();
		for (int i = 0; i < fcLayers.length; i++) {
			FcLayer fc = new FcLayer("fc" + i, inputSize, fcLayers[i]);
			fc.setActivation(i == fcLayers.length - 1 ? new Sigmoid() : new Relu());
			result.add(fc);
			if (i != 0) result.get(i - 1).setNext(fc);
			inputSize = fcLayers[i];
		}
		return result;
	}
	public void clear() {}
	public FloatMatrix forward() {                 // FcLayer.java:74-91; DNN: the last FcLayer's A is P (DNN.java:47)
		this.A = next == null ? GpuStep.current().P() : null;
		return this.A;
	}
	public FloatMatrix backward() {                // FcLayer.java:93-110; DNN: the first call of the reverse loop carries the model's delta
		if (next == null) GpuStep.current().ensureBackward(this.delta);
		return this.delta;
	}
	public void pullWeights() {}                   // weights live in the GPU store; KVStore.get(name + ".weights") snapshots them
	/** Layer.getA() / getDelta() of an inner layer, on request (UI plots, LossSurface): read back from the device */
	public FloatMatrix tapA(int n) { return GpuStep.current().A(name, outputDims, n); }
	public FloatMatrix tapDelta(int n) { return GpuStep.current().delta(name, inputDims, n); }
}

package layer;

import activations.Activation;
import org.jblas.FloatMatrix;

/**
 * Drop-in for layer/FcLayer.java (ctor :34, build :53, forward :74, backward :93, pullWeights :112):
 * same signatures; the arithmetic (W·A + b, activation, dW = δ·Aᵀ/N, db = rowMeans(δ), δ_prev = Wᵀ·δ)
 * runs in gemm_tf32_kernel / gemm_simt_kernel inside the native step.
 */
public class FcLayer extends Layer {
	protected Activation activation;
	public FcLayer(String name, int inputDims, int outputDims) { super(name, inputDims, outputDims); }
	public void setActivation(Activation a) { this.activation = a; }

	public FloatMatrix forward() {                 // FcLayer.java:74-91
		int n = pre.getA().columns;
		this.A = GpuStep.current().A(name, outputDims, n);
		return this.A;
	}
	public FloatMatrix backward() {                // FcLayer.java:93-110
		int n = A.columns;
		this.delta = GpuStep.current().delta(name, inputDims, n);
		return this.delta;
	}
	public void pullWeights() {}                   // weights live in the GPU store; KVStore.get(name + ".weights") snapshots them
}

package layer;

import activations.Activation;
import org.jblas.FloatMatrix;

/**
 * Drop-in for layer/LRLayer.java (ctor :37, clear :56, forward :62, backward :100, pullWeights :122) — the wide branch
 * of WideDeepNN.  The reference calls KVStore.get("wide.weights." + id, init) once per (field, sample) and pushes the
 * batch-mean delta to every key it has ever seen (LRLayer.java:79,110-117); here both happen inside the native step
 * (wide_forward_kernel, wide_update_all_kernel) and this class only hands out the tapped results.
 * SOURCE ONLY: no JDK in the build image.
 */
public class LRLayer extends Layer {
	protected Activation activation;
	public LRLayer(String name, int inputDims) { super(name, inputDims, 1); }
	public void setActivation(Activation a) { this.activation = a; }
	public void clear() {}                         // LRLayer.java:56-60 is never called by the reference either

	public FloatMatrix forward() { this.A = null; return null; }   // LRLayer.java:62-98: z_i = b + sum_j w[W[j,i]] ran inside the native forward loop
	public FloatMatrix backward() { return null; }                  // LRLayer.java:100-120: the pushes happen inside the native reverse loop
	public FloatMatrix tapA(int n) { return GpuStep.current().A("wide", 1, n); }
	public void pullWeights() {}                   // LRLayer.java:122-124: weights live in the GPU wide table
}

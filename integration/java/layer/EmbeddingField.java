package layer;

import activations.Activation;
import nativeps.PsNative;
import org.jblas.FloatMatrix;
import store.KVStore;

/**
 * Drop-in for layer/EmbeddingField.java (ctor :35, preForward :57, forward :66, clear :80, backward :86) for code that
 * holds EmbeddingField objects directly (EmbeddingLayer.setEmbeddingFields, EmbeddingLayer.java:77).  One field = one
 * column block of the GPU table's gather: forward(ids) returns the D x N rows of this field with ReLU applied
 * (EmbeddingField.java:73-76); backward is a no-op because the native step's scatter/update kernels have already consumed
 * the delta of every field at once.
 * SOURCE ONLY: no JDK in the build image.
 */
public class EmbeddingField {
	protected String name;
	protected int inputDims, outputDims;
	protected Activation activation;
	private final int field;

	public EmbeddingField(String name, int inputDims, int outputDims) {
		this.name = name; this.inputDims = inputDims; this.outputDims = outputDims;
		this.field = Integer.parseInt(name.replaceAll("[^0-9]", ""));          // "emF<j>" (EmbeddingLayer.java:52)
	}
	public void setActivation(Activation a) { this.activation = a; }
	public void preForward(float[] nSample) {}     // EmbeddingField.java:57-64: the probe kernel is the batched prefetch

	public FloatMatrix forward(float[] nSample) {  // EmbeddingField.java:66-78
		int n = nSample.length;
		float[] all = PsNative.modelTap(KVStore.ins().model(), "embedding", 0);   // (F*D) x N, column-major
		int rows = all.length / n;
		FloatMatrix out = new FloatMatrix(outputDims, n);
		for (int i = 0; i < n; i++) System.arraycopy(all, i * rows + field * outputDims, out.data, i * outputDims, outputDims);
		return out;
	}
	public void clear() {}
	public void backward(int offset, FloatMatrix delta) {}                      // EmbeddingField.java:86-104
}

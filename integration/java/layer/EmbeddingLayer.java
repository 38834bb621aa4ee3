package layer;

import org.jblas.FloatMatrix;

/**
 * Drop-in for layer/EmbeddingLayer.java (ctor :21, forward :25, build :50, backward :59,
 * pullWeights :71).  forward() is where the native step is kicked: the embedding layer is the
 * first layer of DNN / WideDeepNN (DNN.java:119, WideDeepNN.java:149).
 */
public class EmbeddingLayer extends Layer {
	private int fields, dim;
	private InputLayer number, wide, label;        // wired by the model factory next to `pre` (the category input)
	public EmbeddingLayer(String name, int inputDims, int outputDims) { super(name, inputDims, outputDims); }
	public EmbeddingLayer build(int embeddingFieldNum, int embeddingSize) { fields = embeddingFieldNum; dim = embeddingSize; return this; }
	public void inputs(InputLayer number, InputLayer wide, InputLayer label) { this.number = number; this.wide = wide; this.label = label; }

	public FloatMatrix forward() {                 // EmbeddingLayer.java:25-48 → emb_probe_kernel + emb_gather_kernel
		FloatMatrix E = pre.getA();
		GpuStep.current().ensureRan(E, number.getA(), wide == null ? null : wide.getA(), label.getA());
		this.A = GpuStep.current().A("embedding", fields * dim, E.columns);
		return this.A;
	}
	public FloatMatrix backward() {                // EmbeddingLayer.java:59-69 (called twice per step; both are reads here)
		this.delta = next.getDelta();
		return this.delta;
	}
	public void pullWeights() { GpuStep.current().begin(); }   // EmbeddingLayer.java:71-75 — first call of TrainerThread.call
}

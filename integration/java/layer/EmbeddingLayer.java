package layer;

import org.jblas.FloatMatrix;

/**
 * Drop-in for layer/EmbeddingLayer.java (ctor :21, forward :25, build :50, backward :59, pullWeights :71).  It is the first
 * entry of `layers` in DNN.buildModel / WideDeepNN.buildModel (DNN.java:119, WideDeepNN.java:149), so its forward() is where
 * the native forward loop is kicked.  The three inputs are found through the layer graph the UNCHANGED buildModel wires
 * (WideDeepNN.java:140-147): pre = the category InputLayer; next = the ConcatLayer, whose second input is the number
 * InputLayer; the chain of next pointers ends in the AddLayer whose right operand is the LRLayer fed by the wide InputLayer.
 */
public class EmbeddingLayer extends Layer {
	private int fields, dim;
	public EmbeddingLayer(String name, int inputDims, int outputDims) { super(name, inputDims, outputDims); }
	public EmbeddingLayer build(int embeddingFieldNum, int embeddingSize) { fields = embeddingFieldNum; dim = embeddingSize; return this; }
	public void setEmbeddingFields(java.util.List<EmbeddingField> f) {}    // EmbeddingLayer.java:77: the fields live in the GPU table

	public FloatMatrix forward() {                 // EmbeddingLayer.java:25-48 -> emb_lookup_kernel (resolve + gather in one kernel)
		FloatMatrix E = pre.getA();
		FloatMatrix X = ((ConcatLayer) next).numberInput().getA();
		FloatMatrix W = null;
		for (Layer l = next; l != null; l = l.getNext())
			if (l instanceof AddLayer) { W = ((AddLayer) l).right().getPre().getA(); break; }
		GpuStep.current().ensureForward(E, X, W);
		this.A = null;                             // (F*D) x N stays on the device; GpuStep.A("embedding", ..) reads it back on request
		return this.A;
	}
	public FloatMatrix backward() { return null; } // EmbeddingLayer.java:59-69 (called twice per step): already applied by the native reverse loop
	public void pullWeights() { GpuStep.current().begin(); }   // EmbeddingLayer.java:71-75 — first call of TrainerThread.call
}

package layer;

import org.jblas.FloatMatrix;

import java.util.List;

/**
 * Drop-in for layer/ConcatLayer.java (ctors :14/:26, forward :30, backward :39): the embedding rows and the numeric features
 * are written side by side by the lookup kernel itself (act[0] = [emb | X | 1]); nothing is concatenated on the host.
 * backward() keeps the reference's behaviour of calling every input's backward() (ConcatLayer.java:41-46), which is how
 * EmbeddingLayer.backward gets its second call per step.
 */
public class ConcatLayer extends Layer {
	private List<Layer> inputs;
	public ConcatLayer(String name, List<Layer> layers) {
		int out = 0;
		for (Layer l : layers) out += l.getOutputDims();
		this.inputs = layers; this.name = name; this.inputDims = out; this.outputDims = out;
	}
	public ConcatLayer(String name, int inputDims, int outputDims) { super(name, inputDims, outputDims); }
	Layer numberInput() { return inputs.get(1); }
	public FloatMatrix forward() { this.A = null; return null; }
	public FloatMatrix backward() { for (Layer l : inputs) l.backward(); return null; }
	public void pullWeights() {}
}

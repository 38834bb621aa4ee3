package layer;

import activations.Activation;
import org.jblas.FloatMatrix;

/**
 * Drop-in for layer/AddLayer.java (ctor :25, forward :33, backward :49): deep + wide, then the output Sigmoid.  It is the LAST
 * layer of WideDeepNN (WideDeepNN.java:153): its getA() is the P the model hands to loss.forward / loss.backward, and its
 * backward() — the first call of the reverse loop — receives the model's delta through setDelta (WideDeepNN.java:76) and
 * starts the native reverse loop with it.
 */
public class AddLayer extends Layer {
	private final Layer left, right;
	protected Activation activation;
	public AddLayer(String name, Layer left, Layer right) { super(name, 0, 0); this.left = left; this.right = right; }
	Layer right() { return right; }
	public void setActivation(Activation a) { this.activation = a; }
	public void clear() {}
	public FloatMatrix forward() { this.A = GpuStep.current().P(); return this.A; }
	public FloatMatrix backward() { GpuStep.current().ensureBackward(next == null ? this.delta : next.getDelta()); return this.delta; }
	public void pullWeights() {}
}

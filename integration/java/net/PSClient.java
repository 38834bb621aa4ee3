package net;

import nativeps.PsNative;
import org.jblas.FloatMatrix;
import store.KVStore;

import java.util.ArrayList;
import java.util.HashMap;
import java.util.List;
import java.util.Map;

/**
 * Drop-in for net/PSClient.java (ctors :31/:35, close :42, get :47, getList :72, update :100, updateList :128, push :154,
 * barrier :177): the same public methods, served by the GPU-resident store of this process instead of a gRPC PServer.
 * getList / updateList are ONE batched native call each (ps_model_get_list / ps_model_update_list — embedding keys go through
 * one lookup / insert kernel).  push / barrier are no-ops: a worker's gradients never leave the device — KVStore.update's
 * per-key client.push (KVStore.java:257-260) and the server-side sum + psUpdate (PServer.java:164-214) are what
 * ps_model_backward_update (one GPU) or ps_model_p2p_submit (key-hash sharded over the GPUs of the box) do in their kernels.
 * SOURCE ONLY: no JDK in the build image.
 */
public class PSClient {
	static final int STRIDE = 1 << 16;             // floats reserved per listed key on the wire to the native call (embedding rows use D of them)
	final long model;                              // 0: the process-wide store (KVStore.ins()); else one shard's native model (PSRouterClient)
	public PSClient() { this.model = 0; }
	public PSClient(String host, int port) { this.model = 0; }
	public PSClient(long model) { this.model = model; }
	public void close() {}
	long handle() { return model != 0 ? model : KVStore.ins().model(); }

	public FloatMatrix get(String key) {
		if (model == 0) return KVStore.ins().get(key);
		float[] v = PsNative.modelGet(model, key);
		return v == null ? null : new FloatMatrix(v.length, 1, v);
	}
	public Map<String, FloatMatrix> getList(List<String> keys) {
		Map<String, FloatMatrix> out = new HashMap<String, FloatMatrix>();
		float[][] rows = PsNative.modelGetList(handle(), keys.toArray(new String[0]));
		for (int i = 0; i < keys.size(); i++) out.put(keys.get(i), rows[i] == null ? null : new FloatMatrix(rows[i].length, 1, rows[i]));
		return out;
	}
	public FloatMatrix update(String key, FloatMatrix weights, boolean replace) {
		Map<String, FloatMatrix> one = new HashMap<String, FloatMatrix>();
		one.put(key, weights);
		return updateList(one, replace).get(key);
	}
	public Map<String, FloatMatrix> updateList(Map<String, FloatMatrix> updates, boolean replace) {
		List<String> keys = new ArrayList<String>(updates.keySet());
		float[][] offered = new float[keys.size()][];
		for (int i = 0; i < keys.size(); i++) offered[i] = updates.get(keys.get(i)).data;
		float[][] winners = PsNative.modelUpdateList(handle(), keys.toArray(new String[0]), offered, replace);
		Map<String, FloatMatrix> out = new HashMap<String, FloatMatrix>();
		for (int i = 0; i < keys.size(); i++) {
			FloatMatrix o = updates.get(keys.get(i));
			out.put(keys.get(i), new FloatMatrix(o.rows, o.columns, winners[i]));
		}
		return out;
	}
	public void push(String key, FloatMatrix gradient, String updaterKey, boolean async) {}
	public void barrier() {}
}

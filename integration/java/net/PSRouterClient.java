package net;

import org.jblas.FloatMatrix;

import java.util.ArrayList;
import java.util.HashMap;
import java.util.List;
import java.util.Map;

/**
 * Drop-in for net/PSRouterClient.java (:33 ctor, :55 get, :60 getList, :88 update, :93 updateList, :125 push, :131 barrier) for a JVM
 * that holds one native model per GPU of the box (a checkpoint / inspection tool, or a single worker process driving several GPUs).
 * Host-side list access follows the reference's pattern — bucket the keys by router.shard(key), ONE batched call per shard
 * (PSClient.getList → ps_model_get_list: one lookup kernel), merge (PSRouterClient.java:60-85, 93-122) — with the shards being the
 * GPUs' tables instead of gRPC servers.  Replicated keys (dense, wide) are read from shard 0 and written to every shard.
 * Training does not come through here: the route / lookup / push of a step run on the devices (ps_model_p2p_submit, p2p.cu), which
 * is why push and barrier have nothing left to do.  With the no-argument constructor (one process per GPU, the "-Dmode=dist" worker)
 * the client serves the local shard only, like PSClient.
 * SOURCE ONLY: no JDK in the build image.
 */
public class PSRouterClient extends PSClient {
	final List<PSClient> clients = new ArrayList<PSClient>();
	final Router router;

	public PSRouterClient(Router router, long[] shardModels) {
		super();
		this.router = router;
		for (long m : shardModels) clients.add(new PSClient(m));
	}
	public PSRouterClient(long[] shardModels) { this(new NativeRouter(shardModels.length), shardModels); }
	public PSRouterClient(Router router) { super(); this.router = router; }
	public PSRouterClient() { super(); this.router = null; }

	boolean local() { return clients.isEmpty(); }
	boolean replicated(String key) { return router instanceof NativeRouter && ((NativeRouter) router).replicated(key); }

	@Override public void close() { for (PSClient c : clients) c.close(); }

	@Override public FloatMatrix get(String key) {
		if (local()) return super.get(key);
		return clients.get(router.shard(key)).get(key);
	}

	@Override public Map<String, FloatMatrix> getList(List<String> keys) {
		if (local()) return super.getList(keys);
		List<List<String>> buckets = new ArrayList<List<String>>();
		for (int s = 0; s < clients.size(); s++) buckets.add(new ArrayList<String>());
		for (String k : keys) buckets.get(router.shard(k)).add(k);
		Map<String, FloatMatrix> merged = new HashMap<String, FloatMatrix>();
		for (int s = 0; s < clients.size(); s++)
			if (!buckets.get(s).isEmpty()) merged.putAll(clients.get(s).getList(buckets.get(s)));
		return merged;
	}

	@Override public FloatMatrix update(String key, FloatMatrix weights, boolean replace) {
		Map<String, FloatMatrix> one = new HashMap<String, FloatMatrix>();
		one.put(key, weights);
		return updateList(one, replace).get(key);
	}

	@Override public Map<String, FloatMatrix> updateList(Map<String, FloatMatrix> updates, boolean replace) {
		if (local()) return super.updateList(updates, replace);
		List<Map<String, FloatMatrix>> buckets = new ArrayList<Map<String, FloatMatrix>>();
		for (int s = 0; s < clients.size(); s++) buckets.add(new HashMap<String, FloatMatrix>());
		for (Map.Entry<String, FloatMatrix> e : updates.entrySet()) {
			if (replicated(e.getKey())) for (Map<String, FloatMatrix> b : buckets) b.put(e.getKey(), e.getValue());
			else buckets.get(router.shard(e.getKey())).put(e.getKey(), e.getValue());
		}
		Map<String, FloatMatrix> merged = new HashMap<String, FloatMatrix>();
		for (int s = clients.size() - 1; s >= 0; s--)              // shard 0 last: its answer is the one kept for replicated keys
			if (!buckets.get(s).isEmpty()) merged.putAll(clients.get(s).updateList(buckets.get(s), replace));
		return merged;
	}

	@Override public void push(String key, FloatMatrix gradient, String updaterKey, boolean async) {}
	@Override public void barrier() {}
}

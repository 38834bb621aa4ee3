package net;

import io.grpc.Server;
import io.grpc.ServerBuilder;
import io.grpc.stub.StreamObserver;
import nativeps.PsNative;
import store.KVStore;

import java.io.IOException;
import java.util.HashMap;
import java.util.Map;

/**
 * Drop-in for net/PServer.java (:54 ctor, :60 start, :75 get, :102 getList, :119 upsert, :144 upsertList, :164 push, :197 psUpdate, :238 barrier):
 * the same gRPC service (the generated net.PSGrpc of ps.proto:7-14, unchanged) over the GPU-resident store of this process, for worker JVMs
 * that keep the reference's Java layers and its PSClient.  Every call is one native call on the store under one lock (a model is not
 * thread-safe; each executor thread makes the context current first); push is ps_model_push — one step of the updater the request names, on
 * the device.  The gradient sums follow the reference's server, which never clears them (only Trainer.java:95 calls KVStore.clear): the k-th
 * push of a key applies (s_{k-1} + g_k) / k (KVStore.java:192-208).  barrier keeps the BSP meaning of :238-283 with a monitor instead of the
 * sleep-poll loops.  ps_b200/wire.py is the same server in Python (tests/test_wire.py pins its behaviour).
 * SOURCE ONLY: no JDK, grpc-java or protoc in the build image.
 */
public class PServer implements net.PSGrpc.PS {
	static final Resp success = Resp.newBuilder().setEc(200).setEm("").build();
	private final Server server;
	private final long ctx, model;
	private final int workerNum;
	private final boolean async;
	private final Map<String, float[]> sum = new HashMap<String, float[]>();
	private final Map<String, Integer> cnt = new HashMap<String, Integer>();
	private final Map<String, String> updateKeys = new HashMap<String, String>();
	private long globalStep = 0, workerStep = 0;

	public PServer(int port, int workerNum) {
		this.server = ServerBuilder.forPort(port).addService(net.PSGrpc.bindService(this)).build();
		this.ctx = KVStore.ins().nativeCtx();
		this.model = KVStore.ins().model();
		this.workerNum = workerNum;
		this.async = context.Context.isPsAsync;
	}
	public void start() {
		try { server.start(); server.awaitTermination(); }
		catch (IOException e) { e.printStackTrace(); }
		catch (InterruptedException e) { e.printStackTrace(); }
		finally { close(); }
	}
	public void close() { server.shutdown(); }

	private static Resp error(int ec, String em) { return Resp.newBuilder().setEc(ec).setEm(em).build(); }
	private static Matrix.Builder toProto(String key, float[] v, int rows, int cols) {
		Matrix.Builder m = Matrix.newBuilder().setKey(key);
		if (v == null) return m;
		for (float x : v) m.addData(x);
		return m.setRow(rows).setCols(cols);
	}
	private static float[] data(Matrix m) {
		float[] d = new float[m.getDataCount()];
		for (int i = 0; i < d.length; i++) d[i] = m.getData(i);
		return d;
	}
	/** rows x cols of a stored key as the reference's FloatMatrix has them (FcLayer weights out x in; everything else n x 1) */
	private int[] shape(String key, int n) {
		int[] s = KVStore.ins().shapeOf(key);
		return s != null ? s : new int[]{n, 1};
	}

	public synchronized void get(GetMessage req, StreamObserver<GetMessage> out) {
		PsNative.ctxMakeCurrent(ctx);
		String key = req.getWeights().getKey();
		float[] v = PsNative.modelGet(model, key);
		GetMessage.Builder r = GetMessage.newBuilder();
		if (v == null) r.setResp(error(204, "null weights"));
		else { int[] s = shape(key, v.length); r.setWeights(toProto(key, v, s[0], s[1])).setResp(success); }
		out.onNext(r.build()); out.onCompleted();
	}
	public synchronized void getList(GetListMessage req, StreamObserver<GetListMessage> out) {
		PsNative.ctxMakeCurrent(ctx);
		String[] keys = new String[req.getWeightsCount()];
		for (int i = 0; i < keys.length; i++) keys[i] = req.getWeights(i).getKey();
		float[][] rows = PsNative.modelGetList(model, keys);                 // one batched native call (embedding keys: one lookup kernel)
		GetListMessage.Builder r = GetListMessage.newBuilder();
		for (int i = 0; i < keys.length; i++) {
			if (rows[i] == null) r.addWeights(toProto(keys[i], null, 0, 0));  // unknown key: the key alone (PServer.java:106-111)
			else { int[] s = shape(keys[i], rows[i].length); r.addWeights(toProto(keys[i], rows[i], s[0], s[1])); }
		}
		out.onNext(r.setResp(success).build()); out.onCompleted();
	}
	private Matrix upsertOne(Matrix m, boolean replace) {
		float[] exists = PsNative.modelGet(model, m.getKey());
		boolean update = true;
		int rows = m.getRow(), cols = m.getCols();
		if (exists == null || replace) {
			update = false;
			exists = data(m);
			KVStore.ins().rememberShape(m.getKey(), rows, cols);
			PsNative.modelPut(model, m.getKey(), exists);
		} else { int[] s = shape(m.getKey(), exists.length); rows = s[0]; cols = s[1]; }
		return toProto(m.getKey(), exists, rows, cols).setUpdate(update).build();
	}
	public synchronized void upsert(UpdateMessage req, StreamObserver<UpdateMessage> out) {
		PsNative.ctxMakeCurrent(ctx);
		out.onNext(UpdateMessage.newBuilder().setWeights(upsertOne(req.getWeights(), req.getReplace())).setResp(success).build());
		out.onCompleted();
	}
	public synchronized void upsertList(UpdateListMessage req, StreamObserver<UpdateListMessage> out) {
		PsNative.ctxMakeCurrent(ctx);
		UpdateListMessage.Builder r = UpdateListMessage.newBuilder();
		for (int i = 0; i < req.getWeightsCount(); i++) r.addWeights(upsertOne(req.getWeights(i), req.getReplace()));
		out.onNext(r.setResp(success).build()); out.onCompleted();
	}

	/** KVStore.sum then KVStore.update(updater, key) as a server runs them: in-place mean of a sum that is never cleared */
	private boolean sumAndUpdate(String key, float[] g, String updaterKey, boolean applyNow) {
		float[] s = sum.get(key);
		if (s == null) { sum.put(key, g.clone()); cnt.put(key, 1); }
		else { for (int i = 0; i < s.length; i++) s[i] += g[i]; cnt.put(key, cnt.get(key) + 1); }
		if (!applyNow) { if (!updateKeys.containsKey(key)) updateKeys.put(key, updaterKey); return true; }
		return apply(key, updaterKey);
	}
	private boolean apply(String key, String updaterKey) {
		float[] s = sum.get(key);
		int n = cnt.get(key);
		for (int i = 0; i < s.length; i++) s[i] /= n;
		return PsNative.modelPush(model, key, s, updaterKey);
	}
	public synchronized void push(GradientMessage req, StreamObserver<GradientMessage> out) {
		PsNative.ctxMakeCurrent(ctx);
		GradientMessage.Builder r = GradientMessage.newBuilder();
		float[] spec = null;
		try { spec = PsNative.updaterParse(req.getUpdaterKey()); } catch (RuntimeException e) { spec = null; }
		if (spec == null) r.setResp(error(500, "updater is null"));           // PServer.java:169-174
		else if (!sumAndUpdate(req.getGradient().getKey(), data(req.getGradient()), req.getUpdaterKey(), req.getIsAsync()))
			r.setResp(error(500, "null weights"));
		out.onNext(r.build()); out.onCompleted();
	}
	public void barrier(BarrierMessage req, StreamObserver<BarrierMessage> out) {
		synchronized (this) {
			workerStep++;
			if (async) globalStep++;                                          // PServer.java:241-247: does not block
			else {
				long mine = (workerStep - 1) / workerNum;                     // the meeting this arrival belongs to
				if (workerStep % workerNum == 0) {                            // the last arrival runs psUpdate (:197-214) and releases the others
					PsNative.ctxMakeCurrent(ctx);
					for (Map.Entry<String, String> e : updateKeys.entrySet()) apply(e.getKey(), e.getValue());
					updateKeys.clear();
					globalStep++;
					notifyAll();
				} else {
					while (globalStep <= mine) {
						try { wait(100); } catch (InterruptedException e) { Thread.currentThread().interrupt(); break; }
					}
				}
			}
		}
		out.onNext(BarrierMessage.newBuilder().setResp(success).build()); out.onCompleted();
	}
}

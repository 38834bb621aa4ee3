"""GPU parity tests of the round-2 surface, all through the C ABI (ctypes):
  - the library default FcLayer mode is the tcgen05 tensor-core path (3xTF32) and agrees with the oracle at cfg2 widths
  - ps_model_forward / ps_model_backward_update: DNN.train call by call with the loss computed by the caller
    (DNN.java:44-68, loss/CrossEntropy.java:10-28), against the oracle's whole Trainer step
  - a full embedding table skips the step: nothing is applied, the error surfaces, dense parameters keep their bits
  - a forward that is followed by a larger forward (never by a backward) leaves no counts behind (EmbeddingField.java:66-104)
  - ps_model_submit_text: libsvm text parsed on the device == the host-parsed batch; a bad line drops the batch
  - the TMA-staged rows of the lookup are the same bits as the directly loaded ones
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol
from ps_b200.synth import Synth

pytestmark = pytest.mark.gpu
SEED = 20261017


def rel_err(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(1e-12, float(np.abs(b).max())))


def java_cross_entropy(P, Y):
    """loss/CrossEntropy.java:10-28 as the caller (Java) would run it: float accumulation of per-sample terms computed in double."""
    s = np.float32(0.0)
    for p, l in zip(P.astype(np.float32), Y.astype(np.float32)):
        t = -np.float64(l) * np.log(np.float64(p)) - np.float64(np.float32(1) - l) * np.log(np.float64(np.float32(1) - p))
        s = np.float32(s + np.float32(t))
    loss = np.float32(s / np.float32(len(P)))
    P = P.astype(np.float32)
    delta = (P - Y.astype(np.float32)) / (P * (np.float32(1) - P))
    return float(loss), delta.astype(np.float32)


def test_default_fc_mode_is_tensor_core_and_fp32_grade(ps):
    ctx = ps.Context(0, seed=SEED)
    assert ctx.fc_precision() == ps.PS_FC_TF32X3
    F, D, Xn, fc, N = 23, 16, 45, [256, 256, 256, 1], 512        # BASELINE config 2 widths
    m = ps.Model(ctx, "widedeep", F, D, Xn, fc, emb_capacity=1 << 16, max_batch=N)
    o = ol.OracleModel(ol.KIND_WIDEDEEP, F, D, Xn, fc, SEED)
    syn = Synth(F=F, Xn=Xn, V=20000, seed=3)
    for _ in range(3):
        b = syn.batch(N)
        lg = m.train_step(b["E"], b["X"], b["W"], b["Y"])
        lo = o.train_step(b["E"], b["X"], b["W"], b["Y"])
        assert abs(lg - lo) <= 1e-4 * max(1.0, abs(lo)), (lg, lo)
    for l in range(4):
        assert rel_err(m.get(f"fc{l}.weights"), o.get(f"fc{l}.weights")) <= 2e-4, l
    m.close()
    ctx.close()


@pytest.mark.parametrize("kind,mode", [("dnn", "fp32"), ("widedeep", "fp32"), ("widedeep", "tf32x3")])
def test_forward_backward_split_matches_oracle(ps, ctx, kind, mode):
    ctx.set_fc_precision(ps.PS_FC_FP32 if mode == "fp32" else ps.PS_FC_TF32X3)
    tol = 3e-5 if mode == "fp32" else 1e-4
    F, D, Xn, fc, N = 23, 16, 45, [64, 32, 1], 384
    m = ps.Model(ctx, kind, F, D, Xn, fc, emb_capacity=1 << 16, max_batch=N)
    o = ol.OracleModel(ol.KIND_WIDEDEEP if kind == "widedeep" else ol.KIND_DNN, F, D, Xn, fc, SEED)
    syn = Synth(F=F, Xn=Xn, V=8000, seed=5)
    W = lambda b: b["W"] if kind == "widedeep" else None
    for it in range(3):
        b = syn.batch(N)
        P = m.forward(b["E"], b["X"], W(b))                   # the forward loop
        loss, delta = java_cross_entropy(P, b["Y"])          # Loss.forward / Loss.backward stay in the caller
        m.backward_update(delta, loss)                       # the reverse loop + KVStore.update
        lo = o.train_step(b["E"], b["X"], W(b), b["Y"])
        assert abs(loss - lo) <= tol * max(1.0, abs(lo)), (it, loss, lo)
        assert not m.step_info()["skipped"]
    for l in range(len(fc)):
        assert rel_err(m.get(f"fc{l}.weights"), o.get(f"fc{l}.weights")) <= 5 * tol, l
        assert rel_err(m.get(f"fc{l}.bias"), o.get(f"fc{l}.bias")) <= 5 * tol, l
    for n in range(0, N, 37):
        for j in range(F):
            k = ol.key_string(0, j, int(b["E"][n, j]))
            assert np.allclose(m.get(k), o.get(k), rtol=2e-3 if mode != "fp32" else 5e-4, atol=2e-6), k
    if kind == "widedeep":
        k = ol.key_string(1, 0, int(b["W"][0, 0]))
        assert np.allclose(m.get(k), o.get(k), rtol=1e-3, atol=1e-6)
    # a forward that is never followed by backward_update (DNN.java:58-63 early exit in the caller) is simply forgotten
    b2 = syn.batch(N)
    m.forward(b2["E"], b2["X"], W(b2))
    l3 = m.train_step(b["E"], b["X"], W(b), b["Y"])
    lo3 = o.train_step(b["E"], b["X"], W(b), b["Y"])
    assert abs(l3 - lo3) <= 5 * tol * max(1.0, abs(lo3))
    with pytest.raises(ps.PsError):
        m.backward_update(delta, loss)                       # no pending forward
    m.close()


def test_full_table_skips_the_step(ps, ctx):
    F, D, Xn, fc, N = 4, 8, 3, [16, 1], 256
    m = ps.Model(ctx, "dnn", F, D, Xn, fc, emb_capacity=64, max_batch=N)
    syn = Synth(F=F, Xn=Xn, V=100000, dist="uniform", seed=9)
    small = syn.batch(8)                                        # 32 keys: fits
    m.train_step(small["E"], small["X"], None, small["Y"])
    before = {k: m.get(k).copy() for k in ("fc0.weights", "fc0.bias", "fc1.weights", "fc1.bias")}
    row_key = ol.key_string(0, 0, int(small["E"][0, 0]))
    row_before = m.get(row_key).copy()
    big = syn.batch(N)                                          # ~1000 new keys: the table overflows
    big["E"][0] = small["E"][0]                                 # ... and one sample reuses known keys
    with pytest.raises(ps.PsError) as e:
        m.train_step(big["E"], big["X"], None, big["Y"])
    assert e.value.code == 507
    assert m.step_info()["skipped"]
    for k, v in before.items():
        assert np.array_equal(m.get(k), v), k                   # no update was applied with garbage activations
    assert np.array_equal(m.get(row_key), row_before)
    m.close()


@pytest.mark.parametrize("opt", ["adam", "ftrl"])
def test_forward_without_backward_then_larger_forward(ps, ctx, opt):
    """forward(N=64) never followed by backward, then forward(N=256) + backward must behave as if the first batch's
    occurrence counts never existed (its keys stay created)."""
    F, D = 5, 12
    upd = ps.UpdaterSpec.ftrl() if opt == "ftrl" else None
    emb = ps.EmbeddingLayer(ctx, F, D, capacity=1 << 12, updater=upd)
    o = ol.lib().pso_emb_create(F, D, SEED, 1 if opt == "ftrl" else 0)
    rng = np.random.default_rng(4)
    E1 = rng.integers(0, 40, (64, F)).astype(np.int64)
    E2 = rng.integers(0, 40, (256, F)).astype(np.int64)
    out_o = np.zeros((64, F * D), np.float32)
    ol.lib().pso_emb_forward(o, E1, 64, out_o)
    assert np.array_equal(emb.forward(E1), out_o)
    out_o2 = np.zeros((256, F * D), np.float32)
    ol.lib().pso_emb_forward(o, E2, 256, out_o2)
    assert np.array_equal(emb.forward(E2), out_o2)
    delta = rng.standard_normal((256, F * D)).astype(np.float32)
    emb.backward_update(delta, calls=2)
    ol.lib().pso_emb_backward_update(o, np.ascontiguousarray(delta), F * D, 256, 2)
    out_o3 = np.zeros((256, F * D), np.float32)
    ol.lib().pso_emb_forward(o, E2, 256, out_o3)
    got = emb.forward(E2)
    assert np.allclose(got, out_o3, rtol=2e-5, atol=1e-7)
    emb.backward_update(np.zeros_like(delta), calls=2)
    emb.close()


def _libsvm_text(b, rows):
    F, Xn = b["E"].shape[1], b["X"].shape[1]
    lines = []
    for n in range(rows):
        cols = ["%d" % int(b["Y"][n])] + ["%d:1" % int(b["E"][n, j]) for j in range(F)] + ["%d:%.2f" % (33895 + x, b["X"][n, x]) for x in range(Xn)]
        lines.append(" ".join(cols))
    return ("\n".join(lines) + "\n").encode()


def test_submit_text_equals_host_parsed_batch(ps, ctx):
    F, D, Xn, fc, N = 23, 16, 45, [64, 32, 1], 256
    syn = Synth(F=F, Xn=Xn, V=30000, seed=21)
    ma = ps.Model(ctx, "widedeep", F, D, Xn, fc, emb_capacity=1 << 16, max_batch=N)
    mb = ps.Model(ctx, "widedeep", F, D, Xn, fc, emb_capacity=1 << 16, max_batch=N)
    for it in range(3):
        b = syn.batch(N)
        text = _libsvm_text(b, N)
        buf = ps.PinnedArray((len(text),), np.uint8)
        buf.array[:] = np.frombuffer(text, np.uint8)
        ma.submit_text(buf.ptr, len(text), N)
        la = ma.collect()
        info = ma.step_info()
        assert info["bad_lines"] == 0 and not info["skipped"]
        lb = mb.train_step(b["E"], b["X"], b["W"], b["Y"])       # CTR.parseFeature on the host: W = E % 100000
        assert abs(la - lb) <= 2e-6 * max(1.0, abs(lb)), (it, la, lb)   # same kernels, same inputs (the scatter's atomics may reorder sums)
        buf.free()
    assert np.allclose(ma.get("fc0.weights"), mb.get("fc0.weights"), rtol=1e-4, atol=1e-6)
    # a malformed line drops the whole batch: nothing is applied
    w0 = ma.get("fc0.weights").copy()
    b = syn.batch(N)
    text = _libsvm_text(b, N).replace(b" 33900:", b" 33900;", 1)
    buf = ps.PinnedArray((len(text),), np.uint8)
    buf.array[:] = np.frombuffer(text, np.uint8)
    ma.submit_text(buf.ptr, len(text), N)
    ma.collect()
    info = ma.step_info()
    assert info["skipped"] and info["bad_lines"] >= 1
    assert np.array_equal(ma.get("fc0.weights"), w0)
    # and the next ordinary batch trains normally again
    ma.train_step(b["E"], b["X"], b["W"], b["Y"])
    assert not ma.step_info()["skipped"]
    assert not np.array_equal(ma.get("fc0.weights"), w0)
    buf.free()
    ma.close()
    mb.close()


@pytest.mark.parametrize("D", [10, 16, 64, 128])
def test_tma_staged_rows_are_the_same_bits(ps, D):
    """Rows shared by >= 4 lookups of a warp task are staged in shared memory by cp.async.bulk; PS_HOT_TMA=0 loads every row
    directly.  Low-cardinality fields (2 and 5 ids) make most tasks take the staged path."""
    F, N = 6, 1000
    rng = np.random.default_rng(D)
    E = np.stack([rng.integers(0, v, N) for v in (1, 2, 5, 21, 3000, 7)], axis=1).astype(np.int64)
    outs = []
    for flag in ("1", "0"):
        os.environ["PS_HOT_TMA"] = flag
        try:
            c = ps.Context(0, seed=SEED)
        finally:
            os.environ.pop("PS_HOT_TMA", None)
        emb = ps.EmbeddingLayer(c, F, D, capacity=1 << 14)
        first = emb.forward(E)                                    # rows are created here (initialiser path) ...
        emb.backward_update(np.ones((N, F * D), np.float32), calls=2)
        again = emb.forward(E)                                    # ... and read back from memory here (staged / direct path)
        emb.backward_update(np.zeros((N, F * D), np.float32), calls=2)
        outs.append((first, again))
        emb.close()
        c.close()
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])
    o = ol.lib().pso_emb_create(F, D, SEED, 0)
    ref = np.zeros((N, F * D), np.float32)
    ol.lib().pso_emb_forward(o, E, N, ref)
    assert np.array_equal(outs[0][0], ref)


def test_get_list_and_update_list(ps, ctx):
    """PSClient.getList / updateList by reference key strings (PSClient.java:72-97,128-151; PServer.java:102-117,144-162)."""
    F, D, Xn, fc, N = 23, 16, 45, [32, 1], 128
    m = ps.Model(ctx, "widedeep", F, D, Xn, fc, emb_capacity=1 << 14, max_batch=N)
    o = ol.OracleModel(ol.KIND_WIDEDEEP, F, D, Xn, fc, SEED)
    b = Synth(F=F, Xn=Xn, V=5000, seed=8).batch(N)
    m.train_step(b["E"], b["X"], b["W"], b["Y"])
    o.train_step(b["E"], b["X"], b["W"], b["Y"])
    keys = [ol.key_string(0, j, int(b["E"][n, j])) for n in range(0, N, 9) for j in range(F)]
    keys += ["fc0.bias", "wide.bias", ol.key_string(1, 0, int(b["W"][3, 2])), "emF3.999999.0", "no.such.key"]
    got = m.get_list(keys, stride=64)
    for k in keys:
        ref = o.get(k)
        if ref is None:
            assert got[k] is None, k
        else:
            assert got[k] is not None and np.allclose(got[k], ref, rtol=5e-4, atol=2e-6), k
    # updateList with replace = false: existing keys keep their value and the caller receives it; absent keys are created
    known = ol.key_string(0, 0, int(b["E"][0, 0]))
    fresh = "emF5.424242.0"
    offered = {known: np.full(D, 7.0, np.float32), fresh: np.arange(D, dtype=np.float32)}
    back = m.update_list(offered, replace=False)
    assert np.array_equal(back[known], got[known]) and np.array_equal(m.get(known), got[known])
    assert np.array_equal(back[fresh], offered[fresh]) and np.array_equal(m.get(fresh), offered[fresh])
    back = m.update_list({known: np.full(D, 7.0, np.float32)}, replace=True)
    assert np.array_equal(m.get(known), np.full(D, 7.0, np.float32))
    # a row installed through updateList trains like any other (its ready flag is set)
    l1 = m.train_step(b["E"], b["X"], b["W"], b["Y"])
    assert np.isfinite(l1) and not np.array_equal(m.get(known), np.full(D, 7.0, np.float32))
    m.close()


def test_out_of_domain_id_refuses_the_batch(ps, ctx):
    F, D, Xn, fc, N = 3, 8, 2, [4, 1], 32
    m = ps.Model(ctx, "dnn", F, D, Xn, fc, emb_capacity=1 << 10, max_batch=N)
    b = Synth(F=F, Xn=Xn, V=100, seed=2).batch(N)
    m.train_step(b["E"], b["X"], None, b["Y"])
    w0 = m.get("fc0.weights").copy()
    bad = {k: v.copy() for k, v in b.items()}
    bad["E"][5, 1] = -3                                         # the reference would keep "emF1.-3.0" apart; the packed key cannot
    with pytest.raises(ps.PsError) as e:
        m.train_step(bad["E"], bad["X"], None, bad["Y"])
    assert e.value.code == 400
    assert np.array_equal(m.get("fc0.weights"), w0)
    m.close()


def test_fcnn_forward_backward_split_matches_oracle(ps, ctx):
    """FullConnectedNN.train call by call (FullConnectedNN.java:37-70): forward loop -> SoftmaxLoss in the caller
    (loss/SoftmaxLoss.java:9-28) -> Softmax.backward + reverse loop + KVStore.update in the library."""
    Xn, fc, N, Cn = 784, [150, 50, 10], 96, 10
    m = ps.Model(ctx, "fcnn", 0, 0, Xn, fc, max_batch=N)
    o = ol.OracleModel(ol.KIND_FCNN, 0, 0, Xn, fc, SEED)
    syn = Synth(F=0, Xn=Xn, V=0, seed=23, n_classes=Cn)
    for it in range(3):
        b = syn.batch(N)
        P = np.zeros((N, Cn), np.float32)
        X = np.ascontiguousarray(b["X"], np.float32)
        ps.check(ps.lib().ps_model_forward(m.h, None, X.ctypes.data_as(C.c_void_p), None, N, P.ctypes.data_as(C.c_void_p)))
        hot = b["Y"].astype(np.int64)
        ph = P[np.arange(N), hot]
        loss = float(np.float32(np.sum(-np.log(ph.astype(np.float64)).astype(np.float32), dtype=np.float32) / np.float32(N)))
        delta = np.zeros((N, Cn), np.float32)
        delta[np.arange(N), hot] = np.float32(-1.0) / ph
        ps.check(ps.lib().ps_model_backward_update(m.h, delta.ctypes.data_as(C.c_void_p), N, loss))
        lo = o.train_step(None, b["X"], None, b["Y"])
        assert abs(loss - lo) <= 5e-5 * max(1.0, abs(lo)), (it, loss, lo)
    for l in range(3):
        assert rel_err(m.get(f"fc{l}.weights"), o.get(f"fc{l}.weights")) <= 2e-4, l
        assert rel_err(m.get(f"fc{l}.bias"), o.get(f"fc{l}.bias")) <= 2e-4, l
    m.close()


def test_push_is_one_updater_step_on_every_kind_of_key(ps, ctx):
    """ps_model_push = PServer.push → KVStore.update(updater, key) (PServer.java:164-184, KVStore.java:202-208): the named updater's step on the
    stored weight and state with the pushed gradient — bit for bit what ps_updater_apply (checked against the oracle elsewhere) gives."""
    F, D, Xn, fc, N = 23, 16, 45, [32, 1], 128
    m = ps.Model(ctx, "widedeep", F, D, Xn, fc, emb_capacity=1 << 14, max_batch=N)
    b = Synth(F=F, Xn=Xn, V=5000, seed=12).batch(N)
    m.train_step(b["E"], b["X"], b["W"], b["Y"])
    rng = np.random.default_rng(4)
    keys = [ol.key_string(0, 2, int(b["E"][5, 2])), "fc0.weights", "fc0.bias", ol.key_string(1, 0, int(b["W"][3, 2])), "wide.bias"]
    for spec in (ps.UpdaterSpec.adam(0.01, 0.9, 0.999, 1e-8), ps.UpdaterSpec.ftrl(0.05, 1.0, 0.001, 0.001)):
        for k in keys:
            w, s1, s2 = m.get(k).copy(), m.get_state(k, 0).copy(), m.get_state(k, 1).copy()
            g = rng.standard_normal(w.size).astype(np.float32)
            assert m.push(k, g, spec), k
            ctx.updater_apply(spec, w, s1, s2, g)
            assert np.array_equal(m.get(k), w) and np.array_equal(m.get_state(k, 0), s1) and np.array_equal(m.get_state(k, 1), s2), k
    # FtrlUpdater.java:52: a gradient whose first element is zero leaves the key alone
    k = keys[0]
    before = m.get(k).copy()
    g = rng.standard_normal(D).astype(np.float32)
    g[0] = 0.0
    assert m.push(k, g, ps.UpdaterSpec.ftrl()) and np.array_equal(m.get(k), before)
    assert not m.push("emF3.999999.0", np.zeros(D, np.float32), ps.UpdaterSpec.adam())       # unknown key
    # the FcLayer's transposed / residual copies follow a pushed weight: the model still trains and predicts like one whose weights were put
    m2 = ps.Model(ctx, "widedeep", F, D, Xn, fc, emb_capacity=1 << 14, max_batch=N)
    m2.train_step(b["E"], b["X"], b["W"], b["Y"])
    for k in keys:
        m2.put(k, m.get(k))
    assert np.allclose(m.predict(b["E"], b["X"], b["W"], N), m2.predict(b["E"], b["X"], b["W"], N), rtol=1e-6, atol=1e-7)
    m.close()
    m2.close()


def test_wire_server_over_the_gpu_store(ps, ctx):
    """ps_b200/wire.py: a legacy worker's getList / push / barrier against the tables the native step trains."""
    pytest.importorskip("grpc")
    from ps_b200 import wire
    F, D, Xn, fc, N = 23, 16, 45, [32, 1], 128
    m = ps.Model(ctx, "widedeep", F, D, Xn, fc, emb_capacity=1 << 14, max_batch=N)
    b = Synth(F=F, Xn=Xn, V=5000, seed=13).batch(N)
    m.train_step(b["E"], b["X"], b["W"], b["Y"])
    srv = wire.PsWireServer(wire.ModelStore(ps, m), worker_num=1, is_async=True)
    port = srv.start(0)
    cl = wire.WireClient(f"127.0.0.1:{port}")
    try:
        k = ol.key_string(0, 1, int(b["E"][7, 1]))
        got = cl.get_list([k, "fc0.bias", "emF3.999999.0"])
        assert got["emF3.999999.0"] is None and np.array_equal(got[k][2], m.get(k)) and np.array_equal(got["fc0.bias"][2], m.get("fc0.bias"))
        assert cl.get("emF3.999999.0") is None
        assert cl.get("fc0.weights")[:2] == (fc[0], F * D + Xn) and cl.get(k)[:2] == (D, 1)
        name = ps.updater_name(ps.UpdaterSpec.adam(0.01, 0.9, 0.999, 1e-8))        # the updaterKey a worker sends: Updater.getName()
        w, s1, s2 = m.get(k).copy(), m.get_state(k, 0).copy(), m.get_state(k, 1).copy()
        g = np.linspace(-1, 1, D).astype(np.float32)
        assert cl.push(k, g, name).resp.ec == 0
        ctx.updater_apply(ps.updater_parse(name), w, s1, s2, g)                     # first push of the key: the server's running sum is g itself
        assert np.array_equal(m.get(k), w)
        assert cl.push(k, g, "nosuch@").resp.ec == 500
        fresh = "emF4.31337.0"
        back = cl.update_list({fresh: (D, 1, np.arange(D, dtype=np.float32))}, replace=False)
        assert not back[fresh][1] and np.array_equal(m.get(fresh), np.arange(D, dtype=np.float32))
        assert cl.barrier().resp.ec == 200
    finally:
        cl.close()
        srv.stop()
        m.close()

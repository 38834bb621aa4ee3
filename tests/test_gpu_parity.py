"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle
(oracle/ps_oracle.cpp, a restatement of the reference's standalone Java path) on the same
seeded inputs.  Bars:
  * keys / routing / gather output (copy + ReLU): bit-exact
  * updater arithmetic given the same gradient: bit-exact
  * embedding rows + optimiser state after the fused scatter/update: <= 1e-5 relative
    (fp32 reassociation of <= n-term sums through L2 reductions)
  * FcLayer path, PS_FC_FP32: <= 2e-5 relative to max|x| per matrix (FFMA + tiled order vs the
    oracle's ordered loops)
  * FcLayer path, PS_FC_TF32 (tcgen05, operands truncated to 10 mantissa bits by the tensor core):
    every GEMM is within the truncation bound 2.5e-3 * sum|a||b| of fp64 on its own inputs;
    against the fp32 oracle the chained network agrees to <= 2e-2 (activations, max-relative),
    <= 5e-3 / 5e-2 (activations / deltas, Frobenius-relative) and 2e-2 on the loss
"""
import numpy as np
import pytest

import oracle_lib as ol
from ps_b200.synth import Synth

pytestmark = pytest.mark.gpu

SEED = 20261017


def fro_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(1e-30, np.linalg.norm(b)))


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


# --------------------------------------------------------------------------- updaters
@pytest.mark.parametrize("kind", ["adam", "ftrl", "simple"])
def test_updater_bit_exact(ps, ctx, kind):
    rng = np.random.default_rng(1)
    n = 4096
    w = rng.standard_normal(n).astype(np.float32)
    g = (rng.standard_normal(n) * rng.choice([1e-6, 1e-3, 1.0, 30.0], n)).astype(np.float32)
    g[0] = 0.25
    L = ol.lib()
    for step in range(3):
        if kind == "adam":
            s1 = np.abs(rng.standard_normal(n)).astype(np.float32) * (step > 0)
            s2 = np.abs(rng.standard_normal(n)).astype(np.float32) * (step > 0)
            spec = ps.UpdaterSpec.adam()
            wo, ao, bo = w.copy(), s1.copy(), s2.copy()
            L.pso_adam_update(wo, ao, bo, g, n, *[spec.p[i] for i in range(4)])
        elif kind == "ftrl":
            s1 = (rng.standard_normal(n) * 0.01).astype(np.float32) * (step > 0)
            s2 = np.abs(rng.standard_normal(n)).astype(np.float32) * (step > 0)
            spec = ps.UpdaterSpec.ftrl()
            wo, ao, bo = w.copy(), s1.copy(), s2.copy()
            L.pso_ftrl_update(wo, ao, bo, g, n, *[spec.p[i] for i in range(4)])
        else:
            s1 = np.zeros(n, np.float32)
            s2 = np.zeros(n, np.float32)
            spec = ps.UpdaterSpec.simple(0.05)
            wo, ao, bo = w - np.float32(0.05) * g, s1.copy(), s2.copy()
            wo = (w + g * np.float32(-0.05)).astype(np.float32)
        wg, ag, bg = w.copy(), s1.copy(), s2.copy()
        ctx.updater_apply(spec, wg, ag, bg, g)
        assert np.array_equal(wg.view(np.uint32), wo.view(np.uint32))
        assert np.array_equal(ag.view(np.uint32), ao.view(np.uint32))
        assert np.array_equal(bg.view(np.uint32), bo.view(np.uint32))
        w = wg


def test_ftrl_skips_on_zero_first_element(ps, ctx):
    n = 8
    w = np.ones(n, np.float32)
    z = np.full(n, 0.5, np.float32)
    nn = np.ones(n, np.float32)
    g = np.ones(n, np.float32)
    g[0] = 0.0
    ctx.updater_apply(ps.UpdaterSpec.ftrl(), w, z, nn, g)
    assert np.all(w == 1) and np.all(z == 0.5) and np.all(nn == 1)   # FtrlUpdater.java:52


# --------------------------------------------------------------------------- embedding layer
@pytest.mark.parametrize("F,D,N,V,dist", [(23, 16, 512, 5000, "zipf"), (23, 10, 333, 2000, "zipf"), (3, 32, 64, 50, "uniform"),
                                          (5, 64, 1, 10, "uniform"), (23, 4, 1000, 100000, "uniform"), (2, 128, 17, 9, "uniform")])
def test_embedding_forward_bit_exact(ps, ctx, F, D, N, V, dist):
    emb = ps.EmbeddingLayer(ctx, F, D, capacity=max(1024, 4 * N * F))
    o = ol.lib().pso_emb_create(F, D, SEED, 0)
    syn = Synth(F=F, Xn=1, V=V, dist=dist, seed=3)
    for it in range(3):
        E = syn.batch(N)["E"]
        out_g = emb.forward(E)
        out_o = np.zeros((N, F * D), np.float32)
        ol.lib().pso_emb_forward(o, np.ascontiguousarray(E), N, out_o.reshape(-1))
        assert np.array_equal(out_g.view(np.uint32), out_o.view(np.uint32)), f"iteration {it}"
    # float-carried ids (the reference's FloatMatrix "E") give the same rows
    out_f = emb.forward(E.astype(np.float32))
    assert np.array_equal(out_f.view(np.uint32), out_o.view(np.uint32))
    ol.lib().pso_model_destroy(o)
    emb.close()


@pytest.mark.parametrize("opt", ["adam", "ftrl"])
@pytest.mark.parametrize("F,D,N,V,calls", [(23, 16, 512, 3000, 2), (23, 10, 200, 500, 2), (4, 32, 256, 40, 1), (23, 16, 1024, 200000, 2),
                                           # hot keys (thousands of occurrences of one row: the per-block shared-memory pre-sum), ragged dims
                                           (3, 64, 2048, 7, 2), (23, 16, 4096, 1000, 2), (5, 10, 700, 11, 1), (2, 128, 300, 4, 2), (1, 4, 5000, 2, 2)])
@pytest.mark.parametrize("exact", [False, True])
def test_embedding_backward_update(ps, ctx, opt, F, D, N, V, calls, exact):
    ctx.set_exact_updaters(exact)             # fast forms (default) and the IEEE operation sequence of the Java updaters
    spec = ps.UpdaterSpec.adam() if opt == "adam" else ps.UpdaterSpec.ftrl()
    emb = ps.EmbeddingLayer(ctx, F, D, capacity=max(1024, 4 * N * F), updater=spec)
    o = ol.lib().pso_emb_create(F, D, SEED, 1 if opt == "ftrl" else 0)
    syn = Synth(F=F, Xn=1, V=V, dist="zipf", seed=5)
    rng = np.random.default_rng(9)
    ld = F * D + 7
    seen = {}
    for it in range(4):
        E = syn.batch(N)["E"]
        out_g = emb.forward(E)
        out_o = np.zeros((N, F * D), np.float32)
        ol.lib().pso_emb_forward(o, np.ascontiguousarray(E), N, out_o.reshape(-1))
        assert rel_err(out_g, out_o) <= 2e-5, f"forward drifted at iteration {it}"
        delta = rng.standard_normal((N, ld)).astype(np.float32)
        emb.backward_update(delta, calls=calls)
        ol.lib().pso_emb_backward_update(o, delta.reshape(-1), ld, N, calls)
        for j in range(F):
            for v in np.unique(E[:, j]):
                seen[(j, int(v))] = 1
    keys = list(seen)
    fields = np.array([k[0] for k in keys], np.int32)
    ids = np.array([k[1] for k in keys], np.int64)
    w, s1, s2, found = emb.get_rows(fields, ids, state=True)
    assert found.all() and emb.size() == len(keys)
    om = ol.OracleModel.__new__(ol.OracleModel)
    om.L, om.h = ol.lib(), o
    worst = 0.0
    for i, (j, v) in enumerate(keys):
        key = ol.key_string(0, j, v)
        wo = om.get(key)
        worst = max(worst, float(np.abs(w[i] - wo).max() / max(1e-6, np.abs(wo).max())))
        s1o, s2o = om.get_state(key, 0), om.get_state(key, 1)
        if s1o is not None:
            assert np.allclose(s1[i], s1o, rtol=2e-5, atol=1e-7), key
            assert np.allclose(s2[i], s2o, rtol=2e-5, atol=1e-9), key
    assert worst <= 2e-5, worst
    om.h = None
    ol.lib().pso_model_destroy(o)
    emb.close()


@pytest.mark.parametrize("opt", ["adam", "ftrl"])
def test_embedding_update_exact_mode_is_bit_exact(ps, ctx, opt):
    """With every key occurring once the gradient sum has one term (no reassociation), g_eff = S exactly, and the exact-mode
    sparse update must reproduce the Java updaters' bits: weights and both optimiser states, over several steps."""
    ctx.set_exact_updaters(True)
    F, D, N = 3, 16, 96
    spec = ps.UpdaterSpec.adam() if opt == "adam" else ps.UpdaterSpec.ftrl()
    emb = ps.EmbeddingLayer(ctx, F, D, capacity=4096, updater=spec)
    o = ol.lib().pso_emb_create(F, D, SEED, 1 if opt == "ftrl" else 0)
    rng = np.random.default_rng(21)
    E = np.stack([rng.permutation(1000)[:N] + 1000 * j for j in range(F)], 1).astype(np.int64)     # all distinct within a field
    for it in range(4):
        out_g = emb.forward(E)
        out_o = np.zeros((N, F * D), np.float32)
        ol.lib().pso_emb_forward(o, np.ascontiguousarray(E), N, out_o.reshape(-1))
        assert np.array_equal(out_g.view(np.uint32), out_o.view(np.uint32)), it
        delta = rng.standard_normal((N, F * D)).astype(np.float32)
        emb.backward_update(delta, calls=2)
        ol.lib().pso_emb_backward_update(o, delta.reshape(-1), F * D, N, 2)
    fields = np.repeat(np.arange(F, dtype=np.int32), N)
    ids = E.T.reshape(-1).copy()
    w, s1, s2, found = emb.get_rows(fields, ids, state=True)
    om = ol.OracleModel.__new__(ol.OracleModel)
    om.L, om.h = ol.lib(), o
    for i in range(len(ids)):
        key = ol.key_string(0, int(fields[i]), int(ids[i]))
        assert np.array_equal(w[i].view(np.uint32), om.get(key).view(np.uint32)), key
        for mine, which in ((s1[i], 0), (s2[i], 1)):
            so = om.get_state(key, which)                      # None until the updater first touches the key (Ftrl may skip: FtrlUpdater.java:52)
            so = np.zeros(D, np.float32) if so is None else so
            assert np.array_equal(mine.view(np.uint32), so.view(np.uint32)), (key, which)
    om.h = None
    ol.lib().pso_model_destroy(o)
    emb.close()


def test_embedding_geff_closed_form(ps, ctx):
    """key with n occurrences and gradient sum S receives S(n+1)/(2n^2); first Adam step
    then moves each element by -alfa*g/(|g|+eps)."""
    F, D, N = 1, 4, 6
    emb = ps.EmbeddingLayer(ctx, F, D, capacity=64)
    E = np.array([[5], [5], [5], [7], [9], [9]], np.int64)
    a0 = emb.forward(E)
    w0, _ = emb.get_rows(np.zeros(3, np.int32), np.array([5, 7, 9]))
    delta = np.arange(1, N * D + 1, dtype=np.float32).reshape(N, D) / 10
    emb.backward_update(delta, calls=2)
    w1, _ = emb.get_rows(np.zeros(3, np.int32), np.array([5, 7, 9]))
    for r, (key, rows) in enumerate([(5, [0, 1, 2]), (7, [3]), (9, [4, 5])]):
        n = len(rows)
        S = (delta[rows] * (a0[rows] > 0)).sum(0)
        g = S * (n + 1) / (2 * n * n)
        exp = w0[r] - 0.005 * g / (np.abs(g) + 1e-8)
        assert np.allclose(w1[r], exp, rtol=1e-5, atol=1e-7)
    emb.close()


def test_embedding_get_put_rows(ps, ctx):
    emb = ps.EmbeddingLayer(ctx, 3, 8, capacity=256)
    f = np.array([0, 1, 2, 2], np.int32)
    i = np.array([11, 11, 2 ** 40 + 3, 5], np.int64)
    w, found = emb.get_rows(f, i)
    assert not found.any()                                   # KVStore.get(String) on an absent key → null
    rows = np.arange(32, dtype=np.float32).reshape(4, 8)
    emb.put_rows(f, i, rows, replace=True)                   # KVStore.put
    w, found = emb.get_rows(f, i)
    assert found.all() and np.array_equal(w, rows)
    back = emb.put_rows(f, i, np.full((4, 8), -1, np.float32), replace=False)   # upsert(replace=false): server copy wins
    assert np.array_equal(back, rows)
    E = np.array([[11, 11, 5]], np.int64)
    assert np.array_equal(emb.forward(E), np.concatenate([rows[0], rows[1], rows[3]])[None])
    emb.close()


def test_embedding_capacity_error(ps, ctx):
    emb = ps.EmbeddingLayer(ctx, 1, 4, capacity=8)
    with pytest.raises(ps.PsError) as e:
        emb.forward(np.arange(64, dtype=np.int64).reshape(64, 1))
    assert e.value.code == 507
    emb.close()


# --------------------------------------------------------------------------- whole models
def _compare_models(m, o, F, fc, tol, keys_sample, kind, err=None):
    err = err or rel_err
    for l in range(len(fc)):
        for nm in (f"fc{l}.weights", f"fc{l}.bias"):
            assert err(m.get(nm), o.get(nm)) <= (tol if err is rel_err else 10 * tol), nm
            for which in (0, 1):
                so = o.get_state(nm, which)
                if so is not None:
                    assert err(m.get_state(nm, which), so) <= (10 * tol if err is rel_err else 50 * tol), (nm, which)
    bad = 0
    for key in keys_sample:
        wo = o.get(key)
        wg = m.get(key)
        assert (wo is None) == (wg is None), key
        if wo is not None:
            if err is rel_err:
                assert np.allclose(wg, wo, rtol=20 * tol, atol=20 * tol * 1e-2), key
            else:
                bad += int(not np.allclose(wg, wo, rtol=20 * tol, atol=20 * tol * 1e-2))
    assert bad <= max(1, len(keys_sample) // 50), bad
    if kind == "widedeep":
        assert np.allclose(m.get("wide.bias"), o.get("wide.bias"), rtol=20 * tol, atol=1e-7)


@pytest.mark.parametrize("kind,F,D,Xn,fc,N,V", [
    ("dnn", 23, 10, 45, [150, 10, 1], 250, 3000),          # CTR.java:91 shape
    ("widedeep", 23, 16, 45, [64, 32, 1], 512, 20000),
    ("widedeep", 5, 8, 3, [16, 1], 37, 60),                 # ragged batch, heavy key reuse
    ("dnn", 2, 4, 1, [1], 5, 4),
])
@pytest.mark.parametrize("mode", ["fp32", "tf32x3"])
def test_model_steps_match_oracle_fp32(ps, ctx, kind, F, D, Xn, fc, N, V, mode):
    """PS_FC_FP32 (FFMA) and PS_FC_TF32X3 (tcgen05, error-compensated) both meet the fp32 bar."""
    tol = 2e-5 if mode == "fp32" else 4e-5
    ctx.set_fc_precision(ps.PS_FC_FP32 if mode == "fp32" else ps.PS_FC_TF32X3)
    m = ps.Model(ctx, kind, F, D, Xn, fc, emb_capacity=1 << 16, max_batch=N)
    o = ol.OracleModel(ol.KIND_WIDEDEEP if kind == "widedeep" else ol.KIND_DNN, F, D, Xn, fc, SEED)
    syn = Synth(F=F, Xn=Xn, V=V, seed=11)
    last = None
    for it in range(4):
        b = syn.batch(N)
        lg = m.train_step(b["E"], b["X"], b["W"], b["Y"])
        lo = o.train_step(b["E"], b["X"], b["W"], b["Y"])
        assert abs(lg - lo) <= 1e-4 * max(1.0, abs(lo)), (it, lg, lo)
        assert m.skipped_backward() == o.skipped_backward()
        last = b
    # activations and deltas of the last step
    # (Adam's first steps move a row element by ~alfa whatever |g| is, so a gradient element that is ~0 may
    #  land on the other side under a different summation order: compare in norm)
    assert np.array_equal(m.tap("embedding", 0).view(np.uint32), o.tap("embedding", 0).view(np.uint32)) or \
        fro_err(m.tap("embedding", 0), o.tap("embedding", 0)) <= 50 * tol
    # after 4 Adam steps a gradient element that is ~0 may have stepped the other way under a different
    # summation order (Adam moves by ~alfa whatever |g|): the FFMA mode happens to stay element-wise close,
    # the tensor-core mode is compared in norm
    err = rel_err if mode == "fp32" else fro_err
    for l in range(len(fc)):
        assert err(m.tap(f"fc{l}", 0), o.tap(f"fc{l}", 0)) <= 5 * tol, f"fc{l}.A"
        assert err(m.tap(f"fc{l}", 1), o.tap(f"fc{l}", 1)) <= 20 * tol, f"fc{l}.delta"
    if kind == "widedeep":
        assert err(m.tap("wide", 0), o.tap("wide", 0)) <= 5 * tol
        assert err(m.tap("addWideDeep", 0), o.tap("addWideDeep", 0)) <= 5 * tol
        assert err(m.tap("addWideDeep", 1), o.tap("addWideDeep", 1)) <= 20 * tol
    E = last["E"]
    keys = [ol.key_string(0, j, int(E[n, j])) for n in range(min(N, 8)) for j in range(F)]
    if kind == "widedeep":
        keys += [ol.key_string(1, 0, int(last["W"][n, j])) for n in range(min(N, 4)) for j in range(F)]
    keys += ["emF0.123456789.0", "wide.weights.99999.0"]     # absent keys
    _compare_models(m, o, F, fc, tol, keys, kind, err)
    assert m.num_keys() == o.num_keys()
    # predict (PredictThread): forward only, nothing updated
    b = syn.batch(N)
    pg = m.predict(b["E"], b["X"], b["W"], N)
    po = o.predict(b["E"], b["X"], b["W"], N)
    if mode == "fp32":
        assert np.allclose(pg, po, rtol=1e-4, atol=1e-6)
    else:
        assert fro_err(pg, po) <= 50 * tol
    m.close()


def test_widedeep_ftrl_on_embeddings(ps, ctx):
    """BASELINE config 3: updaters.put("emF", ftrl) through KVStore's prefix rule (KVStore.java:244-248)."""
    F, D, Xn, fc, N = 23, 32, 45, [64, 64, 1], 256
    m = ps.Model(ctx, "widedeep", F, D, Xn, fc, emb_capacity=1 << 16, emb_updater=ps.UpdaterSpec.ftrl(), max_batch=N)
    o = ol.OracleModel(ol.KIND_WIDEDEEP, F, D, Xn, fc, SEED, emb_opt=1)
    syn = Synth(F=F, Xn=Xn, V=5000, seed=13)
    for it in range(3):
        b = syn.batch(N)
        lg = m.train_step(b["E"], b["X"], b["W"], b["Y"])
        lo = o.train_step(b["E"], b["X"], b["W"], b["Y"])
        assert abs(lg - lo) <= 5e-5 * max(1.0, abs(lo))
    keys = [ol.key_string(0, j, int(b["E"][n, j])) for n in range(6) for j in range(F)]
    for key in keys:
        wo, wg = o.get(key), m.get(key)
        assert np.allclose(wg, wo, rtol=5e-4, atol=2e-6), key
        for which in (0, 1):
            so = o.get_state(key, which)
            if so is not None:
                assert np.allclose(m.get_state(key, which), so, rtol=5e-4, atol=2e-6), (key, which)
    m.close()


def test_fcnn_softmax_matches_oracle(ps, ctx):
    """BASELINE config 5 shape (Mnist.java:95): 784 -> 150 -> 50 -> 10, Softmax + SoftmaxLoss."""
    Xn, fc, N = 784, [150, 50, 10], 128
    m = ps.Model(ctx, "fcnn", 0, 0, Xn, fc, max_batch=N)
    o = ol.OracleModel(ol.KIND_FCNN, 0, 0, Xn, fc, SEED)
    syn = Synth(F=0, Xn=Xn, V=0, seed=17, n_classes=10)
    for it in range(3):
        b = syn.batch(N)
        lg = m.train_step(None, b["X"], None, b["Y"])
        lo = o.train_step(None, b["X"], None, b["Y"])
        assert abs(lg - lo) <= 5e-5 * max(1.0, abs(lo)), (it, lg, lo)
    for l in range(3):
        assert rel_err(m.tap(f"fc{l}", 0), o.tap(f"fc{l}", 0)) <= 1e-4
        assert rel_err(m.get(f"fc{l}.weights"), o.get(f"fc{l}.weights")) <= 1e-4
        assert rel_err(m.get(f"fc{l}.bias"), o.get(f"fc{l}.bias")) <= 1e-4
    pg = m.predict(None, b["X"], None, N, out_rows=10)
    po = o.predict(None, b["X"], None, N, out_rows=10)
    assert np.allclose(pg, po, rtol=1e-4, atol=1e-6)
    m.close()


def test_model_put_get_roundtrip(ps, ctx):
    m = ps.Model(ctx, "widedeep", 3, 4, 2, [8, 1], emb_capacity=1024, max_batch=16)
    w = np.arange(8 * 14, dtype=np.float32)              # fc0: out=8, in=3*4+2=14, column-major
    m.put("fc0.weights", w)
    assert np.array_equal(m.get("fc0.weights"), w)
    m.put("emF1.42.0", np.array([1, 2, 3, 4], np.float32))
    assert np.array_equal(m.get("emF1.42.0"), [1, 2, 3, 4])
    m.put("wide.weights.7.0", np.array([0.5], np.float32))
    assert m.get("wide.weights.7.0")[0] == 0.5
    assert m.get("emF1.43.0") is None and m.get("nope") is None
    m.close()


def test_pipelined_submit_equals_sync(ps, ctx):
    F, D, Xn, fc, N = 23, 16, 45, [32, 1], 128
    syn = Synth(F=F, Xn=Xn, V=3000, seed=21)
    batches = [syn.batch(N) for _ in range(9)]
    losses = []
    for mode in ("sync", "pipe", "pipe4"):
        c2 = ps.Context(0, seed=SEED)
        m = ps.Model(c2, "widedeep", F, D, Xn, fc, emb_capacity=1 << 15, max_batch=N)
        out = []
        if mode == "sync":
            for b in batches:
                out.append(m.train_step(b["E"], b["X"], b["W"], b["Y"]))
        else:
            import ctypes as C
            keep = [{k: np.ascontiguousarray(v) for k, v in b.items()} for b in batches]
            ptr = lambda a: a.ctypes.data_as(C.c_void_p)
            depth = 2 if mode == "pipe" else 4                  # the library stages four batches
            for i, b in enumerate(keep):
                m.submit_ptrs(ptr(b["E"]), ptr(b["X"]), ptr(b["W"]), ptr(b["Y"]), N)
                if i >= depth - 1:
                    out.append(m.collect())
                if mode == "pipe4" and i == 3:                  # a fifth step in flight is refused, nothing is enqueued
                    m.submit_ptrs(ptr(keep[4]["E"]), ptr(keep[4]["X"]), ptr(keep[4]["W"]), ptr(keep[4]["Y"]), N)
                    with pytest.raises(ps.PsError):
                        m.submit_ptrs(ptr(keep[5]["E"]), ptr(keep[5]["X"]), ptr(keep[5]["W"]), ptr(keep[5]["Y"]), N)
                    out.append(m.collect())
                    break
            while len(out) < (5 if mode == "pipe4" else len(keep)):
                out.append(m.collect())
        losses.append(out)
        m.close()
        c2.close()
    assert np.allclose(losses[0], losses[1], rtol=1e-5, atol=1e-7)
    assert np.allclose(losses[0][:5], losses[2], rtol=1e-5, atol=1e-7)


# --------------------------------------------------------------------------- TF32 tcgen05 path
@pytest.mark.parametrize("M,N,K", [(128, 64, 32), (256, 256, 413), (4096, 256, 256), (100, 1, 256), (37, 10, 50), (300, 150, 784), (1, 8, 7),
                                   (4096, 413, 256), (8192, 200, 96)])     # the last two take the 128-column tiles (more than one wave of 64-wide ones)
def test_tf32_gemm_matches_fp64(ps, ctx, M, N, K):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    Kp = (K + 3) // 4 * 4                                   # TMA needs 16 B-aligned rows
    Ap, Bp = np.zeros((M, Kp), np.float32), np.zeros((N, Kp), np.float32)
    Ap[:, :K], Bp[:, :K] = A, B
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    c32 = ctx.gemm_nt(ps.PS_FC_FP32, Ap, Bp)
    assert rel_err(c32, ref) <= 1e-5
    ctf = ctx.gemm_nt(ps.PS_FC_TF32, Ap, Bp)
    # TF32 keeps 10 mantissa bits of each operand (truncated): |err| <~ 2 * 2^-10 * sum|a||b|
    bound = 2.5e-3 * (np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T)
    assert np.all(np.abs(ctf - ref) <= bound + 1e-6), float(np.abs(ctf - ref).max())
    # and exact when the operands are representable in TF32
    At = (Ap.view(np.uint32) & 0xFFFFE000).view(np.float32)
    Bt = (Bp.view(np.uint32) & 0xFFFFE000).view(np.float32)
    ref_t = At.astype(np.float64) @ Bt.astype(np.float64).T
    assert rel_err(ctx.gemm_nt(ps.PS_FC_TF32, At, Bt), ref_t) <= 1e-5   # fp32 accumulation over K
    # 3xTF32 (error-compensated operand split on the tensor cores): fp32-grade on arbitrary operands
    c3 = ctx.gemm_nt(ps.PS_FC_TF32X3, Ap, Bp)
    assert np.all(np.abs(c3 - ref) <= 4e-6 * (np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T) + 1e-6), float(np.abs(c3 - ref).max())


def fro_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(1e-30, np.linalg.norm(b)))


@pytest.mark.parametrize("kind,F,D,Xn,fc,N,V", [
    ("widedeep", 23, 16, 45, [256, 256, 256, 1], 1024, 50000),     # BASELINE config 2 network
    ("dnn", 23, 10, 45, [150, 10, 1], 250, 3000),
    ("widedeep", 5, 8, 3, [16, 1], 37, 60),
])
def test_model_steps_match_oracle_tf32(ps, ctx, kind, F, D, Xn, fc, N, V):
    ctx.set_fc_precision(ps.PS_FC_TF32)
    m = ps.Model(ctx, kind, F, D, Xn, fc, emb_capacity=1 << 17, max_batch=N)
    o = ol.OracleModel(ol.KIND_WIDEDEEP if kind == "widedeep" else ol.KIND_DNN, F, D, Xn, fc, SEED)
    syn = Synth(F=F, Xn=Xn, V=V, seed=11)
    for it in range(3):
        b = syn.batch(N)
        W_before = [m.get(f"fc{l}.weights") for l in range(len(fc))] if it == 0 else None
        lg = m.train_step(b["E"], b["X"], b["W"], b["Y"])
        lo = o.train_step(b["E"], b["X"], b["W"], b["Y"])
        assert abs(lg - lo) <= 2e-2 * max(1.0, abs(lo)), (it, lg, lo)
        if it == 0:
            # same parameters on both sides: activations agree to TF32 accuracy element-wise; deltas are
            # compared in Frobenius norm (a ReLU whose pre-activation is ~0 may flip and move one element)
            for l in range(len(fc)):
                assert rel_err(m.tap(f"fc{l}", 0), o.tap(f"fc{l}", 0)) <= 2e-2, f"fc{l}.A"
                assert fro_err(m.tap(f"fc{l}", 0), o.tap(f"fc{l}", 0)) <= 5e-3, f"fc{l}.A"
                assert fro_err(m.tap(f"fc{l}", 1), o.tap(f"fc{l}", 1)) <= 5e-2, f"fc{l}.delta"
            # and each dgrad GEMM against fp64 on ITS OWN inputs, to the TF32 truncation bound
            for l in range(len(fc)):
                d_in = (m.tap(f"fc{l + 1}", 1) if l + 1 < len(fc) else m.tap("addWideDeep", 1) if kind == "widedeep" else None)
                if d_in is None:
                    continue
                out, inn = fc[l], W_before[l].size // fc[l]
                Wm = W_before[l].reshape(inn, out).T.astype(np.float64)          # out x in (column-major on the wire)
                d_in = d_in.reshape(N, out).astype(np.float64)
                exp = d_in @ Wm
                if l > 0:
                    exp = exp * (m.tap(f"fc{l - 1}", 0).reshape(N, inn) > 0)
                got = m.tap(f"fc{l}", 1).reshape(N, inn)
                bound = 2.5e-3 * (np.abs(d_in) @ np.abs(Wm)) + 1e-9
                assert np.all(np.abs(got - exp) <= bound), (l, float(np.abs(got - exp).max()))
    # Adam's first steps move every weight by ~alfa regardless of |g|, so sign flips of tiny
    # gradients are visible: compare the bulk, not the worst element
    for l in range(len(fc)):
        wg, wo = m.get(f"fc{l}.weights"), o.get(f"fc{l}.weights")
        assert np.mean(np.abs(wg - wo)) <= 2e-3 * np.mean(np.abs(wo)) + 1e-4, f"fc{l}.weights"
    m.close()


def test_fcnn_tf32(ps, ctx):
    ctx.set_fc_precision(ps.PS_FC_TF32)
    Xn, fc, N = 784, [150, 50, 10], 1024               # BASELINE config 5
    m = ps.Model(ctx, "fcnn", 0, 0, Xn, fc, max_batch=N)
    o = ol.OracleModel(ol.KIND_FCNN, 0, 0, Xn, fc, SEED)
    b = Synth(F=0, Xn=Xn, V=0, seed=17, n_classes=10).batch(N)
    lg = m.train_step(None, b["X"], None, b["Y"])
    lo = o.train_step(None, b["X"], None, b["Y"])
    assert abs(lg - lo) <= 1e-3 * max(1.0, abs(lo))
    for l in range(3):
        assert rel_err(m.tap(f"fc{l}", 0), o.tap(f"fc{l}", 0)) <= 2e-2
        assert fro_err(m.tap(f"fc{l}", 1), o.tap(f"fc{l}", 1)) <= 5e-2
    m.close()


# --------------------------------------------------------------------------- store dump / load
def test_model_save_load_roundtrip(ps, ctx, tmp_path):
    """Every key of the store with its updater state survives save -> load into a fresh model (of a different table capacity),
    bit for bit, and the two models then take the same next step."""
    kind, F, D, Xn, fc, N = "widedeep", 23, 16, 45, [64, 32, 1], 256
    ctx.set_fc_precision(ps.PS_FC_FP32)
    a = ps.Model(ctx, kind, F, D, Xn, fc, emb_capacity=1 << 15, max_batch=N)
    syn = Synth(F=F, Xn=Xn, V=5000, seed=23)
    batches = [syn.batch(N) for _ in range(4)]
    for b in batches[:3]:
        a.train_step(b["E"], b["X"], b["W"], b["Y"])
    path = str(tmp_path / "store.psb")
    a.save(path)
    b2 = ps.Model(ctx, kind, F, D, Xn, fc, emb_capacity=1 << 16, max_batch=N)
    with pytest.raises(ps.PsError) as e:
        b2.load(str(tmp_path / "missing.psb"))
    assert e.value.code == 204
    b2.load(path)
    assert b2.num_keys() == a.num_keys()
    names = [f"fc{l}.{p}" for l in range(len(fc)) for p in ("weights", "bias")] + ["wide.bias"]
    last = batches[2]
    names += [ol.key_string(0, j, int(last["E"][n, j])) for n in range(0, N, 5) for j in range(0, F, 3)]
    names += [ol.key_string(1, 0, int(last["W"][n, 2])) for n in range(0, N, 9)]
    for k in names:
        va, vb = a.get(k), b2.get(k)
        assert va is not None and np.array_equal(va.view(np.uint32), vb.view(np.uint32)), k
        for which in (0, 1):
            sa, sb = a.get_state(k, which), b2.get_state(k, which)
            assert (sa is None) == (sb is None) and (sa is None or np.array_equal(sa.view(np.uint32), sb.view(np.uint32))), (k, which)
    nb = batches[3]
    la = a.train_step(nb["E"], nb["X"], nb["W"], nb["Y"])
    lb = b2.train_step(nb["E"], nb["X"], nb["W"], nb["Y"])
    assert abs(la - lb) <= 1e-6 * max(1.0, abs(la)), (la, lb)
    assert rel_err(b2.get("fc0.weights"), a.get("fc0.weights")) <= 1e-6
    with pytest.raises(ps.PsError):                       # a model that already holds rows refuses to load
        b2.load(path)
    c = ps.Model(ctx, "dnn", F, D, Xn, fc, emb_capacity=1 << 15, max_batch=N)
    with pytest.raises(ps.PsError):                       # shape mismatch
        c.load(path)
    for m in (a, b2, c):
        m.close()

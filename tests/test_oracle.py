"""Pins the CPU oracle (oracle/ps_oracle.cpp).  The reference cannot run here and its own tests
assert nothing, so the pins are: derived known answers (Java String.hashCode
values, the TestAuc vector's AUC, the updater-name grammar of T/TestPs.java:26-27), closed forms
that follow from the cited Java lines, and an INDEPENDENT numpy re-derivation of a whole
WideDeepNN / DNN step written from the Java sources, not from the C++ restatement."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from ps_b200.synth import Synth

HERE = os.path.dirname(os.path.abspath(__file__))


def test_java_hash_known_answers():
    L = ol.lib()
    kat = {"emF2.15757.0": 1476472925, "emF13.28305.0": -999230970, "fc0.weights": -1943542496,
           "fc0.bias": 761250996, "wide.bias": -534120204, "": 0, "a": 97}
    for s, h in kat.items():
        assert L.pso_java_hash(s.encode()) == h
    # net/Mod.java:14: Java % truncates toward zero → negative shard for negative hashCode
    assert L.pso_router_mod(b"emF13.28305.0", 8) == -(999230970 % 8)
    assert L.pso_router_floormod(b"emF13.28305.0", 8) == (-999230970) % 8


def test_key_spelling_matches_java_float_to_string():
    assert ol.key_string(0, 2, 15757) == "emF2.15757.0"
    assert ol.key_string(0, 0, 0) == "emF0.0.0"
    assert ol.key_string(0, 7, 9999999) == "emF7.9999999.0"
    assert ol.key_string(0, 7, 10000000) == "emF7.1.0E7"
    assert ol.key_string(0, 7, 16777216) == "emF7.1.6777216E7"
    assert ol.key_string(1, 0, 99999) == "wide.weights.99999.0"


def test_auc_known_answer():
    g = np.load(os.path.join(HERE, "golden", "testauc_vector.npz"))
    auc = ol.lib().pso_auc(g["p"].astype(np.float32), g["y"].astype(np.float32), len(g["p"]))
    assert abs(auc - float(g["auc"])) < 1e-12


def test_updater_names():
    import ctypes as C
    buf = C.create_string_buffer(256)
    ol.lib().pso_updater_name(0, 0.005, 0.9, 0.999, 1e-8, buf, 256)
    assert buf.value.decode().startswith("adam@alfa:0.005@beta1:0.9@beta2:0.999@epsilon:1e-08@")


def test_adam_first_step_identity():
    """m/(1-b1) = g and v/(1-b2) = g^2 on the first step, so dw = -alfa * g / (|g| + eps) (AdamUpdater.java:61-69)."""
    rng = np.random.default_rng(0)
    n = 1000
    w = rng.standard_normal(n).astype(np.float32)
    g = rng.standard_normal(n).astype(np.float32)
    m, v, w1 = np.zeros(n, np.float32), np.zeros(n, np.float32), w.copy()
    ol.lib().pso_adam_update(w1, m, v, g, n, 0.005, 0.9, 0.999, 1e-8)
    assert np.allclose(w1 - w, -0.005 * g / (np.abs(g) + 1e-8), rtol=2e-3, atol=1e-7)


def test_ftrl_formula():
    """FtrlUpdater.java:64-74 written out in numpy (float32 arithmetic)."""
    rng = np.random.default_rng(1)
    n = 512
    f = np.float32
    w = rng.standard_normal(n).astype(f)
    z = (rng.standard_normal(n) * 0.01).astype(f)
    nn = np.abs(rng.standard_normal(n)).astype(f)
    g = rng.standard_normal(n).astype(f)
    a, b, l1, l2 = f(0.005), f(1.0), f(0.001), f(0.001)
    sign = np.where(z >= 0, f(1), f(-1))
    wn = np.where(np.abs(z) <= l1, f(0), -(z - sign * l1) / ((l2 + (b + np.sqrt(nn))) / a)).astype(f)
    s = (np.sqrt(nn + g * g) - np.sqrt(nn / a)).astype(f)
    zn = (z + (g - s * wn)).astype(f)
    n2 = (nn + g * g).astype(f)
    wo, zo, no = w.copy(), z.copy(), nn.copy()
    ol.lib().pso_ftrl_update(wo, zo, no, g, n, a, b, l1, l2)
    assert np.allclose(wo, wn, rtol=1e-6, atol=1e-9) and np.allclose(zo, zn, rtol=1e-6, atol=1e-7) and np.allclose(no, n2, rtol=1e-6)


def test_embedding_geff_closed_form_in_oracle():
    """two EmbeddingLayer.backward calls per step + aliasing ⇒ g_eff = S(n+1)/(2n^2)."""
    F, D, N = 2, 4, 7
    L = ol.lib()
    o = L.pso_emb_create(F, D, 5, 0)
    E = np.array([[3, 10], [3, 11], [3, 10], [4, 10], [3, 12], [4, 10], [3, 10]], np.int64)
    A = np.zeros((N, F * D), np.float32)
    L.pso_emb_forward(o, E, N, A.reshape(-1))
    om = ol.OracleModel.__new__(ol.OracleModel)
    om.L, om.h = L, o
    before = {k: om.get(k).copy() for k in ("emF0.3.0", "emF0.4.0", "emF1.10.0", "emF1.11.0", "emF1.12.0")}
    delta = np.random.default_rng(2).standard_normal((N, F * D)).astype(np.float32)
    L.pso_emb_backward_update(o, delta.reshape(-1), F * D, N, 2)
    for key, w0 in before.items():
        j = int(key[3])
        idv = int(float(key.split(".", 1)[1]))
        rows = np.where(E[:, j] == idv)[0]
        n = len(rows)
        S = (delta[rows, j * D:(j + 1) * D] * (A[rows, j * D:(j + 1) * D] > 0)).sum(0)
        g = S * (n + 1) / (2.0 * n * n)
        exp = w0 - 0.005 * g / (np.abs(g) + 1e-8)
        assert np.allclose(om.get(key), exp, rtol=2e-4, atol=1e-6), key
    om.h = None
    L.pso_model_destroy(o)


# --------------------------------------------------------------------------- independent numpy step
def _sigmoid(x):
    return (np.float64(np.float32(0.001)) + np.float64(np.float32(.999) - np.float32(0.001)) / (1.0 + np.exp(-x.astype(np.float64)))).astype(np.float32)


def _numpy_widedeep_step(o, kind, F, D, Xn, fc, b):
    """One Trainer step written from the Java sources with numpy matrices (features x N), reading
    the CURRENT parameters from the oracle store and returning what the next parameters must be
    for the dense keys, plus loss and the top-level activations."""
    N = b["Y"].shape[0]
    E, X, Y = b["E"], b["X"].T.astype(np.float32), b["Y"].astype(np.float32)
    emb = np.zeros((F * D, N), np.float32)
    for j in range(F):                                      # EmbeddingField.forward
        for n in range(N):
            emb[j * D:(j + 1) * D, n] = o.get(ol.key_string(0, j, int(E[n, j])))
    emb = np.maximum(emb, 0)
    A = [np.concatenate([emb, X], 0)]                       # ConcatLayer.forward
    Ws, bs = [], []
    for l, out in enumerate(fc):                            # FcLayer.forward
        W = o.get(f"fc{l}.weights").reshape(A[-1].shape[0], out).T      # column-major out x in
        bias = o.get(f"fc{l}.bias")
        Z = (W.astype(np.float64) @ A[-1].astype(np.float64)).astype(np.float32) + bias[:, None]
        last = l == len(fc) - 1
        if not last:
            Z = np.maximum(Z, 0)
        elif kind == "dnn":
            Z = _sigmoid(Z)
        A.append(Z.astype(np.float32))
        Ws.append(W)
        bs.append(bias)
    if kind == "widedeep":
        wz = np.zeros(N, np.float32)
        for n in range(N):                                  # LRLayer.forward
            s = np.float32(0)
            for j in range(F):
                w = o.get(ol.key_string(1, 0, int(b["W"][n, j])))
                s = np.float32(s + (w[0] if w is not None else np.float32(0)))
            wz[n] = s
        wz = wz + o.get("wide.bias")[0]
        P = _sigmoid(A[-1][0] + wz)                         # AddLayer + Sigmoid
    else:
        P = A[-1][0]
    loss = np.float32(np.sum((-Y * np.log(P.astype(np.float64)) - (1 - Y) * np.log((1 - P).astype(np.float64))).astype(np.float32)) / N)
    d = ((P - Y) / (P * (1 - P))) * (P * (1 - P))           # CrossEntropy.backward then Sigmoid.backward
    d = d[None, :].astype(np.float32)
    grads = {}
    for l in range(len(fc) - 1, -1, -1):                    # FcLayer.backward
        grads[f"fc{l}.bias"] = d.mean(1)
        grads[f"fc{l}.weights"] = (d.astype(np.float64) @ A[l].T.astype(np.float64) / N).astype(np.float32)
        dprev = (Ws[l].T.astype(np.float64) @ d.astype(np.float64)).astype(np.float32)
        if l > 0:
            dprev = dprev * (A[l] > 0)
        d = dprev
    return loss, P, grads, d, A


@pytest.mark.parametrize("kind", ["dnn", "widedeep"])
def test_oracle_step_against_independent_numpy(kind):
    F, D, Xn, fc, N = 4, 3, 2, [6, 5, 1], 9
    o = ol.OracleModel(ol.KIND_WIDEDEEP if kind == "widedeep" else ol.KIND_DNN, F, D, Xn, fc, 42)
    syn = Synth(F=F, Xn=Xn, V=30, seed=3)
    b0 = syn.batch(N)
    o.train_step(b0["E"], b0["X"], b0["W"], b0["Y"])        # creates every key; makes Adam state non-trivial
    for _ in range(2):
        b = dict(b0)                                        # same ids so every key exists before the numpy forward
        b["X"], b["Y"] = syn.batch(N)["X"], syn.batch(N)["Y"]
        loss_np, P, grads, d0, A = _numpy_widedeep_step(o, kind, F, D, Xn, fc, b)
        before = {k: o.get(k).copy() for k in grads}
        m_before = {k: o.get_state(k, 0).copy() for k in grads}
        v_before = {k: o.get_state(k, 1).copy() for k in grads}
        loss_o = o.train_step(b["E"], b["X"], b["W"], b["Y"])
        assert abs(loss_o - loss_np) < 1e-5
        top = "addWideDeep" if kind == "widedeep" else f"fc{len(fc) - 1}"
        assert np.allclose(o.tap(top, 0), P, rtol=1e-5, atol=1e-7)
        assert np.allclose(o.tap("fc0", 1).reshape(N, -1).T, d0, rtol=1e-4, atol=1e-6)
        for k, g in grads.items():                          # AdamUpdater.update with the numpy gradient
            g = g.reshape(-1, order="F").astype(np.float64)
            m = 0.9 * m_before[k] + (1 - np.float32(0.9)) * g
            v = 0.999 * v_before[k] + (1 - np.float32(0.999)) * g * g
            exp = before[k] - 0.005 * (m / (1 - np.float32(0.9))) / (np.sqrt(v / (1 - np.float32(0.999))) + 1e-8)
            assert np.allclose(o.get(k), exp, rtol=2e-4, atol=2e-6), k


def test_fcnn_runs_and_learns():
    """Mnist.java-shaped plumbing: loss decreases on a fixed batch (README.md:29 anchors only the trend here)."""
    Xn, fc, N = 20, [16, 8, 4], 64
    o = ol.OracleModel(ol.KIND_FCNN, 0, 0, Xn, fc, 3)
    rng = np.random.default_rng(0)
    X = rng.random((N, Xn)).astype(np.float32) * 255
    Y = rng.integers(0, 4, N).astype(np.float32)
    losses = [o.train_step(None, X, None, Y) for _ in range(30)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]


def _numpy_fcnn_step(o, Xn, fc, X, Y):
    """One FullConnectedNN Trainer step (FullConnectedNN.java:37-68) written from the Java sources: FcLayer/ReLU stack, Softmax(10000)
    (Softmax.java:21-45: divide by the scale in place, shift by the column max, exp in double, clip exact 0/1), SoftmaxLoss
    (SoftmaxLoss.java:9-28), Softmax.backward (:47-67, which does NOT undo the 1/scale) and FcLayer.backward."""
    N = Y.shape[0]
    A = [X.T.astype(np.float32)]
    Ws = []
    for l, out in enumerate(fc):
        W = o.get(f"fc{l}.weights").reshape(A[-1].shape[0], out).T
        bias = o.get(f"fc{l}.bias")
        Z = (W.astype(np.float64) @ A[-1].astype(np.float64)).astype(np.float32) + bias[:, None]
        if l < len(fc) - 1:
            Z = np.maximum(Z, 0)
        else:
            Z = (Z / np.float32(10000)).astype(np.float32)
            ex = np.exp((Z - Z.max(0, keepdims=True)).astype(np.float64)).astype(np.float32)
            Z = (ex / ex.sum(0, keepdims=True, dtype=np.float32)).astype(np.float32)
            Z = np.where(Z == 0, np.float32(0.001), np.where(Z == 1, np.float32(0.999), Z)).astype(np.float32)
        A.append(Z.astype(np.float32))
        Ws.append(W)
    P = A[-1]
    hot = Y.astype(np.int64)
    p = P[hot, np.arange(N)]
    loss = np.float32(np.sum(-np.log(p.astype(np.float64))) / N)
    d = np.zeros_like(P, dtype=np.float64)
    for i in range(N):                                          # Softmax.backward on dy = -1/p at the hot class only
        j = hot[i]
        dy = -1.0 / np.float64(p[i])
        for k in range(P.shape[0]):
            d[k, i] = (P[k, i] * (1 - P[k, i]) if k == j else -P[j, i] * P[k, i]) * dy
    d = d.astype(np.float32)
    grads = {}
    for l in range(len(fc) - 1, -1, -1):
        grads[f"fc{l}.bias"] = d.mean(1)
        grads[f"fc{l}.weights"] = (d.astype(np.float64) @ A[l].T.astype(np.float64) / N).astype(np.float32)
        dprev = (Ws[l].T.astype(np.float64) @ d.astype(np.float64)).astype(np.float32)
        if l > 0:
            dprev = dprev * (A[l] > 0)
        d = dprev
    return loss, P, grads


def test_oracle_fcnn_step_against_independent_numpy():
    """Pins the cfg5 (Mnist.java) path of the oracle the way the DNN / WideDeepNN step is pinned above."""
    Xn, fc, N = 7, [6, 5, 4], 11
    o = ol.OracleModel(ol.KIND_FCNN, 0, 0, Xn, fc, 9)
    rng = np.random.default_rng(4)
    X0 = (rng.random((N, Xn)) * 255).astype(np.float32)
    Y0 = rng.integers(0, 4, N).astype(np.float32)
    o.train_step(None, X0, None, Y0)                            # non-trivial Adam state
    for _ in range(2):
        X = (rng.random((N, Xn)) * 255).astype(np.float32)
        Y = rng.integers(0, 4, N).astype(np.float32)
        loss_np, P, grads = _numpy_fcnn_step(o, Xn, fc, X, Y)
        before = {k: o.get(k).copy() for k in grads}
        m_before = {k: o.get_state(k, 0).copy() for k in grads}
        v_before = {k: o.get_state(k, 1).copy() for k in grads}
        loss_o = o.train_step(None, X, None, Y)
        assert abs(loss_o - loss_np) < 1e-5, (loss_o, loss_np)
        assert np.allclose(o.tap(f"fc{len(fc) - 1}", 0).reshape(N, -1).T, P, rtol=1e-5, atol=1e-7)
        for k, g in grads.items():
            g = g.reshape(-1, order="F").astype(np.float64)
            m = 0.9 * m_before[k] + (1 - np.float32(0.9)) * g
            v = 0.999 * v_before[k] + (1 - np.float32(0.999)) * g * g
            exp = before[k] - 0.005 * (m / (1 - np.float32(0.9))) / (np.sqrt(v / (1 - np.float32(0.999))) + 1e-8)
            assert np.allclose(o.get(k), exp, rtol=2e-4, atol=2e-6), k


def test_early_exit_freezes_the_store():
    """DNN.java:58-63: when loss <= CrossEntropy.slim (0.01) train() returns before backward — nothing reaches KVStore.sum, so
    Trainer's update() changes nothing.  All-positive labels drive the clipped sigmoid to 0.999 (loss -> 0.001)."""
    F, D, Xn, fc, N = 1, 2, 1, [4, 1], 8
    o = ol.OracleModel(ol.KIND_DNN, F, D, Xn, fc, 5)
    E = np.arange(N, dtype=np.int64).reshape(N, 1) % 3
    X = np.full((N, Xn), 0.5, np.float32)
    Y = np.ones(N, np.float32)
    hit = None
    for step in range(4000):
        loss = o.train_step(E, X, E % 100000, Y)
        if o.skipped_backward():
            hit = step
            break
    assert hit is not None and loss <= 0.01
    snap = {k: o.get(k).copy() for k in ("fc0.weights", "fc0.bias", "fc1.weights", "fc1.bias", ol.key_string(0, 0, 1))}
    loss2 = o.train_step(E, X, E % 100000, Y)
    assert o.skipped_backward() and loss2 == loss
    for k, v in snap.items():
        assert np.array_equal(o.get(k), v), k


def test_wide_gradient_reaches_every_key_ever_seen():
    """SURVEY quirk 7 (LRLayer.java:79,110-117): LRLayer.weights is never cleared, and backward pushes the SAME batch-mean delta
    to every key in it — a wide key absent from the current batch still moves."""
    F, D, Xn, fc, N = 2, 2, 1, [3, 1], 6
    o = ol.OracleModel(ol.KIND_WIDEDEEP, F, D, Xn, fc, 8)
    rng = np.random.default_rng(2)
    X = rng.random((N, Xn)).astype(np.float32)
    Y = (rng.random(N) < 0.5).astype(np.float32)
    E1 = np.array([[1, 2]] * N, np.int64)
    E2 = np.array([[3, 4]] * N, np.int64)
    o.train_step(E1, X, E1, Y)                                  # creates wide.weights.1.0 and .2.0
    k_old, k_new = ol.key_string(1, 0, 1), ol.key_string(1, 0, 3)
    assert o.get(k_new) is None
    before = o.get(k_old).copy()
    z_before = o.get_state(k_old, 0)
    o.train_step(E2, X, E2, Y)                                  # batch 2 does not contain id 1
    assert o.get(k_new) is not None
    z_after = o.get_state(k_old, 0)
    moved = (not np.array_equal(o.get(k_old), before)) or (z_before is None) != (z_after is None) or (z_before is not None and not np.array_equal(z_before, z_after))
    assert moved, "a key seen only in batch 1 must still receive batch 2's gradient (Ftrl state Z accumulates it)"
    assert o.num_keys() == 2 * len(fc) + 1 + 4 + 4             # dense keys + wide.bias + 4 wide keys + 4 embedding rows


@pytest.mark.parametrize("R", [2, 4, 8])
def test_router_balance_keeps_buckets_within_capacity(R):
    """The sharded exchange sizes every per-owner bucket as slack x L / R with slack = 2 (ps_b200/sharded.py): the production router
    (ps_owner_of: splitmix64 of the packed key, Lemire reduction) must spread the UNIQUE keys of a bench batch evenly enough —
    senders de-duplicate, so unique keys are what travels."""
    from ps_b200.synth import CONFIGS
    c = CONFIGS["cfg2"]
    syn = Synth(F=c["F"], Xn=c["Xn"], V=c["V"], dist="zipf", seed=20261017 + 2)
    L = ol.lib()
    for _ in range(3):
        E = syn.batch(c["B"])["E"]
        keys = {int(L.pso_pack_key(j, int(v))) for j in range(c["F"]) for v in np.unique(E[:, j])}
        owners = np.array([L.pso_owner_of(k, R) for k in keys])
        counts = np.bincount(owners, minlength=R)
        assert counts.min() > 0 and counts.max() <= 1.15 * len(keys) / R, counts
        assert counts.max() <= 2 * c["B"] * c["F"] / R                  # the bucket capacity at slack 2

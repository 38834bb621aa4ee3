"""Parity at BASELINE.json's FULL sizes (the other GPU tests use sizes the per-key Python comparisons finish quickly on):
whole-model steps against the oracle where the oracle is fast enough (cfg2, cfg5), closed forms / spot checks where it is not
(cfg3 and cfg4 embedding shapes)."""
import numpy as np
import pytest

import oracle_lib as ol
from ps_b200.synth import CONFIGS, Synth

pytestmark = pytest.mark.gpu
SEED = 20261017


def rel_err(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(1e-12, np.abs(np.asarray(b, np.float64)).max()))


def test_cfg2_full_size_steps_match_oracle(ps, ctx):
    """configs[1]: WideDeepNN, 1 M-key vocab, D = 16, 3 x FC[256], batch 4096, Adam — the bench workload itself, 3 Trainer steps."""
    c = CONFIGS["cfg2"]
    ctx.set_fc_precision(ps.PS_FC_TF32X3)
    m = ps.Model(ctx, c["kind"], c["F"], c["D"], c["Xn"], c["fc"], emb_capacity=2 * c["V"] + (1 << 16), max_batch=c["B"])
    o = ol.OracleModel(ol.KIND_WIDEDEEP, c["F"], c["D"], c["Xn"], c["fc"], SEED)
    if ol.openblas_path():
        ol.lib().pso_set_gemm(2, ol.openblas_path().encode())
    try:
        syn = Synth(F=c["F"], Xn=c["Xn"], V=c["V"], dist="zipf", seed=SEED + 2)
        for it in range(3):
            b = syn.batch(c["B"])
            lg = m.train_step(b["E"], b["X"], b["W"], b["Y"])
            lo = o.train_step(b["E"], b["X"], b["W"], b["Y"])
            assert abs(lg - lo) <= 1e-4 * max(1.0, abs(lo)), (it, lg, lo)
    finally:
        ol.lib().pso_set_gemm(0, None)
    assert m.num_keys() == o.num_keys()
    # Adam's early steps are sign-like, dw = -alfa * g / (|g| + 1e-8): where |g| is comparable to epsilon a reassociated sum over the
    # 4096 samples moves the step by O(alfa).  So: every element within the 3 steps' reach, and all but a sliver within fp32 noise.
    for l in range(len(c["fc"])):
        for nm in (f"fc{l}.weights", f"fc{l}.bias"):
            wg, wo = m.get(nm), o.get(nm)
            d = np.abs(wg.astype(np.float64) - wo)
            assert d.max() <= 3 * 2 * 0.005, nm
            assert np.mean(d > 1e-4 * np.abs(wo).max()) <= 5e-3, (nm, float(np.mean(d > 1e-4 * np.abs(wo).max())))
    rng = np.random.default_rng(3)
    bad = 0
    for n in rng.integers(0, c["B"], 200):
        j = int(rng.integers(0, c["F"]))
        key = ol.key_string(0, j, int(b["E"][n, j]))
        wg, wo = m.get(key), o.get(key)
        bad += int(not np.allclose(wg, wo, rtol=2e-3, atol=2e-5))
    assert bad <= 4, bad                                              # Adam's sign-like first steps amplify 1-ulp gradient differences near zero
    m.close()


def test_cfg5_full_size_steps_match_oracle(ps, ctx):
    """configs[4]: Mnist.java FullConnectedNN 784 -> 150 -> 50 -> 10, batch 1024, dense only (FcLayer tcgen05 path)."""
    c = CONFIGS["cfg5"]
    ctx.set_fc_precision(ps.PS_FC_TF32X3)
    m = ps.Model(ctx, "fcnn", 0, 0, c["Xn"], c["fc"], max_batch=c["B"])
    o = ol.OracleModel(ol.KIND_FCNN, 0, 0, c["Xn"], c["fc"], SEED)
    syn = Synth(F=0, Xn=c["Xn"], V=0, seed=17, n_classes=10)
    for it in range(3):
        b = syn.batch(c["B"])
        lg = m.train_step(None, b["X"], None, b["Y"])
        lo = o.train_step(None, b["X"], None, b["Y"])
        assert abs(lg - lo) <= 1e-4 * max(1.0, abs(lo)), (it, lg, lo)
    for l in range(3):
        wg, wo = m.get(f"fc{l}.weights"), o.get(f"fc{l}.weights")
        d = np.abs(wg.astype(np.float64) - wo)
        assert d.max() <= 3 * 2 * 0.005 and np.mean(d > 1e-4 * np.abs(wo).max()) <= 5e-3, l
    m.close()


def test_cfg4_shape_embedding_closed_form(ps, ctx):
    """configs[3] embedding shapes (10 M-key vocab, D = 64, batch 16384): gather bit-exact against the stored rows, and the
    scatter-add + g_eff + update against the closed form w' = w - eta * S(n+1)/(2n^2) computed in float64."""
    c = CONFIGS["cfg4"]
    F, D, N, V = c["F"], c["D"], c["B"], c["V"]
    eta = 0.5
    emb = ps.EmbeddingLayer(ctx, F, D, capacity=1 << 22, updater=ps.UpdaterSpec.simple(eta))
    syn = Synth(F=F, Xn=1, V=V, dist="zipf", seed=SEED + 4)
    rng = np.random.default_rng(5)
    for it in range(2):
        E = syn.batch(N)["E"]
        out = emb.forward(E)                                            # (N, F*D), ReLU applied
        keys = E + (np.arange(F, dtype=np.int64) << 44)[None, :]
        uniq, inv, cnt = np.unique(keys.reshape(-1), return_inverse=True, return_counts=True)
        fields, ids = (uniq >> 44).astype(np.int32), uniq & ((1 << 44) - 1)
        w0, found = emb.get_rows(fields, ids)
        assert found.all()
        assert np.array_equal(out.reshape(N * F, D).view(np.uint32), np.maximum(w0, 0)[inv].view(np.uint32))   # copy + max: bit-exact
        delta = rng.standard_normal((N, F * D)).astype(np.float32)
        emb.backward_update(delta, calls=2)
        g = (delta.reshape(N * F, D) * (out.reshape(N * F, D) > 0)).astype(np.float64)
        S = np.zeros((len(uniq), D))
        np.add.at(S, inv, g)
        n = cnt[:, None].astype(np.float64)
        exp = w0 - eta * S * (n + 1) / (2 * n * n)
        w1, _ = emb.get_rows(fields, ids)
        assert np.abs(w1 - exp).max() <= 2e-5 * max(1.0, np.abs(exp).max()), it
        assert emb.size() >= len(uniq)
    emb.close()


def test_cfg3_shape_embedding_ftrl_matches_oracle(ps, ctx):
    """configs[2] embedding shapes on one shard (100 M-key vocab, D = 32, batch 8192, Ftrl on the rows): two steps against
    the oracle, every touched row and both Ftrl states."""
    c = CONFIGS["cfg3"]
    F, D, N, V = c["F"], c["D"], c["B"], c["V"]
    emb = ps.EmbeddingLayer(ctx, F, D, capacity=1 << 20, updater=ps.UpdaterSpec.ftrl())
    o = ol.lib().pso_emb_create(F, D, SEED, 1)
    syn = Synth(F=F, Xn=1, V=V, dist="zipf", seed=SEED + 3)
    rng = np.random.default_rng(7)
    seen = set()
    for it in range(2):
        E = syn.batch(N)["E"]
        out_g = emb.forward(E)
        out_o = np.zeros((N, F * D), np.float32)
        ol.lib().pso_emb_forward(o, np.ascontiguousarray(E), N, out_o.reshape(-1))
        assert rel_err(out_g, out_o) <= 2e-5, it
        delta = rng.standard_normal((N, F * D)).astype(np.float32)
        emb.backward_update(delta, calls=2)
        ol.lib().pso_emb_backward_update(o, delta.reshape(-1), F * D, N, 2)
        for j in range(F):
            seen.update((j, int(v)) for v in np.unique(E[::8, j]))
    keys = sorted(seen)
    fields, ids = np.array([k[0] for k in keys], np.int32), np.array([k[1] for k in keys], np.int64)
    w, s1, s2, found = emb.get_rows(fields, ids, state=True)
    assert found.all()
    om = ol.OracleModel.__new__(ol.OracleModel)
    om.L, om.h = ol.lib(), o
    worst = 0.0
    for i, (j, v) in enumerate(keys):
        key = ol.key_string(0, j, v)
        wo = om.get(key)
        worst = max(worst, float(np.abs(w[i] - wo).max() / max(1e-6, np.abs(wo).max())))
        s1o = om.get_state(key, 0)
        if s1o is not None:
            assert np.allclose(s1[i], s1o, rtol=5e-5, atol=1e-6), key
    assert worst <= 5e-5, worst
    om.h = None
    ol.lib().pso_model_destroy(o)
    emb.close()

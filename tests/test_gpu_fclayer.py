"""layer.FcLayer as a standalone operator (ps_fc_*) against a numpy restatement of
FcLayer.java:74-110 and the oracle's updater (update/AdamUpdater.java:57-70)."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def act_fwd(kind, z):
    if kind == 1:
        return np.maximum(z, 0)
    if kind == 2:                                      # activations/Sigmoid.java:9-14 (exp in double)
        return (0.001 + (np.float32(.999) - np.float32(0.001)).astype(np.float64) / (1.0 + np.exp(-z.astype(np.float64)))).astype(np.float32)
    return z


def act_bwd(kind, d, y):
    if kind == 1:
        return d * (y > 0)
    if kind == 2:
        return d * (y * (1 - y))
    return d


@pytest.mark.parametrize("mode", ["fp32", "tf32x3"])
@pytest.mark.parametrize("in_dims,out_dims,N,act", [(413, 256, 512, 1), (7, 3, 5, 2), (784, 150, 300, 1), (50, 10, 64, 0), (256, 1, 100, 2)])
def test_fclayer_forward_backward_update(ps, ctx, mode, in_dims, out_dims, N, act):
    tol = 2e-5 if mode == "fp32" else 5e-5
    ctx.set_fc_precision(ps.PS_FC_FP32 if mode == "fp32" else ps.PS_FC_TF32X3)
    fc = ps.FcLayer(ctx, "fc0", in_dims, out_dims, act=act, max_batch=N)
    W, b = fc.get(0).astype(np.float64), fc.get(1).astype(np.float64)
    bound = 4 * np.sqrt(6) / np.sqrt(in_dims + out_dims)            # FcLayer.java:39
    assert W.shape == (out_dims, in_dims) and np.abs(W).max() <= bound and np.abs(W).max() > 0.5 * bound
    rng = np.random.default_rng(in_dims + out_dims)
    for step in range(2):
        A_prev = rng.random((N, in_dims)).astype(np.float32)
        A = fc.forward(A_prev)
        Z = A_prev.astype(np.float64) @ W.T + b
        A_ref = act_fwd(act, Z.astype(np.float32))
        assert np.abs(A - A_ref).max() <= tol * max(1.0, np.abs(A_ref).max())
        delta = rng.standard_normal((N, out_dims)).astype(np.float32)
        dprev = fc.backward(delta)
        d = act_bwd(act, delta, A).astype(np.float64)
        dprev_ref = d @ W
        assert np.abs(dprev - dprev_ref).max() <= tol * max(1.0, np.abs(dprev_ref).max())
        dW, db = fc.gradients()
        dW_ref, db_ref = d.T @ A_prev.astype(np.float64) / N, d.mean(0)
        assert np.abs(dW - dW_ref).max() <= tol * max(1.0, np.abs(dW_ref).max())
        assert np.abs(db - db_ref).max() <= tol * max(1.0, np.abs(db_ref).max())
        with pytest.raises(ps.PsError):                               # one pending KVStore.sum per update
            fc.backward(delta)
        fc.update()
        W1, b1 = fc.get(0), fc.get(1)
        if step == 0:                                                 # Adam's first step from zero state: the oracle's updater on the GPU's own gradient
            wo, m, v = W.astype(np.float32).reshape(-1).copy(), np.zeros(W.size, np.float32), np.zeros(W.size, np.float32)
            ol.lib().pso_adam_update(wo, m, v, np.ascontiguousarray(dW.reshape(-1)), W.size, 0.005, 0.9, 0.999, 1e-8)
            assert np.abs(W1.reshape(-1) - wo).max() <= 1e-6
            bo, m, v = b.astype(np.float32).copy(), np.zeros(out_dims, np.float32), np.zeros(out_dims, np.float32)
            ol.lib().pso_adam_update(bo, m, v, np.ascontiguousarray(db), out_dims, 0.005, 0.9, 0.999, 1e-8)
            assert np.abs(b1 - bo).max() <= 1e-6
        W, b = W1.astype(np.float64), b1.astype(np.float64)
    with pytest.raises(ps.PsError):
        fc.update()                                                   # nothing pending
    fc.close()


def test_fclayer_put_get_roundtrip(ps, ctx):
    fc = ps.FcLayer(ctx, "fcX", 5, 3, act=ps.PS_ACT_NONE, max_batch=4)
    W = np.arange(15, dtype=np.float32).reshape(3, 5)
    fc.put(0, W)
    fc.put(1, np.array([1, 2, 3], np.float32))
    assert np.array_equal(fc.get(0), W) and np.array_equal(fc.get(1), [1, 2, 3])
    A = fc.forward(np.eye(4, 5, dtype=np.float32))
    assert np.allclose(A, W.T[:4] + np.array([1, 2, 3], np.float32), atol=1e-6)
    fc.close()

"""world_size-2 gloo test of the multi-GPU host logic (ps_b200/sharded.py): the bucket-by-owner /
all-to-all / merge arithmetic of the sharded step, driven on CPU tensors with a numpy stand-in for
the rank-local kernels (this file's FakeOps — test infrastructure, not a product fallback).  The
checks: every lookup gets the row its key maps to regardless of which rank owns it; every owner
sees the GLOBAL occurrence count and gradient sum of its keys; the dense buffer is all-reduced."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

F, D, N, SEED = 5, 4, 24, 77


def pack(field, idv):
    return ((field + 1) << 44) | int(idv)


class FakeOps:
    """numpy restatement of the rank-local calls' CONTRACT (see include/ps_b200.h, sharded group)."""
    has_wide = True

    def __init__(self, rank):
        import contextlib
        import oracle_lib as ol
        self.L = ol.lib()
        self.rank, self.table, self.cnt, self.maxv = rank, {}, {}, self.L.pso_xavier(1, D)
        self.scope = contextlib.nullcontext
        self.gbuf = torch.zeros(6)
        self.finished = None

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype)

    def route(self, E, R):
        E = E.numpy()
        keys = np.array([pack(l % F, E.reshape(-1)[l]) for l in range(E.size)], np.int64)
        owner = np.array([self.L.pso_owner_of(int(k), R) for k in keys])
        order = np.argsort(owner, kind="stable")
        pos = np.empty(E.size, np.int32)
        pos[order] = np.arange(E.size, dtype=np.int32)
        return torch.from_numpy(keys[order]), torch.from_numpy(pos), torch.from_numpy(np.bincount(owner, minlength=R).astype(np.int32))

    def lookup(self, keys):
        self.recv_keys = [int(k) for k in keys.tolist()]
        self.cnt = {}
        rows = np.zeros((len(self.recv_keys), D), np.float32)
        for i, k in enumerate(self.recv_keys):
            assert self.L.pso_owner_of(k, dist.get_world_size()) == self.rank, "key routed to the wrong owner"
            if k not in self.table:
                self.table[k] = np.array([self.L.pso_init_value(SEED, k, j, self.maxv) for j in range(D)], np.float32)
            self.cnt[k] = self.cnt.get(k, 0) + 1
            rows[i] = np.maximum(self.table[k], 0)
        return torch.from_numpy(rows)

    def unpack(self, rows, send_pos, n):
        self.act0 = rows.numpy()[send_pos.numpy()].reshape(n, F * D)

    def dense_step(self, X, W, W_all, Y, n):
        self.W_all = None if W_all is None else W_all.clone()
        self.delta0 = ((self.act0 + 1) * (Y.numpy()[:, None] * 2 - 1) * 0.1).astype(np.float32)
        self.gbuf[:] = torch.tensor([self.act0.sum(), 1.0, float(n), float(self.rank), float(Y.mean()), 0.5])

    def grad_buffer(self):
        return self.gbuf

    def pack_grads(self, send_pos, n):
        g = (self.delta0 * (self.act0 > 0)).reshape(n * F, D)
        out = np.zeros_like(g)
        out[send_pos.numpy()] = g
        return torch.from_numpy(out)

    def finish(self, n_global, R):
        self.finished = (n_global, R)

    def apply(self, grads):
        S = {}
        for k, g in zip(self.recv_keys, grads.numpy()):
            S[k] = S.get(k, 0) + g.astype(np.float64)
        for k, s in S.items():
            n = self.cnt[k]
            g = s * (n + 1) / (2.0 * n * n)
            self.table[k] = (self.table[k] - 0.005 * g / (np.abs(g) + 1e-8)).astype(np.float32)

    def loss(self):
        return float(self.gbuf[4]) / dist.get_world_size()


def global_batch(R):
    rng = np.random.default_rng(5)
    E = rng.integers(0, 9, (R * N, F)).astype(np.int64)          # heavy key reuse across ranks
    Y = (rng.random(R * N) < 0.4).astype(np.float32)
    X = rng.random((R * N, 3)).astype(np.float32)
    return E, X, Y


def worker(rank, R, port, out):
    dist.init_process_group("gloo", init_method=f"file://{out}/rendezvous", rank=rank, world_size=R)   # no port to lose a race for
    from ps_b200.sharded import ShardedTrainer
    E, X, Y = global_batch(R)
    sl = slice(rank * N, (rank + 1) * N)
    ops = FakeOps(rank)
    tr = ShardedTrainer(ops, rank, R)
    loss = tr.step(torch.from_numpy(E[sl]), torch.from_numpy(X[sl]), torch.from_numpy(E[sl] % 7), torch.from_numpy(Y[sl]))
    torch.save({"table": ops.table, "act0": ops.act0, "gbuf": ops.gbuf, "loss": loss, "finished": ops.finished, "W_all": ops.W_all},
               os.path.join(out, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("R", [2, 3])
def test_sharded_step_host_logic(tmp_path, R):
    port = 0
    mp.spawn(worker, args=(R, port, str(tmp_path)), nprocs=R, join=True)
    res = [torch.load(os.path.join(tmp_path, f"r{r}.pt"), weights_only=False) for r in range(R)]
    # single-process reference on the concatenated batch
    import oracle_lib as ol
    L = ol.lib()
    E, X, Y = global_batch(R)
    maxv = L.pso_xavier(1, D)
    table = {}
    act0 = np.zeros((R * N, F * D), np.float32)
    for n in range(R * N):
        for j in range(F):
            k = pack(j, E[n, j])
            if k not in table:
                table[k] = np.array([L.pso_init_value(SEED, k, d, maxv) for d in range(D)], np.float32)
            act0[n, j * D:(j + 1) * D] = np.maximum(table[k], 0)
    delta0 = ((act0 + 1) * (Y[:, None] * 2 - 1) * 0.1).astype(np.float32)
    g = delta0 * (act0 > 0)
    S, cnt = {}, {}
    for n in range(R * N):
        for j in range(F):
            k = pack(j, E[n, j])
            S[k] = S.get(k, 0) + g[n, j * D:(j + 1) * D].astype(np.float64)
            cnt[k] = cnt.get(k, 0) + 1
    for r in range(R):
        assert np.array_equal(res[r]["act0"], act0[r * N:(r + 1) * N])            # rows came back to the right lookups
        assert res[r]["finished"] == (R * N, R)
        assert torch.equal(res[r]["W_all"], torch.from_numpy(E % 7))               # every replica sees every wide id
        assert abs(float(res[r]["gbuf"][0]) - float(act0.sum())) < 1e-2            # all-reduced
        assert float(res[r]["gbuf"][2]) == R * N
        assert abs(res[r]["loss"] - float(Y.reshape(R, N).mean(1).sum()) / R) < 1e-6
    merged = {}
    for r in range(R):
        for k, v in res[r]["table"].items():
            assert k not in merged and L.pso_owner_of(k, R) == r                   # disjoint shards, owned by hash
            merged[k] = v
    assert set(merged) == set(table)
    for k, w0 in table.items():
        n = cnt[k]
        ge = S[k] * (n + 1) / (2.0 * n * n)
        exp = (w0 - 0.005 * ge / (np.abs(ge) + 1e-8)).astype(np.float32)
        assert np.allclose(merged[k], exp, rtol=1e-5, atol=1e-7), k                # owners saw GLOBAL sums and counts

"""The gRPC compatibility door (ps_b200/wire.py) against net/PServer.java's semantics — host logic, no GPU:
an in-memory store with a plain gradient-descent updater stands in for the GPU tables."""
import threading

import numpy as np
import pytest

grpc = pytest.importorskip("grpc")

from ps_b200 import wire  # noqa: E402


class DictStore:
    """get / put / push / has_updater over a dict; "sgd@alfa:<lr>@" is the only updater (w -= lr * g)."""

    def __init__(self):
        self.d = {}

    def get(self, key):
        return self.d.get(key)

    def put(self, key, value):
        rows, cols, data = value
        self.d[key] = (rows, cols, np.array(data, np.float32, copy=True))

    def has_updater(self, k):
        return k.startswith("sgd@alfa:")

    def push(self, key, g, updater_key):
        if key not in self.d:
            return False
        lr = np.float32(updater_key.split(":")[1].rstrip("@"))
        r, c, w = self.d[key]
        self.d[key] = (r, c, (w - lr * np.asarray(g, np.float32)).astype(np.float32))
        return True


@pytest.fixture()
def served():
    store = DictStore()
    srv = wire.PsWireServer(store, worker_num=1, is_async=True)
    port = srv.start(0)
    cl = wire.WireClient(f"127.0.0.1:{port}")
    yield store, srv, cl
    cl.close()
    srv.stop()


def test_get_getlist_upsert_follow_pserver(served):
    store, srv, cl = served
    assert cl.get("emF1.7.0") is None                                   # Resp 204 "null weights" (PServer.java:80-86)
    w = np.arange(6, dtype=np.float32)
    r = cl.update_list({"fc0.weights": (2, 3, w), "emF1.7.0": (4, 1, np.ones(4, np.float32))}, replace=False)
    assert all(not upd for (_, upd) in r.values())                      # newly inserted: Matrix.update = false (PServer.java:150-158)
    assert store.get("fc0.weights")[:2] == (2, 3)
    # insert-if-absent: the stored value wins and is what the caller gets back, update = true
    r = cl.update_list({"fc0.weights": (2, 3, w + 100)}, replace=False)
    (rows, cols, data), upd = r["fc0.weights"]
    assert upd and (rows, cols) == (2, 3) and np.array_equal(data, w)
    r = cl.update_list({"fc0.weights": (2, 3, w + 100)}, replace=True)
    assert not r["fc0.weights"][1] and np.array_equal(store.get("fc0.weights")[2], w + 100)
    got = cl.get_list(["emF1.7.0", "nope", "fc0.weights"])
    assert got["nope"] is None and np.array_equal(got["emF1.7.0"][2], np.ones(4, np.float32)) and got["fc0.weights"][:2] == (2, 3)
    assert cl.get("fc0.weights")[:2] == (2, 3)


def test_push_applies_the_servers_running_sum(served):
    """KVStore.sum / update on a server that never clears (KVStore.java:192-208): push k applies s_k = (s_{k-1} + g_k) / k."""
    store, srv, cl = served
    cl.update_list({"emF0.3.0": (4, 1, np.zeros(4, np.float32))}, replace=False)
    rng = np.random.default_rng(3)
    w, s = np.zeros(4, np.float32), None
    for k in range(1, 6):
        g = rng.standard_normal(4).astype(np.float32)
        resp = cl.push("emF0.3.0", g, "sgd@alfa:0.5@")
        assert resp.resp.ec == 0                                        # the reference sets no Resp on a successful push
        s = g.copy() if s is None else ((s + g) / np.float32(k)).astype(np.float32)
        if k == 1:
            s = (s / np.float32(1)).astype(np.float32)
        w = (w - np.float32(0.5) * s).astype(np.float32)
        assert np.allclose(store.get("emF0.3.0")[2], w, rtol=1e-6, atol=1e-7)
    assert cl.push("emF0.3.0", np.ones(4, np.float32), "nosuch@").resp.ec == 500          # updater is null (PServer.java:169-174)
    assert cl.push("emF9.9.0", np.ones(4, np.float32), "sgd@alfa:0.5@").resp.ec == 500    # unknown key


def test_push_with_cleared_sums_applies_each_gradient():
    store = DictStore()
    srv = wire.PsWireServer(store, clear_after_update=True)
    store.put("k", (2, 1, np.zeros(2, np.float32)))
    G = wire.message("GradientMessage")
    for g in ([1.0, 2.0], [3.0, 4.0]):
        srv.push(G(gradient=wire.to_matrix("k", (2, 1, np.array(g, np.float32))), isAsync=True, updaterKey="sgd@alfa:1@"))
    assert np.allclose(store.get("k")[2], [-4.0, -6.0])


def test_synchronous_pushes_meet_in_the_barrier():
    """isAsync = false: sums wait for psUpdate, which the barrier runs once every worker has arrived (PServer.java:186-214, 238-283)."""
    store = DictStore()
    srv = wire.PsWireServer(store, worker_num=2, is_async=False)
    port = srv.start(0)
    store.put("fc1.bias", (2, 1, np.zeros(2, np.float32)))
    out = []

    def worker(g):
        cl = wire.WireClient(f"127.0.0.1:{port}")
        cl.push("fc1.bias", np.array(g, np.float32), "sgd@alfa:1@", is_async=False)
        out.append(cl.barrier().resp.ec)
        cl.close()
    ts = [threading.Thread(target=worker, args=(g,)) for g in ([2.0, 0.0], [0.0, 4.0])]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=20)
    srv.stop()
    assert out == [200, 200] and srv.global_step == 1
    assert np.allclose(store.get("fc1.bias")[2], [-1.0, -2.0])           # ONE step with the mean of the two pushes


def test_async_barrier_does_not_block(served):
    store, srv, cl = served
    assert cl.barrier().resp.ec == 200 and srv.global_step == 1


def test_messages_use_the_reference_field_numbers():
    """ps.proto:16-75 on the wire: a hand-encoded GradientMessage parses, and our encoding has the reference's tags."""
    G = wire.message("GradientMessage")
    m = G(gradient=wire.to_matrix("k", (1, 1, np.array([1.5], np.float32))), isAsync=True, updaterKey="u")
    b = m.SerializeToString()
    # field 2 (gradient, LEN) = 0x12, field 3 (isAsync, VARINT) = 0x18 0x01, field 4 (updaterKey, LEN) = 0x22 0x01 'u'
    assert b[0] == 0x12 and b.endswith(b"\x18\x01\x22\x01u")
    inner = b[2:2 + b[1]]
    # Matrix: key = 1 (0x0a), row = 2 (0x10), cols = 3 (0x18), data = 4 packed (0x22), 1.5f = 00 00 c0 3f
    assert inner == b"\x0a\x01k\x10\x01\x18\x01\x22\x04\x00\x00\xc0\x3f"
    assert wire.message("GetMessage").DESCRIPTOR.fields_by_name["resp"].number == 4       # ps.proto:44 (not 3)


def test_model_store_shapes_and_locking_without_a_gpu():
    """ModelStore over a stand-in for binding.Model: FcLayer weights travel as out x in (FcLayer.java:40-47), everything else n x 1; the
    context is made current and calls are serialised (gRPC serves from a thread pool)."""
    calls = []

    class Ctx:
        def make_current(self):
            calls.append("current")

    class FakeModel:
        F, D, Xn, fc, ctx = 3, 4, 2, [5, 1], Ctx()

        def __init__(self):
            self.d = {"fc0.weights": np.arange(5 * 14, dtype=np.float32), "fc1.weights": np.zeros(5, np.float32), "emF1.7.0": np.ones(4, np.float32)}

        def get(self, k):
            return self.d.get(k)

        def put(self, k, v):
            self.d[k] = np.array(v, np.float32)

        def push(self, k, g, spec):
            if k not in self.d:
                return False
            self.d[k] = self.d[k] - spec * np.asarray(g, np.float32)
            return True

    class FakePs:
        @staticmethod
        def updater_parse(name):
            if not name.startswith("sgd@alfa:"):
                raise ValueError(name)
            return np.float32(name.split(":")[1].rstrip("@"))

    m = FakeModel()
    st = wire.ModelStore(FakePs, m)
    assert st.get("fc0.weights")[:2] == (5, 3 * 4 + 2) and st.get("fc1.weights")[:2] == (1, 5) and st.get("emF1.7.0")[:2] == (4, 1)
    assert st.get("nope") is None and st.has_updater("sgd@alfa:0.5@") and not st.has_updater("adam@")
    st.put("emF2.9.0", (4, 1, np.full(4, 2.0, np.float32)))
    assert st.push("emF2.9.0", np.ones(4, np.float32), "sgd@alfa:0.5@") and np.allclose(m.d["emF2.9.0"], 1.5)
    assert not st.push("emF2.10.0", np.ones(4, np.float32), "sgd@alfa:0.5@")
    assert calls.count("current") >= 6
    srv = wire.PsWireServer(st)
    G = wire.message("GradientMessage")
    assert srv.push(G(gradient=wire.to_matrix("emF2.10.0", (4, 1, np.ones(4, np.float32))), isAsync=True, updaterKey="sgd@alfa:1@")).resp.ec == 500

"""ps_b200/wire.py — the reference's parameter-server gRPC protocol served from the GPU-resident store.

A legacy worker of the reference (`-Dmode=dist`, net/PSClient.java) pulls weights with `get` / `getList`, registers them with
`upsert` / `upsertList` (insert-if-absent unless `replace`), pushes ONE gradient per key per step with `push`, and meets the
other workers in `barrier`.  `PsWireServer` answers those six calls (service `net.PS`, src/main/resources/proto/ps.proto:7-14)
with the semantics of net/PServer.java on top of a store that lives on the GPU (`ModelStore` over `binding.Model`), so such
workers can train against tables the native step also trains.  It is a compatibility door, not a fast path: the protocol is
one RPC per key.

Nothing is generated: the message types of ps.proto:16-75 are declared below as a FileDescriptorProto (field numbers and types
are the wire contract), grpc's generic handlers do the rest.  Semantics restated from PServer.java, including two the reference
has whether it meant them or not:
  * a `get` of an unknown key answers `Resp{ec: 204, em: "null weights"}` (PServer.java:80-86); `getList` answers an EMPTY matrix
    for it and 200 (PServer.java:106-111);
  * the server's KVStore never clears its gradient sums (only Trainer.java:95 calls `KVStore.clear`, and a server runs no Trainer):
    `push` does `sum += g; cnt += 1; sum /= cnt` IN PLACE and applies that (KVStore.java:192-208), so the k-th push of a key applies
    (s_{k-1} + g_k) / k.  `clear_after_update=True` switches this off (each push applies its own gradient).
Synchronous mode (PServer.java:186-195, 197-214, 238-283): pushes that are not `isAsync` are summed and applied by `barrier` once
`worker_num` workers have arrived; the reference's workers always push with isAsync = true (KVStore.java:210,225), so in practice
`barrier` is the BSP meeting point only.
"""
import threading

import grpc
import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_T = descriptor_pb2.FieldDescriptorProto


def _file_descriptor():
    """ps.proto:16-75 as a descriptor (package `net`): the wire contract of the six calls."""
    f = descriptor_pb2.FileDescriptorProto(name="ps_b200/ps_wire.proto", package="net", syntax="proto3")

    def msg(name, *fields):
        m = f.message_type.add(name=name)
        for fname, num, ftype, tname, rep in fields:
            m.field.add(name=fname, number=num, type=ftype, type_name=tname or None,
                        label=_T.LABEL_REPEATED if rep else _T.LABEL_OPTIONAL)

    msg("Matrix", ("key", 1, _T.TYPE_STRING, "", 0), ("row", 2, _T.TYPE_INT32, "", 0), ("cols", 3, _T.TYPE_INT32, "", 0),
        ("data", 4, _T.TYPE_FLOAT, "", 1), ("update", 5, _T.TYPE_BOOL, "", 0))
    msg("Resp", ("ec", 1, _T.TYPE_INT32, "", 0), ("em", 2, _T.TYPE_STRING, "", 0))
    msg("RequestMeta", ("host", 1, _T.TYPE_STRING, "", 0))
    meta = ("meta", 1, _T.TYPE_MESSAGE, ".net.RequestMeta", 0)
    msg("GetListMessage", meta, ("weights", 2, _T.TYPE_MESSAGE, ".net.Matrix", 1), ("resp", 3, _T.TYPE_MESSAGE, ".net.Resp", 0))
    msg("GetMessage", meta, ("weights", 2, _T.TYPE_MESSAGE, ".net.Matrix", 0), ("resp", 4, _T.TYPE_MESSAGE, ".net.Resp", 0))
    msg("UpdateMessage", meta, ("weights", 2, _T.TYPE_MESSAGE, ".net.Matrix", 0), ("resp", 3, _T.TYPE_MESSAGE, ".net.Resp", 0),
        ("replace", 4, _T.TYPE_BOOL, "", 0))
    msg("UpdateListMessage", meta, ("weights", 2, _T.TYPE_MESSAGE, ".net.Matrix", 1), ("resp", 3, _T.TYPE_MESSAGE, ".net.Resp", 0),
        ("replace", 4, _T.TYPE_BOOL, "", 0))
    msg("GradientMessage", meta, ("gradient", 2, _T.TYPE_MESSAGE, ".net.Matrix", 0), ("isAsync", 3, _T.TYPE_BOOL, "", 0),
        ("updaterKey", 4, _T.TYPE_STRING, "", 0), ("resp", 5, _T.TYPE_MESSAGE, ".net.Resp", 0))
    msg("BarrierMessage", meta, ("resp", 2, _T.TYPE_MESSAGE, ".net.Resp", 0))
    return f


_POOL = descriptor_pool.DescriptorPool()
_POOL.Add(_file_descriptor())


def message(name):
    """The message class `net.<name>` (Matrix, Resp, GetMessage, GetListMessage, UpdateMessage, UpdateListMessage, GradientMessage, BarrierMessage)."""
    return message_factory.GetMessageClass(_POOL.FindMessageTypeByName("net." + name))


Matrix, Resp = message("Matrix"), message("Resp")
METHODS = {"get": "GetMessage", "getList": "GetListMessage", "upsert": "UpdateMessage", "upsertList": "UpdateListMessage",
           "push": "GradientMessage", "barrier": "BarrierMessage"}        # ps.proto:7-14: request and response share the type


def to_matrix(key, value):
    """MatrixUtil.FloatMatrix_2_ProtoMatrix (util/MatrixUtil.java:84-96): value = (rows, cols, float32 array) or None → key only."""
    m = Matrix(key=key)
    if value is not None:
        rows, cols, data = value
        m.row, m.cols = int(rows), int(cols)
        m.data.extend(np.asarray(data, np.float32).ravel().tolist())
    return m


def from_matrix(m):
    """MatrixUtil.ProtoMatrix_2_FloatMatrix (util/MatrixUtil.java:98-109)."""
    return int(m.row), int(m.cols), np.asarray(m.data, np.float32)


class ModelStore:
    """The store behind the server: a `binding.Model` (its GPU tables).  Values are (rows, cols, column-major float32 data)."""

    def __init__(self, ps, model):
        self.ps, self.model, self.shapes = ps, model, {}
        self.mutex = threading.Lock()                 # gRPC serves from a thread pool; calls on one library context must not overlap
        widths = [model.F * model.D + model.Xn] + list(model.fc)                  # FcLayer weights are out x in (FcLayer.java:40-47), biases out x 1
        for l, out in enumerate(model.fc):
            self.shapes[f"fc{l}.weights"] = (out, widths[l])

    def get(self, key):
        with self.mutex:
            self.model.ctx.make_current()
            v = self.model.get(key)
        if v is None:
            return None
        rows, cols = self.shapes.get(key, (v.size, 1))                            # embedding rows D x 1 (EmbeddingField.java:40), wide weights 1 x 1
        return rows, cols, v

    def put(self, key, value):
        rows, cols, data = value
        with self.mutex:
            self.model.ctx.make_current()
            self.shapes[key] = (rows, cols)
            self.model.put(key, data)

    def push(self, key, grad, updater_key):
        """One step of the updater the request names on `key`; False when the key is unknown."""
        with self.mutex:
            self.model.ctx.make_current()
            return self.model.push(key, grad, self.ps.updater_parse(updater_key))

    def has_updater(self, updater_key):
        try:
            self.ps.updater_parse(updater_key)
            return True
        except Exception:
            return False


class PsWireServer:
    """net/PServer.java over a store with get / put / push / has_updater (ModelStore; the tests also use an in-memory one)."""

    def __init__(self, store, worker_num=1, is_async=True, clear_after_update=False):
        self.store, self.worker_num, self.is_async, self.clear_after_update = store, int(worker_num), bool(is_async), bool(clear_after_update)
        self.lock = threading.Condition()
        self.sum, self.cnt = {}, {}                   # KVStore.sum / sumCnt on the server (never cleared: see the module docstring)
        self.update_keys = {}                         # key -> updaterKey of the pending synchronous pushes (PServer.java:40,188-190)
        self.global_step, self.worker_step = 0, 0     # PServer.java:36-38
        self.server = None

    # ---- KVStore.sum / KVStore.update(updater, key) as the server runs them (KVStore.java:192-208)
    def _sum(self, key, g):
        if key not in self.sum:
            self.sum[key], self.cnt[key] = np.array(g, np.float32, copy=True), 1
        else:
            self.sum[key] += g
            self.cnt[key] += 1

    def _update(self, key, updater_key):
        self.sum[key] /= np.float32(self.cnt[key])    # divi: in place, the divided sum stays in the map
        ok = self.store.push(key, self.sum[key], updater_key)
        if self.clear_after_update:
            del self.sum[key], self.cnt[key]
        return ok

    @staticmethod
    def _ok():
        return Resp(ec=200, em="")

    # ---- the six calls
    def get(self, req, ctx=None):
        cls = message("GetMessage")
        v = self.store.get(req.weights.key)
        if v is None:
            return cls(resp=Resp(ec=204, em="null weights"))                       # PServer.java:80-86
        return cls(weights=to_matrix(req.weights.key, v), resp=self._ok())

    def getList(self, req, ctx=None):
        out = message("GetListMessage")(resp=self._ok())
        for m in req.weights:                                                      # unknown key: a matrix with the key only (PServer.java:106-111)
            out.weights.append(to_matrix(m.key, self.store.get(m.key)))
        return out

    def _upsert_one(self, m, replace):
        """PServer.java:121-141: the stored value wins unless absent or `replace`; Matrix.update says whether it existed."""
        exists = self.store.get(m.key)
        update = True
        if exists is None or replace:
            update = False
            exists = from_matrix(m)
            self.store.put(m.key, exists)
        r = to_matrix(m.key, exists)
        r.update = update
        return r

    def upsert(self, req, ctx=None):
        with self.lock:
            return message("UpdateMessage")(weights=self._upsert_one(req.weights, req.replace), resp=self._ok())

    def upsertList(self, req, ctx=None):
        out = message("UpdateListMessage")(resp=self._ok())
        with self.lock:
            for m in req.weights:
                out.weights.append(self._upsert_one(m, req.replace))
        return out

    def push(self, req, ctx=None):
        cls = message("GradientMessage")
        key = req.gradient.key
        if not self.store.has_updater(req.updaterKey):
            return cls(resp=Resp(ec=500, em="updater is null"))                    # PServer.java:169-174
        _, _, g = from_matrix(req.gradient)
        with self.lock:
            self._sum(key, g)
            if req.isAsync:                                                        # PServer.java:176-184
                if not self._update(key, req.updaterKey):
                    return cls(resp=Resp(ec=500, em="null weights"))               # (the reference's updater exits the JVM here: AdamUpdater.java:65-68)
            else:
                self.update_keys.setdefault(key, req.updaterKey)                   # PServer.java:186-191
        return cls()                                                               # (an empty response: no Resp is set on this path)

    def _ps_update(self):
        """PServer.psUpdate (PServer.java:197-214); the caller holds the lock."""
        for key, uk in list(self.update_keys.items()):
            self._update(key, uk)
        self.update_keys.clear()
        self.global_step += 1

    def barrier(self, req, ctx=None):
        cls = message("BarrierMessage")
        with self.lock:
            self.worker_step += 1
            if self.is_async:                                                      # PServer.java:241-247: does not block
                self.global_step += 1
                return cls(resp=self._ok())
            mine = (self.worker_step - 1) // self.worker_num                       # the meeting this arrival belongs to
            if self.worker_step % self.worker_num == 0:                            # the last worker of the meeting runs the update
                self._ps_update()
                self.lock.notify_all()
            else:
                while self.global_step <= mine:
                    self.lock.wait(0.1)
        return cls(resp=self._ok())

    # ---- gRPC plumbing
    def handler(self):
        rpcs = {}
        for name, mtype in METHODS.items():
            cls = message(mtype)
            rpcs[name] = grpc.unary_unary_rpc_method_handler(getattr(self, name), request_deserializer=cls.FromString,
                                                             response_serializer=lambda m: m.SerializeToString())
        return grpc.method_handlers_generic_handler("net.PS", rpcs)

    def start(self, port=0, max_workers=8, host="127.0.0.1"):
        """Serves on host:port (0 = any free port); returns the bound port.  PServer.java:54-58."""
        from concurrent import futures
        self.server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers))
        self.server.add_generic_rpc_handlers((self.handler(),))
        bound = self.server.add_insecure_port(f"{host}:{port}")
        self.server.start()
        return bound

    def stop(self):
        if self.server is not None:
            self.server.stop(0)
            self.server = None


class WireClient:
    """What net/PSClient.java does on the wire (for tests and tools): the six calls with the reference's argument shapes."""

    def __init__(self, target):
        self.channel = grpc.insecure_channel(target)
        self.calls = {}
        for name, mtype in METHODS.items():
            cls = message(mtype)
            self.calls[name] = self.channel.unary_unary(f"/net.PS/{name}", request_serializer=lambda m: m.SerializeToString(),
                                                        response_deserializer=cls.FromString)

    def close(self):
        self.channel.close()

    def get(self, key):                                                            # PSClient.java:47-70
        r = self.calls["get"](message("GetMessage")(weights=Matrix(key=key)))
        return None if r.resp.ec != 200 else from_matrix(r.weights)

    def get_list(self, keys):                                                      # PSClient.java:72-97
        req = message("GetListMessage")()
        for k in keys:
            req.weights.append(Matrix(key=k))
        r = self.calls["getList"](req)
        return {m.key: (from_matrix(m) if len(m.data) else None) for m in r.weights}

    def update_list(self, values, replace):                                        # PSClient.java:128-151
        req = message("UpdateListMessage")(replace=replace)
        for k, v in values.items():
            req.weights.append(to_matrix(k, v))
        r = self.calls["upsertList"](req)
        return {m.key: (from_matrix(m), m.update) for m in r.weights}

    def push(self, key, grad, updater_key, is_async=True):                         # PSClient.java:154-174
        g = np.asarray(grad, np.float32)
        return self.calls["push"](message("GradientMessage")(gradient=to_matrix(key, (g.size, 1, g)), isAsync=is_async, updaterKey=updater_key))

    def barrier(self):                                                             # PSClient.java:177-186
        return self.calls["barrier"](message("BarrierMessage")())

"""ctypes binding of libps_b200.so (include/ps_b200.h) — what the JNI shim of
integration/java does from Java, done from Python for the tests and the bench.

There is NO fallback: if the shared library is missing or no CUDA device is present every
entry point raises.  Nothing here touches oracle/.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("PS_B200_LIB") or os.path.join(ROOT, "ps_b200", "lib", "libps_b200.so")   # (the override is for A/B builds of one kernel)

PS_OK, PS_NOT_FOUND = 0, 204
PS_FC_FP32, PS_FC_TF32, PS_FC_TF32X3 = 0, 1, 2
PS_UPD_ADAM, PS_UPD_FTRL, PS_UPD_SIMPLE = 0, 1, 2
PS_MODEL_DNN, PS_MODEL_WIDEDEEP, PS_MODEL_FCNN = 0, 1, 2
KINDS = {"dnn": PS_MODEL_DNN, "widedeep": PS_MODEL_WIDEDEEP, "fcnn": PS_MODEL_FCNN}


class UpdaterSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("p", C.c_float * 4)]

    @staticmethod
    def adam(alfa=0.005, beta1=0.9, beta2=0.999, epsilon=1e-8):
        s = UpdaterSpec()
        s.kind = PS_UPD_ADAM
        s.p[:] = [alfa, beta1, beta2, epsilon]
        return s

    @staticmethod
    def ftrl(alfa=0.005, beta=1.0, l1=0.001, l2=0.001):
        s = UpdaterSpec()
        s.kind = PS_UPD_FTRL
        s.p[:] = [alfa, beta, l1, l2]
        return s

    @staticmethod
    def simple(eta):
        s = UpdaterSpec()
        s.kind = PS_UPD_SIMPLE
        s.p[:] = [eta, 0, 0, 0]
        return s


# every symbol include/ps_b200.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_pp = C.POINTER(C.c_void_p)
SYMBOLS = {
    "ps_last_error": (C.c_char_p, []),
    "ps_abi_version": (_i, []),
    "ps_ctx_create": (_i, [_i, C.c_uint64, _pp]),
    "ps_ctx_destroy": (_i, [_vp]),
    "ps_ctx_set_fc_precision": (_i, [_vp, _i]),
    "ps_ctx_get_fc_precision": (_i, [_vp, C.POINTER(_i)]),
    "ps_ctx_set_exact_updaters": (_i, [_vp, _i]),
    "ps_ctx_synchronize": (_i, [_vp]),
    "ps_ctx_make_current": (_i, [_vp]),
    "ps_ctx_launch_count": (_i, [_vp, C.POINTER(_i64)]),
    "ps_ctx_device_info": (_i, [_vp, C.c_char_p, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "ps_ctx_stream": (_i, [_vp, _pp]),
    "ps_host_alloc": (_i, [C.c_size_t, _pp]),
    "ps_host_free": (_i, [_vp]),
    "ps_updater_parse": (_i, [C.c_char_p, C.POINTER(UpdaterSpec)]),
    "ps_updater_name": (_i, [C.POINTER(UpdaterSpec), C.c_char_p, _i]),
    "ps_updater_apply": (_i, [_vp, C.POINTER(UpdaterSpec), _vp, _vp, _vp, _vp, _i]),
    "ps_emb_create": (_i, [_vp, _i, _i, _i64, C.POINTER(UpdaterSpec), _pp]),
    "ps_emb_destroy": (_i, [_vp]),
    "ps_emb_forward": (_i, [_vp, _vp, _i, _vp]),
    "ps_emb_forward_f32ids": (_i, [_vp, _vp, _i, _vp]),
    "ps_emb_backward_update": (_i, [_vp, _vp, _i, _i, _i]),
    "ps_emb_get_rows": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "ps_emb_put_rows": (_i, [_vp, _vp, _vp, _i, _vp, _i]),
    "ps_emb_size": (_i, [_vp, C.POINTER(_i64)]),
    "ps_model_create": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i64, C.POINTER(UpdaterSpec), _i, _pp]),
    "ps_model_destroy": (_i, [_vp]),
    "ps_model_train_step": (_i, [_vp, _vp, _vp, _vp, _vp, _i, C.POINTER(_f)]),
    "ps_model_submit": (_i, [_vp, _vp, _vp, _vp, _vp, _i]),
    "ps_model_collect": (_i, [_vp, C.POINTER(_f)]),
    "ps_model_forward": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "ps_model_backward_update": (_i, [_vp, _vp, _i, _f]),
    "ps_model_submit_text": (_i, [_vp, _vp, C.c_size_t, _i]),
    "ps_model_step_info": (_i, [_vp, C.POINTER(_i), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "ps_model_shape": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "ps_model_train_step_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _i]),
    "ps_model_read_loss": (_i, [_vp, C.POINTER(_f)]),
    "ps_model_loss_dev": (_i, [_vp, _pp]),
    "ps_model_predict": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "ps_model_get": (_i, [_vp, C.c_char_p, _vp, _i, C.POINTER(_i)]),
    "ps_model_put": (_i, [_vp, C.c_char_p, _vp, _i]),
    "ps_model_push": (_i, [_vp, C.c_char_p, _vp, _i, _vp]),
    "ps_model_get_list": (_i, [_vp, C.POINTER(C.c_char_p), _i, _vp, _i, _vp]),
    "ps_model_update_list": (_i, [_vp, C.POINTER(C.c_char_p), _i, _vp, _i, _vp, _i]),
    "ps_model_get_state": (_i, [_vp, C.c_char_p, _i, _vp, _i, C.POINTER(_i)]),
    "ps_model_tap": (_i, [_vp, C.c_char_p, _i, _vp, _i, C.POINTER(_i)]),
    "ps_model_num_keys": (_i, [_vp, C.POINTER(_i64)]),
    "ps_model_skipped_backward": (_i, [_vp, C.POINTER(_i)]),
    "ps_model_save": (_i, [_vp, C.c_char_p]),
    "ps_model_load": (_i, [_vp, C.c_char_p]),
    "ps_model_profile": (_i, [_vp, _i]),
    "ps_model_phase_times": (_i, [_vp, _vp, _i, C.POINTER(_i), C.c_char_p, _i]),
    "ps_model_kernel_times": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "ps_model_gemm_times": (_i, [_vp, _i, _i, _vp, _i]),
    "ps_key_owner": (_i, [C.c_char_p, _i, C.POINTER(_i)]),
    "ps_shard_route_dev": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ps_shard_route_padded_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ps_model_shard_lookup_dev": (_i, [_vp, _vp, _i, _vp]),
    "ps_model_shard_row_stride": (_i, [_vp, C.POINTER(_i)]),
    "ps_model_shard_unpack_dev": (_i, [_vp, _vp, _vp, _i]),
    "ps_model_shard_dense_step_dev": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i]),
    "ps_model_shard_grad_buffer": (_i, [_vp, _pp, C.POINTER(_i64)]),
    "ps_model_shard_pack_grads_dev": (_i, [_vp, _vp, _i, _vp]),
    "ps_model_shard_finish_dev": (_i, [_vp, _i, _i]),
    "ps_model_shard_apply_dev": (_i, [_vp, _vp, _i]),
    "ps_model_p2p_init": (_i, [_vp, _i, _i, _i, _vp]),
    "ps_model_p2p_connect": (_i, [_vp, _vp]),
    "ps_model_p2p_step_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _i]),
    "ps_model_p2p_submit": (_i, [_vp, _vp, _vp, _vp, _vp, _i]),
    "ps_model_p2p_overflowed": (_i, [_vp, C.POINTER(_i)]),
    "ps_fc_create": (_i, [_vp, C.c_char_p, _i, _i, _i, C.POINTER(UpdaterSpec), _i, _pp]),
    "ps_fc_destroy": (_i, [_vp]),
    "ps_fc_forward": (_i, [_vp, _vp, _i, _vp]),
    "ps_fc_backward": (_i, [_vp, _vp, _i, _vp]),
    "ps_fc_gradients": (_i, [_vp, _vp, _vp]),
    "ps_fc_update": (_i, [_vp]),
    "ps_fc_get": (_i, [_vp, _i, _vp, _i, C.POINTER(_i)]),
    "ps_fc_put": (_i, [_vp, _i, _vp, _i]),
    "ps_libsvm_parse_line": (_i, [C.c_char_p, C.c_size_t, _i, _i, _i64, _vp, _vp, _vp, _vp, C.POINTER(_i)]),
    "ps_libsvm_parse_dev": (_i, [_vp, _vp, C.c_size_t, _i, _i, _i64, _i, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i)]),
    "ps_reader_open": (_i, [C.c_char_p, _i, _i, _i64, _i, _i, _i, _i, _pp]),
    "ps_reader_next": (_i, [_vp, _vp, _vp, _vp, _vp, C.POINTER(_i)]),
    "ps_reader_shape": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "ps_reader_reset": (_i, [_vp]),
    "ps_reader_stats": (_i, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "ps_reader_close": (_i, [_vp]),
    "ps_test_gemm_nt": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i]),
}

_lib = None


class PsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libps_b200 error {code}: {msg}")
        self.code = code


def lib():
    """Loads libps_b200.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != PS_OK:
        raise PsError(rc, lib().ps_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dt)


class PinnedArray:
    """numpy view over page-locked memory from ps_host_alloc."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = C.c_void_p()
        check(lib().ps_host_alloc(self.nbytes, C.byref(self.ptr)))
        buf = (C.c_byte * self.nbytes).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().ps_host_free(self.ptr)
            self.ptr = None


def parse_libsvm_line(line, F=23, Xn=45, wide_size=100000):
    """data.LibsvmParser.parse + one column of CTR.parseFeature.  Returns (status, E, X, W, Y): status 0 ok, 1 blank/short, 2 unparsable."""
    if isinstance(line, str):
        line = line.encode()
    E, W = np.zeros(F, np.int64), np.zeros(F, np.int64)
    X, Y = np.zeros(Xn, np.float32), np.zeros(1, np.float32)
    st = C.c_int()
    check(lib().ps_libsvm_parse_line(line, len(line), F, Xn, wide_size, _p(E), _p(X), _p(W), _p(Y), C.byref(st)))
    return st.value, E, X, W, Y[0]


class LibsvmReader:
    """data.DataSet over data.FileSource with data.LibsvmParser + CTR.parseFeature (DataSet.java:37-100): batches of E, X, W, Y
    in the layout Model.train_step / submit take.  `buffers` may hold PinnedArray-backed arrays to parse straight into pinned memory."""

    def __init__(self, path, F=23, Xn=45, wide_size=100000, batch=1000, offset=0, step=1, threads=1):
        self.F, self.Xn, self.batch = F, Xn, batch
        self.h = C.c_void_p()
        check(lib().ps_reader_open(os.fsencode(path), F, Xn, wide_size, batch, offset, step, threads, C.byref(self.h)))

    def next(self, buffers=None):
        """dict(E, X, W, Y) of the next batch (views trimmed to its rows), or None at end of data."""
        b = buffers or dict(E=np.empty((self.batch, self.F), np.int64), W=np.empty((self.batch, self.F), np.int64),
                            X=np.empty((self.batch, self.Xn), np.float32), Y=np.empty(self.batch, np.float32))
        n = C.c_int()
        check(lib().ps_reader_next(self.h, _p(b["E"]), _p(b["X"]), _p(b["W"]), _p(b["Y"]), C.byref(n)))
        if n.value == 0:
            return None
        return {k: v[: n.value] for k, v in b.items()}

    def __iter__(self):
        while True:
            b = self.next()
            if b is None:
                return
            yield b

    def reset(self):
        check(lib().ps_reader_reset(self.h))

    def stats(self):
        a, b, c = _i64(), _i64(), _i64()
        check(lib().ps_reader_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(lines=a.value, batches=b.value, dropped_batches=c.value)

    def close(self):
        if self.h:
            lib().ps_reader_close(self.h)
            self.h = None


PS_ACT_NONE, PS_ACT_RELU, PS_ACT_SIGMOID = 0, 1, 2


class FcLayer:
    """layer.FcLayer as a standalone operator (FcLayer.java:34-115).  Matrices as numpy (N, features) arrays = jblas features x N."""

    def __init__(self, ctx, name, in_dims, out_dims, act=PS_ACT_RELU, updater=None, max_batch=1024):
        self.in_dims, self.out_dims = in_dims, out_dims
        self.h = C.c_void_p()
        check(lib().ps_fc_create(ctx.h, name.encode(), in_dims, out_dims, act, C.byref(updater) if updater is not None else None, max_batch, C.byref(self.h)))

    def forward(self, A_prev):
        A_prev = _c(A_prev, np.float32)
        out = np.zeros((A_prev.shape[0], self.out_dims), np.float32)
        check(lib().ps_fc_forward(self.h, _p(A_prev), A_prev.shape[0], _p(out)))
        return out

    def backward(self, delta):
        delta = _c(delta, np.float32)
        dprev = np.zeros((delta.shape[0], self.in_dims), np.float32)
        check(lib().ps_fc_backward(self.h, _p(delta), delta.shape[0], _p(dprev)))
        return dprev

    def gradients(self):
        dW = np.zeros(self.out_dims * self.in_dims, np.float32)
        db = np.zeros(self.out_dims, np.float32)
        check(lib().ps_fc_gradients(self.h, _p(dW), _p(db)))
        return dW.reshape(self.in_dims, self.out_dims).T.copy(), db     # (out, in)

    def update(self):
        check(lib().ps_fc_update(self.h))

    def get(self, which):
        n = self.out_dims * (self.in_dims if which == 0 else 1)
        out = np.zeros(n, np.float32)
        got = C.c_int()
        check(lib().ps_fc_get(self.h, which, _p(out), n, C.byref(got)))
        return out.reshape(self.in_dims, self.out_dims).T.copy() if which == 0 else out

    def put(self, which, value):
        v = _c(np.asarray(value, np.float32).T if which == 0 else value, np.float32).reshape(-1)
        check(lib().ps_fc_put(self.h, which, _p(v), v.size))

    def close(self):
        if self.h:
            lib().ps_fc_destroy(self.h)
            self.h = None


class Context:
    """store.KVStore.ins() + the device it lives on."""

    def __init__(self, device=0, seed=0):
        self.h = C.c_void_p()
        check(lib().ps_ctx_create(device, seed, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().ps_ctx_destroy(self.h)
            self.h = None

    def set_fc_precision(self, mode):
        check(lib().ps_ctx_set_fc_precision(self.h, mode))

    def fc_precision(self):
        v = C.c_int()
        check(lib().ps_ctx_get_fc_precision(self.h, C.byref(v)))
        return v.value

    def set_exact_updaters(self, on):
        check(lib().ps_ctx_set_exact_updaters(self.h, 1 if on else 0))

    def synchronize(self):
        check(lib().ps_ctx_synchronize(self.h))

    def make_current(self):
        """For host threads other than the one that created the context (see ps_ctx_make_current)."""
        check(lib().ps_ctx_make_current(self.h))

    def launch_count(self):
        v = C.c_int64()
        check(lib().ps_ctx_launch_count(self.h, C.byref(v)))
        return v.value

    def stream(self):
        s = C.c_void_p()
        check(lib().ps_ctx_stream(self.h, C.byref(s)))
        return s.value

    def device_info(self):
        name = C.create_string_buffer(128)
        sms, ma, mi = C.c_int(), C.c_int(), C.c_int()
        check(lib().ps_ctx_device_info(self.h, name, 128, C.byref(sms), C.byref(ma), C.byref(mi)))
        return dict(name=name.value.decode(), sms=sms.value, cc=(ma.value, mi.value))

    def parse_libsvm_dev(self, text_dev, nbytes, F, Xn, wide_size, max_rows, E_dev, X_dev, W_dev, Y_dev, status_dev):
        """GPU-side libsvm parse (device pointers as ints); returns the number of lines found."""
        rows = C.c_int()
        check(lib().ps_libsvm_parse_dev(self.h, C.c_void_p(text_dev), nbytes, F, Xn, wide_size, max_rows, C.c_void_p(E_dev), C.c_void_p(X_dev),
                                        C.c_void_p(W_dev), C.c_void_p(Y_dev), C.c_void_p(status_dev), C.byref(rows)))
        return rows.value

    def updater_apply(self, spec, w, s1, s2, g):
        check(lib().ps_updater_apply(self.h, C.byref(spec), _p(w), _p(s1), _p(s2), _p(g), w.size))

    def gemm_nt(self, mode, A, B):
        A, B = _c(A, np.float32), _c(B, np.float32)
        M, K = A.shape
        N = B.shape[0]
        Cm = np.zeros((M, N), np.float32)
        check(lib().ps_test_gemm_nt(self.h, mode, M, N, K, _p(A), K, _p(B), K, _p(Cm), N))
        return Cm


def updater_parse(name):
    s = UpdaterSpec()
    check(lib().ps_updater_parse(name.encode(), C.byref(s)))
    return s


def key_owner(key, n_shards):
    """net/Router.java:5 for the native store: the shard of an embedding key string, -1 for keys every shard holds."""
    o = C.c_int()
    check(lib().ps_key_owner(key.encode(), int(n_shards), C.byref(o)))
    return o.value


def updater_name(spec):
    buf = C.create_string_buffer(256)
    check(lib().ps_updater_name(C.byref(spec), buf, 256))
    return buf.value.decode()


class EmbeddingLayer:
    """layer.EmbeddingLayer (+ its EmbeddingFields) over the GPU-resident table."""

    def __init__(self, ctx, F, D, capacity, updater=None):
        self.ctx, self.F, self.D = ctx, F, D
        self.h = C.c_void_p()
        check(lib().ps_emb_create(ctx.h, F, D, capacity, C.byref(updater) if updater is not None else None, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().ps_emb_destroy(self.h)
            self.h = None

    def forward(self, E):
        if E.dtype == np.float32:
            E = _c(E, np.float32)
            fn = lib().ps_emb_forward_f32ids
        else:
            E = _c(E, np.int64)
            fn = lib().ps_emb_forward
        N = E.shape[0]
        out = np.empty((N, self.F * self.D), np.float32)
        check(fn(self.h, _p(E), N, _p(out)))
        return out

    def backward_update(self, delta, calls=2):
        delta = _c(delta, np.float32)
        N, ld = delta.shape
        check(lib().ps_emb_backward_update(self.h, _p(delta), ld, N, calls))

    def get_rows(self, fields, ids, state=False):
        fields, ids = _c(fields, np.int32), _c(ids, np.int64)
        n = ids.size
        w = np.zeros((n, self.D), np.float32)
        s1 = np.zeros((n, self.D), np.float32) if state else None
        s2 = np.zeros((n, self.D), np.float32) if state else None
        found = np.zeros(n, np.int32)
        check(lib().ps_emb_get_rows(self.h, _p(fields), _p(ids), n, _p(w), _p(s1), _p(s2), _p(found)))
        return (w, s1, s2, found) if state else (w, found)

    def put_rows(self, fields, ids, w, replace=True):
        fields, ids, w = _c(fields, np.int32), _c(ids, np.int64), np.array(w, np.float32, order="C")
        check(lib().ps_emb_put_rows(self.h, _p(fields), _p(ids), ids.size, _p(w), 1 if replace else 0))
        return w

    def size(self):
        v = C.c_int64()
        check(lib().ps_emb_size(self.h, C.byref(v)))
        return v.value


class Model:
    """model.DNN / WideDeepNN / FullConnectedNN stepped like train.Trainer with thread = 1."""

    def __init__(self, ctx, kind, F, D, Xn, fc, emb_capacity=1 << 20, emb_updater=None, max_batch=4096):
        self.ctx, self.kind, self.F, self.D, self.Xn, self.fc = ctx, kind, F, D, Xn, list(fc)
        self.h = C.c_void_p()
        fca = np.asarray(fc, np.int32)
        check(lib().ps_model_create(ctx.h, KINDS[kind] if isinstance(kind, str) else kind, F, D, Xn, _p(fca), len(fc), emb_capacity,
                                    C.byref(emb_updater) if emb_updater is not None else None, max_batch, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().ps_model_destroy(self.h)
            self.h = None

    def train_step(self, E, X, W, Y):
        E, W, X, Y = _c(E, np.int64), _c(W, np.int64), _c(X, np.float32), _c(Y, np.float32)
        loss = C.c_float()
        check(lib().ps_model_train_step(self.h, _p(E), _p(X), _p(W), _p(Y), Y.shape[0], C.byref(loss)))
        return loss.value

    def forward(self, E, X, W):
        """The forward loop of DNN.train / WideDeepNN.train: returns P (N,), keeps the batch pending for backward_update."""
        E, W, X = _c(E, np.int64), _c(W, np.int64), _c(X, np.float32)
        N = X.shape[0]
        P = np.zeros(N, np.float32)
        check(lib().ps_model_forward(self.h, _p(E), _p(X), _p(W), N, _p(P)))
        return P

    def backward_update(self, delta_top, loss=float("nan")):
        """The reverse loop + KVStore.update given delta_top = loss.backward(P, Y) computed by the caller."""
        d = _c(delta_top, np.float32)
        check(lib().ps_model_backward_update(self.h, _p(d), d.shape[0], loss))

    def submit_text(self, text_ptr, nbytes, N):
        check(lib().ps_model_submit_text(self.h, text_ptr, nbytes, N))

    def step_info(self):
        sk, bad, nu = C.c_int(), C.c_uint32(), C.c_uint32()
        check(lib().ps_model_step_info(self.h, C.byref(sk), C.byref(bad), C.byref(nu)))
        return dict(skipped=bool(sk.value), bad_lines=bad.value, n_unique=nu.value)

    def submit_ptrs(self, E, X, W, Y, N):
        check(lib().ps_model_submit(self.h, E, X, W, Y, N))

    def collect(self):
        loss = C.c_float()
        check(lib().ps_model_collect(self.h, C.byref(loss)))
        return loss.value

    def p2p_submit_ptrs(self, E, X, W, Y, N):
        check(lib().ps_model_p2p_submit(self.h, E, X, W, Y, N))

    def train_step_dev(self, E, X, W, Y, N):
        check(lib().ps_model_train_step_dev(self.h, E, X, W, Y, N))

    def read_loss(self):
        loss = C.c_float()
        check(lib().ps_model_read_loss(self.h, C.byref(loss)))
        return loss.value

    def loss_dev(self):
        """Device address of the step's loss (for asynchronous copies on the library's stream)."""
        p = C.c_void_p()
        check(lib().ps_model_loss_dev(self.h, C.byref(p)))
        return p.value

    def predict(self, E, X, W, N, out_rows=1):
        E, W, X = _c(E, np.int64), _c(W, np.int64), _c(X, np.float32)
        out = np.zeros(N * out_rows, np.float32)
        check(lib().ps_model_predict(self.h, _p(E), _p(X), _p(W), N, _p(out)))
        return out

    def _fetch(self, fn, key, *extra):
        n = C.c_int()
        rc = fn(self.h, key.encode(), *extra, None, 0, C.byref(n))
        if rc == PS_NOT_FOUND:
            return None
        check(rc)
        out = np.zeros(n.value, np.float32)
        check(fn(self.h, key.encode(), *extra, _p(out), n.value, C.byref(n)))
        return out

    def get(self, key):
        return self._fetch(lib().ps_model_get, key)

    def get_list(self, keys, stride):
        """PSClient.getList: dict key -> array (None when absent)."""
        arr = (C.c_char_p * len(keys))(*[k.encode() for k in keys])
        out = np.zeros((len(keys), stride), np.float32)
        found = np.zeros(len(keys), np.int32)
        check(lib().ps_model_get_list(self.h, arr, len(keys), _p(out), stride, _p(found)))
        return {k: (out[i, : found[i]].copy() if found[i] else None) for i, k in enumerate(keys)}

    def update_list(self, updates, replace=False):
        """PSClient.updateList: offers {key: array}; returns {key: winning array} (insert-if-absent unless replace)."""
        keys = list(updates)
        stride = max(len(np.ravel(v)) for v in updates.values())
        io = np.zeros((len(keys), stride), np.float32)
        lens = np.zeros(len(keys), np.int32)
        for i, k in enumerate(keys):
            v = np.ravel(np.asarray(updates[k], np.float32))
            io[i, : v.size] = v
            lens[i] = v.size
        arr = (C.c_char_p * len(keys))(*[k.encode() for k in keys])
        check(lib().ps_model_update_list(self.h, arr, len(keys), _p(io), stride, _p(lens), 1 if replace else 0))
        return {k: io[i, : lens[i]].copy() for i, k in enumerate(keys)}

    def get_state(self, key, which):
        return self._fetch(lib().ps_model_get_state, key, which)

    def tap(self, layer, what=0):
        return self._fetch(lib().ps_model_tap, layer, what)

    def push(self, key, grad, spec):
        """PServer.push: one step of the updater `spec` (an UpdaterSpec, e.g. updater_parse(updaterKey)) on an existing key with a pushed
        gradient in the reference's layout.  False when the key does not exist."""
        g = _c(grad, np.float32)
        rc = lib().ps_model_push(self.h, key.encode(), _p(g), g.size, C.byref(spec))
        if rc == PS_NOT_FOUND:
            return False
        check(rc)
        return True

    def put(self, key, v):
        v = _c(v, np.float32)
        check(lib().ps_model_put(self.h, key.encode(), _p(v), v.size))

    def num_keys(self):
        v = C.c_int64()
        check(lib().ps_model_num_keys(self.h, C.byref(v)))
        return v.value

    def save(self, path):
        check(lib().ps_model_save(self.h, os.fsencode(path)))

    def load(self, path):
        check(lib().ps_model_load(self.h, os.fsencode(path)))

    def skipped_backward(self):
        v = C.c_int()
        check(lib().ps_model_skipped_backward(self.h, C.byref(v)))
        return bool(v.value)

    def kernel_times(self, E_ptrs, N, reps=64):
        """device us of {key resolution alone, the fused lookup kernel (resolve + gather), scatter + update, clear_batch};
        E_ptrs: device addresses of [N][F] int64 id batches"""
        arr = (C.c_void_p * len(E_ptrs))(*E_ptrs)
        us = np.zeros(4, np.float32)
        check(lib().ps_model_kernel_times(self.h, arr, len(E_ptrs), N, reps, _p(us)))
        return dict(zip(["emb_resolve", "emb_lookup", "emb_scatter_update", "emb_clear"], us.tolist()))

    def gemm_times(self, N, reps=64):
        us = np.zeros(3 * len(self.fc), np.float32)
        check(lib().ps_model_gemm_times(self.h, N, reps, _p(us), us.size))
        return {f"fc{l}.{n}": float(us[3 * l + i]) for l in range(len(self.fc)) for i, n in enumerate(("forward", "dgrad", "wgrad"))}

    def profile(self, enable=True):
        check(lib().ps_model_profile(self.h, 1 if enable else 0))

    def phase_times(self):
        n = C.c_int()
        ms = np.zeros(64, np.float32)
        names = C.create_string_buffer(1024)
        check(lib().ps_model_phase_times(self.h, _p(ms), 64, C.byref(n), names, 1024))
        nm = names.value.decode().split(";") if names.value else []
        return dict(zip(nm, ms[: n.value].tolist()))

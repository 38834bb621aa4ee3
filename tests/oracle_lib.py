"""ctypes loader for the CPU oracle (oracle/ps_oracle.cpp).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "libps_oracle.so")

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build():
    src = os.path.join(ROOT, "oracle", "ps_oracle.cpp")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    L.pso_java_hash.restype = C.c_int32
    L.pso_java_hash.argtypes = [C.c_char_p]
    L.pso_router_mod.restype = C.c_int32
    L.pso_router_mod.argtypes = [C.c_char_p, C.c_int32]
    L.pso_router_floormod.restype = C.c_int32
    L.pso_router_floormod.argtypes = [C.c_char_p, C.c_int32]
    L.pso_key_string.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_char_p, C.c_int]
    L.pso_pack_key.restype = C.c_uint64
    L.pso_pack_key.argtypes = [C.c_uint32, C.c_uint64]
    L.pso_name_key.restype = C.c_uint64
    L.pso_name_key.argtypes = [C.c_char_p]
    L.pso_init_value.restype = C.c_float
    L.pso_init_value.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_float]
    L.pso_owner_of.restype = C.c_uint32
    L.pso_owner_of.argtypes = [C.c_uint64, C.c_uint32]
    L.pso_xavier.restype = C.c_float
    L.pso_xavier.argtypes = [C.c_int, C.c_int]
    L.pso_auc.restype = C.c_double
    L.pso_auc.argtypes = [f32p, f32p, C.c_int]
    L.pso_adam_update.argtypes = [f32p, f32p, f32p, f32p, C.c_int] + [C.c_float] * 4
    L.pso_ftrl_update.argtypes = [f32p, f32p, f32p, f32p, C.c_int] + [C.c_float] * 4
    L.pso_updater_name.argtypes = [C.c_int] + [C.c_float] * 4 + [C.c_char_p, C.c_int]
    L.pso_set_gemm.argtypes = [C.c_int, C.c_char_p]
    L.pso_set_blas_threads.restype = C.c_int
    L.pso_set_blas_threads.argtypes = [C.c_int]
    L.pso_model_create.restype = C.c_void_p
    L.pso_model_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, i32p, C.c_int, C.c_uint64, C.c_int]
    L.pso_model_destroy.argtypes = [C.c_void_p]
    L.pso_train_step.restype = C.c_float
    L.pso_train_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.pso_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, f32p]
    L.pso_skipped_backward.argtypes = [C.c_void_p]
    L.pso_num_keys.restype = C.c_int64
    L.pso_num_keys.argtypes = [C.c_void_p]
    for fn in (L.pso_get, L.pso_get_init):
        fn.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
    L.pso_get_state.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int]
    L.pso_layer_tap.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int]
    L.pso_emb_create.restype = C.c_void_p
    L.pso_emb_create.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int]
    L.pso_emb_forward.argtypes = [C.c_void_p, i64p, C.c_int, f32p]
    L.pso_emb_backward_update.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int, C.c_int]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def key_string(kind, field, idv):
    buf = C.create_string_buffer(64)
    n = lib().pso_key_string(kind, field, int(idv), buf, 64)
    assert n > 0
    return buf.value.decode()


def openblas_path():
    d = os.path.join(os.path.dirname(np.__file__), "..", "numpy.libs")
    if os.path.isdir(d):
        for f in os.listdir(d):
            if "openblas" in f:
                return os.path.abspath(os.path.join(d, f))
    return None


KIND_DNN, KIND_WIDEDEEP, KIND_FCNN = 0, 1, 2


class OracleModel:
    """model.DNN / model.WideDeepNN / model.FullConnectedNN driven by train.Trainer with thread=1."""

    def __init__(self, kind, F, D, Xn, fc, seed, emb_opt=0):
        self.L = lib()
        self.kind, self.F, self.D, self.Xn, self.fc = kind, F, D, Xn, list(fc)
        fc_a = np.asarray(fc, np.int32)
        self.h = self.L.pso_model_create(kind, F, D, Xn, fc_a, len(fc), seed, emb_opt)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.pso_model_destroy(self.h)
            self.h = None

    def train_step(self, E, X, W, Y):
        N = Y.shape[-1]
        E = None if E is None else np.ascontiguousarray(E, np.int64)
        W = None if W is None else np.ascontiguousarray(W, np.int64)
        X = np.ascontiguousarray(X, np.float32)
        Y = np.ascontiguousarray(Y, np.float32)
        return float(self.L.pso_train_step(self.h, _ptr(E), _ptr(X), _ptr(W), _ptr(Y), N))

    def predict(self, E, X, W, N, out_rows=1):
        E = None if E is None else np.ascontiguousarray(E, np.int64)
        W = None if W is None else np.ascontiguousarray(W, np.int64)
        X = np.ascontiguousarray(X, np.float32)
        out = np.zeros(out_rows * N, np.float32)
        self.L.pso_predict(self.h, _ptr(E), _ptr(X), _ptr(W), N, out)
        return out

    def _fetch(self, fn, key, *extra):
        n = fn(self.h, key.encode(), *extra, None, 0)
        if n < 0:
            return None
        out = np.zeros(n, np.float32)
        fn(self.h, key.encode(), *extra, _ptr(out), n)
        return out

    def get(self, key):
        return self._fetch(self.L.pso_get, key)

    def get_init(self, key):
        return self._fetch(self.L.pso_get_init, key)

    def get_state(self, key, which):
        return self._fetch(self.L.pso_get_state, key, which)

    def tap(self, layer, what=0):
        return self._fetch(self.L.pso_layer_tap, layer, what)

    def num_keys(self):
        return int(self.L.pso_num_keys(self.h))

    def skipped_backward(self):
        return bool(self.L.pso_skipped_backward(self.h))

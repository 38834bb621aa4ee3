"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports
every symbol include/ps_b200.h declares, and fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ps_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(ps):
    names = declared_symbols()
    assert len(names) >= 35
    L = C.CDLL(ps.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in ps_b200.h but not exported"
    assert set(names) == set(ps.SYMBOLS), set(names) ^ set(ps.SYMBOLS)
    assert L.ps_abi_version() == 1


def test_library_is_sm100a_native(ps):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", ps.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(ps):
    from conftest import HAS_GPU
    if HAS_GPU:
        pytest.skip("device present")
    with pytest.raises(ps.PsError) as e:
        ps.Context(0, seed=1)
    assert e.value.code == 500


def test_updater_name_grammar(ps):
    """T/TestPs.java:26-27 pins `name@k:v@k:v@`; AdamUpdater(String)/getName round-trip (AdamUpdater.java:50-55,72-74)."""
    s = ps.updater_parse("adam@alfa:0.005@beta1:0.9@beta2:0.999@epsilon:1.0E-8@")
    assert s.kind == ps.PS_UPD_ADAM and abs(s.p[3] - 1e-8) < 1e-12
    assert ps.updater_name(s) == "adam@alfa:0.005@beta1:0.9@beta2:0.999@epsilon:1.0E-8@"
    f = ps.updater_parse("adam@alfa:0.005@beta:1.0@l1:0.001@l2:0.001@")     # FtrlUpdater.getName says "adam@" (sic)
    assert f.kind == ps.PS_UPD_FTRL and ps.updater_name(f) == "adam@alfa:0.005@beta:1.0@l1:0.001@l2:0.001@"
    assert ps.updater_name(ps.UpdaterSpec.adam(1, 2, 3, 1e7)) == "adam@alfa:1.0@beta1:2.0@beta2:3.0@epsilon:1.0E7@"
    with pytest.raises(ps.PsError):
        ps.updater_parse("bogus")


def test_key_owner_is_the_route_kernels_owner(ps):
    """ps_key_owner (host, key strings: what PSRouterClient asks its Router, PSRouterClient.java:55-57) = ps_owner_of of the packed key,
    the function the device-side route kernels and the oracle's router use; dense and wide keys are replicated (-1)."""
    import oracle_lib as ol
    L = ol.lib()
    rng = np.random.default_rng(5)
    for R in (1, 2, 3, 8):
        seen = set()
        for _ in range(300):
            f, v = int(rng.integers(0, 26)), int(rng.integers(0, 1 << 24))     # the reference's ids are floats: exact below 2^24
            key = ol.key_string(0, f, v)
            o = ps.key_owner(key, R)
            assert o == L.pso_owner_of((f + 1) << 44 | v, R) and 0 <= o < R
            seen.add(o)
        assert len(seen) == R                                   # every shard gets keys
    assert ps.key_owner("fc0.weights", 4) == -1 and ps.key_owner("wide.weights.77.0", 4) == -1 and ps.key_owner("wide.bias", 4) == -1
    assert ps.key_owner("emF3.15757.0", 1) == 0
    with pytest.raises(ps.PsError):
        ps.key_owner("emF0.5.0", 0)


def test_product_never_imports_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, "ps_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_lib" not in text and "ps_oracle" not in text and "libps_oracle" not in text, f


def test_jni_shim_typechecks_and_covers_every_native_method():
    """No JDK in this image: the shim is type-checked against include/ps_b200.h with a stub jni.h that carries the JNI
    specification's signatures for the functions it uses, and every `native` method of PsNative.java must have its
    Java_nativeps_PsNative_<name> definition."""
    import re
    import subprocess
    shim = os.path.join(ROOT, "integration", "jni", "ps_jni.c")
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-I" + os.path.join(ROOT, "integration", "jni", "stub"),
                        "-I" + os.path.join(ROOT, "include"), shim], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    java = open(os.path.join(ROOT, "integration", "java", "nativeps", "PsNative.java")).read()
    natives = set(re.findall(r"public static native [\w\[\]]+ (\w+)\(", java))
    impl = set(re.findall(r"Java_nativeps_PsNative_(\w+)\(", open(shim).read()))
    assert natives and natives == impl, (sorted(natives - impl), sorted(impl - natives))

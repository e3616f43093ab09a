"""CPU tests of the boundary: the shared library loads, exports every symbol include/uwvkf.h declares, refuses to
run without a GPU (no CPU fallback), and the host-side helpers (synthetic inputs, sharding) behave."""
import ctypes as C
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "uwvkf.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(uwvkf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    names = declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/uwvkf.h but not exported: {missing}"
    assert b"sm_100a" in lib.uwvkf_version()


def test_library_contains_sm100a_code(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", pkg.library_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_no_cpu_fallback(pkg):
    """Without a CUDA device every constructor fails loudly; nothing routes to a CPU path."""
    lib = pkg.load_library()
    h = C.c_void_p()
    rc = lib.uwvkf_create(0, 8, 0, C.byref(h))
    if rc == 0:      # running on a GPU box
        lib.uwvkf_destroy(h)
        pytest.skip("a CUDA device is present")
    assert rc == -2 and b"no CPU fallback" in lib.uwvkf_last_error()
    syn = importlib.import_module("slam-uwv_kalman_filters_b200.synthetic")
    with pytest.raises(pkg.UwvkfError) as ei:
        pkg.PoseUKFBatch(4, syn.pose_params(syn.pose_config()))
    assert ei.value.code == -2
    for cls in (pkg.VelocityUKFBatch, pkg.BottomUKFBatch):
        with pytest.raises(pkg.UwvkfError):
            cls(4)


def test_missing_library_raises(pkg, monkeypatch):
    capi = importlib.import_module("slam-uwv_kalman_filters_b200.capi")
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "library_path", lambda: "/nonexistent/libuwvkf.so")
    with pytest.raises(capi.UwvkfError, match="no CPU path"):
        capi.load_library()


def test_argument_validation_without_device(pkg):
    lib = pkg.load_library()
    assert lib.uwvkf_create(0, 0, 0, C.byref(C.c_void_p())) == -1          # batch must be positive
    assert lib.uwvkf_create(99, 1, 0, C.byref(C.c_void_p())) == -1         # unknown filter kind
    assert lib.uwvkf_predict(None, C.c_double(0.01)) == -1                 # null handle
    assert lib.uwvkf_destroy(None) == 0


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no file of the package may reference it."""
    pkg_dir = os.path.join(ROOT, "slam-uwv_kalman_filters_b200")
    for dp, _, files in os.walk(pkg_dir):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dp, fn)).read()
                assert "oracle_lib" not in text and "libuwvkf_oracle" not in text and "oracle/" not in text.replace("oracle/manifold.hpp", ""), fn


def test_synthetic_inputs_are_deterministic_and_shardable(syn):
    a = syn.pose_inputs(np.arange(0, 64), 7)
    b = syn.pose_inputs(np.arange(32, 64), 7)
    for k in a:
        assert np.array_equal(a[k][32:], b[k]), k            # a rank regenerates exactly its slice
    assert not np.array_equal(a["dvl"][0], a["dvl"][1])
    n = syn.noise(np.arange(200000), 3, 5)
    assert abs(n.mean()) < 0.01 and abs(n.std() - 1.0) < 0.01 and np.abs(n).max() <= 2 * np.sqrt(3) + 1e-12
    mu, cov = syn.pose_initial_batch(np.arange(8))
    assert np.abs(np.linalg.norm(mu[:, 3:7], axis=1) - 1).max() < 1e-15 and cov.shape == (8, 53, 53)


def test_shard_ranges_and_ensemble_reduction_identity(syn):
    par = importlib.import_module("slam-uwv_kalman_filters_b200.parallel")
    for total, world in [(10, 3), (1 << 20, 8), (7, 8)]:
        r = [par.shard_range(total, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == total and all(r[k][1] == r[k + 1][0] for k in range(world - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    mu, _ = syn.pose_initial_batch(np.arange(100), with_cov=False)
    parts = sum(par.local_ensemble_stats(mu[slice(*par.shard_range(100, k, 4))]) for k in range(4))
    full = par.local_ensemble_stats(mu)
    assert np.allclose(parts, full, rtol=1e-13)
    n, mean, cov = par.ensemble_mean_cov(full, 54)
    assert n == 100 and np.allclose(mean, mu.mean(0)) and np.allclose(cov, np.cov(mu.T, bias=True), atol=1e-9)


def test_converted_temporaries_outlive_the_pointer(pkg):
    """`_ptr(_f64(list))` used inline in an argument list must not hand out a dangling pointer: the pointer object keeps the
    converted temporary alive (numpy's small-block cache would otherwise give the next temporary the same address)."""
    capi = importlib.import_module("slam-uwv_kalman_filters_b200.capi")
    a, b = capi._ptr(capi._f64([0, 0, 1])), capi._ptr(capi._f64([5, 6, 7]))
    assert a.value != b.value
    assert np.array_equal(np.ctypeslib.as_array(C.cast(a, C.POINTER(C.c_double)), (3,)), [0, 0, 1])
    assert np.array_equal(np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_double)), (3,)), [5, 6, 7])
    c = capi._ptr(capi._f64(np.arange(6, dtype=np.float32).reshape(2, 3)[:, ::2]))      # float32 + non-contiguous -> copy
    assert np.array_equal(np.ctypeslib.as_array(C.cast(c, C.POINTER(C.c_double)), (4,)), [0, 2, 3, 5])
    assert capi._ptr(None) is None and capi._ptr(1234).value == 1234


def test_parity_gates_catch_small_block_errors():
    """tests/parity.py: the per-component and scaled gates must fail where the global (north_star) metrics are blind."""
    import parity
    n = 53
    sd = np.full(n, 1.0); sd[12:15] = 1e-5            # a gyro-bias-like block
    cov = np.diag(sd ** 2); cov[20, 20] = 10.0
    mu = np.zeros(54); mu[6] = 1.0; mu[53] = 1025.0; mu[13:16] = 1e-5
    bad_mu = mu.copy(); bad_mu[13] *= 1.1             # 10 % error on a 1e-5 bias
    bad_cov = cov.copy(); bad_cov[12, 12] *= 2.0      # small variance entirely wrong
    e = parity.all_errs(bad_mu, bad_cov, mu, cov)
    assert e[0] <= parity.TOL_STATE and e[1] <= parity.TOL_COV            # the global metrics pass ...
    assert e[2] > 1e-2 and e[3] > 0.5                                     # ... the new gates do not
    with pytest.raises(AssertionError):
        parity.check_all(bad_mu[None], bad_cov[None], mu[None], cov[None], "synthetic")
    assert parity.check_all(mu[None], cov[None], mu[None], cov[None], "identity") == (0.0, 0.0, 0.0, 0.0)

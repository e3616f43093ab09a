"""Device math of the CUDA engine (through uwvkf_debug_eval) against the oracle's double-precision restatements:
the small-argument fast paths of the transverse Mercator inverse, the SO3 chart series branches, the Newton reciprocal
square root and the tensor-path warp sum."""
import numpy as np
import pytest

import oracle_lib as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ukf(pkg, syn):
    u = pkg.PoseUKFBatch(1, syn.pose_params(syn.pose_config()))
    yield u
    u.close()


def rows8(*cols):
    n = len(cols[0])
    r = np.zeros((n, 8))
    for k, c in enumerate(cols):
        r[:, k] = c
    return r


def test_tmerc_inverse_fast_and_generic_paths(ukf, syn):
    rng = np.random.default_rng(1)
    # inside the Taylor range (|x|,|y| < 398 km), at its edge, and far outside (generic sincos / exp path)
    x = np.concatenate([rng.uniform(-2e3, 2e3, 200), rng.uniform(-3.9e5, 3.9e5, 200), rng.uniform(-2e6, 2e6, 100), [0.0, 3.98e5, -3.99e5]])
    y = np.concatenate([rng.uniform(-2e3, 2e3, 200), rng.uniform(-3.9e5, 3.9e5, 200), rng.uniform(-1.5e6, 1.5e6, 100), [0.0, -3.98e5, 3.99e5]])
    out = ukf.debug_eval(ukf.DEBUG_TMERC_SINCOS_LAT, rows8(x, y))
    lat = np.array([orc.tmerc_nav_to_world(syn.LAT0, syn.LON0, a, b)[0] for a, b in zip(x, y)])
    assert np.abs(out[:, 2] - lat).max() < 5e-15 * 2        # generic device path vs oracle (both libm-accurate)
    assert np.abs(out[:, 0] - np.sin(lat)).max() < 1e-15 * 4
    assert np.abs(out[:, 1] - np.cos(lat)).max() < 1e-15 * 4


@pytest.mark.parametrize("scale", [1e-9, 1e-3, 1.1e-2, 0.3, 0.6, 2.5])
def test_so3_charts_match_oracle(ukf, scale):
    rng = np.random.default_rng(int(scale * 1e9) % 1000 + 3)
    n = 256
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    v = rng.normal(size=(n, 3)) * scale
    r = np.zeros((n, 8)); r[:, 0:4] = q; r[:, 4:7] = v; r[:, 7] = 1.0
    gp = ukf.debug_eval(ukf.DEBUG_SO3_BOXPLUS, r)
    op = np.stack([orc.so3_boxplus(q[i], v[i], 1.0) for i in range(n)])
    assert np.abs(gp - op).max() < 1e-15
    r2 = np.zeros((n, 8)); r2[:, 0:4] = op; r2[:, 4:8] = q
    gm = ukf.debug_eval(ukf.DEBUG_SO3_BOXMINUS, r2)[:, 0:3]
    om = np.stack([orc.so3_boxminus(op[i], q[i]) for i in range(n)])
    assert np.abs(gm - om).max() < 4e-15 * max(1.0, scale), "device boxminus vs oracle"
    if scale < 1.0:   # rotation below pi: the chart round trip is the identity
        assert np.abs(gm - v).max() < 1e-14, "(q (+) v) (-) q = v"


def test_reciprocal_square_root_is_correctly_rounded_to_an_ulp(ukf):
    rng = np.random.default_rng(7)
    d = np.concatenate([10.0 ** rng.uniform(-30, 30, 2000), [1.0, 4.0, 2.0, 1e-300 * 1e10, 1e300]])
    out = ukf.debug_eval(ukf.DEBUG_RSQRT, rows8(d))
    ref = 1.0 / np.sqrt(d.astype(np.longdouble))
    for k in (0, 1):
        assert (np.abs(out[:, k] - ref) / ref).max() < 2.3e-16
    assert (np.abs(out[:, 2] - 1.0 / d.astype(np.longdouble)) * d).max() < 2.3e-16


def test_tensor_path_warp_sum(ukf):
    rng = np.random.default_rng(11)
    v = rng.normal(size=32 * 64) * 10.0 ** rng.integers(-3, 4, size=32 * 64)
    out = ukf.debug_eval(ukf.DEBUG_WARP_SUM, rows8(v))
    ref = np.repeat(v.reshape(-1, 32).sum(axis=1), 32)
    scale = np.repeat(np.abs(v).reshape(-1, 32).sum(axis=1), 32)
    assert (np.abs(out[:, 0] - ref) / scale).max() < 1e-15
    assert (np.abs(out[:, 1] - ref) / scale).max() < 1e-15
    # every lane of a warp receives the same value
    assert np.array_equal(out[:, 0].reshape(-1, 32), np.repeat(out[::32, 0], 32).reshape(-1, 32))


def test_phase_probe_is_absent_from_the_product_build(ukf, pkg):
    """uwvkf_debug_phase_cycles only works in a -DUWVKF_PHASE_TIMING build (tools/phase_timing.py); the product library says so."""
    with pytest.raises(pkg.UwvkfError):
        ukf.debug_phase_cycles()

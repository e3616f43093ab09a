"""CPU tests that pin the oracle (it has no reference golden vectors to lean on: test/suite.cpp:1-6 is empty):
closed-form Kalman filter, LAPACK dpotrf, manifold identities, high-precision geodesy, committed fixtures."""
import os

import mpmath as mp
import numpy as np
import pytest
import scipy.linalg

import oracle_lib as orc
import structured_ref as sr

rng = np.random.default_rng(20181018)
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def spd(n, cond=1e3):
    A = rng.normal(size=(n, n))
    Qm, _ = np.linalg.qr(A)
    return (Qm * np.logspace(0, -np.log10(cond), n)) @ Qm.T


# ---- UKF core: for linear models the ukfom recipe IS the Kalman filter (pins weights 1/(2n+1), 1/2 and the unscaled sqrt)
@pytest.mark.parametrize("n,m", [(4, 3), (6, 2), (12, 2), (20, 6)])
def test_linear_kalman_known_answer(n, m):
    P, mu = spd(n), rng.normal(size=n)
    F, c, Q = rng.normal(size=(n, n)), rng.normal(size=n), spd(n, 10) * 0.1
    mu1, P1, st = orc.linear_predict(mu, P, F, c, Q)
    assert st == 0
    assert np.allclose(mu1, F @ mu + c, rtol=0, atol=1e-13)
    assert np.allclose(P1, F @ P @ F.T + Q, rtol=1e-12, atol=1e-13)
    H, R, z = rng.normal(size=(m, n)), spd(m, 10) * 0.5, rng.normal(size=m)
    mu2, P2, st = orc.linear_update(mu, P, H, z, R)
    S = H @ P @ H.T + R
    K = P @ H.T @ np.linalg.inv(S)
    assert np.allclose(mu2, mu + K @ (z - H @ mu), atol=1e-12)
    assert np.allclose(P2, P - K @ S @ K.T, atol=1e-12)


@pytest.mark.parametrize("n", [3, 6, 53])
@pytest.mark.parametrize("cond", [1e2, 1e8, 1e12])
def test_cholesky_matches_lapack_dpotrf(n, cond):
    P = spd(n, cond)
    L, st = orc.cholesky_lower(P)
    Lref, info = scipy.linalg.lapack.dpotrf(P, lower=1)
    assert st == 0 and info == 0
    assert np.abs(L - np.tril(Lref)).max() <= 1e-9 * np.abs(Lref).max() * max(1, cond * 1e-12)
    assert np.abs(L @ L.T - P).max() <= 1e-14 * n * np.abs(P).max()


def test_cholesky_flags_non_spd():
    P = spd(5)
    P[2, 2] = -1.0
    assert orc.cholesky_lower(P)[1] == 1


# ---- manifolds
def rand_quat():
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


@pytest.mark.parametrize("scale", [1e-9, 1e-3, 1.1e-2, 0.5, 2.5])
def test_so3_chart_identities(scale):
    for _ in range(20):
        q, d = rand_quat(), rng.normal(size=3)
        d *= scale / np.linalg.norm(d)
        y = orc.so3_boxplus(q, d)
        assert abs(np.linalg.norm(y) - 1) < 1e-14
        assert np.abs(orc.so3_boxminus(y, q) - d).max() < 1e-13 * max(1.0, scale)      # (x + d) - x = d
        p = rand_quat()
        back = orc.so3_boxplus(q, orc.so3_boxminus(p, q))                              # x + (y - x) = y  (up to sign)
        assert min(np.abs(back - p).max(), np.abs(back + p).max()) < 1e-13


def test_so3_exp_against_high_precision_and_branch_continuity():
    mp.mp.dps = 40
    q = np.array([0.0, 0.0, 0.0, 1.0])
    for ang in [1e-8, 1e-3, 0.0220, 0.0221, 0.0222, 0.3, 3.0]:        # 2*sqrt(eps^(1/4)) = 0.0221 is the series/libm switch
        v = np.array([ang, 0.0, 0.0])
        e = orc.so3_boxplus(q, v)
        assert abs(e[3] - float(mp.cos(mp.mpf(ang) / 2))) < 2e-16
        assert abs(e[0] - float(mp.sin(mp.mpf(ang) / 2))) < 2e-16 * max(1.0, ang)


def test_quaternion_rotation_matches_matrix():
    for _ in range(10):
        q, v = rand_quat(), rng.normal(size=3)
        x, y, z, w = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.abs(orc.quat_rot(q, v) - R @ v).max() < 1e-14


@pytest.mark.parametrize("scale", [1e-8, 1e-2, 1.0, 3.0])
def test_s2_chart_identities(scale):
    for _ in range(20):
        x = rng.normal(size=3)
        x /= np.linalg.norm(x)
        d = rng.normal(size=2)
        d *= scale / np.linalg.norm(d)
        y = orc.s2_boxplus(x, d)
        assert abs(np.linalg.norm(y) - 1) < 1e-14
        assert abs(np.arccos(np.clip(x @ y, -1, 1)) - scale) < 1e-7 + 1e-12 * scale     # geodesic distance = |d|
        assert np.abs(orc.s2_boxminus(y, x) - d).max() < 1e-12 * max(1.0, scale)
    ex = np.array([1.0, 0.0, 0.0])   # the chart's reference axis itself
    assert np.abs(orc.s2_boxminus(orc.s2_boxplus(ex, np.array([0.1, -0.2])), ex) - [0.1, -0.2]).max() < 1e-15


# ---- geodesy
def test_transverse_mercator_against_high_precision_meridian_arc():
    """On the central meridian the northing must equal the meridian arc between lat0 and lat (mpmath quadrature)."""
    mp.mp.dps = 30
    a, f = mp.mpf(6378137), 1 / mp.mpf("298.257223563")
    e2 = f * (2 - f)
    arc = lambda p0, p1: mp.quad(lambda p: a * (1 - e2) / (1 - e2 * mp.sin(p) ** 2) ** mp.mpf(1.5), [p0, p1])
    lat0, lon0 = 0.92677, 0.15446
    for dlat in [1e-4, 1e-3, 1e-2, -5e-3]:
        x, y = orc.tmerc_world_to_nav(lat0, lon0, lat0 + dlat, lon0)
        assert abs(x - float(arc(mp.mpf(lat0), mp.mpf(lat0 + dlat)))) < 2e-9 and abs(y) < 1e-9
    # west of the meridian y is positive (NWU) and the local scale is ~1
    x, y = orc.tmerc_world_to_nav(lat0, lon0, lat0, lon0 - 1e-5)
    nu = float(a / mp.sqrt(1 - e2 * mp.sin(lat0) ** 2))
    assert abs(y - 1e-5 * nu * np.cos(lat0)) < 1e-6
    assert abs(x - y * y * np.tan(lat0) / (2 * nu)) < 1e-7       # meridian convergence: second-order northing


def test_transverse_mercator_round_trip():
    lat0, lon0 = 0.92677, 0.15446
    for x, y in [(0, 0), (1000, -2000), (5e4, 8e4), (-1e5, 1e5)]:
        lat, lon = orc.tmerc_nav_to_world(lat0, lon0, x, y)
        assert np.abs(orc.tmerc_world_to_nav(lat0, lon0, lat, lon) - [x, y]).max() < 5e-9


def test_gravity_model():
    assert abs(orc.gravity_wgs84(0.0, 0.0) - 9.7803267714) < 1e-12
    assert abs(orc.gravity_wgs84(np.pi / 2, 0.0) - 9.8321849378) < 2e-6   # published WGS-84 polar gravity
    assert orc.gravity_wgs84(0.9, 1000.0) < orc.gravity_wgs84(0.9, 0.0)


# ---- PoseUKF pieces
@pytest.fixture(scope="module")
def scen(syn):
    cfg = syn.pose_config()
    return dict(cfg=cfg, par=syn.pose_params(cfg), Q=syn.pose_process_noise(cfg, 0.01))


def test_config_builders_match_constructor_restatement(syn, scen):
    """synthetic.pose_initial / pose_process_noise (host side) == the oracle's restatement of src/PoseUKF.cpp:288-439."""
    M, Dl, Dq = syn.uwv_matrices()
    mu0, cov0 = syn.pose_initial(scen["cfg"])
    mu_o, cov_o = orc.pose_init_from_config(np.zeros(3), np.eye(3), np.array([0, 0, 0, 1.0]), np.diag([1e-4, 1e-4, 1e-2]), scen["cfg"],
                                            M, Dl, Dq)
    assert np.array_equal(mu_o, mu0) and np.array_equal(cov_o, cov0)
    assert np.allclose(orc.pose_process_noise_from_config(scen["cfg"], 0.01), scen["Q"], rtol=1e-15, atol=0)
    assert abs(mu0[19] - orc.gravity_wgs84(syn.LAT0, 0.0)) < 1e-15
    # a rotated IMU mounting rotates the bias blocks
    c, s = np.cos(0.3), np.sin(0.3)
    Rib = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    cfg2 = scen["cfg"].copy()
    cfg2[syn.CFG_GYRO_BIAS_INSTABILITY:syn.CFG_GYRO_BIAS_INSTABILITY + 3] = [1e-6, 2e-6, 3e-6]
    _, cov_r = orc.pose_init_from_config(np.zeros(3), np.eye(3), np.array([0, 0, 0, 1.0]), np.eye(3) * 1e-4, cfg2, M, Dl, Dq, Rib)
    assert np.allclose(cov_r[12:15, 12:15], Rib @ np.diag(np.array([1e-6, 2e-6, 3e-6]) ** 2) @ Rib.T, rtol=1e-14)


def test_process_model_fixed_point(syn, scen):
    """Gyro reading exactly the earth rate, zero velocity / acceleration, states at their Markov offsets:
    the mean must not move (attitude stationary in the rotating frame)."""
    mu0, cov0 = syn.pose_initial(scen["cfg"])
    off = np.concatenate([mu0[20:47], mu0[53:54]])
    earth = np.array([syn.EARTHW * np.cos(syn.LAT0), 0.0, syn.EARTHW * np.sin(syn.LAT0)])
    cov = cov0 * 1e-6
    mu1, cov1, st, _ = orc.pose_predict(mu0, cov, scen["Q"], earth, scen["par"], off, 0.01)
    assert st == 0
    assert np.abs(mu1 - mu0).max() < 1e-12
    assert np.all(np.diag(cov1) >= np.diag(cov) * (1 - 1e-3))


def test_structured_transform_equals_literal_recipe(syn, scen):
    """The O(n^2) prefix-structured evaluation used by the CUDA engine == the literal 107-sigma-point recipe."""
    filt = np.array([3, 11])
    mu, cov = syn.pose_initial_batch(filt, scen["cfg"])
    par, Q = scen["par"], scen["Q"]
    nav_lat = lambda x, y: orc.tmerc_nav_to_world(par[0], par[1], x, y)[0]
    for f in range(2):
        m, c = mu[f], cov[f]
        off = np.concatenate([m[20:47], m[53:54]])
        for step in range(1, 6):
            inp = syn.pose_inputs(filt[f:f + 1], step)
            w, z = inp["gyro"][0], inp["dvl"][0]
            m1, c1, st, _ = orc.pose_predict(m, c, Q, w, par, off, 0.01)
            m2, c2 = sr.pose_predict(m, c, Q, w, par, off, 0.01, nav_lat)
            assert np.abs(m2 - m1).max() < 1e-13 and np.abs(c2 - c1).max() < 1e-12 * np.abs(c1).max()
            m3, c3, st, _ = orc.pose_update(orc.VELOCITY, m1, c1, z, syn.R_DVL, None, w, par)
            m4, c4, ok = sr.pose_update(m1, c1, z, syn.R_DVL, sr.h_velocity, 9)
            assert ok and np.abs(m4 - m3).max() < 1e-13 and np.abs(c4 - c3).max() < 1e-12 * np.abs(c3).max()
            m, c = m3, c3


def test_affine_tail_transform_equals_literal_recipe(syn, scen):
    """Pressure (prefix 19 + tail {water density}) and ADCP (prefix 9 + tail {water velocities, ADCP bias}) evaluated with
    the affine-tail Schur-complement identities == the literal recipe over all 107 sigma points."""
    filt = np.array([5, 17])
    mu, cov = syn.pose_initial_batch(filt, scen["cfg"])
    par, Q = scen["par"], scen["Q"]
    for f in range(2):
        m, c = mu[f], cov[f]
        off = np.concatenate([m[20:47], m[53:54]])
        for step in range(1, 5):
            inp = syn.pose_inputs(filt[f:f + 1], step)
            w = inp["gyro"][0]
            m, c, st, _ = orc.pose_predict(m, c, Q, w, par, off, 0.01)
            m, c, st, _ = orc.pose_update(orc.VELOCITY, m, c, inp["dvl"][0], syn.R_DVL, None, w, par)
            aux = np.array([0.1, -0.05, 0.2])
            m1, c1, st, _ = orc.pose_update(orc.PRESSURE, m, c, inp["pressure"][0], syn.R_PRESSURE, aux, w, par)
            h, jac = sr.make_h_pressure(par[17], aux)
            m2, c2, ok = sr.pose_update_tail(m, c, inp["pressure"][0], syn.R_PRESSURE, h, 19, [52], jac)
            assert st == 0 and ok
            assert np.abs(m2 - m1).max() < 1e-12 * max(1.0, np.abs(m1).max()) and np.abs(c2 - c1).max() < 1e-12 * np.abs(c1).max()
            wgt = 0.3
            m3, c3, st, _ = orc.pose_update(orc.WATER_VELOCITY, m1, c1, inp["adcp"][0], syn.R_ADCP, np.array([wgt]), w, par)
            h, jac = sr.make_h_adcp(wgt)
            m4, c4, ok = sr.pose_update_tail(m1, c1, inp["adcp"][0], syn.R_ADCP, h, 9, range(46, 52), jac, gate=lambda d2: d2 <= 5.991)
            assert ok == (st & 4 == 0)
            assert np.abs(m4 - m3).max() < 1e-12 * max(1.0, np.abs(m3).max()) and np.abs(c4 - c3).max() < 1e-12 * np.abs(c3).max()
            m, c = m3, c3


def test_gate_and_nan_semantics(syn, scen):
    mu, cov = syn.pose_initial(scen["cfg"])
    w = np.zeros(3)
    # d2p95 gate (src/PoseUKF.cpp:275-286): a 10 m/s ADCP outlier is refused, state untouched
    m, c, st, d2 = orc.pose_update(orc.WATER_VELOCITY, mu, cov, np.array([10.0, 10.0]), 4e-4 * np.eye(2), np.array([0.5]), w, scen["par"])
    assert st & 4 and d2 > 5.991 and np.array_equal(m, mu) and np.array_equal(c, cov)
    m, c, st, d2 = orc.pose_update(orc.WATER_VELOCITY, mu, cov, np.array([0.01, 0.0]), 4e-4 * np.eye(2), np.array([0.5]), w, scen["par"])
    assert st == 0 and d2 <= 5.991 and not np.array_equal(m, mu)
    # checkMeasurment: NaN is refused
    _, _, st, _ = orc.pose_update(orc.VELOCITY, mu, cov, np.array([np.nan, 0, 0]), np.eye(3), None, w, scen["par"])
    assert st == 2


def test_fossen_model_consistency(syn):
    """VelocityUKF process model: constant thrust drives the surge velocity to the damping equilibrium."""
    uwv = syn.uwv_params_flat()
    mu, cov = np.array([0.0, 0, 0, -10.0]), np.eye(4) * 1e-4
    sim = orc.vel_setup_motion_model(mu, np.zeros(3))
    tau = np.array([90.0, 0, 0, 0, 0, 0])
    for _ in range(3000):
        mu, cov, sim, st = orc.vel_predict(mu, cov, np.diag([1e-4] * 3 + [0.0]), sim, np.zeros(3), tau, uwv, 0.01)
    v = mu[0]                                     # 30 v + 60 v|v| = 90  ->  v = 1
    assert abs(v - 1.0) < 1e-4 and abs(sim[4] - 1.0) < 1e-9    # the UKF mean sits slightly below (E[v|v|] > E[v]|E[v]|)


# ---- committed fixtures
def test_oracle_reproduces_golden_pose_sequence(syn, scen):
    g = np.load(os.path.join(GOLD, "pose_sequence.npz"))
    mus, covs = g["mu"], g["cov"]
    B = mus.shape[1]
    filt = np.arange(B)
    mu, cov = syn.pose_initial_batch(filt, scen["cfg"])
    assert np.array_equal(mu, mus[0]) and np.array_equal(cov, covs[0])
    off = np.concatenate([mu[:, 20:47], mu[:, 53:54]], axis=1)
    par, Q = scen["par"], scen["Q"]
    for s in range(1, int(g["steps"]) + 1):
        inp = syn.pose_inputs(filt, s)
        for f in range(B):
            w = inp["gyro"][f]
            m, c, st, _ = orc.pose_predict(mu[f], cov[f], Q, w, par, off[f], 0.01)
            m, c, st, _ = orc.pose_update(orc.VELOCITY, m, c, inp["dvl"][f], syn.R_DVL, None, w, par)
            m, c, st, _ = orc.pose_update(orc.ACCELERATION, m, c, inp["acc"][f], syn.R_ACC, None, w, par)
            if s % 4 == 0:
                m, c, st, _ = orc.pose_update(orc.PRESSURE, m, c, inp["pressure"][f], syn.R_PRESSURE, None, w, par)
                m, c, st, _ = orc.pose_update(orc.WATER_VELOCITY, m, c, inp["adcp"][f], syn.R_ADCP, np.array([0.5]), w, par)
            mu[f], cov[f] = m, c
        assert np.abs(mu - mus[s]).max() <= 1e-12 * max(1.0, np.abs(mus[s]).max())
        assert np.abs(cov - covs[s]).max() <= 1e-11 * np.abs(covs[s]).max()


def test_oracle_reproduces_golden_small_sequences(syn):
    g = np.load(os.path.join(GOLD, "small_sequences.npz"))
    B = 3
    filt = np.arange(B)
    mu, cov = syn.bottom_initial_batch(filt)
    for s in range(1, 11):
        inp = syn.bottom_inputs(filt, s)
        for f in range(B):
            mu[f], cov[f], _ = orc.bottom_predict(mu[f], cov[f], np.eye(3), inp["velocity"][f], 0.1)
            for k in range(4):
                mu[f], cov[f], _ = orc.bottom_update_range(mu[f], cov[f], inp["ranges"][f, k:k + 1], np.array([4e-4]), syn.BEAMS[k],
                                                           syn.BEAM_ORIGIN)
            if s % 5 == 0:
                mu[f], cov[f], _ = orc.bottom_update_normal(mu[f], cov[f], inp["normal"][f], 1e-4 * np.eye(2))
        assert np.abs(np.concatenate([mu, cov.reshape(B, -1)], 1) - g["bottom"][s]).max() < 1e-12

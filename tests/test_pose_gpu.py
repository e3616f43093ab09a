"""GPU parity tests of the PoseUKF hot path: CUDA engine (through the C ABI) vs the CPU oracle.

Tolerances are BASELINE.json's north_star: <= 1e-9 relative on the state and <= 1e-8 on the covariance per step.
"""
import numpy as np
import pytest

import oracle_lib as orc
from parity import TOL_COV, TOL_STATE, check_all, cov_err, cov_err_scaled, state_err

pytestmark = pytest.mark.gpu

DT = 0.01


@pytest.fixture(scope="module")
def scen(syn):
    cfg = syn.pose_config()
    par = syn.pose_params(cfg)
    Q = syn.pose_process_noise(cfg, DT)
    return dict(cfg=cfg, par=par, Q=Q)


def make_batch(pkg, syn, scen, B, first_filter=0):
    filt = np.arange(first_filter, first_filter + B)
    mu, cov = syn.pose_initial_batch(filt, scen["cfg"])
    ukf = pkg.PoseUKFBatch(B, scen["par"])
    ukf.set_process_noise(scen["Q"])
    ukf.set_state(mu, cov)
    offsets = np.concatenate([mu[:, 20:47], mu[:, 53:54]], axis=1)
    return ukf, filt, mu, cov, offsets


def oracle_predict(mu, cov, scen, w, offsets, dt=DT):
    out_mu, out_cov = np.empty_like(mu), np.empty_like(cov)
    for f in range(mu.shape[0]):
        out_mu[f], out_cov[f], st, _ = orc.pose_predict(mu[f], cov[f], scen["Q"], w[f], scen["par"], offsets[f], dt)
        assert st == 0
    return out_mu, out_cov


def oracle_update(kind, mu, cov, scen, z, R, aux, w, valid=None):
    out_mu, out_cov = mu.copy(), cov.copy()
    sts = np.zeros(mu.shape[0], dtype=np.int64)
    for f in range(mu.shape[0]):
        if valid is not None and not valid[f]:
            continue
        Rf = R[f] if R.ndim == 3 else R
        out_mu[f], out_cov[f], sts[f], _ = orc.pose_update(kind, mu[f], cov[f], z[f], Rf, aux, w[f], scen["par"])
    return out_mu, out_cov, sts


def assert_parity(gmu, gcov, omu, ocov, what):
    worst = check_all(gmu, gcov, omu, ocov, what)
    assert np.allclose(gcov, np.swapaxes(gcov, 1, 2)), "covariance getter must return a symmetric matrix"
    return worst


def test_set_get_roundtrip(pkg, syn, scen):
    ukf, _, mu, cov, _ = make_batch(pkg, syn, scen, 37)
    rng = np.random.default_rng(1)
    A = rng.normal(size=(37, 53, 53))
    cov = A @ np.swapaxes(A, 1, 2)
    ukf.set_state(mu, cov)
    gmu, gcov = ukf.get_state()
    assert np.array_equal(gmu, mu)
    assert np.array_equal(gcov, np.tril(cov) + np.swapaxes(np.tril(cov, -1), 1, 2))
    # partial ranges and the shared-covariance form
    ukf.set_state(mu[5:9] + 1.0, cov[0], first=5)
    gmu2, gcov2 = ukf.get_state(first=4, count=6)
    assert np.array_equal(gmu2[0], mu[4]) and np.array_equal(gmu2[1:5], mu[5:9] + 1.0) and np.array_equal(gmu2[5], mu[9])
    assert np.array_equal(gcov2[2], gcov[0])
    ukf.close()


def test_predict_parity(pkg, syn, scen):
    B = 48
    ukf, filt, mu, cov, off = make_batch(pkg, syn, scen, B)
    w = syn.pose_inputs(filt, 1)["gyro"]
    ukf.integrate_rotation_rate(w)
    ukf.prediction_step(DT)
    gmu, gcov = ukf.get_state()
    omu, ocov = oracle_predict(mu, cov, scen, w, off)
    assert_parity(gmu, gcov, omu, ocov, "predict from block-diagonal Sigma0")
    assert not ukf.get_status().any()
    # second predict from a now dense covariance
    ukf.prediction_step(DT)
    gmu2, gcov2 = ukf.get_state()
    omu2, ocov2 = oracle_predict(gmu, gcov, scen, w, off)
    assert_parity(gmu2, gcov2, omu2, ocov2, "predict from dense Sigma")
    ukf.close()


def _dense_state(pkg, syn, scen, B, steps=3):
    """A batch a few fused steps into the run, so that Sigma is dense and correlations are realistic."""
    ukf, filt, mu, cov, off = make_batch(pkg, syn, scen, B)
    for s in range(1, steps + 1):
        inp = syn.pose_inputs(filt, s)
        ukf.predict_update(DT, inp["gyro"], pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
        ukf.integrate_measurement(pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC)
    return ukf, filt, off


@pytest.mark.parametrize("kind", ["velocity", "acceleration", "xy", "z", "pressure", "adcp", "geographic", "delayed_xy"])
def test_update_parity(pkg, syn, scen, kind):
    B = 24
    ukf, filt, off = _dense_state(pkg, syn, scen, B)
    mu, cov = ukf.get_state()
    inp = syn.pose_inputs(filt, 4)
    w = inp["gyro"]
    ukf.integrate_rotation_rate(w)
    truth = syn.truth(4 * DT)
    rng = np.random.default_rng(5)
    aux = None
    if kind == "velocity":
        k, z, R = pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL
    elif kind == "acceleration":
        k, z, R = pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC
    elif kind == "xy":
        k, z, R = pkg.POSE_XY, mu[:, 0:2] + 0.1 * rng.normal(size=(B, 2)), np.array([[0.04, 0.01], [0.01, 0.09]])
    elif kind == "z":
        k, z, R = pkg.POSE_Z, (truth["z"] + 0.05 * rng.normal(size=(B, 1))), np.array([[0.0025]])
    elif kind == "pressure":
        k, z, R, aux = pkg.POSE_PRESSURE, inp["pressure"], syn.R_PRESSURE, np.array([0.3, -0.1, 0.2])
    elif kind == "adcp":
        k, z, R, aux = pkg.POSE_WATER_VELOCITY, inp["adcp"], syn.R_ADCP, np.array([0.5])
    elif kind == "geographic":
        ll = np.stack([orc.tmerc_nav_to_world(syn.LAT0, syn.LON0, mu[f, 0] + 0.3, mu[f, 1] - 0.2) for f in range(B)])
        k, z, R, aux = pkg.POSE_GEOGRAPHIC, ll, np.array([[1.0, 0.0], [0.0, 1.0]]), np.array([0.5, 0.1, -0.3])
    else:
        k, z, R, aux = pkg.POSE_DELAYED_XY, mu[:, 0:2] + 0.05, np.array([[0.04, 0.0], [0.0, 0.04]]), np.array([0.1, -0.2])
    z = np.ascontiguousarray(z, dtype=np.float64)
    ukf.integrate_measurement(k, z, R, aux=aux)
    gmu, gcov = ukf.get_state()
    omu, ocov, sts = oracle_update(k, mu, cov, scen, z, R, aux, w)
    gst = ukf.get_status()
    assert np.array_equal(gst & 4 != 0, sts & 4 != 0), "gate decisions differ"
    assert not (gst & ~np.uint32(4)).any()
    assert_parity(gmu, gcov, omu, ocov, f"update {kind}")
    ukf.close()


def test_fused_step_equals_separate_calls(pkg, syn, scen):
    B = 40
    a, filt, mu, cov, off = make_batch(pkg, syn, scen, B)
    b, *_ = make_batch(pkg, syn, scen, B)
    for s in range(1, 4):
        inp = syn.pose_inputs(filt, s)
        a.predict_update(DT, inp["gyro"], pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
        b.integrate_rotation_rate(inp["gyro"])
        b.prediction_step(DT)
        b.integrate_measurement(pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
    ma, ca = a.get_state()
    mb, cb = b.get_state()
    assert np.array_equal(ma, mb) and np.array_equal(ca, cb)
    a.close(); b.close()


def _random_states(syn, scen, B, seed):
    """Random filters far from the synthetic scenario: arbitrary attitudes, dense SPD covariances with the scales of a working
    filter (Sigma = D C D with a random correlation matrix C of moderate condition number), perturbed parameter states."""
    rng = np.random.default_rng(seed)
    mu0, cov0 = syn.pose_initial(scen["cfg"])
    sd = np.sqrt(np.diag(cov0))
    mu = np.tile(mu0, (B, 1))
    mu[:, 0:3] += rng.normal(size=(B, 3)) * [50.0, 50.0, 5.0]
    q = rng.normal(size=(B, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    mu[:, 3:7] = q
    mu[:, 7:10] = rng.normal(size=(B, 3)) * 0.5
    mu[:, 10:13] = rng.normal(size=(B, 3)) * 0.2
    mu[:, 13:19] += rng.normal(size=(B, 6)) * 1e-3
    mu[:, 20:47] *= 1.0 + 0.05 * rng.normal(size=(B, 27))
    mu[:, 47:53] = rng.normal(size=(B, 6)) * 0.1
    cov = np.empty((B, 53, 53))
    for f in range(B):
        A = rng.normal(size=(53, 80))
        Cm = A @ A.T / 80.0                                   # random correlation structure, condition number ~ 30
        dg = np.sqrt(np.diag(Cm))
        Cm = Cm / np.outer(dg, dg)
        scale = sd * np.exp(0.5 * rng.normal(size=53))
        cov[f] = Cm * np.outer(scale, scale)
    off = np.concatenate([mu[:, 20:47], mu[:, 53:54]], axis=1)
    return mu, cov, off, rng


@pytest.mark.parametrize("seed", [11, 12])
def test_random_states_parity(pkg, syn, scen, seed):
    """Parity on RANDOM states (not the synthetic trajectory): predict, every vector update, the fused forms."""
    B = 24
    mu, cov, off, rng = _random_states(syn, scen, B, seed)
    ukf = pkg.PoseUKFBatch(B, scen["par"])
    ukf.set_process_noise(scen["Q"])
    ukf.set_state(mu, cov)
    ukf.set_offsets(off)
    w = rng.normal(size=(B, 3)) * 0.2
    ukf.integrate_rotation_rate(w)
    # predict
    ukf.prediction_step(DT)
    gmu, gcov = ukf.get_state()
    omu, ocov = oracle_predict(mu, cov, scen, w, off)
    assert_parity(gmu, gcov, omu, ocov, f"random predict (seed {seed})")
    # every vector update from the predicted state, each from the same start
    import structured_ref as sr

    def h_mean(name, m_, aux):   # the measurement model at the mean: where a plausible sample lies
        q, p, v = m_[3:7], m_[0:3], m_[7:10]
        if name == "velocity":
            return sr.qrot(sr.qconj(q), v)
        if name == "acceleration":
            return sr.qrot(sr.qconj(q), m_[10:13] + np.array([0.0, 0.0, m_[19]])) + m_[16:19]
        if name == "xy":
            return p[0:2].copy()
        if name == "z":
            return p[2:3].copy()
        if name == "pressure":
            return np.array([syn.P_ATM - (p + sr.qrot(q, aux))[2] * m_[19] * m_[53]])
        wgt = aux[0]
        wv, wvb = np.array([m_[47], m_[48], 0.0]), np.array([m_[49], m_[50], 0.0])
        return (wgt * sr.qrot(sr.qconj(q), v - wvb) + (1 - wgt) * sr.qrot(sr.qconj(q), v - wv))[0:2] + m_[51:53]

    kinds = {
        "velocity": (pkg.POSE_VELOCITY, orc.VELOCITY, 3, syn.R_DVL, None),
        "acceleration": (pkg.POSE_ACCELERATION, orc.ACCELERATION, 3, syn.R_ACC, None),
        "xy": (pkg.POSE_XY, orc.XY, 2, np.array([[0.04, 0.01], [0.01, 0.09]]), None),
        "z": (pkg.POSE_Z, orc.Z, 1, np.array([[0.0025]]), None),
        "pressure": (pkg.POSE_PRESSURE, orc.PRESSURE, 1, syn.R_PRESSURE, np.array([0.3, -0.1, 0.2])),
        "adcp": (pkg.POSE_WATER_VELOCITY, orc.WATER_VELOCITY, 2, syn.R_ADCP, np.array([0.4])),
    }
    for name, (gk, ok, m, R, aux) in kinds.items():
        ukf.set_state(omu, ocov)
        ukf.clear_status()
        zc = np.stack([h_mean(name, omu[f], aux) for f in range(B)])
        z = np.ascontiguousarray(zc + rng.normal(size=(B, m)) * np.sqrt(np.diag(R)))
        ukf.integrate_measurement(gk, z, R, aux=aux)
        g2, c2 = ukf.get_state()
        o2, oc2, sts = oracle_update(ok, omu, ocov, scen, z, R, aux, w)
        gst = ukf.get_status()
        assert np.array_equal(gst & 4 != 0, sts & 4 != 0), f"{name}: gate decisions differ"
        assert not (gst & ~np.uint32(4)).any(), name
        assert_parity(g2, c2, o2, oc2, f"random update {name} (seed {seed})")
    # fused predict + DVL from the random start
    ukf.set_state(mu, cov)
    ukf.clear_status()
    zc = np.stack([h_mean("velocity", omu[f], None) for f in range(B)])
    z = np.ascontiguousarray(zc + 0.01 * rng.normal(size=(B, 3)))
    ukf.predict_update(DT, w, pkg.POSE_VELOCITY, z, syn.R_DVL)
    g3, c3 = ukf.get_state()
    o3, oc3, _ = oracle_update(orc.VELOCITY, omu, ocov, scen, z, syn.R_DVL, None, w)
    assert_parity(g3, c3, o3, oc3, f"random fused predict + DVL (seed {seed})")
    assert not ukf.get_status().any()
    ukf.close()


def test_two_update_step_equals_separate_calls(pkg, syn, scen):
    """predictionStep + DVL + accelerometer in ONE kernel (uwvkf_pose_predict_update2) is bit-identical to the three separate
    calls -- also with dropout masks on either update, without a prediction (dt = 0), and for a pair that is not fused."""
    B = 96
    a, filt, mu, cov, off = make_batch(pkg, syn, scen, B)
    b, *_ = make_batch(pkg, syn, scen, B)
    va = (np.arange(B) % 3 != 0).astype(np.uint8)          # DVL dropouts
    vb = (np.arange(B) % 5 != 1).astype(np.uint8)          # accelerometer dropouts
    for s in range(1, 6):
        inp = syn.pose_inputs(filt, s)
        ma_, mb_ = (va, vb) if s >= 3 else (None, None)
        dt = 0.0 if s == 5 else DT
        a.predict_update2(dt, inp["gyro"], pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL, pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC,
                          valid_a=ma_, valid_b=mb_)
        b.integrate_rotation_rate(inp["gyro"])
        if dt > 0:
            b.prediction_step(dt)
        b.integrate_measurement(pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL, valid=ma_)
        b.integrate_measurement(pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC, valid=mb_)
        ma, ca = a.get_state()
        mb, cb = b.get_state()
        assert np.array_equal(ma, mb) and np.array_equal(ca, cb), f"step {s}"
    assert not a.get_status().any() and not b.get_status().any()
    # parity of the fused step against the oracle (one more step from the common state)
    inp = syn.pose_inputs(filt, 6)
    mu0, cov0 = a.get_state()
    a.predict_update2(DT, inp["gyro"], pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL, pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC)
    gmu, gcov = a.get_state()
    omu, ocov = oracle_predict(mu0, cov0, scen, inp["gyro"], off)
    omu, ocov, _ = oracle_update(orc.VELOCITY, omu, ocov, scen, inp["dvl"], syn.R_DVL, None, inp["gyro"])
    omu, ocov, _ = oracle_update(orc.ACCELERATION, omu, ocov, scen, inp["acc"], syn.R_ACC, None, inp["gyro"])
    assert_parity(gmu, gcov, omu, ocov, "fused predict + DVL + accelerometer")
    # a NaN sample in the second update: that filter keeps prediction + first update, the others are untouched by it
    zb = inp["acc"].copy(); zb[7, 0] = np.nan
    import torch
    zbd, zad, wd = torch.from_numpy(zb).cuda(), torch.from_numpy(inp["dvl"]).cuda(), torch.from_numpy(inp["gyro"]).cuda()
    b.set_state(mu0, cov0); a.set_state(mu0, cov0)
    a.predict_update2(DT, wd.data_ptr(), pkg.POSE_VELOCITY, zad.data_ptr(), syn.R_DVL, pkg.POSE_ACCELERATION, zbd.data_ptr(), syn.R_ACC,
                      memspace=pkg.DEVICE)
    b.predict_update(DT, inp["gyro"], pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
    m7, c7 = b.get_state(first=7, count=1)
    b.integrate_measurement(pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC)
    ma, ca = a.get_state()
    mb, cb = b.get_state()
    assert np.array_equal(ma[7], m7[0]) and np.array_equal(ca[7], c7[0]) and a.get_status()[7] & pkg.ST_NAN_MEASUREMENT
    keep = np.arange(B) != 7
    assert np.array_equal(ma[keep], mb[keep]) and np.array_equal(ca[keep], cb[keep])
    # an unfused pair (XY + Z) gives the sequential result through the same entry point
    a.clear_status()
    xy, zz = ma[:, 0:2] + 0.05, ma[:, 2:3] - 0.02
    Rxy, Rz = 0.04 * np.eye(2), np.array([[0.0025]])
    a.predict_update2(DT, inp["gyro"], pkg.POSE_XY, xy, Rxy, pkg.POSE_Z, zz, Rz)
    b.set_state(ma, ca)
    b.predict_update(DT, inp["gyro"], pkg.POSE_XY, xy, Rxy)
    b.integrate_measurement(pkg.POSE_Z, zz, Rz)
    assert all(np.array_equal(x, y) for x, y in zip(a.get_state(), b.get_state()))
    a.close(); b.close()


@pytest.mark.parametrize("kind", ["pressure", "adcp", "acceleration", "z"])
def test_fused_predict_update_parity(pkg, syn, scen, kind):
    """predictionStep + integrateMeasurement(kind) in ONE kernel vs the oracle's two calls.  For pressure / ADCP this
    exercises the affine tail on a trailing triangle that still holds the pre-predict covariance in HBM."""
    B = 24
    ukf, filt, off = _dense_state(pkg, syn, scen, B)
    mu, cov = ukf.get_state()
    inp = syn.pose_inputs(filt, 4)
    w = inp["gyro"]
    if kind == "pressure":
        k, z, R, aux = pkg.POSE_PRESSURE, inp["pressure"], syn.R_PRESSURE, np.array([0.3, -0.1, 0.2])
    elif kind == "adcp":
        k, z, R, aux = pkg.POSE_WATER_VELOCITY, inp["adcp"], syn.R_ADCP, np.array([0.25])
    elif kind == "acceleration":
        k, z, R, aux = pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC, None
    else:
        k, z, R, aux = pkg.POSE_Z, np.full((B, 1), syn.truth(4 * DT)["z"]), np.array([[0.0025]]), None
    z = np.ascontiguousarray(z, dtype=np.float64)
    ukf.predict_update(DT, w, k, z, R, aux=aux)
    gmu, gcov = ukf.get_state()
    omu, ocov = oracle_predict(mu, cov, scen, w, off)
    omu, ocov, sts = oracle_update(k, omu, ocov, scen, z, R, aux, w)
    gst = ukf.get_status()
    assert np.array_equal(gst & 4 != 0, sts & 4 != 0), "gate decisions differ"
    assert not (gst & ~np.uint32(4)).any()
    assert_parity(gmu, gcov, omu, ocov, f"fused predict + {kind}")
    ukf.close()


@pytest.mark.parametrize("velocity_only", [False, True])
def test_body_efforts_update_parity(pkg, syn, scen, velocity_only):
    """integrateMeasurement(BodyEffortsMeasurement, only_affect_velocity) (reference src/PoseUKF.cpp:581-602): the Fossen model
    with the inertia / damping blocks of each sigma point (m = 6, tangent prefix 48), or constrainVelocity (prefix 9)."""
    B = 12
    par = scen["par"].copy()
    par[19:22] = [0.2, -0.1, 0.05]   # imu_in_body lever arm
    filt = np.arange(B)
    mu, cov = syn.pose_initial_batch(filt, scen["cfg"])
    ukf = pkg.PoseUKFBatch(B, par)
    ukf.set_process_noise(scen["Q"])
    ukf.set_state(mu, cov)
    off = np.concatenate([mu[:, 20:47], mu[:, 53:54]], axis=1)
    uwv = syn.uwv_params_flat()
    uwv[108:114] = [0.0, 0.0, 0.02, 0.0, 0.0, -0.01]   # cog / cob levers so that the restoring torque is exercised
    uwv[115] = 1490.0
    with pytest.raises(pkg.UwvkfError) as ei:             # model parameters are required first
        ukf.integrate_measurement(pkg.POSE_BODY_EFFORTS, np.zeros((B, 6)), np.eye(6))
    assert ei.value.code == -5
    ukf.set_model_parameters(uwv)
    for s in range(1, 4):   # a few steps so that Sigma is dense
        inp = syn.pose_inputs(filt, s)
        ukf.predict_update(DT, inp["gyro"], pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
    mu, cov = ukf.get_state()
    inp = syn.pose_inputs(filt, 4)
    w = inp["gyro"]
    ukf.integrate_rotation_rate(w)
    z = np.ascontiguousarray(syn.velocity_inputs(filt, 4)["efforts"])
    R = np.diag([4.0, 4.0, 4.0, 0.25, 0.25, 0.25])
    kind = pkg.POSE_BODY_EFFORTS_VELOCITY_ONLY if velocity_only else pkg.POSE_BODY_EFFORTS
    okind = orc.BODY_EFFORTS_VEL if velocity_only else orc.BODY_EFFORTS
    ukf.integrate_measurement(kind, z, R)
    gmu, gcov = ukf.get_state()
    omu, ocov = mu.copy(), cov.copy()
    for f in range(B):
        omu[f], ocov[f], st, _ = orc.pose_update(okind, mu[f], cov[f], z[f], R, uwv, w[f], par)
        assert st == 0
    assert not ukf.get_status().any()
    assert np.abs(gmu - mu).max() > 1e-6, "the update must move the state"
    assert_parity(gmu, gcov, omu, ocov, f"body efforts (velocity_only={velocity_only})")
    ukf.close()


def _project_marker(mu, syn, marker_pose, features, cam_in_imu, camera):
    """pixel coordinates of the marker features seen from the filter's mean pose (numpy restatement of the camera model)."""
    import structured_ref as sr
    cam_t = mu[0:3] + sr.qrot(mu[3:7], cam_in_imu[0:3])
    cam_q = sr.qmul(mu[3:7], cam_in_imu[3:7])
    px = np.zeros((features.shape[0], 2))
    for k, ft in enumerate(features):
        pc = sr.qrot(sr.qconj(cam_q), marker_pose[0:3] + sr.qrot(marker_pose[3:7], ft) - cam_t)
        px[k] = [camera[0] * pc[0] / pc[2] + camera[2], camera[1] * pc[1] / pc[2] + camera[3]]
    return px


def test_visual_marker_update_parity(pkg, syn, scen):
    """integrateMeasurement(marker_features, ...) on the 59-DOF marker-augmented state (reference src/PoseUKF.cpp:613-654):
    119 sigma points, S2 innovation, one sequential update per corner, leading 53 x 53 block kept."""
    import structured_ref as sr
    B = 8
    ukf, filt, off = _dense_state(pkg, syn, scen, B)
    mu, cov = ukf.get_state()
    rng = np.random.default_rng(3)
    # a marker 4 m ahead of the vehicle (seen from the mean pose of filter 0), slightly rotated, uncertain pose
    mq = sr.so3_exp(np.array([0.05, -0.1, 0.2]))
    marker_pose = np.concatenate([mu[0, 0:3] + sr.qrot(mu[0, 3:7], np.array([4.0, 0.3, -0.2])), sr.qmul(mu[0, 3:7], mq)])
    covm = np.diag([4e-4, 4e-4, 9e-4, 1e-4, 1e-4, 4e-4])
    covm[0, 1] = covm[1, 0] = 1e-4
    pxcov = np.tile(np.array([0.25, 0.05, 0.05, 0.36]), (4, 1))
    px = np.stack([_project_marker(mu[f], syn, marker_pose, syn.MARKER_FEATURES, syn.CAM_IN_BODY, syn.CAMERA) for f in range(B)])
    px += 0.7 * rng.normal(size=px.shape)
    valid = np.ones(B, dtype=np.uint8)
    valid[5] = 0
    ukf.integrate_features(px, pxcov, syn.MARKER_FEATURES, marker_pose, covm, syn.CAMERA, syn.CAM_IN_BODY, valid=valid)
    gmu, gcov = ukf.get_state()
    omu, ocov = mu.copy(), cov.copy()
    for f in range(B):
        if valid[f]:
            omu[f], ocov[f], st = orc.pose_update_features(mu[f], cov[f], px[f], pxcov, syn.MARKER_FEATURES, marker_pose, covm,
                                                           syn.CAMERA, syn.CAM_IN_BODY)
            assert st == 0
    assert not ukf.get_status().any()
    assert np.array_equal(gmu[5], mu[5]) and np.array_equal(gcov[5], cov[5])
    assert np.abs(gmu[0] - mu[0]).max() > 1e-6, "the update must move the state"
    assert_parity(gmu, gcov, omu, ocov, "visual marker update (4 corners)")
    # a NaN corner refuses the whole measurement for that filter (checkMeasurment()); small host batches raise like the reference
    bad = px.copy(); bad[2, 1, 0] = np.nan
    with pytest.raises(pkg.UwvkfError) as ei:
        ukf.integrate_features(bad, pxcov, syn.MARKER_FEATURES, marker_pose, covm, syn.CAMERA, syn.CAM_IN_BODY)
    assert ei.value.code == -4
    ukf.close()


def test_sample_stream_replays_like_direct_calls(pkg, syn, scen):
    """The caller-side loop (uwvkf_stream_*): timestamped gyro / DVL / acceleration / pressure sample sets pushed from host memory
    through a small ring and replayed in time order == the same sequence issued call by call."""
    B, STEPS = 16, 25
    a, filt, mu, cov, off = make_batch(pkg, syn, scen, B)
    b, *_ = make_batch(pkg, syn, scen, B)
    stream = pkg.PoseSampleStream(a, slots=8)
    t0 = 100.0
    for s in range(0, STEPS + 1):
        inp = syn.pose_inputs(filt, max(s, 1))
        t = t0 + s * DT
        valid = (np.arange(B) + s) % 4 != 0
        samples = [(pkg.STREAM_ROTATION_RATE, inp["gyro"], None, None, None),
                   (pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL, None, valid),
                   (pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC, None, None)]
        if s % 10 == 0:
            samples.append((pkg.POSE_PRESSURE, inp["pressure"], syn.R_PRESSURE, np.array([0.3, -0.1, 0.2]), None))
        for kind, z, R, aux, v in samples:
            stream.push(t, kind, z, R, aux, v)
            if stream.pending >= 6:          # the ring has 8 slots: drain before it fills
                stream.run()
        # the same thing call by call: the first measurement of a new timestamp carries the prediction step
        b.integrate_rotation_rate(inp["gyro"])
        if s == 0:
            b.integrate_measurement(pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL, valid=valid)
        else:
            b.predict_update(DT, None, pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL, valid=valid)
        b.integrate_measurement(pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC)
        if s % 10 == 0:
            b.integrate_measurement(pkg.POSE_PRESSURE, inp["pressure"], syn.R_PRESSURE, aux=np.array([0.3, -0.1, 0.2]))
    stream.run(t0 + (STEPS - 1) * DT)        # replay up to the second-to-last timestamp only
    assert stream.pending > 0 and abs(stream.time - (t0 + (STEPS - 1) * DT)) < 1e-12
    stream.run()
    assert stream.pending == 0 and abs(stream.time - (t0 + STEPS * DT)) < 1e-9
    with pytest.raises(pkg.UwvkfError):      # time must not run backwards
        stream.push(t0, pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
    znan = inp["dvl"].copy(); znan[-1, 2] = np.nan
    with pytest.raises(pkg.UwvkfError) as ei:   # a NaN sample is refused before the slot is touched: the ring is unchanged
        stream.push(t0 + (STEPS + 1) * DT, pkg.POSE_VELOCITY, znan, syn.R_DVL)
    assert ei.value.code == -4 and stream.pending == 0
    ma, ca = a.get_state()
    mb, cb = b.get_state()
    # predict dt: the stream derives it from timestamps (t0 + s DT differences), the direct path passes DT itself
    assert np.abs(ma - mb).max() < 1e-9 and np.abs(ca - cb).max() < 1e-9 * np.abs(cb).max()
    assert not a.get_status().any() and not b.get_status().any()
    stream.close(); a.close(); b.close()


def test_pipelined_host_step_equals_direct_calls(pkg, syn, scen):
    """uwvkf_pose_step_host (one call per step, HOST sample buffers in, navigation solution out, copies overlapped by the
    library) gives bit-identical results to rotation rate + predictionStep + integrateMeasurement, and honours its buffer rules."""
    import torch
    B, STEPS = 3000, 7
    a, filt, mu, cov, off = make_batch(pkg, syn, scen, B)
    b, *_ = make_batch(pkg, syn, scen, B)
    ring_in = [(torch.empty((B, 3), dtype=torch.float64).pin_memory().numpy(), torch.empty((B, 3), dtype=torch.float64).pin_memory().numpy())
               for _ in range(2)]
    ring_nav = [torch.empty((B, 10), dtype=torch.float64).pin_memory().numpy() for _ in range(3)]
    expect = []
    for s in range(1, STEPS + 1):
        inp = syn.pose_inputs(filt, s)
        g, z = ring_in[s % 2]                          # inputs of call i may be overwritten once call i+1 has returned
        g[:] = inp["gyro"]; z[:] = inp["dvl"]
        a.step_host(DT, g, pkg.POSE_VELOCITY, z, syn.R_DVL, nav_out=ring_nav[s % 3])
        b.integrate_rotation_rate(inp["gyro"])
        b.prediction_step(DT)
        b.integrate_measurement(pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
        expect.append(b.get_navigation())
        if s >= 3:                                     # the output of call i is complete once call i+2 has returned
            assert np.array_equal(ring_nav[(s - 2) % 3], expect[s - 3]), f"navigation output of step {s - 2}"
    a.synchronize()
    assert np.array_equal(ring_nav[STEPS % 3], expect[-1]) and np.array_equal(ring_nav[(STEPS - 1) % 3], expect[-2])
    ma, ca = a.get_state()
    mb, cb = b.get_state()
    assert np.array_equal(ma, mb) and np.array_equal(ca, cb)
    # NaN samples are refused per filter on the device (the pipelined call does not scan host memory)
    g, z = ring_in[0]
    z[5, 1] = np.nan
    a.step_host(DT, g, pkg.POSE_VELOCITY, z, syn.R_DVL, nav_out=ring_nav[0])
    a.synchronize()
    st = a.get_status()
    assert st[5] & pkg.ST_NAN_MEASUREMENT and np.count_nonzero(st) == 1
    # the navigation solution comes out of the fused kernel itself: also for the filter whose update was refused
    assert np.array_equal(ring_nav[0], a.get_navigation())
    a.close(); b.close()


def test_two_handles_on_two_streams_are_independent(pkg, syn, scen):
    """Distinct handles are independent (include/uwvkf.h): two batches stepping concurrently on two non-blocking streams give
    the results of running them one after the other."""
    import torch
    B = 20000
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    filt_a, filt_b = np.arange(B), np.arange(B, 2 * B)
    ref = []
    for filt in (filt_a, filt_b):
        u, *_ = make_batch(pkg, syn, scen, B, first_filter=int(filt[0]))
        for s in range(1, 4):
            inp = syn.pose_inputs(filt, s)
            u.predict_update(DT, inp["gyro"], pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
            u.integrate_measurement(pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC)
        ref.append(u.get_state(with_cov=False)); u.close()
    mu_a, cov_a = syn.pose_initial_batch(filt_a, scen["cfg"])
    mu_b, cov_b = syn.pose_initial_batch(filt_b, scen["cfg"])
    a = pkg.PoseUKFBatch(B, scen["par"], stream=sa.cuda_stream); a.set_process_noise(scen["Q"]); a.set_state(mu_a, cov_a)
    b = pkg.PoseUKFBatch(B, scen["par"], stream=sb.cuda_stream); b.set_process_noise(scen["Q"]); b.set_state(mu_b, cov_b)
    dev = {}
    for s in range(1, 4):
        for name, filt in (("a", filt_a), ("b", filt_b)):
            inp = syn.pose_inputs(filt, s)
            dev[name, s] = {k: torch.from_numpy(np.ascontiguousarray(inp[k])).cuda() for k in ("gyro", "dvl", "acc")}
    torch.cuda.synchronize()
    for s in range(1, 4):                                   # interleaved submission, nothing synchronises in between
        for u, name in ((a, "a"), (b, "b")):
            d_ = dev[name, s]
            u.predict_update(DT, d_["gyro"].data_ptr(), pkg.POSE_VELOCITY, d_["dvl"].data_ptr(), syn.R_DVL, memspace=pkg.DEVICE)
            u.integrate_measurement(pkg.POSE_ACCELERATION, d_["acc"].data_ptr(), syn.R_ACC, memspace=pkg.DEVICE)
    a.synchronize(); b.synchronize()
    assert np.array_equal(a.get_state(with_cov=False), ref[0]) and np.array_equal(b.get_state(with_cov=False), ref[1])
    assert not a.get_status().any() and not b.get_status().any()
    a.close(); b.close()


def test_host_step_argument_checks(pkg, syn, scen):
    B = 64
    ukf, filt, mu, cov, off = make_batch(pkg, syn, scen, B)
    inp = syn.pose_inputs(filt, 1)
    g, z = np.ascontiguousarray(inp["gyro"]), np.ascontiguousarray(inp["dvl"])
    ukf.step_host(DT, g, pkg.POSE_VELOCITY, z, syn.R_DVL)                 # no read-back requested
    ukf.synchronize()
    other, *_ = make_batch(pkg, syn, scen, B)
    other.predict_update(DT, g, pkg.POSE_VELOCITY, z, syn.R_DVL)
    assert all(np.array_equal(x, y) for x, y in zip(ukf.get_state(), other.get_state()))
    with pytest.raises(pkg.UwvkfError):                                   # six-dimensional kinds are not part of the pipelined step
        ukf.step_host(DT, g, pkg.POSE_BODY_EFFORTS, np.zeros((B, 6)), np.eye(6))
    with pytest.raises(pkg.UwvkfError):
        ukf.step_host(float("nan"), g, pkg.POSE_VELOCITY, z, syn.R_DVL)
    ukf.close(); other.close()


def test_cuda_graph_replay_of_a_schedule(pkg, syn, scen):
    """A K-step schedule (rotation rate + fused predict/DVL + acceleration per step, inputs resident in HBM) captured once into a
    CUDA graph and replayed == the same calls issued one by one."""
    import torch
    B, K = 32, 6
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    filt = np.arange(B)
    mu, cov = syn.pose_initial_batch(filt, scen["cfg"])
    mk = lambda: pkg.PoseUKFBatch(B, scen["par"], stream=stream.cuda_stream)
    a, b = mk(), mk()
    for u in (a, b):
        u.set_process_noise(scen["Q"]); u.set_state(mu, cov)
    inputs = [syn.pose_inputs(filt, s) for s in range(1, K + 1)]
    gy = torch.from_numpy(np.stack([i["gyro"] for i in inputs])).to(dev)
    dv = torch.from_numpy(np.stack([i["dvl"] for i in inputs])).to(dev)
    ac = torch.from_numpy(np.stack([i["acc"] for i in inputs])).to(dev)
    torch.cuda.synchronize()

    def schedule(u):
        for k in range(K):
            u.predict_update(DT, gy[k].data_ptr(), pkg.POSE_VELOCITY, dv[k].data_ptr(), syn.R_DVL, memspace=pkg.DEVICE)
            u.integrate_measurement(pkg.POSE_ACCELERATION, ac[k].data_ptr(), syn.R_ACC, memspace=pkg.DEVICE)

    schedule(b)                      # reference: call by call (also warms up every kernel of the schedule on this device)
    schedule(b)
    b.synchronize()
    n0 = a.kernel_launches
    a.graph_begin()
    schedule(a)
    g = a.graph_end()
    assert a.kernel_launches == n0, "nothing runs while capturing"
    ma0, _ = a.get_state()
    assert np.array_equal(ma0, mu)
    a.graph_launch(g)
    a.graph_launch(g)
    a.synchronize()
    assert a.kernel_launches == n0 + 2 * 3 * K + 1      # 2 replays x (latitude pre-pass + fused step + update) x K, + the getter above
    ma, ca = a.get_state()
    mb, cb = b.get_state()
    assert np.array_equal(ma, mb) and np.array_equal(ca, cb)
    assert not a.get_status().any()
    a.graph_destroy(g)
    a.close(); b.close()


def test_lockstep_sequence_parity(pkg, syn, scen):
    """100 Hz predict + accelerometer + DVL, pressure and ADCP at 10 Hz; GPU and oracle are re-synchronised every
    step so that the asserted error is the per-step error of north_star."""
    B, STEPS = 8, 40
    ukf, filt, mu, cov, off = make_batch(pkg, syn, scen, B)
    worst = np.zeros(4)
    for s in range(1, STEPS + 1):
        inp = syn.pose_inputs(filt, s)
        w = inp["gyro"]
        mu, cov = ukf.get_state()
        ukf.predict_update(DT, w, pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
        ukf.integrate_measurement(pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC)
        omu, ocov = oracle_predict(mu, cov, scen, w, off)
        omu, ocov, _ = oracle_update(orc.VELOCITY, omu, ocov, scen, inp["dvl"], syn.R_DVL, None, w)
        omu, ocov, _ = oracle_update(orc.ACCELERATION, omu, ocov, scen, inp["acc"], syn.R_ACC, None, w)
        if s % 10 == 0:
            ukf.integrate_measurement(pkg.POSE_PRESSURE, inp["pressure"], syn.R_PRESSURE)
            ukf.integrate_measurement(pkg.POSE_WATER_VELOCITY, inp["adcp"], syn.R_ADCP, aux=np.array([0.5]))
            omu, ocov, _ = oracle_update(orc.PRESSURE, omu, ocov, scen, inp["pressure"], syn.R_PRESSURE, None, w)
            omu, ocov, _ = oracle_update(orc.WATER_VELOCITY, omu, ocov, scen, inp["adcp"], syn.R_ADCP, np.array([0.5]), w)
        gmu, gcov = ukf.get_state()
        worst = np.maximum(worst, assert_parity(gmu, gcov, omu, ocov, f"step {s}"))
    print(f"[parity] lockstep worst over {STEPS} steps: state {worst[0]:.2e} cov {worst[1]:.2e} state(comp) {worst[2]:.2e} "
          f"cov(scaled) {worst[3]:.2e}")
    assert not (ukf.get_status() & ~np.uint32(4)).any()
    ukf.close()


def test_free_running_drift(pkg, syn, scen):
    """GPU and oracle run independently for 300 steps from the same start: drift is reported and loosely bounded."""
    B, STEPS = 4, 300
    ukf, filt, mu, cov, off = make_batch(pkg, syn, scen, B)
    omu, ocov = mu.copy(), cov.copy()
    for s in range(1, STEPS + 1):
        inp = syn.pose_inputs(filt, s)
        w = inp["gyro"]
        ukf.predict_update(DT, w, pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
        omu, ocov = oracle_predict(omu, ocov, scen, w, off)
        omu, ocov, _ = oracle_update(orc.VELOCITY, omu, ocov, scen, inp["dvl"], syn.R_DVL, None, w)
    gmu, gcov = ukf.get_state()
    # free-running over 300 steps: every gate (also the per-component and the scaled-covariance one) loosened 100x
    check_all(gmu, gcov, omu, ocov, f"[drift] {STEPS} free-running steps", scale=100.0)
    ukf.close()


def test_valid_mask_and_per_filter_R(pkg, syn, scen):
    B = 16
    ukf, filt, off = _dense_state(pkg, syn, scen, B)
    mu, cov = ukf.get_state()
    inp = syn.pose_inputs(filt, 4)
    valid = (np.arange(B) % 3 != 0).astype(np.uint8)
    R = np.stack([syn.R_DVL * (1 + f) for f in range(B)])
    ukf.integrate_rotation_rate(inp["gyro"])
    ukf.integrate_measurement(pkg.POSE_VELOCITY, inp["dvl"], R, valid=valid)
    gmu, gcov = ukf.get_state()
    omu, ocov, _ = oracle_update(orc.VELOCITY, mu, cov, scen, inp["dvl"], R, None, inp["gyro"], valid=valid)
    assert np.array_equal(gmu[valid == 0], mu[valid == 0]) and np.array_equal(gcov[valid == 0], cov[valid == 0])
    assert_parity(gmu, gcov, omu, ocov, "masked update with per-filter R")
    ukf.close()


def test_gate_rejects_outlier(pkg, syn, scen):
    B = 8
    ukf, filt, off = _dense_state(pkg, syn, scen, B)
    mu, cov = ukf.get_state()
    inp = syn.pose_inputs(filt, 4)
    z = inp["adcp"].copy()
    z[::2] += 5.0  # gross outliers on the even filters
    ukf.integrate_measurement(pkg.POSE_WATER_VELOCITY, z, syn.R_ADCP, aux=np.array([0.5]))
    gmu, gcov = ukf.get_state()
    st = ukf.get_status()
    assert (st[::2] & pkg.ST_GATE_REJECTED).all() and not (st[1::2] & pkg.ST_GATE_REJECTED).any()
    assert ukf.status_counts() == {"not_spd": 0, "nan_measurement": 0, "gate_rejected": B // 2, "mean_not_converged": 0}
    assert np.array_equal(gmu[::2], mu[::2]) and np.array_equal(gcov[::2], cov[::2])
    assert not np.array_equal(gmu[1::2], mu[1::2])
    ukf.clear_status()
    assert not ukf.get_status().any()
    ukf.close()


def test_nan_measurement_is_refused(pkg, syn, scen):
    B = 4
    ukf, filt, off = _dense_state(pkg, syn, scen, B)
    mu, cov = ukf.get_state()
    z = syn.pose_inputs(filt, 4)["dvl"].copy()
    z[1, 2] = np.nan
    with pytest.raises(pkg.UwvkfError) as ei:   # host check == checkMeasurment() throwing
        ukf.integrate_measurement(pkg.POSE_VELOCITY, z, syn.R_DVL)
    assert ei.value.code == -4
    Rbad = syn.R_DVL.copy(); Rbad[0, 0] = np.nan
    with pytest.raises(pkg.UwvkfError):
        ukf.integrate_measurement(pkg.POSE_VELOCITY, np.zeros((B, 3)), Rbad)
    gmu, gcov = ukf.get_state()
    assert np.array_equal(gmu, mu) and np.array_equal(gcov, cov)
    ukf.close()


def test_nan_checks_for_large_batches_and_device_inputs(pkg, syn, scen):
    """HOST inputs are scanned whatever the batch size (the reference throws, src/PoseUKF.cpp:478,494); DEVICE inputs are
    refused per filter on the device -- including the rotation-rate control, whose NaN would otherwise poison every predict."""
    import torch
    B = 3000                                           # above the old 1,024-filter threshold of the host scan
    ukf, filt, mu, cov, off = make_batch(pkg, syn, scen, B)
    inp = syn.pose_inputs(filt, 1)
    z = inp["dvl"].copy(); z[2222, 1] = np.nan
    with pytest.raises(pkg.UwvkfError) as ei:
        ukf.integrate_measurement(pkg.POSE_VELOCITY, z, syn.R_DVL)
    assert ei.value.code == -4                       # UWVKF_ERR_NAN_MEASUREMENT
    w = inp["gyro"].copy(); w[2999, 0] = np.nan
    with pytest.raises(pkg.UwvkfError):
        ukf.integrate_rotation_rate(w)
    with pytest.raises(pkg.UwvkfError):
        ukf.predict_update(DT, w, pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
    g0, c0 = ukf.get_state()
    assert np.array_equal(g0, mu) and not ukf.get_status().any()
    # the same NaNs in DEVICE memory: filters 2222 / 2999 are refused, every other filter steps normally
    wd, zd = torch.from_numpy(w).cuda(), torch.from_numpy(z).cuda()
    ukf.predict_update(DT, wd.data_ptr(), pkg.POSE_VELOCITY, zd.data_ptr(), syn.R_DVL, memspace=pkg.DEVICE)
    ukf.synchronize()
    st = ukf.get_status()
    assert st[2999] & pkg.ST_NAN_MEASUREMENT and st[2222] & pkg.ST_NAN_MEASUREMENT and np.count_nonzero(st) == 2
    g1, c1 = ukf.get_state()
    assert np.isfinite(g1).all() and np.isfinite(c1).all()
    # filter 2999 (NaN rotation rate): the prediction is refused, its valid DVL sample is still integrated
    m0, cc0, _, _ = orc.pose_update(orc.VELOCITY, mu[2999], cov[2999], z[2999], syn.R_DVL, None, inp["gyro"][2999], scen["par"])
    check_all(g1[2999:3000], c1[2999:3000], m0[None], cc0[None], "NaN rotation rate: update only")
    m1, cc1, _, _ = orc.pose_predict(mu[2222], cov[2222], scen["Q"], w[2222], scen["par"], off[2222], DT)
    check_all(g1[2222:2223], c1[2222:2223], m1[None], cc1[None], "NaN DVL sample: prediction only")
    # the dense engine (separate predict call) guards the control input as well
    ukf.clear_status()
    ukf.prediction_step(DT)
    assert ukf.get_status()[2999] & pkg.ST_NAN_MEASUREMENT
    ukf.close()


def test_list_and_float32_arguments(pkg, syn, scen):
    """Python lists / float32 / non-contiguous arrays are converted to temporaries that must stay alive through the call."""
    B = 8
    ukf, filt, off = _dense_state(pkg, syn, scen, B)
    mu, cov = ukf.get_state()
    inp = syn.pose_inputs(filt, 4)
    z32 = inp["pressure"].astype(np.float32)
    lever = [0.3, -0.1, 0.2]
    ukf.integrate_measurement(pkg.POSE_PRESSURE, z32, syn.R_PRESSURE.tolist(), aux=lever)
    gmu, gcov = ukf.get_state()
    omu, ocov, _ = oracle_update(orc.PRESSURE, mu, cov, scen, z32.astype(np.float64), syn.R_PRESSURE, np.array(lever), inp["gyro"])
    assert_parity(gmu, gcov, omu, ocov, "pressure update from float32 z, list R, list lever arm")
    pose = [[1.0, 2.0, 3.0, 0.0, 0.0, float(np.sin(0.2)), float(np.cos(0.2))]] * B
    ukf.reset_with_external_pose(pose)
    assert np.array_equal(ukf.get_state()[0][:, :7], np.asarray(pose))
    ukf.close()


def test_handle_is_usable_right_after_create_on_a_nonblocking_stream(pkg, syn, scen):
    """uwvkf_create zeroes its buffers on the legacy stream; a non-blocking user stream does not order behind that, so create
    must have finished them (a late memset would wipe the state just written)."""
    import torch
    B = 200000                                         # 2.4 GB of state: the memset takes a while
    s = torch.cuda.Stream()
    mu1, cov1 = syn.pose_initial_batch(np.arange(1), scen["cfg"])
    mus = np.tile(mu1, (B, 1))
    torch.cuda.synchronize()
    ukf = pkg.PoseUKFBatch(B, scen["par"], stream=s.cuda_stream)
    ukf.set_state(mus, np.ascontiguousarray(cov1[0]))       # first work on the user stream: uploads + pack kernels
    ukf.synchronize()
    torch.cuda.synchronize()
    g = ukf.get_state(first=B - 3, count=3, with_cov=False)
    assert np.array_equal(g, np.tile(mu1, (3, 1)))
    assert not ukf.get_status().any()
    ukf.close()


def test_not_spd_freezes_filter(pkg, syn, scen):
    B = 4
    ukf, filt, mu, cov, off = make_batch(pkg, syn, scen, B)
    cov = cov.copy()
    cov[2, 4, 4] = -1.0  # filter 2: negative orientation variance
    ukf.set_state(mu, cov)
    inp = syn.pose_inputs(filt, 1)
    ukf.predict_update(DT, inp["gyro"], pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
    st = ukf.get_status()
    assert st[2] & pkg.ST_NOT_SPD and not (np.delete(st, 2) & pkg.ST_NOT_SPD).any()
    gmu, gcov = ukf.get_state()
    assert np.array_equal(gmu[2], mu[2]) and np.array_equal(gcov[2], cov[2])
    omu, ocov, st_o, _ = orc.pose_predict(mu[2], cov[2], scen["Q"], inp["gyro"][2], scen["par"], off[2], DT)
    assert st_o & 1, "oracle flags the same filter"
    ukf.close()


def test_rotation_rate_getter_and_pose_reset(pkg, syn, scen):
    B = 8
    ukf, filt, off = _dense_state(pkg, syn, scen, B)
    mu, cov = ukf.get_state()
    w = syn.pose_inputs(filt, 4)["gyro"]
    ukf.integrate_rotation_rate(w)
    got = ukf.get_rotation_rate()
    ref = np.stack([orc.pose_rotation_rate(mu[f], w[f], scen["par"]) for f in range(B)])
    assert np.abs(got - ref).max() < 1e-15
    pose = np.tile(np.array([1.0, 2.0, 3.0, 0.0, 0.0, np.sin(0.2), np.cos(0.2)]), (B, 1))
    ukf.reset_with_external_pose(pose)
    mu2, cov2 = ukf.get_state()
    assert np.array_equal(mu2[:, :7], pose) and np.array_equal(mu2[:, 7:], mu[:, 7:]) and np.array_equal(cov2, cov)
    ukf.close()


def test_ensemble_stats(pkg, syn, scen):
    B = 300
    ukf, filt, mu, cov, off = make_batch(pkg, syn, scen, B)
    out = ukf.ensemble_stats()
    assert out[0] == B
    assert np.allclose(out[1:55], mu.sum(0), rtol=1e-12, atol=1e-9)
    M2 = mu.T @ mu
    il = np.tril_indices(54)
    assert np.allclose(out[55:], M2[il], rtol=1e-12, atol=1e-7)
    ukf.close()


def test_mean_policy_is_honoured(pkg, syn, scen):
    """A loose tolerance stops the manifold-mean iteration earlier on both sides and the results still agree."""
    B = 8
    ukf, filt, mu, cov, off = make_batch(pkg, syn, scen, B)
    ukf.set_mean_policy(1e-3, 3)
    w = syn.pose_inputs(filt, 1)["gyro"]
    ukf.integrate_rotation_rate(w)
    ukf.prediction_step(DT)
    gmu, gcov = ukf.get_state()
    pol = np.array([1e-3, 3.0])
    for f in range(B):
        omu, ocov, st, _ = orc.pose_predict(mu[f], cov[f], scen["Q"], w[f], scen["par"], off[f], DT, policy=pol)
        assert state_err(gmu[f], omu) <= TOL_STATE and cov_err(gcov[f], ocov) <= TOL_COV
    ukf.close()


def test_matches_committed_golden_sequence(pkg, syn, scen):
    """CUDA path vs the committed fixture tests/golden/pose_sequence.npz (free-running, no re-synchronisation)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pose_sequence.npz"))
    mus, covs = g["mu"], g["cov"]
    B = mus.shape[1]
    ukf, filt, mu, cov, off = make_batch(pkg, syn, scen, B)
    for s in range(1, int(g["steps"]) + 1):
        inp = syn.pose_inputs(filt, s)
        ukf.predict_update(DT, inp["gyro"], pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
        ukf.integrate_measurement(pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC)
        if s % 4 == 0:
            ukf.integrate_measurement(pkg.POSE_PRESSURE, inp["pressure"], syn.R_PRESSURE)
            ukf.integrate_measurement(pkg.POSE_WATER_VELOCITY, inp["adcp"], syn.R_ADCP, aux=np.array([0.5]))
        gmu, gcov = ukf.get_state()
        assert_parity(gmu, gcov, mus[s], covs[s], f"golden step {s}")
    ukf.close()


def test_update_rolls_back_when_the_downdate_loses_definiteness(pkg, syn, scen):
    """A (numerically accepted, physically invalid) negative R makes K S K^T overshoot: the downdated position block is not
    positive definite, apply_delta cannot draw sigma points, and the engine must leave the filter as it was (status bit set), alone
    and after a predict in the same launch.  The oracle flags the same filters."""
    B = 12
    ukf, filt, off = _dense_state(pkg, syn, scen, B)
    ref, _, _ = _dense_state(pkg, syn, scen, B)
    mu, cov = ukf.get_state()
    bad = np.zeros(B, dtype=bool); bad[[0, 5, 11]] = True
    R = np.stack([(-0.9 * cov[f][0:2, 0:2]) if bad[f] else 0.25 * np.eye(2) for f in range(B)])
    z = mu[:, 0:2] + 0.05
    w = syn.pose_inputs(filt, 4)["gyro"]
    ukf.integrate_measurement(pkg.POSE_XY, z, R)
    gmu, gcov = ukf.get_state()
    st = ukf.get_status()
    assert (st[bad] & pkg.ST_NOT_SPD).all() and not (st[~bad] & pkg.ST_NOT_SPD).any()
    assert np.array_equal(gmu[bad], mu[bad])
    for f in np.nonzero(bad)[0]:
        assert cov_err(gcov[f], cov[f]) < 1e-13                      # panel restored by the inverse rank-2 update
        assert np.array_equal(gcov[f][15:, 15:], cov[f][15:, 15:])   # trailing triangle never touched
    omu, ocov, sts = oracle_update(orc.XY, mu, cov, scen, z, R, None, w)
    assert (sts[bad] & 1).all() and not (sts[~bad] & 1).any()
    assert_parity(gmu[~bad], gcov[~bad], omu[~bad], ocov[~bad], "XY update next to rolled-back filters")
    # the same inside a fused launch: predict + refused update == predict alone
    ukf.set_state(mu, cov); ref.set_state(mu, cov)
    ukf.clear_status()
    ukf.predict_update(DT, w, pkg.POSE_XY, z, R)
    ref.integrate_rotation_rate(w); ref.prediction_step(DT)
    gmu, gcov = ukf.get_state()
    pmu, pcov = ref.get_state()
    assert (ukf.get_status()[bad] & pkg.ST_NOT_SPD).all()
    assert np.array_equal(gmu[bad], pmu[bad])
    for f in np.nonzero(bad)[0]:
        assert cov_err(gcov[f], pcov[f]) < 1e-13
    ukf.close(); ref.close()


@pytest.mark.parametrize("B", [1, 13, 1776 + 29])
def test_ragged_batch_sizes(pkg, syn, scen, B):
    """Batches that do not fill the persistent grid (12 warps x #SMs) evenly: first / last filters against the oracle, and the
    fused step against separate launches for every filter."""
    a, filt, mu, cov, off = make_batch(pkg, syn, scen, B)
    b, *_ = make_batch(pkg, syn, scen, B)
    for s in range(1, 3):
        inp = syn.pose_inputs(filt, s)
        a.predict_update(DT, inp["gyro"], pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
        b.integrate_rotation_rate(inp["gyro"]); b.prediction_step(DT)
        b.integrate_measurement(pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
    ma, ca = a.get_state()
    mb, cb = b.get_state()
    assert np.array_equal(ma, mb) and np.array_equal(ca, cb)
    assert not a.get_status().any()
    pick = sorted(set([0, B - 1, B // 2]))
    inp = syn.pose_inputs(filt, 1)
    omu, ocov = oracle_predict(mu[pick], cov[pick], scen, inp["gyro"][pick], off[pick])
    omu, ocov, _ = oracle_update(orc.VELOCITY, omu, ocov, scen, inp["dvl"][pick], syn.R_DVL, None, inp["gyro"][pick])
    inp2 = syn.pose_inputs(filt, 2)
    omu, ocov = oracle_predict(omu, ocov, scen, inp2["gyro"][pick], off[pick])
    omu, ocov, _ = oracle_update(orc.VELOCITY, omu, ocov, scen, inp2["dvl"][pick], syn.R_DVL, None, inp2["gyro"][pick])
    assert_parity(ma[pick], ca[pick], omu, ocov, f"two fused steps, batch of {B}")
    a.close(); b.close()

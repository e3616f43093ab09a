"""GPU parity tests of VelocityUKF, BottomUKF and IndirectPoseUKF (BASELINE.json configs 1, 2, 5) vs the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as orc
from parity import check_all

pytestmark = pytest.mark.gpu


def check(gmu, gcov, omu, ocov, what):
    check_all(gmu, gcov, omu, ocov, what)


def test_velocity_ukf_sequence(pkg, syn):
    """config 1/2: 100 Hz thrust + gyro predict, DVL and pressure at 10 Hz; lockstep per-step parity."""
    B, STEPS, DT = 32, 60, 0.01
    filt = np.arange(B)
    mu, cov = syn.velocity_initial_batch(filt)
    uwv = syn.uwv_params_flat()
    ukf = pkg.VelocityUKFBatch(B)
    ukf.set_state(mu, cov)
    with pytest.raises(pkg.UwvkfError) as ei:       # reference: throw "Motion model is not initialized!"
        ukf.prediction_step(DT)
    assert ei.value.code == -5 and "Motion model is not initialized" in str(ei.value)
    ukf.integrate_gyro(syn.velocity_inputs(filt, 0)["gyro"])
    ukf.setup_motion_model(uwv)
    sim = np.stack([orc.vel_setup_motion_model(mu[f], syn.velocity_inputs(filt, 0)["gyro"][f]) for f in range(B)])
    assert np.array_equal(ukf.get_motion_model_state(), sim)
    Q = np.diag([1e-4, 1e-4, 1e-4, 0.0])           # constructor default (src/VelocityUKF.cpp:53-55)
    R_dvl, R_p = 1e-4 * np.eye(3), np.array([[2.5e-3]])
    for s in range(1, STEPS + 1):
        inp = syn.velocity_inputs(filt, s)
        mu, cov = ukf.get_state()
        sim = ukf.get_motion_model_state()
        ukf.integrate_gyro(inp["gyro"])
        ukf.integrate_body_efforts(inp["efforts"])
        ukf.prediction_step(DT)
        omu, ocov, osim = mu.copy(), cov.copy(), sim.copy()
        for f in range(B):
            sf = sim[f].copy()
            sf[7:10] = inp["gyro"][f]                  # integrateMeasurement(Gyro) also writes the simulator
            omu[f], ocov[f], osim[f], st = orc.vel_predict(mu[f], cov[f], Q, sf, inp["gyro"][f], inp["efforts"][f], uwv, DT)
            assert st == 0
        if s % 10 == 0:
            ukf.integrate_measurement(ukf.DVL, inp["dvl"], R_dvl)
            ukf.integrate_measurement(ukf.PRESSURE, inp["pressure"], R_p)
            for f in range(B):
                omu[f], ocov[f], _ = orc.vel_update(0, omu[f], ocov[f], inp["dvl"][f], R_dvl)
                omu[f], ocov[f], _ = orc.vel_update(1, omu[f], ocov[f], inp["pressure"][f], R_p)
        gmu, gcov = ukf.get_state()
        check(gmu, gcov, omu, ocov, f"velocity step {s}")
        assert np.abs(ukf.get_motion_model_state() - osim).max() < 1e-12
    assert not ukf.get_status().any()
    ukf.close()


def test_velocity_nan_and_mask(pkg, syn):
    B = 8
    filt = np.arange(B)
    mu, cov = syn.velocity_initial_batch(filt)
    ukf = pkg.VelocityUKFBatch(B)
    ukf.set_state(mu, cov)
    z = syn.velocity_inputs(filt, 1)["dvl"].copy()
    z[3, 0] = np.nan
    with pytest.raises(pkg.UwvkfError):
        ukf.integrate_measurement(ukf.DVL, z, 1e-4 * np.eye(3))
    z[3, 0] = 0.0
    valid = (np.arange(B) % 2).astype(np.uint8)
    ukf.integrate_measurement(ukf.DVL, z, 1e-4 * np.eye(3), valid=valid)
    gmu, gcov = ukf.get_state()
    assert np.array_equal(gmu[::2], mu[::2]) and not np.array_equal(gmu[1::2], mu[1::2])
    ukf.close()


def test_bottom_ukf_sequence(pkg, syn):
    """config 5 (BottomUKF half): velocity control + 4 beam ranges at 10 Hz, optional normal measurement."""
    B, STEPS, DT = 32, 25, 0.1
    filt = np.arange(B)
    mu, cov = syn.bottom_initial_batch(filt)
    ukf = pkg.BottomUKFBatch(B)
    ukf.set_state(mu, cov)
    Q, R, Rn = np.eye(3), np.array([4e-4]), 1e-4 * np.eye(2)
    for s in range(1, STEPS + 1):
        inp = syn.bottom_inputs(filt, s)
        mu, cov = ukf.get_state()
        ukf.set_velocity(inp["velocity"])
        ukf.prediction_step(DT)
        omu, ocov = mu.copy(), cov.copy()
        for f in range(B):
            omu[f], ocov[f], st = orc.bottom_predict(mu[f], cov[f], Q, inp["velocity"][f], DT)
            assert st == 0
        for k in range(4):
            ukf.integrate_range(inp["ranges"][:, k].copy(), R, syn.BEAMS[k], syn.BEAM_ORIGIN)
            for f in range(B):
                omu[f], ocov[f], st = orc.bottom_update_range(omu[f], ocov[f], inp["ranges"][f, k:k + 1], R, syn.BEAMS[k],
                                                              syn.BEAM_ORIGIN)
                assert st == 0
        if s % 5 == 0:
            ukf.integrate_normal(inp["normal"], Rn)
            for f in range(B):
                omu[f], ocov[f], st = orc.bottom_update_normal(omu[f], ocov[f], inp["normal"][f], Rn)
        gmu, gcov = ukf.get_state()
        check(gmu, gcov, omu, ocov, f"bottom step {s}")
        assert np.abs(np.linalg.norm(gmu[:, 1:4], axis=1) - 1).max() < 1e-12
    assert not ukf.get_status().any()
    ukf.close()


def test_indirect_pose_ukf_sequence(pkg, syn):
    """config 5 (IndirectPoseUKF half): pose reference at 100 Hz, four marker corners every 50 steps."""
    B, STEPS, DT, TAU = 16, 120, 0.01, 100.0
    filt = np.arange(B)
    pes, oes = np.array([0.01, 0.01, 0.01]), np.array([0.002, 0.002, 0.005])
    ukf = pkg.IndirectPoseUKFBatch(B, pes, oes, TAU, initial_position_error=(0.0, 0.0, 0.0),
                                   initial_position_error_std=(0.5, 0.5, 0.5))
    Q = np.diag(np.concatenate([pes ** 2, oes ** 2]))
    pxcov = np.tile((np.eye(2) * 0.25).ravel(), (4, 1))
    covm = np.diag([1e-4] * 3 + [1e-6] * 3)
    mu0, cov0 = ukf.get_state()
    assert np.array_equal(mu0[0], [0, 0, 0, 0, 0, 0, 1.0]) and np.allclose(np.diag(cov0[0]), [0.25] * 3 + list(oes ** 2))
    for s in range(1, STEPS + 1):
        inp = syn.indirect_inputs(filt, s)
        mu, cov = ukf.get_state()
        ukf.update_pose_reference(inp["pose_ref"])
        ukf.prediction_step(DT)
        omu, ocov = mu.copy(), cov.copy()
        for f in range(B):
            omu[f], ocov[f], st = orc.indirect_predict(mu[f], cov[f], Q, TAU, DT)
            assert st == 0
        if s % 50 == 0:
            ukf.integrate_features(inp["px"], pxcov, syn.MARKER_FEATURES, syn.MARKER_POSE, covm, syn.CAMERA, syn.CAM_IN_BODY)
            for f in range(B):
                omu[f], ocov[f], st = orc.indirect_update_features(omu[f], ocov[f], inp["px"][f], pxcov, syn.MARKER_FEATURES,
                                                                   syn.MARKER_POSE, covm, syn.CAMERA, syn.CAM_IN_BODY,
                                                                   inp["pose_ref"][f])
                assert st == 0
        gmu, gcov = ukf.get_state()
        check(gmu, gcov, omu, ocov, f"indirect step {s}")
    corrected = ukf.get_corrected_pose()
    ref = np.stack([orc.indirect_corrected_pose(gmu[f], inp["pose_ref"][f]) for f in range(B)])
    assert np.abs(corrected - ref).max() < 1e-14
    # the marker updates must have pulled the position error towards the simulated offset (0.15, -0.1, 0.05)
    assert np.abs(gmu[:, 0:3].mean(0) - np.array([0.15, -0.1, 0.05])).max() < 0.05
    assert not ukf.get_status().any()
    ukf.close()


def test_small_filter_ensemble_stats(pkg, syn):
    B = 500
    mu, cov = syn.velocity_initial_batch(np.arange(B))
    ukf = pkg.VelocityUKFBatch(B)
    ukf.set_state(mu, cov[0])
    gmu, gcov = ukf.get_state()
    assert np.array_equal(gmu, mu) and np.array_equal(gcov[7], cov[0])
    out = ukf.ensemble_stats()
    assert out[0] == B and np.allclose(out[1:5], mu.sum(0)) and np.allclose(out[5:], (mu.T @ mu)[np.tril_indices(4)])
    ukf.close()


def test_bottom_matches_committed_golden_sequence(pkg, syn):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "small_sequences.npz"))["bottom"]
    B = 3
    filt = np.arange(B)
    mu, cov = syn.bottom_initial_batch(filt)
    ukf = pkg.BottomUKFBatch(B)
    ukf.set_state(mu, cov)
    for s in range(1, 11):
        inp = syn.bottom_inputs(filt, s)
        ukf.set_velocity(inp["velocity"])
        ukf.prediction_step(0.1)
        for k in range(4):
            ukf.integrate_range(inp["ranges"][:, k].copy(), np.array([4e-4]), syn.BEAMS[k], syn.BEAM_ORIGIN)
        if s % 5 == 0:
            ukf.integrate_normal(inp["normal"], 1e-4 * np.eye(2))
        gmu, gcov = ukf.get_state()
        check(gmu, gcov, g[s][:, :4], g[s][:, 4:].reshape(B, 3, 3), f"bottom golden step {s}")
    ukf.close()


def test_list_arguments_do_not_alias(pkg, syn):
    """integrate_range(z, R, [0, 0, 1], [0, 0, 0]): the two converted list temporaries used to share one freed buffer, so the
    beam direction silently became the origin (advisor finding).  Lists, tuples and float32 arrays must behave like float64 arrays."""
    B = 16
    filt = np.arange(B)
    mu, cov = syn.bottom_initial_batch(filt)
    a, b = pkg.BottomUKFBatch(B), pkg.BottomUKFBatch(B)
    a.set_state(mu, cov); b.set_state(mu, cov)
    z = syn.bottom_inputs(filt, 1)["ranges"][:, 0].copy()
    direction, origin = [0.0, 0.0, -1.0], [0.1, -0.2, 0.3]
    a.integrate_range(z.astype(np.float32), 4e-4, direction, tuple(origin))
    b.integrate_range(z.astype(np.float32).astype(np.float64), np.array([4e-4]), np.array(direction), np.array(origin))
    ma, ca = a.get_state()
    mb, cb = b.get_state()
    assert np.array_equal(ma, mb) and np.array_equal(ca, cb)
    assert not np.array_equal(ma, mu)
    om, oc, st = orc.bottom_update_range(mu[0], cov[0], z[:1].astype(np.float32).astype(np.float64), np.array([4e-4]), np.array(direction),
                                         np.array(origin))
    check(ma[:1], ca[:1], om[None], oc[None], "range update from list arguments")
    a.close(); b.close()

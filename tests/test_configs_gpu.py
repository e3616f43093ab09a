"""BASELINE.json configs 2-5 at their full per-GPU batch sizes.

The oracle cannot run 10^5..10^6 filters in a test, so each config runs the whole batch on the GPU and checks (i) a SAMPLE
of filters spread over the batch against the oracle, free-running over the whole sequence (inputs are a hash of the global
filter index, so a filter's trajectory does not depend on the batch it runs in), and (ii) size-independent properties of
the whole batch: no fault bits, symmetric positive-definite covariances on a second sample, unit quaternions / unit normals,
shard invariance (a filter computes the same bits whatever slice of the batch it sits in) and the ensemble statistics
against a host sum.
"""
import numpy as np
import pytest

import oracle_lib as orc
from parity import all_errs, cov_err, state_err

pytestmark = pytest.mark.gpu

# free-running comparisons over tens of steps: the north_star per-step tolerances still hold for (state, cov); the
# per-component and scaled-covariance gates (tests/parity.py) get 10x for the accumulated steps
FREE_TOL = (1e-9, 1e-8, 1e-8, 1e-7)


def _gate(errs, what):
    print(f"{what}: state {errs[0]:.2e} cov {errs[1]:.2e} state(comp) {errs[2]:.2e} cov(scaled) {errs[3]:.2e}")
    assert all(e < t for e, t in zip(errs, FREE_TOL)), (what, tuple(errs))

DT = 0.01


def _sample(B, n=6):
    return np.unique(np.concatenate([[0, 1, B - 1], np.linspace(0, B - 1, n).astype(np.int64)]))


def _pose_batch(pkg, syn, B, first=0):
    cfg = syn.pose_config()
    par, Q = syn.pose_params(cfg), syn.pose_process_noise(cfg, DT)
    ukf = pkg.PoseUKFBatch(B, par)
    ukf.set_process_noise(Q)
    _, cov0 = syn.pose_initial(cfg)
    CH = 1 << 16
    for o in range(0, B, CH):
        c = min(CH, B - o)
        mu_c, _ = syn.pose_initial_batch(np.arange(first + o, first + o + c), cfg, with_cov=False)
        ukf.set_state(mu_c, cov0, first=o)
    return ukf, cfg, par, Q, cov0


def _pose_inputs(syn, first, B, step):
    CH = 1 << 16
    parts = [syn.pose_inputs(np.arange(first + o, first + min(o + CH, B)), step) for o in range(0, B, CH)]
    return {k: np.ascontiguousarray(np.concatenate([p[k] for p in parts])) for k in parts[0]}


def test_config2_velocity_ukf_65536(pkg, syn):
    """config 2: VelocityUKF batch of 65,536 independent filters, synthetic force/torque + DVL."""
    B, STEPS = 65536, 30
    filt = np.arange(B)
    mu, cov = syn.velocity_initial_batch(filt)
    uwv = syn.uwv_params_flat()
    ukf = pkg.VelocityUKFBatch(B)
    ukf.set_state(mu, cov)
    g0 = syn.velocity_inputs(filt, 0)["gyro"]
    ukf.integrate_gyro(g0)
    ukf.setup_motion_model(uwv)
    Q, R_dvl, R_p = np.diag([1e-4, 1e-4, 1e-4, 0.0]), 1e-4 * np.eye(3), np.array([[2.5e-3]])
    smp = _sample(B)
    omu, ocov = mu[smp].copy(), cov[smp].copy()
    osim = np.stack([orc.vel_setup_motion_model(mu[f], g0[f]) for f in smp])
    for s in range(1, STEPS + 1):
        inp = syn.velocity_inputs(filt, s)
        ukf.integrate_gyro(inp["gyro"])
        ukf.integrate_body_efforts(inp["efforts"])
        ukf.prediction_step(DT)
        if s % 10 == 0:
            ukf.integrate_measurement(ukf.DVL, inp["dvl"], R_dvl)
            ukf.integrate_measurement(ukf.PRESSURE, inp["pressure"], R_p)
        for k, f in enumerate(smp):
            osim[k][7:10] = inp["gyro"][f]
            omu[k], ocov[k], osim[k], st = orc.vel_predict(omu[k], ocov[k], Q, osim[k], inp["gyro"][f], inp["efforts"][f], uwv, DT)
            assert st == 0
            if s % 10 == 0:
                omu[k], ocov[k], _ = orc.vel_update(0, omu[k], ocov[k], inp["dvl"][f], R_dvl)
                omu[k], ocov[k], _ = orc.vel_update(1, omu[k], ocov[k], inp["pressure"][f], R_p)
    gmu, gcov = ukf.get_state()
    assert not ukf.get_status().any()
    errs = np.max([all_errs(gmu[f], gcov[f], omu[k], ocov[k]) for k, f in enumerate(smp)], axis=0)
    _gate(errs, f"[config 2] {B} filters x {STEPS} steps, free-running sample")
    assert np.isfinite(gmu).all() and np.allclose(gcov, np.swapaxes(gcov, 1, 2), rtol=0, atol=1e-15)
    assert (np.linalg.eigvalsh(gcov[:: B // 64]) > 0).all()
    stats = ukf.ensemble_stats()
    assert stats[0] == B and np.allclose(stats[1:5], gmu.sum(0), rtol=1e-12)
    ukf.close()


def test_config3_pose_ukf_full_state_262144(pkg, syn):
    """config 3: PoseUKF full state (rotation rate, acceleration, DVL at 100 Hz, pressure at 10 Hz), batch 262,144."""
    B, STEPS = 262144, 20
    ukf, cfg, par, Q, cov0 = _pose_batch(pkg, syn, B)
    smp = _sample(B)
    omu, _ = syn.pose_initial_batch(smp, cfg, with_cov=False)
    ocov = np.tile(cov0, (len(smp), 1, 1))
    off = np.concatenate([omu[:, 20:47], omu[:, 53:54]], axis=1)
    for s in range(1, STEPS + 1):
        inp = _pose_inputs(syn, 0, B, s)
        ukf.predict_update(DT, inp["gyro"], pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL)
        ukf.integrate_measurement(pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC)
        if s % 10 == 0:
            ukf.integrate_measurement(pkg.POSE_PRESSURE, inp["pressure"], syn.R_PRESSURE, aux=np.array([0.3, -0.1, 0.2]))
        for k, f in enumerate(smp):
            w = inp["gyro"][f]
            omu[k], ocov[k], st, _ = orc.pose_predict(omu[k], ocov[k], Q, w, par, off[k], DT)
            omu[k], ocov[k], st, _ = orc.pose_update(orc.VELOCITY, omu[k], ocov[k], inp["dvl"][f], syn.R_DVL, None, w, par)
            omu[k], ocov[k], st, _ = orc.pose_update(orc.ACCELERATION, omu[k], ocov[k], inp["acc"][f], syn.R_ACC, None, w, par)
            if s % 10 == 0:
                omu[k], ocov[k], st, _ = orc.pose_update(orc.PRESSURE, omu[k], ocov[k], inp["pressure"][f], syn.R_PRESSURE,
                                                         np.array([0.3, -0.1, 0.2]), w, par)
            assert st == 0
    assert not ukf.get_status().any()
    errs = np.zeros(4)
    for k, f in enumerate(smp):
        gm, gc = ukf.get_state(first=int(f), count=1)
        errs = np.maximum(errs, all_errs(gm[0], gc[0], omu[k], ocov[k]))
    _gate(errs, f"[config 3] {B} filters x {STEPS} steps, free-running sample")
    nav = ukf.get_navigation()
    assert np.isfinite(nav).all() and np.abs(np.linalg.norm(nav[:, 3:7], axis=1) - 1).max() < 1e-12
    for f in np.linspace(0, B - 1, 16).astype(np.int64):
        _, gc = ukf.get_state(first=int(f), count=1)
        assert np.array_equal(gc[0], gc[0].T) and np.linalg.eigvalsh(gc[0]).min() > 0
    stats = ukf.ensemble_stats()
    mu_all = ukf.get_state(with_cov=False)
    assert stats[0] == B and np.allclose(stats[1:55], mu_all.sum(0), rtol=1e-11, atol=1e-6)
    ukf.close()


def test_headline_1m_filters_host_step(pkg, syn):
    """The headline workload at its full size (BASELINE metric: 1,048,576 PoseUKF filters on one GPU, fused predictionStep +
    integrateMeasurement(Velocity)) through the host-buffer entry point `bench.py` times as e2e (uwvkf_pose_step_host): a sample
    of filters spread over the batch against the free-running oracle, and size-independent properties of the whole batch."""
    import torch
    B, STEPS = 1 << 20, 6
    ukf, cfg, par, Q, cov0 = _pose_batch(pkg, syn, B)
    smp = _sample(B, 5)
    omu, _ = syn.pose_initial_batch(smp, cfg, with_cov=False)
    ocov = np.tile(cov0, (len(smp), 1, 1))
    off = np.concatenate([omu[:, 20:47], omu[:, 53:54]], axis=1)
    pin = lambda shape: torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
    ring = [(pin((B, 3)), pin((B, 3))) for _ in range(2)]
    navs = [pin((B, 10)) for _ in range(3)]
    for s in range(1, STEPS + 1):
        inp = _pose_inputs(syn, 0, B, s)
        g, z = ring[s % 2]
        g[:] = inp["gyro"]; z[:] = inp["dvl"]
        ukf.step_host(DT, g, pkg.POSE_VELOCITY, z, syn.R_DVL, nav_out=navs[s % 3])
        for k, f in enumerate(smp):
            w = inp["gyro"][f]
            omu[k], ocov[k], st, _ = orc.pose_predict(omu[k], ocov[k], Q, w, par, off[k], DT)
            omu[k], ocov[k], st, _ = orc.pose_update(orc.VELOCITY, omu[k], ocov[k], inp["dvl"][f], syn.R_DVL, None, w, par)
            assert st == 0
    ukf.synchronize()
    assert not ukf.get_status().any()
    errs = np.zeros(4)
    for k, f in enumerate(smp):
        gm, gc = ukf.get_state(first=int(f), count=1)
        errs = np.maximum(errs, all_errs(gm[0], gc[0], omu[k], ocov[k]))
        assert np.array_equal(gc[0], gc[0].T) and np.linalg.eigvalsh(gc[0]).min() > 0
    _gate(errs, f"[headline] {B} filters x {STEPS} fused host steps, free-running sample")
    nav = navs[STEPS % 3]                                 # written by the fused kernel itself, downloaded on the copy stream
    assert np.array_equal(nav, ukf.get_navigation())
    assert np.isfinite(nav).all() and np.abs(np.linalg.norm(nav[:, 3:7], axis=1) - 1).max() < 1e-12
    stats = ukf.ensemble_stats()
    assert stats[0] == B and np.allclose(stats[1:11], nav.sum(0), rtol=1e-11, atol=1e-6)
    ukf.close()


def test_config4_adcp_dropouts_monte_carlo_shard(pkg, syn):
    """config 4: PoseUKF with ADCP aiding and DVL dropouts, Monte-Carlo sweep (per-filter DVL noise multipliers); one
    131,072-filter shard of the 1M-filter job (= what each of 8 GPUs owns), and shard invariance against a second,
    differently placed handle."""
    SHARD, STEPS = 131072, 12
    first = 5 * SHARD                                   # rank 5 of 8
    ukf, cfg, par, Q, cov0 = _pose_batch(pkg, syn, SHARD, first=first)
    gidx = np.arange(first, first + SHARD)
    mult = 1.0 + (gidx % 7) * 0.5                       # per-filter noise multiplier of the sweep
    Rpf = np.ascontiguousarray(mult[:, None, None] ** 2 * syn.R_DVL[None])
    smp = np.unique(np.concatenate([_sample(SHARD), np.flatnonzero(gidx % 97 == 0)[:2]]))   # include gated outliers
    omu, _ = syn.pose_initial_batch(gidx[smp], cfg, with_cov=False)
    ocov = np.tile(cov0, (len(smp), 1, 1))
    off = np.concatenate([omu[:, 20:47], omu[:, 53:54]], axis=1)
    # a small second handle holding 64 of the same global filters at other local positions
    sub = gidx[1000:1064]
    ukf2 = pkg.PoseUKFBatch(64, par)
    ukf2.set_process_noise(Q)
    mu2, _ = syn.pose_initial_batch(sub, cfg, with_cov=False)
    ukf2.set_state(mu2, cov0)
    rejected = 0
    for s in range(1, STEPS + 1):
        inp = _pose_inputs(syn, first, SHARD, s)
        valid = ((gidx + s) % 5 != 0).astype(np.uint8)  # DVL dropouts: a fifth of the filters each step
        ukf.predict_update(DT, inp["gyro"], pkg.POSE_VELOCITY, inp["dvl"], Rpf, valid=valid)
        ukf2.predict_update(DT, inp["gyro"][1000:1064], pkg.POSE_VELOCITY, inp["dvl"][1000:1064], Rpf[1000:1064], valid=valid[1000:1064])
        adcp = s % 4 == 0
        if adcp:
            z = inp["adcp"].copy()
            z[gidx % 97 == 0] += 3.0                    # outliers the d2p95 gate must refuse
            ukf.integrate_measurement(pkg.POSE_WATER_VELOCITY, z, syn.R_ADCP, aux=np.array([0.5]))
            ukf2.integrate_measurement(pkg.POSE_WATER_VELOCITY, z[1000:1064], syn.R_ADCP, aux=np.array([0.5]))
        for k, j in enumerate(smp):
            w = inp["gyro"][j]
            omu[k], ocov[k], st, _ = orc.pose_predict(omu[k], ocov[k], Q, w, par, off[k], DT)
            if valid[j]:
                omu[k], ocov[k], st, _ = orc.pose_update(orc.VELOCITY, omu[k], ocov[k], inp["dvl"][j], Rpf[j], None, w, par)
            if adcp:
                omu[k], ocov[k], st, _ = orc.pose_update(orc.WATER_VELOCITY, omu[k], ocov[k], z[j], syn.R_ADCP, np.array([0.5]), w, par)
                rejected += int(st & 4 != 0)
    st = ukf.get_status()
    assert not (st & ~np.uint32(4)).any()
    assert (st[gidx % 97 == 0] & 4).all(), "every outlier must have been gated"
    errs = np.zeros(4)
    for k, j in enumerate(smp):
        gm, gc = ukf.get_state(first=int(j), count=1)
        errs = np.maximum(errs, all_errs(gm[0], gc[0], omu[k], ocov[k]))
    _gate(errs, f"[config 4] shard of {SHARD} filters x {STEPS} steps (oracle gate rejections in the sample: {rejected}), "
                "free-running sample")
    assert rejected >= 2
    a_mu, a_cov = ukf.get_state(first=1000, count=64)
    b_mu, b_cov = ukf2.get_state()
    assert np.array_equal(a_mu, b_mu) and np.array_equal(a_cov, b_cov), "a filter's result must not depend on its shard"
    ukf.close(); ukf2.close()


def test_config5_indirect_and_bottom_mixed_rate(pkg, syn):
    """config 5: IndirectPoseUKF + BottomUKF mixed-rate stream (one 131,072-filter shard each) with the ensemble statistics
    that the ranks all-reduce every 100 steps (checked here against a host sum; the 2-rank reduction is a gloo CPU test)."""
    B, STEPS, TAU = 131072, 100, 100.0
    filt = np.arange(B)
    pes, oes = np.array([0.01, 0.01, 0.01]), np.array([0.002, 0.002, 0.005])
    ind = pkg.IndirectPoseUKFBatch(B, pes, oes, TAU, initial_position_error_std=(0.5, 0.5, 0.5))
    bot = pkg.BottomUKFBatch(B)
    bmu, bcov = syn.bottom_initial_batch(filt)
    bot.set_state(bmu, bcov)
    Qi = np.diag(np.concatenate([pes ** 2, oes ** 2]))
    Qb, Rr = np.eye(3), np.array([4e-4])
    pxcov = np.tile((np.eye(2) * 0.25).ravel(), (4, 1))
    covm = np.diag([1e-4] * 3 + [1e-6] * 3)
    smp = _sample(B, 4)
    imu_o, icov_o = ind.get_state(first=0, count=1)
    o_im = np.tile(imu_o[0], (len(smp), 1)); o_ic = np.tile(icov_o[0], (len(smp), 1, 1))
    o_bm, o_bc = bmu[smp].copy(), bcov[smp].copy()
    for s in range(1, STEPS + 1):
        ii = syn.indirect_inputs(filt, s)
        ind.update_pose_reference(ii["pose_ref"])
        ind.prediction_step(DT)
        for k, f in enumerate(smp):
            o_im[k], o_ic[k], st = orc.indirect_predict(o_im[k], o_ic[k], Qi, TAU, DT)
        if s % 50 == 0:                                  # marker corners at 2 Hz
            ind.integrate_features(ii["px"], pxcov, syn.MARKER_FEATURES, syn.MARKER_POSE, covm, syn.CAMERA, syn.CAM_IN_BODY)
            for k, f in enumerate(smp):
                o_im[k], o_ic[k], st = orc.indirect_update_features(o_im[k], o_ic[k], ii["px"][f], pxcov, syn.MARKER_FEATURES, syn.MARKER_POSE,
                                                                    covm, syn.CAMERA, syn.CAM_IN_BODY, ii["pose_ref"][f])
                assert st == 0
        if s % 10 == 0:                                  # DVL beams at 10 Hz
            bi = syn.bottom_inputs(filt, s // 10)
            bot.set_velocity(bi["velocity"])
            bot.prediction_step(0.1)
            for k, f in enumerate(smp):
                o_bm[k], o_bc[k], st = orc.bottom_predict(o_bm[k], o_bc[k], Qb, bi["velocity"][f], 0.1)
            for b in range(4):
                bot.integrate_range(bi["ranges"][:, b].copy(), Rr, syn.BEAMS[b], syn.BEAM_ORIGIN)
                for k, f in enumerate(smp):
                    o_bm[k], o_bc[k], st = orc.bottom_update_range(o_bm[k], o_bc[k], bi["ranges"][f, b:b + 1], Rr, syn.BEAMS[b], syn.BEAM_ORIGIN)
    assert not ind.get_status().any() and not bot.get_status().any()
    gim, gic = ind.get_state()
    gbm, gbc = bot.get_state()
    ei = np.max([all_errs(gim[f], gic[f], o_im[k], o_ic[k]) for k, f in enumerate(smp)], axis=0)
    eb = np.max([all_errs(gbm[f], gbc[f], o_bm[k], o_bc[k]) for k, f in enumerate(smp)], axis=0)
    print(f"[config 5] {B} + {B} filters x {STEPS} steps, free-running samples (state, cov, state(comp), cov(scaled)): "
          f"indirect {ei} bottom {eb}")
    assert (ei < 1e-7).all() and (eb < 1e-7).all()      # 100 free-running steps with 2 Hz marker updates
    assert np.abs(np.linalg.norm(gim[:, 3:7], axis=1) - 1).max() < 1e-12 and np.abs(np.linalg.norm(gbm[:, 1:4], axis=1) - 1).max() < 1e-12
    for ukf, gm in ((ind, gim), (bot, gbm)):
        st = ukf.ensemble_stats()
        d = gm.shape[1]
        assert st[0] == B and np.allclose(st[1:1 + d], gm.sum(0), rtol=1e-11, atol=1e-9)
        tri = np.array([gm[:, i] @ gm[:, j] for i in range(d) for j in range(i + 1)])
        assert np.allclose(st[1 + d:], tri, rtol=1e-10, atol=1e-9)
    ind.close(); bot.close()

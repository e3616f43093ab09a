"""Generates the golden fixtures under tests/golden/ from the CPU oracle (run in the dev container):

    python tests/golden/make_golden.py

The reference ships no golden vectors (test/suite.cpp:1-6 is an empty stub) and cannot be built here (its
dependencies are absent), so these fixtures freeze the ORACLE's outputs on the SYNTH-1 inputs: they guard the
oracle against silent drift and give the GPU tests a second, file-based target.
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as orc  # noqa: E402

syn = importlib.import_module("slam-uwv_kalman_filters_b200.synthetic")
DT = 0.01


def pose_sequence(B=2, steps=12):
    cfg = syn.pose_config()
    par, Q = syn.pose_params(cfg), syn.pose_process_noise(cfg, DT)
    filt = np.arange(B)
    mu, cov = syn.pose_initial_batch(filt, cfg)
    off = np.concatenate([mu[:, 20:47], mu[:, 53:54]], axis=1)
    mus, covs = [mu.copy()], [cov.copy()]
    for s in range(1, steps + 1):
        inp = syn.pose_inputs(filt, s)
        for f in range(B):
            w = inp["gyro"][f]
            m, c, st, _ = orc.pose_predict(mu[f], cov[f], Q, w, par, off[f], DT)
            assert st == 0
            m, c, st, _ = orc.pose_update(orc.VELOCITY, m, c, inp["dvl"][f], syn.R_DVL, None, w, par)
            m, c, st, _ = orc.pose_update(orc.ACCELERATION, m, c, inp["acc"][f], syn.R_ACC, None, w, par)
            if s % 4 == 0:
                m, c, st, _ = orc.pose_update(orc.PRESSURE, m, c, inp["pressure"][f], syn.R_PRESSURE, None, w, par)
                m, c, st, _ = orc.pose_update(orc.WATER_VELOCITY, m, c, inp["adcp"][f], syn.R_ADCP, np.array([0.5]), w, par)
            mu[f], cov[f] = m, c
        mus.append(mu.copy())
        covs.append(cov.copy())
    np.savez_compressed(os.path.join(HERE, "pose_sequence.npz"), mu=np.array(mus), cov=np.array(covs), steps=steps)


def small_sequences():
    # VelocityUKF
    B, steps = 3, 30
    filt = np.arange(B)
    mu, cov = syn.velocity_initial_batch(filt)
    uwv = syn.uwv_params_flat()
    Q = np.diag([1e-4, 1e-4, 1e-4, 0.0])
    sim = np.stack([orc.vel_setup_motion_model(mu[f], syn.velocity_inputs(filt, 0)["gyro"][f]) for f in range(B)])
    out = [np.concatenate([mu, cov.reshape(B, -1), sim], 1)]
    for s in range(1, steps + 1):
        inp = syn.velocity_inputs(filt, s)
        for f in range(B):
            sim[f, 7:10] = inp["gyro"][f]
            mu[f], cov[f], sim[f], _ = orc.vel_predict(mu[f], cov[f], Q, sim[f], inp["gyro"][f], inp["efforts"][f], uwv, DT)
            if s % 10 == 0:
                mu[f], cov[f], _ = orc.vel_update(0, mu[f], cov[f], inp["dvl"][f], 1e-4 * np.eye(3))
                mu[f], cov[f], _ = orc.vel_update(1, mu[f], cov[f], inp["pressure"][f], np.array([[2.5e-3]]))
        out.append(np.concatenate([mu, cov.reshape(B, -1), sim], 1))
    vel = np.array(out)
    # BottomUKF
    mu, cov = syn.bottom_initial_batch(filt)
    out = [np.concatenate([mu, cov.reshape(B, -1)], 1)]
    for s in range(1, 11):
        inp = syn.bottom_inputs(filt, s)
        for f in range(B):
            mu[f], cov[f], _ = orc.bottom_predict(mu[f], cov[f], np.eye(3), inp["velocity"][f], 0.1)
            for k in range(4):
                mu[f], cov[f], _ = orc.bottom_update_range(mu[f], cov[f], inp["ranges"][f, k:k + 1], np.array([4e-4]), syn.BEAMS[k],
                                                           syn.BEAM_ORIGIN)
            if s % 5 == 0:
                mu[f], cov[f], _ = orc.bottom_update_normal(mu[f], cov[f], inp["normal"][f], 1e-4 * np.eye(2))
        out.append(np.concatenate([mu, cov.reshape(B, -1)], 1))
    bot = np.array(out)
    # IndirectPoseUKF
    pes, oes = np.array([0.01] * 3), np.array([0.002, 0.002, 0.005])
    Q = np.diag(np.concatenate([pes ** 2, oes ** 2]))
    mu = np.tile(np.array([0, 0, 0, 0, 0, 0, 1.0]), (B, 1))
    cov = np.tile(np.diag([0.25] * 3 + list(oes ** 2)), (B, 1, 1))
    pxcov = np.tile((np.eye(2) * 0.25).ravel(), (4, 1))
    covm = np.diag([1e-4] * 3 + [1e-6] * 3)
    out = [np.concatenate([mu, cov.reshape(B, -1)], 1)]
    for s in range(1, 61):
        inp = syn.indirect_inputs(filt, s)
        for f in range(B):
            mu[f], cov[f], _ = orc.indirect_predict(mu[f], cov[f], Q, 100.0, DT)
            if s % 20 == 0:
                mu[f], cov[f], _ = orc.indirect_update_features(mu[f], cov[f], inp["px"][f], pxcov, syn.MARKER_FEATURES, syn.MARKER_POSE,
                                                                covm, syn.CAMERA, syn.CAM_IN_BODY, inp["pose_ref"][f])
        out.append(np.concatenate([mu, cov.reshape(B, -1)], 1))
    ind = np.array(out)
    np.savez_compressed(os.path.join(HERE, "small_sequences.npz"), velocity=vel, bottom=bot, indirect=ind)


if __name__ == "__main__":
    pose_sequence()
    small_sequences()
    print("golden fixtures written to", HERE)

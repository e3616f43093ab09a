"""The C++ host classes (reference names / signatures) driven end to end and checked against the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as orc
from parity import TOL_COV, TOL_STATE, cov_err, state_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "dropin_test")


def build_binary():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")], stdout=subprocess.DEVNULL)
    return BIN


def test_cpp_shim_compiles_and_refuses_to_run_without_gpu(pkg):
    build_binary()
    import ctypes as C
    h = C.c_void_p()
    if pkg.load_library().uwvkf_create(0, 1, 0, C.byref(h)) == 0:
        pkg.load_library().uwvkf_destroy(h)
        pytest.skip("a CUDA device is present")
    r = subprocess.run([BIN], input="1\n", capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_classes_match_oracle(syn):
    build_binary()
    steps, DT = 10, 0.01
    filt = np.array([0])
    pose_in = [syn.pose_inputs(filt, s) for s in range(1, steps + 1)]
    vel_in = [syn.velocity_inputs(filt, s) for s in range(1, steps + 1)]
    lines = [str(steps)]
    for i in pose_in:
        lines.append(" ".join(repr(float(x)) for x in np.concatenate([i["gyro"][0], i["dvl"][0], i["acc"][0], i["pressure"][0]])))
    for i in vel_in:
        lines.append(" ".join(repr(float(x)) for x in np.concatenate([i["gyro"][0], i["efforts"][0], i["dvl"][0]])))
    r = subprocess.run([BIN], input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = {ln.split()[0]: np.array(ln.split()[1:], dtype=float) for ln in r.stdout.strip().splitlines()}
    assert out["NAN_THROWS"][0] == 1 and out["MODEL_THROWS"][0] == 1

    # PoseUKF: same sequence through the oracle, state/covariance built by the oracle's restatement of the constructor
    cfg = syn.pose_config()
    par, Q = syn.pose_params(cfg), orc.pose_process_noise_from_config(cfg, DT)
    M, Dl, Dq = syn.uwv_matrices()
    mu, cov = orc.pose_init_from_config(np.zeros(3), np.eye(3), np.array([0, 0, 0, 1.0]), np.diag([1e-4, 1e-4, 1e-2]), cfg, M, Dl, Dq)
    off = np.concatenate([mu[20:47], mu[53:54]])
    for s, i in enumerate(pose_in, 1):
        w = i["gyro"][0]
        mu, cov, st, _ = orc.pose_predict(mu, cov, Q, w, par, off, DT)
        mu, cov, st, _ = orc.pose_update(orc.VELOCITY, mu, cov, i["dvl"][0], 1e-4 * np.eye(3), None, w, par)
        mu, cov, st, _ = orc.pose_update(orc.ACCELERATION, mu, cov, i["acc"][0], 1e-4 * np.eye(3), None, w, par)
        if s % 5 == 0:
            mu, cov, st, _ = orc.pose_update(orc.PRESSURE, mu, cov, i["pressure"][0], np.array([[1e4]]), np.array([0.3, -0.1, 0.2]), w, par)
    for tag in ("POSE2", "POSE3"):
        g2 = out[tag]
        assert state_err(g2[:54], mu) <= TOL_STATE and cov_err(g2[54:].reshape(53, 53), cov) <= TOL_COV, tag
    assert np.array_equal(out["POSE2"], out["POSE3"])
    tau, Re = np.array([12.0, -3.0, 1.5, 0.0, 0.0, 0.4]), np.diag([4.0] * 3 + [0.25] * 3)
    w = pose_in[-1]["gyro"][0]
    mu, cov, st, _ = orc.pose_update(orc.BODY_EFFORTS, mu, cov, tau, Re, syn.uwv_params_flat(), w, par)
    mu, cov, st, _ = orc.pose_update(orc.BODY_EFFORTS_VEL, mu, cov, tau, Re, syn.uwv_params_flat(), w, par)
    g = out["POSE"]
    assert state_err(g[:54], mu) <= TOL_STATE and cov_err(g[54:].reshape(53, 53), cov) <= TOL_COV
    assert np.abs(out["ROTRATE"] - orc.pose_rotation_rate(mu, pose_in[-1]["gyro"][0], par)).max() < 1e-14

    # the batched entry points (one call per IMU sample; pipelined host-buffer step) end where the per-measurement calls end
    assert "POSE2" in out and "POSE3" in out

    # VelocityUKF
    uwv = syn.uwv_params_flat()
    vm, vc = np.array([1.0, 0, 0, -10.0]), np.diag([0.1, 0.1, 0.1, 1.0])
    sim = orc.vel_setup_motion_model(vm, np.zeros(3))
    for s, i in enumerate(vel_in, 1):
        sim[7:10] = i["gyro"][0]
        vm, vc, sim, st = orc.vel_predict(vm, vc, np.diag([1e-4] * 3 + [0.0]), sim, i["gyro"][0], i["efforts"][0], uwv, DT)
        if s % 5 == 0:
            vm, vc, st = orc.vel_update(0, vm, vc, i["dvl"][0], 1e-4 * np.eye(3))
    g = out["VEL"]
    assert state_err(g[:4], vm) <= TOL_STATE and cov_err(g[4:].reshape(4, 4), vc) <= TOL_COV

    # BottomUKF
    bm, bc = np.array([5.0, 0, 0, 1.0]), np.diag([1.0, 0.05, 0.05])
    bm, bc, st = orc.bottom_predict(bm, bc, np.eye(3), np.array([1.0, 0.2, -0.05]), 0.1)
    bm, bc, st = orc.bottom_update_range(bm, bc, np.array([6.4]), np.array([4e-4]),
                                         np.array([0.35355339059327373, 0.35355339059327373, -0.8660254037844387]), np.array([0.1, 0, -0.2]))
    g = out["BOTTOM"]
    assert state_err(g[:4], bm) <= TOL_STATE and cov_err(g[4:].reshape(3, 3), bc) <= TOL_COV

    # IndirectPoseUKF
    pes, oes = np.array([0.01] * 3), np.array([0.002, 0.002, 0.005])
    im, ic = np.array([0, 0, 0, 0, 0, 0, 1.0]), np.diag([0.25] * 3 + list(oes ** 2))
    im, ic, st = orc.indirect_predict(im, ic, np.diag(np.concatenate([pes ** 2, oes ** 2])), 100.0, DT)
    g = out["INDIRECT"]
    assert state_err(g[:7], im) <= TOL_STATE and cov_err(g[7:].reshape(6, 6), ic) <= TOL_COV
    assert np.abs(out["CORRECTED"] - orc.indirect_corrected_pose(im, np.array([0.2, 0.05, 0, 0, 0, 0, 1.0]))[:3]).max() < 1e-14

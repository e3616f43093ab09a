#!/usr/bin/env python
"""Drift of the small filters (VelocityUKF = BASELINE config 1/2, BottomUKF and IndirectPoseUKF = config 5) against the CPU
oracle over long free-running sequences; same idea as tests/drift_report.py.

  python tests/drift_small.py [--steps 10000] [--filters 4] [--out gpurun_out/drift_small.json]
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--filters", type=int, default=4)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import oracle_lib as orc
    from parity import cov_err, state_err
    pkg = importlib.import_module("slam-uwv_kalman_filters_b200")
    syn = importlib.import_module("slam-uwv_kalman_filters_b200.synthetic")
    B, N, DT = a.filters, a.steps, 0.01
    filt = np.arange(B)
    cps = sorted({int(round(10 ** e)) for e in np.linspace(0, np.log10(N), 9)} | {N})
    res = {"steps": N, "filters": B}

    def err(gm, gc, om, oc):
        return max(state_err(gm[f], om[f]) for f in range(B)), max(cov_err(gc[f], oc[f]) for f in range(B))

    # ---- VelocityUKF: config 1 sequence (100 Hz thrust + gyro, DVL and pressure at 10 Hz)
    mu, cov = syn.velocity_initial_batch(filt)
    uwv = syn.uwv_params_flat()
    v = pkg.VelocityUKFBatch(B)
    v.set_state(mu, cov)
    g0 = syn.velocity_inputs(filt, 0)["gyro"]
    v.integrate_gyro(g0)
    v.setup_motion_model(uwv)
    Q, R_dvl, R_p = np.diag([1e-4, 1e-4, 1e-4, 0.0]), 1e-4 * np.eye(3), np.array([[2.5e-3]])
    om, oc = mu.copy(), cov.copy()
    osim = np.stack([orc.vel_setup_motion_model(mu[f], g0[f]) for f in range(B)])
    rows = []
    for s in range(1, N + 1):
        inp = syn.velocity_inputs(filt, s)
        v.integrate_gyro(inp["gyro"]); v.integrate_body_efforts(inp["efforts"]); v.prediction_step(DT)
        if s % 10 == 0:
            v.integrate_measurement(v.DVL, inp["dvl"], R_dvl); v.integrate_measurement(v.PRESSURE, inp["pressure"], R_p)
        for f in range(B):
            osim[f][7:10] = inp["gyro"][f]
            om[f], oc[f], osim[f], st = orc.vel_predict(om[f], oc[f], Q, osim[f], inp["gyro"][f], inp["efforts"][f], uwv, DT)
            if s % 10 == 0:
                om[f], oc[f], _ = orc.vel_update(0, om[f], oc[f], inp["dvl"][f], R_dvl)
                om[f], oc[f], _ = orc.vel_update(1, om[f], oc[f], inp["pressure"][f], R_p)
        if s in cps:
            gm, gc = v.get_state()
            e = err(gm, gc, om, oc)
            rows.append({"step": s, "state_rel_err": e[0], "cov_rel_err": e[1]})
            print(f"velocity step {s:6d}: state {e[0]:.3e} cov {e[1]:.3e}", flush=True)
    res["velocity"] = {"checkpoints": rows, "status_or": int(np.bitwise_or.reduce(v.get_status()))}
    v.close()

    # ---- BottomUKF (10 Hz) and IndirectPoseUKF (100 Hz, marker corners at 2 Hz): config 5 stream
    bm, bc = syn.bottom_initial_batch(filt)
    bot = pkg.BottomUKFBatch(B)
    bot.set_state(bm, bc)
    pes, oes, TAU = np.array([0.01] * 3), np.array([0.002, 0.002, 0.005]), 100.0
    ind = pkg.IndirectPoseUKFBatch(B, pes, oes, TAU, initial_position_error_std=(0.5, 0.5, 0.5))
    Qi, Qb, Rr = np.diag(np.concatenate([pes ** 2, oes ** 2])), np.eye(3), np.array([4e-4])
    pxcov = np.tile((np.eye(2) * 0.25).ravel(), (4, 1))
    covm = np.diag([1e-4] * 3 + [1e-6] * 3)
    obm, obc = bm.copy(), bc.copy()
    oim, oic = ind.get_state()
    rows_b, rows_i = [], []
    for s in range(1, N + 1):
        ii = syn.indirect_inputs(filt, s)
        ind.update_pose_reference(ii["pose_ref"]); ind.prediction_step(DT)
        for f in range(B):
            oim[f], oic[f], st = orc.indirect_predict(oim[f], oic[f], Qi, TAU, DT)
        if s % 50 == 0:
            ind.integrate_features(ii["px"], pxcov, syn.MARKER_FEATURES, syn.MARKER_POSE, covm, syn.CAMERA, syn.CAM_IN_BODY)
            for f in range(B):
                oim[f], oic[f], st = orc.indirect_update_features(oim[f], oic[f], ii["px"][f], pxcov, syn.MARKER_FEATURES, syn.MARKER_POSE, covm,
                                                                  syn.CAMERA, syn.CAM_IN_BODY, ii["pose_ref"][f])
        if s % 10 == 0:
            bi = syn.bottom_inputs(filt, s // 10)
            bot.set_velocity(bi["velocity"]); bot.prediction_step(0.1)
            for f in range(B):
                obm[f], obc[f], st = orc.bottom_predict(obm[f], obc[f], Qb, bi["velocity"][f], 0.1)
            for k in range(4):
                bot.integrate_range(bi["ranges"][:, k].copy(), Rr, syn.BEAMS[k], syn.BEAM_ORIGIN)
                for f in range(B):
                    obm[f], obc[f], st = orc.bottom_update_range(obm[f], obc[f], bi["ranges"][f, k:k + 1], Rr, syn.BEAMS[k], syn.BEAM_ORIGIN)
        if s in cps:
            gm, gc = bot.get_state(); e = err(gm, gc, obm, obc)
            rows_b.append({"step": s, "state_rel_err": e[0], "cov_rel_err": e[1]})
            gm, gc = ind.get_state(); e2 = err(gm, gc, oim, oic)
            rows_i.append({"step": s, "state_rel_err": e2[0], "cov_rel_err": e2[1]})
            print(f"step {s:6d}: bottom state {e[0]:.3e} cov {e[1]:.3e} | indirect state {e2[0]:.3e} cov {e2[1]:.3e}", flush=True)
    res["bottom"] = {"checkpoints": rows_b, "status_or": int(np.bitwise_or.reduce(bot.get_status()))}
    res["indirect"] = {"checkpoints": rows_i, "status_or": int(np.bitwise_or.reduce(ind.get_status()))}
    print(json.dumps(res))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()

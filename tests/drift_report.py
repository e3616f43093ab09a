#!/usr/bin/env python
"""Drift of the CUDA engine against the CPU oracle over a long free-running sequence (north_star: "drift over 10k steps
reported").  Both start from the same states and consume the same synthetic inputs; neither is re-synchronised.

  python tests/drift_report.py [--steps 10000] [--filters 8] [--out gpurun_out/drift.json]

Sequence = BASELINE.json config 3/4 mix: 100 Hz predict + DVL + accelerometer, pressure at 10 Hz, ADCP (gated) at 1 Hz,
a DVL dropout window.  Reported: max over filters of the relative state error and the relative covariance error (same
norms as tests/parity.py) at logarithmically spaced checkpoints.
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10000)
    ap.add_argument("--filters", type=int, default=8)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import oracle_lib as orc
    from parity import all_errs
    pkg = importlib.import_module("slam-uwv_kalman_filters_b200")
    syn = importlib.import_module("slam-uwv_kalman_filters_b200.synthetic")
    DT = 0.01
    cfg = syn.pose_config()
    par, Q = syn.pose_params(cfg), syn.pose_process_noise(cfg, DT)
    B = a.filters
    filt = np.arange(B)
    mu, cov = syn.pose_initial_batch(filt, cfg)
    off = np.concatenate([mu[:, 20:47], mu[:, 53:54]], axis=1)
    ukf = pkg.PoseUKFBatch(B, par)
    ukf.set_process_noise(Q)
    ukf.set_state(mu, cov)
    omu, ocov = mu.copy(), cov.copy()
    checkpoints = sorted({int(round(10 ** e)) for e in np.linspace(0, np.log10(a.steps), 13)} | {a.steps})
    rows, t0 = [], time.time()
    gate_g = gate_o = 0
    for s in range(1, a.steps + 1):
        inp = syn.pose_inputs(filt, s)
        w = inp["gyro"]
        dvl_valid = np.array([not (2000 + 50 * (f % 64) <= s < 3500 + 50 * (f % 64)) for f in filt], dtype=np.uint8)
        ukf.predict_update(DT, w, pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL, valid=dvl_valid)
        ukf.integrate_measurement(pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC)
        if s % 10 == 0:
            ukf.integrate_measurement(pkg.POSE_PRESSURE, inp["pressure"], syn.R_PRESSURE)
        if s % 100 == 0:
            ukf.integrate_measurement(pkg.POSE_WATER_VELOCITY, inp["adcp"], syn.R_ADCP, aux=np.array([0.5]))
        for f in range(B):
            omu[f], ocov[f], st, _ = orc.pose_predict(omu[f], ocov[f], Q, w[f], par, off[f], DT)
            if dvl_valid[f]:
                omu[f], ocov[f], st, _ = orc.pose_update(orc.VELOCITY, omu[f], ocov[f], inp["dvl"][f], syn.R_DVL, None, w[f], par)
            omu[f], ocov[f], st, _ = orc.pose_update(orc.ACCELERATION, omu[f], ocov[f], inp["acc"][f], syn.R_ACC, None, w[f], par)
            if s % 10 == 0:
                omu[f], ocov[f], st, _ = orc.pose_update(orc.PRESSURE, omu[f], ocov[f], inp["pressure"][f], syn.R_PRESSURE, None, w[f], par)
            if s % 100 == 0:
                omu[f], ocov[f], st, _ = orc.pose_update(orc.WATER_VELOCITY, omu[f], ocov[f], inp["adcp"][f], syn.R_ADCP, np.array([0.5]), w[f], par)
                gate_o += int(st & 4 != 0)
        if s in checkpoints:
            gmu, gcov = ukf.get_state()
            e = np.max([all_errs(gmu[f], gcov[f], omu[f], ocov[f]) for f in range(B)], axis=0)
            es, ec = float(e[0]), float(e[1])
            rows.append({"step": s, "state_rel_err": es, "cov_rel_err": ec, "state_per_component_err": float(e[2]),
                         "cov_scaled_err": float(e[3])})
            print(f"step {s:6d}: state {es:.3e}  cov {ec:.3e}  state(comp) {e[2]:.3e}  cov(scaled) {e[3]:.3e}   ({time.time() - t0:.0f} s)",
                  flush=True)
    st = ukf.get_status()
    gate_g = int((st & 4 != 0).sum())
    res = {"steps": a.steps, "filters": B, "dt": DT, "sequence": "predict+DVL(dropout 2000..3500)+acc @100Hz, pressure @10Hz, ADCP gated @1Hz",
           "checkpoints": rows, "gpu_status_bits_or": int(np.bitwise_or.reduce(st)), "oracle_gate_rejections": gate_o,
           "gpu_filters_with_gate_rejection": gate_g, "mean_policy": {"tol": 1e-9, "max_it": 100}}
    print(json.dumps(res))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)
    ukf.close()


if __name__ == "__main__":
    main()

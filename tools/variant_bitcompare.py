#!/usr/bin/env python
"""Are two builds of libuwvkf.so bit-identical on the PoseUKF path?  (GPU box.)

  bash tools/build_variant.sh notma -DUWVKF_TMA_PREFIX=0 -DUWVKF_TMA_TRAIL=0      # CPU box
  python tools/variant_bitcompare.py --lib-a slam-uwv_kalman_filters_b200/libuwvkf.so --lib-b build/libuwvkf_notma.so

Each library runs the same mixed schedule (fused predict + DVL every step, accelerometer every step, pressure / ADCP / XY every
10th, a predict-only step every 7th, a DVL dropout window per filter) on the same seeded inputs in its OWN process; the final
(mu, Sigma) of every filter are compared with np.array_equal.  Data movement experiments (cp.async vs TMA bulk copies, staging
buffers) must come out identical; arithmetic experiments need not and report their distance instead.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_one(out, filters, steps):
    sys.path.insert(0, ROOT)
    import numpy as np
    pkg = importlib.import_module("slam-uwv_kalman_filters_b200")
    syn = importlib.import_module("slam-uwv_kalman_filters_b200.synthetic")
    DT = 0.01
    cfg = syn.pose_config()
    ukf = pkg.PoseUKFBatch(filters, syn.pose_params(cfg), device=0)
    ukf.set_process_noise(syn.pose_process_noise(cfg, DT))
    idx = np.arange(filters)
    mu, cov = syn.pose_initial_batch(idx, cfg)
    ukf.set_state(mu, cov)
    for s in range(1, steps + 1):
        inp = syn.pose_inputs(idx, s)
        valid = ((idx + s) % 11 != 0).astype(np.uint8)
        if s % 7 == 0:
            ukf.integrate_rotation_rate(inp["gyro"])
            ukf.prediction_step(DT)
            ukf.integrate_measurement(pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL, valid=valid)
        else:
            ukf.predict_update(DT, inp["gyro"], pkg.POSE_VELOCITY, inp["dvl"], syn.R_DVL, valid=valid)
        ukf.integrate_measurement(pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC)
        if s % 10 == 0:
            ukf.integrate_measurement(pkg.POSE_PRESSURE, inp["pressure"], syn.R_PRESSURE)
            ukf.integrate_measurement(pkg.POSE_WATER_VELOCITY, inp["adcp"], syn.R_ADCP, aux=np.array([0.5]))
            xy = np.full((filters, 2), 0.05) + 1e-3 * (idx % 17)[:, None]
            ukf.integrate_measurement(pkg.POSE_XY, xy, np.array([[0.04, 0.01], [0.01, 0.09]]))
    gmu, gcov = ukf.get_state()
    st = ukf.get_status()
    np.savez(out, mu=gmu, cov=gcov, status=st)
    ukf.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib-a", default=os.path.join(ROOT, "slam-uwv_kalman_filters_b200", "libuwvkf.so"))
    ap.add_argument("--lib-b", required=False)
    ap.add_argument("--filters", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--child", default="")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    if args.child:
        run_one(args.child, args.filters, args.steps)
        return
    import numpy as np
    res = []
    with tempfile.TemporaryDirectory() as td:
        for tag, lib in (("a", args.lib_a), ("b", args.lib_b)):
            env = dict(os.environ, UWVKF_LIBRARY=os.path.abspath(lib))
            f = os.path.join(td, tag + ".npz")
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", f, "--filters", str(args.filters),
                                   "--steps", str(args.steps)], env=env)
            with np.load(f) as z:
                res.append({k: z[k] for k in z.files})
    a, b = res
    rep = {"lib_a": args.lib_a, "lib_b": args.lib_b, "filters": args.filters, "steps": args.steps,
           "mu_identical": bool(np.array_equal(a["mu"], b["mu"])), "cov_identical": bool(np.array_equal(a["cov"], b["cov"])),
           "status_identical": bool(np.array_equal(a["status"], b["status"])),
           "mu_max_abs_diff": float(np.abs(a["mu"] - b["mu"]).max()), "cov_max_abs_diff": float(np.abs(a["cov"] - b["cov"]).max()),
           "faults": int((a["status"] & ~np.uint32(4)).astype(bool).sum()), "finite": bool(np.isfinite(a["mu"]).all() and np.isfinite(a["cov"]).all())}
    print(json.dumps(rep))
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(rep, fh, indent=1)


if __name__ == "__main__":
    main()

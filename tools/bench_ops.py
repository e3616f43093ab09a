#!/usr/bin/env python
"""Per-operation device timing of every kernel family (not the headline: bench.py is).

  python tools/bench_ops.py [--filters 262144] [--reps 5] [--out gpurun_out/ops.json]

For each PoseUKF op (predict, every integrateMeasurement kind, the fused predict+DVL step) and for the three small
filters it reports ms per launch over the batch, filter-ops/s and the achieved algorithmic HBM bandwidth
(read + write of the filter block).  Inputs are resident in HBM; timing = CUDA events on the handle's stream.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--filters", type=int, default=262144)
    ap.add_argument("--small-filters", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    ap.add_argument("--fast-only", action="store_true", help="skip the dense paths (body efforts, visual markers) and the small filters")
    ap.add_argument("--dense-only", action="store_true", help="only the dense PoseUKF paths (body efforts, visual markers): profiling runs")
    args = ap.parse_args()
    import torch
    pkg = importlib.import_module("slam-uwv_kalman_filters_b200")
    syn = importlib.import_module("slam-uwv_kalman_filters_b200.synthetic")
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    DT = 0.01
    res = {}

    def timeit(name, fn, B, bytes_per_filter, ukf):
        for _ in range(2):
            fn()
        ukf.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(args.reps):
                fn()
            e1.record(stream)
        ukf.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        res[name] = {"ms": ms, "filters": B, "filter_ops_per_s": B / (ms * 1e-3), "alg_GBps": bytes_per_filter * B / (ms * 1e-3) / 1e9}
        print(f"{name:34s} {ms:9.3f} ms  {B / (ms * 1e-3):12.4g} filter-ops/s  {res[name]['alg_GBps']:8.1f} GB/s", flush=True)

    # ---------------- PoseUKF
    B = args.filters
    cfg = syn.pose_config()
    par, Q = syn.pose_params(cfg), syn.pose_process_noise(cfg, DT)
    ukf = pkg.PoseUKFBatch(B, par, device=0, stream=stream.cuda_stream)
    ukf.set_process_noise(Q)
    _, cov0 = syn.pose_initial(cfg)
    CH = 1 << 16
    for o in range(0, B, CH):
        c = min(CH, B - o)
        mu_c, _ = syn.pose_initial_batch(np.arange(o, o + c), cfg, with_cov=False)
        ukf.set_state(mu_c, cov0, first=o)
    inp = {k: [] for k in ("gyro", "dvl", "acc", "pressure", "adcp")}
    for o in range(0, B, CH):
        c = min(CH, B - o)
        i = syn.pose_inputs(np.arange(o, o + c), 1)
        for k in inp:
            inp[k].append(i[k])
    d = {k: torch.from_numpy(np.ascontiguousarray(np.concatenate(v))).to(dev) for k, v in inp.items()}
    nav = ukf.get_navigation(0, 1)
    xy = torch.from_numpy(np.tile(nav[0, 0:2], (B, 1))).to(dev)
    zz = torch.from_numpy(np.tile(nav[0, 2:3], (B, 1))).to(dev)
    P = 23760.0
    ukf.integrate_rotation_rate(d["gyro"].data_ptr(), memspace=pkg.DEVICE)
    # a few fused steps first so that Sigma is a working filter's covariance (dense), not the block-diagonal Sigma0
    for _ in range(3):
        ukf.predict_update(DT, d["gyro"].data_ptr(), pkg.POSE_VELOCITY, d["dvl"].data_ptr(), syn.R_DVL, memspace=pkg.DEVICE)
    upd = lambda kind, z, R, aux=None: (lambda: ukf.integrate_measurement(kind, z.data_ptr(), R, aux=aux, memspace=pkg.DEVICE))
    if args.dense_only:
        timeit_fast = lambda *a_, **k_: None
    else:
        timeit_fast = timeit
    timeit_fast("pose predict", lambda: ukf.prediction_step(DT), B, P, ukf)
    timeit_fast("pose update velocity (DVL)", upd(pkg.POSE_VELOCITY, d["dvl"], syn.R_DVL), B, P, ukf)
    timeit_fast("pose update acceleration", upd(pkg.POSE_ACCELERATION, d["acc"], syn.R_ACC), B, P, ukf)
    timeit_fast("pose update xy", upd(pkg.POSE_XY, xy, np.eye(2)), B, P, ukf)
    timeit_fast("pose update z", upd(pkg.POSE_Z, zz, np.eye(1)), B, P, ukf)
    timeit_fast("pose update pressure", upd(pkg.POSE_PRESSURE, d["pressure"], syn.R_PRESSURE, np.zeros(3)), B, P, ukf)
    timeit_fast("pose update adcp", upd(pkg.POSE_WATER_VELOCITY, d["adcp"], syn.R_ADCP, np.array([0.5])), B, P, ukf)
    timeit_fast("pose fused predict+DVL", lambda: ukf.predict_update(DT, d["gyro"].data_ptr(), pkg.POSE_VELOCITY, d["dvl"].data_ptr(),
                                                                  syn.R_DVL, memspace=pkg.DEVICE), B, P, ukf)
    timeit_fast("pose fused predict+DVL+accelerometer", lambda: ukf.predict_update2(DT, d["gyro"].data_ptr(), pkg.POSE_VELOCITY, d["dvl"].data_ptr(),
                                                                               syn.R_DVL, pkg.POSE_ACCELERATION, d["acc"].data_ptr(), syn.R_ACC,
                                                                               memspace=pkg.DEVICE), B, P, ukf)
    timeit_fast("pose fused DVL+accelerometer", lambda: ukf.predict_update2(0.0, None, pkg.POSE_VELOCITY, d["dvl"].data_ptr(),
                                                                       syn.R_DVL, pkg.POSE_ACCELERATION, d["acc"].data_ptr(), syn.R_ACC,
                                                                       memspace=pkg.DEVICE), B, P, ukf)
    if args.fast_only:
        if args.out:
            json.dump(res, open(args.out, "w"), indent=1)
        return
    # dense paths (no prefix structure): body efforts (Fossen model per sigma point) and visual markers (59-DOF augmented state)
    ukf.set_model_parameters(syn.uwv_params_flat())
    eff = torch.from_numpy(np.tile(np.array([12.0, -3.0, 1.5, 0.0, 0.0, 0.4]), (B, 1))).to(dev)
    Re = np.diag([4.0] * 3 + [0.25] * 3)
    timeit("pose update body efforts", upd(pkg.POSE_BODY_EFFORTS, eff, Re), B, P, ukf)
    timeit("pose update efforts (velocity only)", upd(pkg.POSE_BODY_EFFORTS_VELOCITY_ONLY, eff, Re), B, P, ukf)
    nav = ukf.get_navigation(0, 1)[0]
    import structured_ref as sr
    marker_pose = np.concatenate([nav[0:3] + sr.qrot(nav[3:7], np.array([4.0, 0.3, -0.2])), nav[3:7]])
    cam_t = nav[0:3] + sr.qrot(nav[3:7], syn.CAM_IN_BODY[0:3]); cam_q = sr.qmul(nav[3:7], syn.CAM_IN_BODY[3:7])
    px1 = np.zeros((4, 2))
    for k, ft in enumerate(syn.MARKER_FEATURES):
        pc = sr.qrot(sr.qconj(cam_q), marker_pose[0:3] + sr.qrot(marker_pose[3:7], ft) - cam_t)
        px1[k] = [syn.CAMERA[0] * pc[0] / pc[2] + syn.CAMERA[2], syn.CAMERA[1] * pc[1] / pc[2] + syn.CAMERA[3]]
    pxd = torch.from_numpy(np.tile(px1, (B, 1, 1))).to(dev)
    pxcov = np.tile(np.array([25.0, 0.0, 0.0, 25.0]), (4, 1))
    covm = np.diag([4e-2] * 3 + [1e-2] * 3)
    timeit("pose update visual marker (4 corners)", lambda: ukf.integrate_features(pxd.data_ptr(), pxcov, syn.MARKER_FEATURES, marker_pose, covm, syn.CAMERA,
                                                                               syn.CAM_IN_BODY, memspace=pkg.DEVICE), B, P, ukf)
    faults = int((ukf.get_status() & ~np.uint32(4)).astype(bool).sum())
    res["pose_faults"] = faults
    ukf.close()
    if args.dense_only:
        if args.out:
            json.dump(res, open(args.out, "w"), indent=1)
        return

    # ---------------- small filters
    Bs = args.small_filters
    filt = np.arange(Bs)
    v = pkg.VelocityUKFBatch(Bs, device=0, stream=stream.cuda_stream)
    mu, cov = syn.velocity_initial_batch(filt)
    v.set_state(mu, cov)
    v.set_process_noise(np.diag([0.01, 0.01, 0.01, 0.001]))
    v.setup_motion_model(syn.uwv_params_flat())
    vi = syn.velocity_inputs(filt, 1)
    dv = {k: torch.from_numpy(np.ascontiguousarray(x)).to(dev) for k, x in vi.items() if isinstance(x, np.ndarray)}
    v.integrate_gyro(dv["gyro"].data_ptr(), memspace=pkg.DEVICE)
    v.integrate_body_efforts(dv["efforts"].data_ptr(), memspace=pkg.DEVICE)
    timeit("velocity predict", lambda: v.prediction_step(DT), Bs, 2 * 20 * 8, v)
    timeit("velocity update DVL", lambda: v.integrate_measurement(0, dv["dvl"].data_ptr(), syn.R_DVL, memspace=pkg.DEVICE), Bs, 2 * 20 * 8, v)
    v.close()
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

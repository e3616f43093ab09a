#!/usr/bin/env python
"""Device throughput of BASELINE.json configs 2-5 on ONE GPU at their per-GPU batch sizes.

  python tools/bench_configs.py [--steps 200] [--out gpurun_out/configs.json]

Not the headline (bench.py is): each config replays its mixed-rate schedule (what runs at 100 Hz, 10 Hz, 2 Hz, 1 Hz) with the
inputs of ONE synthetic step resident in HBM and reused, CUDA events on the handle's stream around STEPS simulated 100 Hz
steps.  Reported: simulated filter-steps/s (batch x 100 Hz steps / time) and ms per simulated step.  Parity of the same
schedules against the oracle is tests/test_configs_gpu.py.
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
DT = 0.01


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    pkg = importlib.import_module("slam-uwv_kalman_filters_b200")
    syn = importlib.import_module("slam-uwv_kalman_filters_b200.synthetic")
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    D = pkg.DEVICE
    res = {}

    def dev_arrays(d):
        return {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in d.items() if isinstance(v, np.ndarray)}

    def timeit(name, B, step_fn, handles, warm=20):
        for s in range(1, warm + 1):
            step_fn(s)
        for h in handles:
            h.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for s in range(warm + 1, warm + 1 + args.steps):
                step_fn(s)
            e1.record(stream)
        for h in handles:
            h.synchronize()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        faults = int(sum(int((h.get_status() & ~np.uint32(4)).astype(bool).sum()) for h in handles))
        res[name] = {"filters": B, "ms_per_simulated_step": ms, "filter_steps_per_s": B / (ms * 1e-3), "steps": args.steps, "faults": faults}
        print(f"{name:70s} {B:8d} filters  {ms:8.3f} ms/step  {B / (ms * 1e-3):11.4g} filter-steps/s  faults {faults}", flush=True)

    def pose_batch(B, first=0):
        cfg = syn.pose_config()
        ukf = pkg.PoseUKFBatch(B, syn.pose_params(cfg), device=0, stream=stream.cuda_stream)
        ukf.set_process_noise(syn.pose_process_noise(cfg, DT))
        _, cov0 = syn.pose_initial(cfg)
        CH = 1 << 16
        parts = []
        for o in range(0, B, CH):
            idx = np.arange(first + o, first + min(o + CH, B))
            mu_c, _ = syn.pose_initial_batch(idx, cfg, with_cov=False)
            ukf.set_state(mu_c, cov0, first=o)
            parts.append(syn.pose_inputs(idx, 1))
        inp = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
        return ukf, dev_arrays(inp)

    # ---- config 2: VelocityUKF, 65,536 filters: predict @100 Hz (efforts + gyro controls), DVL + pressure @10 Hz
    B = 65536
    filt = np.arange(B)
    v = pkg.VelocityUKFBatch(B, device=0, stream=stream.cuda_stream)
    mu, cov = syn.velocity_initial_batch(filt)
    v.set_state(mu, cov)
    v.set_process_noise(np.diag([1e-4, 1e-4, 1e-4, 0.0]))
    vi = dev_arrays(syn.velocity_inputs(filt, 1))
    v.integrate_gyro(vi["gyro"].data_ptr(), memspace=D)
    v.setup_motion_model(syn.uwv_params_flat())
    R_dvl, R_p = 1e-4 * np.eye(3), np.array([[2.5e-3]])

    def step2(s):
        v.integrate_gyro(vi["gyro"].data_ptr(), memspace=D)
        v.integrate_body_efforts(vi["efforts"].data_ptr(), memspace=D)
        v.prediction_step(DT)
        if s % 10 == 0:
            v.integrate_measurement(v.DVL, vi["dvl"].data_ptr(), R_dvl, memspace=D)
            v.integrate_measurement(v.PRESSURE, vi["pressure"].data_ptr(), R_p, memspace=D)
    timeit("config 2: VelocityUKF, predict @100 Hz, DVL + pressure @10 Hz", B, step2, [v])
    v.close()

    # ---- config 3: PoseUKF full state, 262,144 filters: fused predict + DVL and accelerometer @100 Hz, pressure @10 Hz
    B = 262144
    ukf, d = pose_batch(B)
    lever = np.array([0.3, -0.1, 0.2])

    def step3(s):
        ukf.predict_update(DT, d["gyro"].data_ptr(), pkg.POSE_VELOCITY, d["dvl"].data_ptr(), syn.R_DVL, memspace=D)
        ukf.integrate_measurement(pkg.POSE_ACCELERATION, d["acc"].data_ptr(), syn.R_ACC, memspace=D)
        if s % 10 == 0:
            ukf.integrate_measurement(pkg.POSE_PRESSURE, d["pressure"].data_ptr(), syn.R_PRESSURE, aux=lever, memspace=D)
    timeit("config 3: PoseUKF, predict + DVL + accelerometer @100 Hz, pressure @10 Hz", B, step3, [ukf])
    ukf.close()

    # ---- config 4: PoseUKF + ADCP + DVL dropouts, one 131,072-filter shard (1/8 of the 1M-filter sweep), per-filter R
    B = 131072
    first = 5 * B
    ukf, d = pose_batch(B, first=first)
    gidx = np.arange(first, first + B)
    mult = 1.0 + (gidx % 7) * 0.5
    Rpf = torch.from_numpy(np.ascontiguousarray(mult[:, None, None] ** 2 * syn.R_DVL[None])).to(dev)
    valid = torch.from_numpy(((gidx + 1) % 5 != 0).astype(np.uint8)).to(dev)
    w_adcp = np.array([0.5])

    def step4(s):
        ukf.predict_update(DT, d["gyro"].data_ptr(), pkg.POSE_VELOCITY, d["dvl"].data_ptr(), Rpf.data_ptr(), valid=valid.data_ptr(),
                           memspace=D, per_filter_R=True)
        ukf.integrate_measurement(pkg.POSE_ACCELERATION, d["acc"].data_ptr(), syn.R_ACC, memspace=D)
        if s % 10 == 0:
            ukf.integrate_measurement(pkg.POSE_PRESSURE, d["pressure"].data_ptr(), syn.R_PRESSURE, aux=lever, memspace=D)
        if s % 100 == 0:
            ukf.integrate_measurement(pkg.POSE_WATER_VELOCITY, d["adcp"].data_ptr(), syn.R_ADCP, aux=w_adcp, memspace=D)
    timeit("config 4: config 3 + gated ADCP @1 Hz, DVL dropouts, per-filter R (shard)", B, step4, [ukf])
    ukf.close()

    # ---- config 5: IndirectPoseUKF + BottomUKF, one 131,072-filter shard each: pose reference @100 Hz, marker corners @2 Hz,
    #      4 DVL beam ranges @10 Hz
    filt = np.arange(B)
    pes, oes = np.array([0.01, 0.01, 0.01]), np.array([0.002, 0.002, 0.005])
    ind = pkg.IndirectPoseUKFBatch(B, pes, oes, 100.0, initial_position_error_std=(0.5, 0.5, 0.5), device=0, stream=stream.cuda_stream)
    bot = pkg.BottomUKFBatch(B, device=0, stream=stream.cuda_stream)
    bmu, bcov = syn.bottom_initial_batch(filt)
    bot.set_state(bmu, bcov)
    ii = dev_arrays(syn.indirect_inputs(filt, 1))
    bi = syn.bottom_inputs(filt, 1)
    bvel = torch.from_numpy(np.ascontiguousarray(bi["velocity"])).to(dev)
    branges = [torch.from_numpy(np.ascontiguousarray(bi["ranges"][:, b])).to(dev) for b in range(4)]
    pxcov = np.tile((np.eye(2) * 0.25).ravel(), (4, 1))
    covm = np.diag([1e-4] * 3 + [1e-6] * 3)
    Rr = np.array([4e-4])

    def step5(s):
        ind.update_pose_reference(ii["pose_ref"].data_ptr(), memspace=D)
        ind.prediction_step(DT)
        if s % 50 == 0:
            ind.integrate_features(ii["px"].data_ptr(), pxcov, syn.MARKER_FEATURES, syn.MARKER_POSE, covm, syn.CAMERA, syn.CAM_IN_BODY, memspace=D)
        if s % 10 == 0:
            bot.set_velocity(bvel.data_ptr(), memspace=D)
            bot.prediction_step(0.1)
            for b in range(4):
                bot.integrate_range(branges[b].data_ptr(), Rr, syn.BEAMS[b], syn.BEAM_ORIGIN, memspace=D)
    timeit("config 5: IndirectPoseUKF (ref @100 Hz, markers @2 Hz) + BottomUKF (4 beams @10 Hz)", B, step5, [ind, bot])
    ind.close(); bot.close()
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

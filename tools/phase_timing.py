#!/usr/bin/env python
"""Per-phase latency of the fused PoseUKF kernel (predict + DVL update), from a -DUWVKF_PHASE_TIMING build.

  python tools/phase_timing.py --build                 # CPU box: cross-compile build/libuwvkf_timing.so (travels with gpurun)
  python tools/phase_timing.py [--filters 262144] [--steps 5] [--out gpurun_out/phases.json]     # GPU box

Warp 0 of every CTA brackets each phase of its filters with clock64(); the table is cycles per filter-step per phase, i.e. the
dependent-latency budget of ONE warp (the kernel is latency bound: DESIGN.md section 6).  The probes add a __syncwarp and a
clock read per phase, so the instrumented kernel is a few per cent slower than the product build; never a bench number.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "build", "libuwvkf_timing.so")
PHASES = [
    "wait prefix / prefetch / offsets", "predict: chol15", "predict: sigma points (tmerc, boxplus)", "predict: SO3 mean loop",
    "predict: boxminus + S_oo + A", "-", "predict: A Lp^T", "predict: Q_oo, new mean", "-",
    "predict: panel F Sigma F^T + Q", "predict: ori rows, mu", "update: z, R", "update: chol9", "update: sigma points h(x)",
    "update: mean", "update: S_zz + B", "-", "update: B Lp^T", "update: S^-1, gate, gains, K tables",
    "update: panel downdate", "apply_delta: chol6", "apply_delta: sigma points", "apply_delta: SO3 mean loop",
    "apply_delta: boxminus + S_oo + A", "-", "apply_delta: A Lp^T", "apply_delta: ori rows, mu",
    "(between update and write back)", "write back panel + next prefix", "trailing pass + status",
]


def build(lib=LIB, extra=()):
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    src = os.path.join(ROOT, "slam-uwv_kalman_filters_b200", "csrc")
    cmd = ["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-DUWVKF_PHASE_TIMING", *extra, "-shared", "-o", lib, os.path.join(src, "uwvkf.cu"),
           os.path.join(src, "small_filters.cu"), "-lcudart"]
    subprocess.check_call(cmd)
    print("built", lib)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--filters", type=int, default=262144)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default="")
    ap.add_argument("--lib", default=LIB, help="another build of the library: only ms/step is reported when it has no probes")
    ap.add_argument("--two-update", action="store_true", help="time the predict + DVL + accelerometer kernel (phases of both updates add up)")
    ap.add_argument("--mean-max-it", type=int, default=0, help="experiment: cap the iterated means (uwvkf_set_mean_policy)")
    ap.add_argument("--define", action="append", default=[], help="extra -D flags for --build (e.g. UWVKF_FUSED_MAX_WARPS=8)")
    args = ap.parse_args()
    if args.build:
        build(os.path.abspath(args.lib), [f"-D{d}" for d in args.define])
        return
    os.environ["UWVKF_LIBRARY"] = os.path.abspath(args.lib)
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    pkg = importlib.import_module("slam-uwv_kalman_filters_b200")
    syn = importlib.import_module("slam-uwv_kalman_filters_b200.synthetic")
    capi = importlib.import_module("slam-uwv_kalman_filters_b200.capi")
    lib = capi.load_library()
    lib.uwvkf_debug_phase_cycles.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    dev = torch.device("cuda", 0)
    B, DT = args.filters, 0.01
    cfg = syn.pose_config()
    ukf = pkg.PoseUKFBatch(B, syn.pose_params(cfg), device=0)
    ukf.set_process_noise(syn.pose_process_noise(cfg, DT))
    if args.mean_max_it:
        ukf.set_mean_policy(1e-9, args.mean_max_it)
    _, cov0 = syn.pose_initial(cfg)
    CH = 1 << 16
    gy, dv, ac = [], [], []
    for o in range(0, B, CH):
        idx = np.arange(o, min(o + CH, B))
        mu_c, _ = syn.pose_initial_batch(idx, cfg, with_cov=False)
        ukf.set_state(mu_c, cov0, first=o)
        i = syn.pose_inputs(idx, 1)
        gy.append(i["gyro"]); dv.append(i["dvl"]); ac.append(i["acc"])
    gyro = torch.from_numpy(np.ascontiguousarray(np.concatenate(gy))).to(dev)
    dvl = torch.from_numpy(np.ascontiguousarray(np.concatenate(dv))).to(dev)
    step = lambda: ukf.predict_update(DT, gyro.data_ptr(), pkg.POSE_VELOCITY, dvl.data_ptr(), syn.R_DVL, memspace=pkg.DEVICE)
    if args.two_update:
        acc = torch.from_numpy(np.ascontiguousarray(np.concatenate(ac))).to(dev)
        step = lambda: ukf.predict_update2(DT, gyro.data_ptr(), pkg.POSE_VELOCITY, dvl.data_ptr(), syn.R_DVL, pkg.POSE_ACCELERATION,
                                           acc.data_ptr(), syn.R_ACC, memspace=pkg.DEVICE)
    out = np.zeros(40, dtype=np.uint64)
    for _ in range(3):
        step()
    ukf.synchronize()
    probes = lib.uwvkf_debug_phase_cycles(ukf._h, out.ctypes.data, 1) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    ukf.synchronize()
    torch.cuda.synchronize()
    if not probes:
        ms = e0.elapsed_time(e1) / args.steps
        print(f"{os.path.basename(args.lib)}: {B} filters, {ms:.3f} ms/step, {B / ms * 1e3:.4g} filter-steps/s", flush=True)
        return
    ukf._ck(lib.uwvkf_debug_phase_cycles(ukf._h, out.ctypes.data, 0))
    n = float(out[39])
    tot = float(out[:39].sum())
    res = {"filters": B, "steps": args.steps, "ms_per_step_instrumented": e0.elapsed_time(e1) / args.steps,
           "sampled_filter_steps": n, "cycles_per_filter_step": tot / n, "phases": {}}
    print(f"{B} filters, {res['ms_per_step_instrumented']:.3f} ms/step (instrumented), {tot / n:.0f} cycles per filter-step per warp")
    for k, name in enumerate(PHASES):
        if name == "-":
            continue
        c = float(out[k]) / n
        res["phases"][name] = c
        print(f"  {k:2d} {name:42s} {c:9.0f} cycles  {100 * c * n / tot:5.1f} %")
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

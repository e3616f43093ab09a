#!/usr/bin/env python
"""Headline benchmark: PoseUKF predictionStep + integrateMeasurement(Velocity) filter-steps/s on a 1M-filter batch.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU reference path on the host cores)

One "step" = one fused predict + 3-D DVL update of every filter of the batch (BASELINE.json metric).
Weak scaling: every rank owns `--filters` independent filters (default 1,048,576 = the 1M-filter batch of the
metric); there is no data-path collective, NCCL only reduces the ensemble statistics and the timings.
Prints ONE JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

DT = 0.01
B_ALG = 23760.0          # algorithmic HBM bytes per filter-step: read + write of mu (54) and packed Sigma (1431) doubles
W_ALG = 1.0e6            # FP64 flop per filter-step of the REFERENCE algorithm (symmetric-minimal count)
W_EXEC = 1.441e5         # FP64 flop per filter-step executed (final round-2 ncu capture: 32 x (2 x 1395 DFMA + 485 DMUL + 125 DADD) + 512 x
                         # 66 DMMA in the fused kernel = 1.426e5, + ~1.5e3 in the thread-per-filter latitude pre-pass)
FP64_PEAK = 36.9e12      # measured DFMA peak of this pool's B200 (tools/fp64_peak.cu, profiles/r1_fp64_peak.json)
METRIC = "PoseUKF predict+update filter-steps/sec"
UNIT = "filter-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--filters", type=int, default=1 << 20, help="filters per GPU")
    ap.add_argument("--ref-filters", type=int, default=0, help="reference arm: filters per step (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[2]) for r in rows)}


def scenario():
    syn = importlib.import_module("slam-uwv_kalman_filters_b200.synthetic")
    cfg = syn.pose_config()
    return syn, cfg, syn.pose_params(cfg), syn.pose_process_noise(cfg, DT)


def our_config(filters_per_gpu, world, input_ring=4):
    """The workload both arms are quoted on (the reference arm times a bounded sample of it)."""
    return {"workload": "PoseUKF (n=53) fused predictionStep + integrateMeasurement(Velocity), SYNTH-1 inputs",
            "filters_per_gpu": filters_per_gpu, "filters_total": world * filters_per_gpu, "dt": DT,
            "parallelism": f"filters sharded x{world}",
            "l2": "working set 12.5 GB per GPU >> 126 MB L2 (no flush needed)", "input_ring": input_ring}


def host_cores():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which would hide them)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(syn, cfg, par, Q, nfilt, nsteps, nthreads=0):
    """Times the CPU oracle (literal ukfom recipe = the reference's algorithm) on nfilt filters x nsteps fused steps."""
    import oracle_lib as orc
    nthreads = nthreads or host_cores()
    filt = np.arange(nfilt)
    mu, cov = syn.pose_initial_batch(filt, cfg)
    mu, cov = np.ascontiguousarray(mu), np.ascontiguousarray(cov)
    off = np.ascontiguousarray(np.concatenate([mu[:, 20:47], mu[:, 53:54]], axis=1))
    inp = syn.pose_inputs(filt, 1)
    secs, st = orc.pose_step_batch(mu, cov, Q, inp["gyro"], inp["dvl"], syn.R_DVL, par, off, DT, nsteps=nsteps, nthreads=nthreads)
    assert not st.any(), "oracle reported a filter fault"
    return secs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib as orc
    syn, cfg, par, Q = scenario()
    cores = host_cores()
    nf = args.ref_filters or 128 * cores
    cpu_reference_run(syn, cfg, par, Q, cores, 1)  # touch
    for _ in range(args.warmup):
        cpu_reference_run(syn, cfg, par, Q, nf, 1)
    t = sum(cpu_reference_run(syn, cfg, par, Q, nf, 1) for _ in range(args.steps))
    v = nf * args.steps / t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": our_config(args.filters, args.gpus),
            "reference_sample": {"what": "CPU oracle = literal ukfom recipe (all 107 sigma points), OpenMP over filters",
                                 "filters_per_step": nf},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{nf} filters x {args.steps} fused steps, OpenMP over filters"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("slam-uwv_kalman_filters_b200")
    syn, cfg, par, Q = scenario()

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.filters, args.steps, args.warmup
    stream = torch.cuda.Stream(device=dev)
    ukf = pkg.PoseUKFBatch(B, par, device=local, stream=stream.cuda_stream)
    ukf.set_process_noise(Q)

    # per-filter initial states (global filter index = rank * B + local index), one shared Sigma0
    first = rank * B
    CH = 1 << 17
    _, cov0 = syn.pose_initial(cfg)
    for o in range(0, B, CH):
        c = min(CH, B - o)
        mu_c, _ = syn.pose_initial_batch(np.arange(first + o, first + o + c), cfg, with_cov=False)
        ukf.set_state(mu_c, cov0, first=o)
    # a ring of input sets (gyro, DVL) in pinned host memory and resident copies in HBM
    NR = 4
    host_in, dev_in = [], []
    for s in range(NR):
        g = torch.empty((B, 3), dtype=torch.float64).pin_memory()
        z = torch.empty((B, 3), dtype=torch.float64).pin_memory()
        for o in range(0, B, CH):
            c = min(CH, B - o)
            inp = syn.pose_inputs(np.arange(first + o, first + o + c), s + 1)
            g[o:o + c] = torch.from_numpy(inp["gyro"])
            z[o:o + c] = torch.from_numpy(inp["dvl"])
        host_in.append((g, z))
        dev_in.append((g.to(dev), z.to(dev)))
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident(i):
        g, z = dev_in[i % NR]
        ukf.predict_update(DT, g.data_ptr(), pkg.POSE_VELOCITY, z.data_ptr(), syn.R_DVL, memspace=pkg.DEVICE)

    # end-to-end arm: ONE C-ABI call per step with HOST buffers (uwvkf_pose_step_host, include/uwvkf.h): that step's gyro + DVL
    # samples are read from pinned host memory, the fused step runs, and the navigation solution (p, q, v) of every filter
    # lands in pinned host memory.  The library pipelines the three legs on its own copy streams (upload of step i+1 and
    # download of step i-1 overlap the kernel of step i); the host-to-device and device-to-host copies are inside the timed region.
    host_np = [(g.numpy(), z.numpy()) for g, z in host_in]
    e2e_host = [torch.empty((B, 10), dtype=torch.float64).pin_memory() for _ in range(3)]
    e2e_np = [t_.numpy() for t_ in e2e_host]

    def step_e2e(i):
        g, z = host_np[i % NR]
        ukf.step_host(DT, g, pkg.POSE_VELOCITY, z, syn.R_DVL, nav_out=e2e_np[i % 3])

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)

    def timed(fn, join=False):
        for i in range(W):
            fn(i)
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        l0 = ukf.kernel_launches
        t0 = time.time()
        with torch.cuda.stream(stream):
            ev[0].record(stream)
            for i in range(K):
                fn(W + i)
                ev[i + 1].record(stream)
            if join:                                 # e2e: the timed region ends when the last result has reached the host
                ukf.join()
                ev[K] = torch.cuda.Event(enable_timing=True)
                ev[K].record(stream)
        barrier()
        t1 = time.time()
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(K)]
        total_ms = ev[0].elapsed_time(ev[K])
        return total_ms, per, ukf.kernel_launches - l0, (t0, t1)

    tot_ms, per_ms, launches, win = timed(step_resident)
    e2e_ms, _, e2e_launches, win2 = timed(step_e2e, join=True)
    assert np.isfinite(e2e_np[(W + K - 1) % 3]).all() and abs(np.linalg.norm(e2e_np[(W + K - 1) % 3][:, 3:7], axis=1) - 1).max() < 1e-9
    st = ukf.get_status()
    faults = int((st != 0).sum())

    # ensemble statistics: the one collective of the design (NCCL all-reduce of a 12 KB buffer)
    stats = torch.zeros(ukf.ensemble_stats_size(), dtype=torch.float64, device=dev)
    ukf.ensemble_stats(stats.data_ptr(), memspace=pkg.DEVICE)
    ukf.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if world > 1:
        dist.all_reduce(stats)
    e1.record()
    torch.cuda.synchronize(dev)
    ens_ms = e0.elapsed_time(e1)
    n_total = float(stats[0].item())

    t = torch.tensor([tot_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot_ms, e2e_ms = float(t[0]), float(t[1])
    sampler.stop()
    clocks = sampler.summary(win[0], win[1])

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        kern_ms = sum(per_ms) / len(per_ms)
        # DRAM bytes of one launch from the committed `ncu --set full` capture (profiles/), valid for the same launch size
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if int(tj.get("filters_per_launch", -1)) == B:
                traffic, traffic_src = float(tj["dram_bytes_per_launch"]) / 1e9, tj.get("source")
        achieved = B_ALG * B / (kern_ms * 1e-3) / 1e9
        value = world * B * K / (tot_ms * 1e-3)
        e2e = world * B * K / (e2e_ms * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": tot_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": our_config(B, world, NR),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": peak_src, "kernel": "pose_fused_kernel<predict, DVL update, 12 warps> (kernel_ms is the whole step: the thread-per-filter latitude pre-pass, ~1 % of it, is not subtracted)",
                             "traffic_source": traffic_src,
                             "kernel_ms": kern_ms, "alg_bytes_per_filter_step": B_ALG,
                             "fp64": {"executed_tflops": W_EXEC * B / (kern_ms * 1e-3) / 1e12, "peak_tflops": FP64_PEAK / 1e12,
                                      "frac_executed": W_EXEC * B / (kern_ms * 1e-3) / FP64_PEAK,
                                      "frac_if_reference_flops": W_ALG * B / (kern_ms * 1e-3) / FP64_PEAK,
                                      "note": "secondary roof; executed flop from the committed ncu capture, peak measured"}},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 2 * B * 3 * 8, "d2h_bytes_per_step": B * 10 * 8,
                        "ms_per_step": e2e_ms / K,
                        "what": "per step ONE C-ABI call with HOST buffers (uwvkf_pose_step_host): pinned-host gyro+DVL -> HBM, fused "
                                "step, navigation solution (p,q,v) of every filter -> pinned host; the library overlaps the copies "
                                "with the neighbouring steps' kernels"},
                "gpu_launches": launches, "gpu_launches_e2e": e2e_launches, "clocks": clocks, "filter_faults": faults,
                "ensemble": {"filters_reduced": n_total, "allreduce_ms": ens_ms}}
        if world == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            nf = 64 * cores
            cpu_reference_run(syn, cfg, par, Q, nf, 1)                      # warm-up (thread team, page faults)
            t_probe = cpu_reference_run(syn, cfg, par, Q, nf, 2) / 2        # seconds per pass over nf filters
            nsteps = max(2, min(2000, int(12.0 / max(t_probe, 1e-4))))       # ~12 s of wall time on all cores
            secs = cpu_reference_run(syn, cfg, par, Q, nf, nsteps)
            line["cpu_baseline"] = {"value": nf * nsteps / secs, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{nf} filters x {nsteps} fused steps (oracle = literal ukfom recipe, OpenMP over "
                                              f"filters on {cores} threads, {secs:.1f} s)"}
        print(json.dumps(line), flush=True)
    ukf.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

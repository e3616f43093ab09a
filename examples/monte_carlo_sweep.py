#!/usr/bin/env python
"""Monte-Carlo sweep of a PoseUKF with ADCP aiding and DVL dropouts (BASELINE config 4), sharded over the GPUs of one box.

  python examples/monte_carlo_sweep.py --filters 131072 --steps 200                                    # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
      examples/monte_carlo_sweep.py --filters 1048576 --steps 200                                      # 8 GPUs, 131,072 filters each

Every filter of the sweep is an independent PoseUKF (reference src/PoseUKF.cpp): its own initial state drawn from N(mu0, Sigma0),
its own sensor noise, its own DVL noise multiplier (per-filter R) and its own DVL dropout window.  Rank r owns the contiguous
slice [r B/G, (r+1) B/G) of the global filter indices; inputs are a hash of the global index, so the sweep does not depend on how
many GPUs it runs on.  There is no exchange on the step path; the only collective is the NCCL sum of the ensemble-statistics
buffer [count, sum mu, sum mu mu^T] (12 KB) every --stats-every steps, from which rank 0 prints the ensemble mean and spread of the
navigation solution.

Schedule per simulated 100 Hz step (the calls the reference's task makes, batched):
  integrateMeasurement(RotationRate) + predictionStep(dt) + integrateMeasurement(Velocity)   -- ONE fused kernel
  integrateMeasurement(Acceleration)                                                          @100 Hz
  integrateMeasurement(Pressure)                                                              @10 Hz
  integrateMeasurement(WaterVelocityMeasurement, cell_weighting)  (chi^2 gate)                @1 Hz
"""
import argparse
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--filters", type=int, default=131072, help="filters of the WHOLE sweep (sharded over the ranks)")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--stats-every", type=int, default=100)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("slam-uwv_kalman_filters_b200")
    syn = importlib.import_module("slam-uwv_kalman_filters_b200.synthetic")
    par = importlib.import_module("slam-uwv_kalman_filters_b200.parallel")

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lo, hi = par.shard_range(args.filters, rank, world)
    B, DT = hi - lo, 0.01
    idx = np.arange(lo, hi)

    cfg = syn.pose_config()
    ukf = pkg.PoseUKFBatch(B, syn.pose_params(cfg), device=local)
    ukf.set_process_noise(syn.pose_process_noise(cfg, DT))
    _, cov0 = syn.pose_initial(cfg)
    for o in range(0, B, 1 << 16):                                    # initial states: one draw per global filter index
        c = min(1 << 16, B - o)
        mu_c, _ = syn.pose_initial_batch(idx[o:o + c], cfg, with_cov=False)
        ukf.set_state(mu_c, cov0, first=o)
    # the sweep: DVL noise multiplier 1x .. 4x and a 30-step dropout window per filter, both functions of the global index
    mult = 1.0 + 3.0 * syn.uniform(idx.astype(np.uint64), 0, 900)
    R_dvl = syn.R_DVL[None, :, :] * (mult ** 2)[:, None, None]
    drop_at = 20 + (idx % 150)
    stats = torch.zeros(ukf.ensemble_stats_size(), dtype=torch.float64, device=f"cuda:{local}")

    t0 = time.time()
    for s in range(1, args.steps + 1):
        inp = syn.pose_inputs(idx, s)
        valid = ((s < drop_at) | (s >= drop_at + 30)).astype(np.uint8)
        ukf.predict_update(DT, inp["gyro"], pkg.POSE_VELOCITY, inp["dvl"], R_dvl, valid=valid, per_filter_R=True)
        ukf.integrate_measurement(pkg.POSE_ACCELERATION, inp["acc"], syn.R_ACC)
        if s % 10 == 0:
            ukf.integrate_measurement(pkg.POSE_PRESSURE, inp["pressure"], syn.R_PRESSURE, aux=np.array([0.3, -0.1, 0.2]))
        if s % 100 == 0:
            ukf.integrate_measurement(pkg.POSE_WATER_VELOCITY, inp["adcp"], syn.R_ADCP, aux=np.array([0.5]))
        if s % args.stats_every == 0 or s == args.steps:
            ukf.ensemble_stats(out=stats.data_ptr(), memspace=pkg.DEVICE)      # [count, sum mu, sum mu mu^T] of this shard, on the device
            ukf.synchronize()
            par.allreduce_ensemble_stats(stats)                                 # the one collective: NCCL sum over the ranks
            n, mean, cov = par.ensemble_mean_cov(stats.cpu().numpy(), 54)
            if rank == 0:
                sd = np.sqrt(np.maximum(np.diag(cov), 0))
                print(f"step {s:5d}: {int(n)} filters on {world} GPU(s); ensemble position {mean[0:3].round(4)} +- {sd[0:3].round(4)} m, "
                      f"velocity {mean[7:10].round(4)} +- {sd[7:10].round(4)} m/s", flush=True)
    ukf.synchronize()
    counts = ukf.status_counts()
    if rank == 0:
        print(f"{args.steps} steps in {time.time() - t0:.1f} s (host-side input generation included); status counts of shard 0: {counts}")
    ukf.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""IndirectPoseUKF + BottomUKF on a mixed-rate measurement stream (BASELINE config 5), sharded over the GPUs of one box.

  python examples/mixed_rate_stream.py --filters 131072 --steps 200                                     # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
      examples/mixed_rate_stream.py --filters 1048576 --steps 200                                       # 8 GPUs

Per rank: one batch of IndirectPoseUKF filters (reference src/IndirectPoseUKF.cpp: pose reference @100 Hz, the four corners of a
visual marker @2 Hz) and one batch of BottomUKF filters (src/BottomUKF.cpp: velocity control + four DVL beam ranges @10 Hz), both
over the rank's contiguous slice of the global filter indices.  No exchange on the step path; every --stats-every steps the two
ensemble-statistics buffers [count, sum mu, sum mu mu^T] are summed over the ranks with NCCL and rank 0 prints the ensemble's
corrected position error estimate and bottom distance.
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
    ap.add_argument("--filters", type=int, default=131072, help="filters of the WHOLE job, per filter kind (sharded over the ranks)")
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

    pes, oes = np.array([0.01, 0.01, 0.01]), np.array([0.002, 0.002, 0.005])
    ind = pkg.IndirectPoseUKFBatch(B, pes, oes, 100.0, initial_position_error_std=(0.5, 0.5, 0.5), device=local)
    bot = pkg.BottomUKFBatch(B, device=local)
    bmu, bcov = syn.bottom_initial_batch(idx)
    bot.set_state(bmu, bcov)
    pxcov = np.tile((np.eye(2) * 0.25).ravel(), (4, 1))
    covm = np.diag([1e-4] * 3 + [1e-6] * 3)
    Rr = np.array([4e-4])
    s_ind = torch.zeros(ind.ensemble_stats_size(), dtype=torch.float64, device=f"cuda:{local}")
    s_bot = torch.zeros(bot.ensemble_stats_size(), dtype=torch.float64, device=f"cuda:{local}")

    t0 = time.time()
    for s in range(1, args.steps + 1):
        ii = syn.indirect_inputs(idx, s)
        ind.update_pose_reference(ii["pose_ref"])                      # 100 Hz
        ind.prediction_step(DT)
        if s % 50 == 0:                                                # marker corners @2 Hz
            ind.integrate_features(ii["px"], pxcov, syn.MARKER_FEATURES, syn.MARKER_POSE, covm, syn.CAMERA, syn.CAM_IN_BODY)
        if s % 10 == 0:                                                # DVL beams @10 Hz
            bi = syn.bottom_inputs(idx, s // 10)
            bot.set_velocity(bi["velocity"])
            bot.prediction_step(0.1)
            for b in range(4):
                bot.integrate_range(np.ascontiguousarray(bi["ranges"][:, b]), Rr, syn.BEAMS[b], syn.BEAM_ORIGIN)
        if s % args.stats_every == 0 or s == args.steps:
            ind.ensemble_stats(out=s_ind.data_ptr(), memspace=pkg.DEVICE)
            bot.ensemble_stats(out=s_bot.data_ptr(), memspace=pkg.DEVICE)
            ind.synchronize(); bot.synchronize()
            par.allreduce_ensemble_stats(s_ind)                         # NCCL sum over the ranks (12 KB buffers)
            par.allreduce_ensemble_stats(s_bot)
            ni, mi, ci = par.ensemble_mean_cov(s_ind.cpu().numpy(), 7)
            nb, mb, cb = par.ensemble_mean_cov(s_bot.cpu().numpy(), 4)
            if rank == 0:
                print(f"step {s:5d}: {int(ni)} + {int(nb)} filters on {world} GPU(s); position error estimate {mi[0:3].round(4)} "
                      f"+- {np.sqrt(np.maximum(np.diag(ci)[0:3], 0)).round(4)} m, bottom distance {mb[0]:.4f} +- {np.sqrt(max(cb[0, 0], 0)):.4f} m",
                      flush=True)
    ind.synchronize(); bot.synchronize()
    if rank == 0:
        print(f"{args.steps} steps in {time.time() - t0:.1f} s (host-side input generation included); status of shard 0: "
              f"indirect {ind.status_counts()}, bottom {bot.status_counts()}")
    ind.close(); bot.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

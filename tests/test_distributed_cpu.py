"""world_size-2 gloo test of the N>1 path: shards regenerate their own inputs, the only collective is the
sum-reduction of the ensemble statistics, and the reduced result equals the single-process result."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    syn = importlib.import_module("slam-uwv_kalman_filters_b200.synthetic")
    par = importlib.import_module("slam-uwv_kalman_filters_b200.parallel")
    lo, hi = par.shard_range(total, rank, world)
    mu, _ = syn.pose_initial_batch(np.arange(lo, hi), with_cov=False)       # global filter indices -> identical to 1-process run
    stats = torch.from_numpy(par.local_ensemble_stats(mu))
    par.allreduce_ensemble_stats(stats)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)                # the bench's max-over-ranks timing reduction
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((stats.numpy().copy(), float(t[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ensemble_statistics_match_single_process():
    total, world = 301, 2
    ctx = tmp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    stats, tmax = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    syn = importlib.import_module("slam-uwv_kalman_filters_b200.synthetic")
    par = importlib.import_module("slam-uwv_kalman_filters_b200.parallel")
    mu, _ = syn.pose_initial_batch(np.arange(total), with_cov=False)
    ref = par.local_ensemble_stats(mu)
    assert stats[0] == total and tmax == 2.0
    assert np.allclose(stats, ref, rtol=1e-12, atol=1e-9)

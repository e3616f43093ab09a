"""Multi-GPU plumbing: filters shard across ranks with no data-path exchange; the only collective is the
sum-reduction of the ensemble-statistics buffer (north_star: "NCCL over NVLink used only to gather ensemble
statistics").  One process per GPU; torch.distributed (nccl on GPUs, gloo in the CPU tests) is plumbing only."""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of `total` filters owned by `rank` of `world` (sizes differ by at most one)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_ensemble_stats(mu):
    """Host statement of uwvkf_ensemble_stats: [count, sum mu, sum mu mu^T (lower triangle, row-major)]."""
    mu = np.asarray(mu, dtype=np.float64)
    d = mu.shape[1]
    il = np.tril_indices(d)
    return np.concatenate([[mu.shape[0]], mu.sum(0), (mu.T @ mu)[il]])


def allreduce_ensemble_stats(stats):
    """Sum the ensemble-statistics tensor over all ranks in place (no-op when torch.distributed is not initialised)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def ensemble_mean_cov(stats, dim):
    """[count, sum x, sum x x^T] -> (count, mean[dim], covariance[dim, dim]) of the stored state vectors."""
    stats = np.asarray(stats, dtype=np.float64)
    n = stats[0]
    mean = stats[1:1 + dim] / n
    M2 = np.zeros((dim, dim))
    il = np.tril_indices(dim)
    M2[il] = stats[1 + dim:]
    M2 = M2 + np.tril(M2, -1).T
    cov = M2 / n - np.outer(mean, mean)
    return n, mean, cov

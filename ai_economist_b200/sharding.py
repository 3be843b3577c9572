"""Multi-GPU = embarrassingly parallel env replicas: rank r owns the contiguous global env range
[r*E_g, (r+1)*E_g) with seeds base + global_index; there is no collective on the step path.  torch.distributed is
used only to bracket timing (barrier) and to take the max over ranks."""


def shard_range(rank, world, envs_per_rank):
    lo = rank * envs_per_rank
    return lo, lo + envs_per_rank


def shard_seeds(base_seed, rank, world, envs_per_rank):
    lo, hi = shard_range(rank, world, envs_per_rank)
    return [base_seed + g for g in range(lo, hi)]


def max_over_ranks(value, dist=None, device=None):
    """Max of a python float over all ranks (no-op when not initialised)."""
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

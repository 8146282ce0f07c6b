"""Multi-GPU plumbing: the batch shards by contiguous system ranges, one process per GPU, with
no data-path collective (systems are independent).  The only exchange is the final reduction
of the convergence statistics, done with torch.distributed (NCCL on GPUs, gloo
in the CPU tests)."""
import numpy as np

from ._lib import NLB_STAT_COUNT, NLB_STAT_NAMES

STAT_MAX_INDEX = NLB_STAT_NAMES.index("max_iter")


def shard_range(B, rank, world_size):
    """Contiguous range [lo, hi) of systems owned by `rank`; sizes differ by at most one."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, rem = divmod(int(B), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_soa(a, lo, hi):
    """Slice the system axis (last) of an SoA array and make the slice contiguous."""
    if a is None:
        return None
    return np.ascontiguousarray(a[..., lo:hi])


def combine_stats(per_rank):
    """Combine per-rank int64[16] statistics vectors: SUM everywhere, MAX for max_iter."""
    per_rank = np.asarray(per_rank, dtype=np.int64).reshape(-1, NLB_STAT_COUNT)
    out = per_rank.sum(axis=0)
    out[STAT_MAX_INDEX] = per_rank[:, STAT_MAX_INDEX].max()
    return out


def allreduce_stats(stats, group=None):
    """All-reduce one int64[16] torch tensor (CUDA for NCCL, CPU for gloo) in place and return it.

    One all_gather of 128 bytes per rank, combined locally (SUM, and MAX for max_iter)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    world = dist.get_world_size(group)
    gathered = torch.empty((world, NLB_STAT_COUNT), dtype=torch.int64, device=stats.device)
    dist.all_gather_into_tensor(gathered, stats.reshape(1, NLB_STAT_COUNT).contiguous(), group=group)
    out = gathered.sum(dim=0)
    out[STAT_MAX_INDEX] = gathered[:, STAT_MAX_INDEX].max()
    stats.copy_(out)
    return stats


def stats_dict(stats):
    v = stats.detach().cpu().numpy() if hasattr(stats, "detach") else np.asarray(stats)
    return {k: int(v[i]) for i, k in enumerate(NLB_STAT_NAMES)}

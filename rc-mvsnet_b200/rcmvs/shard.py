"""Multi-GPU plumbing for the view-sharded hot path: one process per GPU, independent reference
views per rank, no data-path collective.  Only bookkeeping crosses ranks (unit assignment, max-over-ranks timing,
result gathering), so it runs identically over NCCL (GPU) and gloo (CPU tests)."""
import torch
import torch.distributed as dist


def shard_units(n_units, rank, world):
    """Round-robin assignment of reference-view indices to ranks (the reference's DistributedSampler semantics,
    train_rcmvsnet.py:593-603, without padding): rank r gets r, r+world, r+2*world, ..."""
    return list(range(rank, n_units, world))


def max_over_ranks(value, device="cpu", group=None):
    """max of a python float over all ranks (device-timed milliseconds are reduced this way in bench.py)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value, device="cpu", group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def gather_results(local, n_units, rank, world, group=None):
    """local: {unit_index: tensor [..]} computed by this rank -> on every rank, list of n_units tensors in unit
    order (used by view-sharded evaluation to reassemble per-view depth maps)."""
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return [local[i] for i in range(n_units)]
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: v.cpu() for k, v in local.items()}, group=group)
    merged = {}
    for g in gathered:
        merged.update(g)
    return [merged[i] for i in range(n_units)]


def throughput(units_per_rank_total, ms_max):
    """whole-job units/s = all units processed by all ranks / the slowest rank's device time."""
    return units_per_rank_total / (ms_max * 1e-3)

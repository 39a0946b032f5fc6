"""Multi-GPU partitioning of the workload: independent streams are dealt out to
ranks, one process per GPU, and there is no collective on the data path.  `torch.distributed` is
used for exactly two things: the barrier around the timed region and the max-over-ranks of the
device time.  bench.py runs these helpers over NCCL; tests/test_shard.py over gloo on the CPU."""


def rank_seeds(rank, lanes, base=100, stride=1000):
    """Writer seeds of the `lanes` streams rank `rank` decodes; disjoint between ranks."""
    if lanes > stride:
        raise ValueError("more lanes per rank than the seed stride")
    return [base + rank * stride + i for i in range(lanes)]


def max_over_ranks(value, dist=None, device="cpu"):
    """MAX of a per-rank scalar (a duration) over all ranks; identity without a process group."""
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, dist=None, device="cpu"):
    """SUM of a per-rank count (units processed) over all ranks."""
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def whole_job_rate(units_this_rank, seconds_this_rank, dist=None, device="cpu"):
    """Units all ranks processed / the slowest rank's time (the contract's `value`)."""
    total = sum_over_ranks(units_this_rank, dist, device)
    worst = max_over_ranks(seconds_this_rank, dist, device)
    return total / worst if worst > 0 else 0.0

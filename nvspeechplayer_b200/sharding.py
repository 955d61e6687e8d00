"""Multi-GPU sharding of the batch path: streams are independent (per-stream noise, no shared state), so a job of
`total` streams is cut into contiguous ranges, one per rank, with NO data-path collective.
torch.distributed is only used for the timing barrier and to combine per-rank scalars (max time, sum of samples)."""


def shard_range(total_streams, rank, world):
    """[first, last) of the streams rank `rank` renders: contiguous, sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, extra = divmod(int(total_streams), int(world))
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def weak_shard(streams_per_rank, rank):
    """Weak scaling (bench.py): every rank renders `streams_per_rank` streams; ids are globally unique."""
    first = rank * int(streams_per_rank)
    return first, first + int(streams_per_rank)


def combine(dist, device, elapsed_ms, samples):
    """(max over ranks of elapsed_ms, sum over ranks of samples) -- the two numbers the benchmark line needs."""
    import torch
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    n = torch.tensor([float(samples)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), float(n.item())

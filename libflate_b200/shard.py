"""Multi-GPU sharding of the path: streams (and DEFLATE blocks) are independent, so ranks never exchange payload bytes.
`assign_streams` is the static round-robin used by multi-stream workloads (BASELINE configs 2 and 4); the only collectives
are the counter reductions below (NCCL on GPUs, gloo in the CPU tests)."""


def assign_streams(n_streams, world, rank):
    """indices of the streams rank `rank` encodes/decodes (round-robin => balanced for equal-size streams)"""
    return list(range(rank, n_streams, world))


def max_over_ranks(x, world, device="cpu"):
    """max of a scalar over all ranks (timing is max-over-ranks by contract)"""
    if world == 1:
        return float(x)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counters(counters, world, device="cpu"):
    """all-gather of a small dict of per-rank counters (bytes, seconds): returns a list of dicts, one per rank"""
    if world == 1:
        return [dict(counters)]
    import torch
    import torch.distributed as dist
    keys = sorted(counters)
    t = torch.tensor([float(counters[k]) for k in keys], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [{k: float(v) for k, v in zip(keys, o.tolist())} for o in out]

"""world_size-2 gloo test of the multi-rank plumbing used by bench.py (no GPU): stream assignment covers every stream exactly
once, the timing reduction is a max over ranks, counters are gathered from every rank, and each rank's shard of a batch
encodes to exactly the bytes the single-process oracle produces for those streams."""
import os
import socket

import pytest
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from libflate_b200 import shard
    from oracle import oracle as orc
    n_streams = 7
    mine = shard.assign_streams(n_streams, world, rank)
    datas = [(b"stream %d " % i) * (200 + 13 * i) for i in range(n_streams)]
    encs = {i: orc.encode(orc.FMT_ZLIB, datas[i]) for i in mine}          # CPU stand-in for the per-rank device work
    t = shard.max_over_ranks(0.5 + rank, world)
    ctr = shard.gather_counters({"bytes": sum(len(datas[i]) for i in mine), "seconds": 0.5 + rank}, world)
    q.put((rank, mine, {i: len(e) for i, e in encs.items()}, t, ctr))
    dist.destroy_process_group()


def test_two_rank_sharding_and_counters():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert sorted(res[0][1] + res[1][1]) == list(range(7)) and not set(res[0][1]) & set(res[1][1])
    assert res[0][3] == res[1][3] == 1.5                                   # max over ranks
    total = sum(c["bytes"] for c in res[0][4])
    assert total == sum(len((b"stream %d " % i) * (200 + 13 * i)) for i in range(7))
    assert res[0][4] == res[1][4] and len(res[0][4]) == 2
    from oracle import oracle as orc
    for r in res:
        for i, n in r[2].items():
            assert n == len(orc.encode(orc.FMT_ZLIB, (b"stream %d " % i) * (200 + 13 * i)))

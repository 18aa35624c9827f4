"""Data-parallel gradient exchange on CPU (gloo, world size 2): the bucketed all-reduce that the hand-written backward
drives (wavjepa_b200/dist.py; reference: Lightning DDP, train.py:174-179).  No GPU and no CUDA library involved."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, bucket_bytes, ready_offsets, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wavjepa_b200.dist import BucketedAllReduce

    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(n, generator=g)
    mine = flat.clone()
    red = BucketedAllReduce(bucket_bytes=bucket_bytes)
    assert red.world_size == world
    red.begin(flat)
    for off in ready_offsets:           # the backward announces "flat[off:] is final" from the end towards 0
        red.ready(off)
    red.finish()
    others = [torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    expect = sum(others)
    ok = torch.allclose(flat, expect, rtol=0, atol=1e-6)
    covered = sorted(red.launched)
    contiguous = covered[0][0] == 0 and covered[-1][1] == n and all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
    # second step re-uses the reducer
    flat2 = mine.clone()
    red.begin(flat2)
    red.finish()
    ok2 = torch.allclose(flat2, expect, rtol=0, atol=1e-6) and red.launched == [(0, n)]
    q.put((rank, bool(ok), bool(contiguous), len(covered), bool(ok2)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,bucket_bytes,offsets", [
    (10_000, 4 * 3000, [9000, 8000, 6500, 6000, 2500, 100, 0]),   # several buckets, flushed remainder
    (777, 1 << 20, [500, 10]),                                     # one bucket, only finish() launches it
])
def test_bucketed_allreduce_gloo_world2(n, bucket_bytes, offsets):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, bucket_bytes, offsets, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, contiguous, n_buckets, ok2 in res:
        assert ok and contiguous and ok2, (rank, ok, contiguous, n_buckets, ok2)
    if bucket_bytes < 4 * n:
        assert res[0][3] > 1

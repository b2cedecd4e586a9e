"""N>1 path on CPU: two processes over the `gloo` backend exercise the batch x head partition and the
timing aggregation bench.py uses under torchrun (max over ranks of the device time, sum of FLOPs).
The attention kernels themselves never communicate, so this is everything the multi-GPU path adds."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import flashattn_b200 as fb
from flashattn_b200.sharding import Aggregator, shard_batch, shard_rows


def test_shard_batch_partitions_exactly():
    for B in (1, 7, 8, 64, 65):
        for W in (1, 2, 3, 4, 8):
            parts = [shard_batch(B, W, r) for r in range(W)]
            assert parts[0][0] == 0 and parts[-1][1] == B
            for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == B
    assert shard_rows(10, 4, 3) == (8, 10)
    with pytest.raises(ValueError):
        shard_batch(8, 2, 2)


def test_aggregator_without_process_group_is_identity():
    agg = Aggregator(None)
    agg.barrier()
    assert agg.max(3.5) == 3.5 and agg.sum(2.0) == 2.0 and agg.gather(7.0) == [7.0]
    tput, ms = agg.whole_job_throughput(1e12, 10.0)
    assert ms == 10.0 and tput == pytest.approx(1e14)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        agg = Aggregator(dist, "cpu")
        b0, b1 = shard_batch(64, world, rank)          # config #5: global batch 64
        local_flops = 14.0 * (b1 - b0) * 32 * 8192 * 8192 * 128
        local_ms = 100.0 + 25.0 * rank                 # rank 1 is the slow one
        agg.barrier()
        tput, ms = agg.whole_job_throughput(local_flops, local_ms)
        out.put((rank, b0, b1, tput, ms, agg.max(float(rank)), agg.sum(1.0), agg.gather(local_ms)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_aggregation():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, a0, a1, t0, ms0, mx0, sm0, g0), (r1, b0, b1, t1, ms1, mx1, sm1, g1) = res
    assert g0 == g1 == [100.0, 125.0]                   # per-rank times in rank order (bench.py `per_rank_ms`)
    assert (a0, a1, b0, b1) == (0, 32, 32, 64)          # disjoint halves of the batch, no overlap
    assert ms0 == ms1 == 125.0                          # max over ranks
    total = 14.0 * 64 * 32 * 8192 * 8192 * 128
    assert t0 == t1 == pytest.approx(total / 0.125)     # whole-job FLOPs / slowest rank
    assert mx0 == mx1 == 1.0 and sm0 == sm1 == 2.0

"""Batch x head sharding of the attention path across GPUs: every (batch, head) pair is
an independent problem, so ranks own disjoint batch slices and never exchange data.  The only
collective is the timing aggregation of the benchmark (max of the per-rank device times, sum of the
per-rank FLOPs) -- a control-plane reduction, done with torch.distributed (NCCL on GPUs, gloo in the
CPU tests)."""
from __future__ import annotations

from typing import Tuple


def shard_batch(global_batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous batch slice [start, stop) of `rank`; remainders go to the first ranks."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world {world_size}")
    base, rem = divmod(global_batch, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_rows(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Row slice for the companion kernels (softmax / layernorm shard over rows the same way)."""
    return shard_batch(n_rows, world_size, rank)


class Aggregator:
    """max / sum over ranks of a python float; identity when no process group is given."""

    def __init__(self, dist=None, device="cpu"):
        self.dist = dist
        self.device = device

    def barrier(self) -> None:
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, x: float, op_name: str) -> float:
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op_name))
        return float(t.item())

    def max(self, x: float) -> float:
        return self._reduce(x, "MAX")

    def gather(self, x: float) -> list:
        """The value of every rank, in rank order (per-rank attribution of a max-over-ranks time)."""
        if self.dist is None:
            return [float(x)]
        import torch
        world = self.dist.get_world_size()
        out = torch.zeros(world, dtype=torch.float64, device=self.device)
        out[self.dist.get_rank()] = float(x)
        self.dist.all_reduce(out, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in out.tolist()]

    def sum(self, x: float) -> float:
        return self._reduce(x, "SUM")

    def whole_job_throughput(self, local_units: float, local_ms: float) -> Tuple[float, float]:
        """(units of all ranks) / (max over ranks of the device time); returns (units_per_second, max_ms)."""
        ms = self.max(local_ms)
        units = self.sum(local_units)
        return units / (ms * 1e-3), ms

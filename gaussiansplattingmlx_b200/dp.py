"""View-parallel data parallelism: the batch of views is split round-robin over
the ranks, every rank holds a full replica of the parameters and Adam state, and the per-Gaussian
gradient block is summed across ranks with ONE collective between backward and Adam.

``L_batch = (1/B) * sum_v L_v`` — each rank scales its views by ``1/B`` (the GLOBAL batch size), so the
all-reduce is a plain SUM and a single rank (world 1) reproduces the reference loop exactly.

torch.distributed is plumbing here (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence


def shard_views(num_views: int, rank: int, world: int) -> List[int]:
    """Round-robin: view v belongs to rank ``v % world``."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return [v for v in range(num_views) if v % world == rank]


def grad_scale(num_views: int) -> float:
    if num_views < 1:
        raise ValueError("a step needs at least one view")
    return 1.0 / float(num_views)


class ViewParallel:
    def __init__(self, rank: int = 0, world: int = 1, group=None):
        self.rank, self.world, self.group = rank, world, group

    @classmethod
    def from_env(cls):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return cls(dist.get_rank(), dist.get_world_size())
        return cls(0, 1)

    def my_views(self, num_views: int) -> List[int]:
        return shard_views(num_views, self.rank, self.world)

    def all_reduce_sum(self, tensor):
        """In-place SUM over ranks of the contiguous gradient block (no-op for world 1)."""
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)
        return tensor

    def all_reduce_max(self, value: float, device=None) -> float:
        if self.world == 1:
            return value
        import torch
        import torch.distributed as dist
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

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

    # ---- step fused with its collective over NVLink peer memory (include/gsb.h: gsb_trainer_peers_*) ----------
    def enable_peers(self, ctx) -> bool:
        """Exchange the replicas' CUDA-IPC blobs and map every replica's trainer slab into ``ctx``.  Returns False (and
        leaves the all-reduce path in charge) for world 1, non-NCCL backends or more than 8 ranks."""
        if self.world == 1 or self.world > 8:
            return False
        import torch
        import torch.distributed as dist
        if dist.get_backend(self.group) != "nccl":
            return False
        blob = ctx.trainer_peers_export()
        mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(ctx.device)
        everyone = torch.empty(self.world * mine.numel(), dtype=torch.uint8, device=ctx.device)
        dist.all_gather_into_tensor(everyone, mine, group=self.group)
        raw = bytes(everyone.cpu().numpy().tobytes())
        n = len(blob)
        ok = 1
        try:
            ctx.trainer_peers_import(self.world, self.rank, [raw[i * n:(i + 1) * n] for i in range(self.world)])
        except Exception:   # no peer access between these devices (IPC refused): every rank falls back together
            ok = 0
        self._token = torch.zeros(1, dtype=torch.float32, device=ctx.device)
        agreed = torch.tensor([ok], dtype=torch.int32, device=ctx.device)
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=self.group)   # also: nobody launches before everyone has mapped
        return bool(agreed.item())

    # ---- the same step through the NVSwitch: symmetric memory + multicast (gsb_trainer_attach_symmetric) -----------
    def enable_multicast(self, ctx) -> bool:
        """Put the replicas' parameter and gradient blocks into torch symmetric memory (one multicast range per block) and
        attach them to ``ctx``.  Call again after every ``ctx.trainer_init`` / densification (the buffers are kept).
        Returns False when the box has no NVLS multicast (or world 1 / non-NCCL / > 8 ranks)."""
        if self.world == 1 or self.world > 8:
            return False
        import torch
        import torch.distributed as dist
        if dist.get_backend(self.group) != "nccl":
            return False
        ok = 1
        try:
            import torch.distributed._symmetric_memory as symm_mem
            floats = int(ctx.trainer_grad_block().numel())
            if getattr(self, "_sym_floats", 0) < floats:
                self._symP = symm_mem.empty(floats, dtype=torch.float32, device=ctx.device)
                self._symG = symm_mem.empty(floats, dtype=torch.float32, device=ctx.device)
                grp = self.group if self.group is not None else dist.group.WORLD
                self._hP = symm_mem.rendezvous(self._symP, grp)
                self._hG = symm_mem.rendezvous(self._symG, grp)
                self._sym_floats = floats
            if not (self._hP.multicast_ptr and self._hG.multicast_ptr):
                ok = 0
            else:
                ctx.trainer_attach_symmetric(self.world, self.rank, self._symP.data_ptr(), self._symG.data_ptr(),
                                             self._hP.multicast_ptr, self._hG.multicast_ptr, self._sym_floats)
        except Exception:
            ok = 0
        self._token = torch.zeros(1, dtype=torch.float32, device=ctx.device)
        agreed = torch.tensor([ok], dtype=torch.int32, device=ctx.device)
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=self.group)
        return bool(agreed.item())

    def multicast_step(self, ctx, iteration: int, total_iterations: int, reset_state: bool = False):
        self.stream_barrier()            # every replica's gradients are complete
        ctx.trainer_apply_multicast(iteration, total_iterations, reset_state)
        self.stream_barrier()            # every replica's parameters are written

    def disable_peers(self, ctx):
        """Unmap the other replicas' slabs on every rank, then synchronise: call before the contexts are destroyed
        (memory exported through CUDA IPC must not be freed while another process still has it open)."""
        if self.world > 1:
            import torch.distributed as dist
            ctx.trainer_peers_close()
            dist.barrier(group=self.group)

    def stream_barrier(self):
        """Stream-ordered barrier: a 4-byte all-reduce enqueued behind the work already on the current stream."""
        import torch.distributed as dist
        dist.all_reduce(self._token, group=self.group)

    def peer_step(self, ctx, iteration: int, total_iterations: int, reset_state: bool = False):
        """What replaces ``all_reduce_sum(grad_block); ctx.trainer_apply(...)`` once ``enable_peers`` succeeded."""
        self.stream_barrier()            # every replica's gradients are complete
        ctx.trainer_apply_peers(iteration, total_iterations, reset_state)
        self.stream_barrier()            # every replica's parameters (and D1 accumulators) are written

    def fused_step(self, ctx, cams, targets, grad_scale: float, iteration: int, total_iterations: int, reset_state: bool = False,
                   want_loss: bool = False, loss_out=None):
        """The whole data-parallel step in the library, replicas synchronised by flags in peer memory - no NCCL call, no host
        barrier (``gsb_trainer_step_peers``): views -> chunked projection backward of the last view -> per-chunk exchange
        (peer loads / stores, or NVLS multimem when ``enable_multicast`` attached symmetric buffers) overlapping it.  Needs
        ``enable_peers``.  ``cams`` / ``targets`` may be empty on a rank without views; returns this rank's part of the loss."""
        return ctx.trainer_step_peers(cams, targets, grad_scale, iteration, total_iterations, reset_state, want_loss, loss_out)

    def all_reduce_scalar_sum(self, value: float, device=None) -> float:
        """SUM over ranks of a host scalar (the reported batch loss: every rank holds the part of its own views)."""
        if self.world == 1:
            return value
        import torch
        import torch.distributed as dist
        if dist.get_backend(self.group) != "nccl":
            device = None
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return float(t.item())

    def all_reduce_max(self, value: float, device=None) -> float:
        if self.world == 1:
            return value
        import torch
        import torch.distributed as dist
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

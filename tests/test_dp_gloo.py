"""The N > 1 path on CPU: world_size-2 gloo.  Each rank computes the gradients of ITS views with the
oracle (weight 1/B), the gradient block is summed with ViewParallel.all_reduce_sum, and the result
must equal the single-process batch gradient — the identity the GPU view-parallel step relies on."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ["OMP_NUM_THREADS"] = "2"
    from gaussiansplattingmlx_b200.dp import ViewParallel, grad_scale
    from gaussiansplattingmlx_b200.scene import make_cameras, make_gaussians, make_targets
    from oracle import pipeline as pl
    from oracle.api import Port
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    vp = ViewParallel.from_env()
    assert (vp.rank, vp.world) == (rank, world)
    n, W, H, B = 150, 32, 24, 4
    params = make_gaussians(n, 77, 2)
    cams = make_cameras(W, H, B); targets = make_targets(W, H, B, 77)
    o = Port()
    keys = list(params)
    block = torch.zeros(sum(params[k].size for k in keys), dtype=torch.float64)
    for v in vp.my_views(B):
        _, _, bw = pl.loss_and_grads(o, params, cams[v], targets[v], 2)
        block += grad_scale(B) * torch.from_numpy(np.concatenate([bw["grads"][k].reshape(-1) for k in keys]).astype(np.float64))
    vp.all_reduce_sum(block)
    t_max = vp.all_reduce_max(float(rank + 1))
    assert t_max == float(world)
    # the peer-memory / multicast steps need NCCL on one NVLink box: on gloo both report "not available" (without touching
    # the context) and the caller stays on all-reduce + apply
    assert vp.enable_peers(None) is False and vp.enable_multicast(None) is False
    np.save(Path(out_dir) / f"block_{rank}.npy", block.numpy())
    dist.destroy_process_group()


def test_owned_slices_cover_every_gaussian_once():
    """The slicing rule of gsb_trainer_apply_peers / _multicast (api.cu): equal slices on multiples of 4 Gaussians."""
    for n in (1, 3, 4, 5, 50001, 1_000_000, 6_000_003):
        for world in (1, 2, 3, 4, 8):
            per = (((n + world - 1) // world) + 3) & ~3
            covered = 0
            for r in range(world):
                g0 = min(r * per, n); g1 = min(g0 + per, n)
                assert g0 == covered and g0 % 4 == 0 or g0 == n
                covered = g1
            assert covered == n


def test_view_parallel_gradient_equals_batch_gradient(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    b0 = np.load(tmp_path / "block_0.npy"); b1 = np.load(tmp_path / "block_1.npy")
    assert np.array_equal(b0, b1)                       # replicas stay identical
    from gaussiansplattingmlx_b200.scene import make_cameras, make_gaussians, make_targets
    from oracle import pipeline as pl
    from oracle.api import Port
    n, W, H, B = 150, 32, 24, 4
    params = make_gaussians(n, 77, 2)
    cams = make_cameras(W, H, B); targets = make_targets(W, H, B, 77)
    o = Port()
    ref = np.zeros_like(b0)
    for v in range(B):
        _, _, bw = pl.loss_and_grads(o, params, cams[v], targets[v], 2)
        ref += np.concatenate([bw["grads"][k].reshape(-1) for k in params]).astype(np.float64) / B
    assert np.abs(b0 - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


# ------------------------------------------------------------------------------------------------
# host logic of GaussianTrainer.startTrain on two ranks: the early-stopping decision must be taken from the
# batch loss summed over ranks (a rank-local loss would let one rank leave the loop while the other blocks in
# the next collective) and before the optimiser update, like the reference (GaussianTrainer.swift:1003-1058)
# ------------------------------------------------------------------------------------------------
class _FakeCtx:
    """Duck-typed stand-in for context.Context: records the calls the trainer makes, no GPU."""
    device = None

    def __init__(self, losses):
        self.losses, self.applied, self.accumulated = list(losses), [], []
        self.block = torch.zeros(8)

    def trainer_init(self, params): pass
    def trainer_grad_block(self): return self.block
    def trainer_tensors(self):
        z = {k: torch.zeros(1) for k in ("_xyz", "_features_dc", "_features_rest", "_scales", "_rotation", "_opacity")}
        return {"params": z, "m": z, "v": z}

    def trainer_accumulate(self, cams, targets, zero_grads=True, grad_scale=None, want_loss=True, **kw):
        self.accumulated.append(len(cams))
        return self.losses.pop(0) if want_loss else None

    def trainer_apply(self, iteration, total, reset_state=False): self.applied.append(iteration)
    def trainer_densify(self, *a, **k): raise AssertionError("densification is outside [500, 15000] in this test")


def _trainer_worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    from types import SimpleNamespace
    from gaussiansplattingmlx_b200.dp import ViewParallel
    from gaussiansplattingmlx_b200.model import GaussModel
    from gaussiansplattingmlx_b200.scene import make_cameras, make_gaussians
    from gaussiansplattingmlx_b200.trainer import GaussianTrainer, TrainData
    from gaussiansplattingmlx_b200 import _lib
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    _lib.make_camera = lambda c: c                       # no ctypes camera needed by the fake context
    cams = make_cameras(16, 16, 3)
    data = TrainData(cams, [np.zeros((16, 16, 3), np.float32)] * 3)
    # one view per step: rank 0 renders it, rank 1 has no view and contributes 0 to the batch loss.
    # loss read-backs happen on iterations 0, 9, 19, 20: 0.5, 0.25, then 5e-5 < 1e-4 on iteration 19
    ctx = _FakeCtx([0.5, 0.25, 5e-5] if rank == 0 else [])
    params = make_gaussians(10, 1, 1)
    model = GaussModel.__new__(GaussModel)
    for k, v in params.items():
        setattr(model, k, v)
    tr = GaussianTrainer(model, data, SimpleNamespace(ctx=ctx), iterationCount=100, views_per_step=1,
                         parallel=ViewParallel.from_env(), seed=0)
    assert tr._peers is False                            # gloo: all-reduce + apply
    tr.startTrain(earlyStoppingThreshold=1e-4)
    np.save(Path(out_dir) / f"trainer_{rank}.npy", np.array([len(ctx.applied), int(tr.stopped_early), len(tr.losses),
                                                              len(ctx.accumulated)] + [round(l, 9) for l in tr.losses]))
    dist.destroy_process_group()


def test_trainer_early_stop_is_a_collective_decision(tmp_path):
    world = 2
    mp.spawn(_trainer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "trainer_0.npy"); r1 = np.load(tmp_path / "trainer_1.npy")
    # both ranks applied iterations 0..18 and stopped on iteration 19 before its update, with the same reported losses
    assert r0[0] == r1[0] == 19 and r0[1] == r1[1] == 1 and r0[2] == r1[2] == 3
    assert np.array_equal(r0[4:], r1[4:]) and abs(r0[-1] - 5e-5) < 1e-12
    assert r0[3] == 20 and r1[3] == 0                    # only rank 0 rendered

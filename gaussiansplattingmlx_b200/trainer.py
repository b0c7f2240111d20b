"""``GaussianTrainer`` — host-side mirror of the reference train loop
(``Trainer/GaussianTrainer.swift:448-477`` init, ``:934-1114`` startTrain) over the fused CUDA step.

One iteration = ``gsb_trainer_accumulate`` (per view: render → loss → backward, gradients summed with
weight 1/B) → ``gsb_trainer_apply`` (Adam + D1 grad-norm accumulation); view-parallel: the step fused with its
collective over NVLink peer memory (``dp.ViewParallel.peer_step``), or all-reduce + apply on non-NCCL backends.
With ``views_per_step == 1`` and one rank this is the reference iteration, including ``split_and_prune`` every
``split_and_prune_per_iteration`` iterations (``gsb_trainer_densify``) and the optimiser-state reset cadence.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .camera import Camera
from .dp import ViewParallel
from .model import GaussModel, PARAM_ORDER
from .renderer import GaussianRenderer


def _pin(t: torch.Tensor) -> torch.Tensor:
    """Page-locked staging memory for the per-view H2D copies (plain memory where no CUDA device exists: CPU tests)."""
    return t.pin_memory() if torch.cuda.is_available() else t


@dataclass
class TrainData:
    """``GaussianTrainer.swift`` TrainData: cameras + rgbArray[V,H,W,3] (f32) and, when the dataset carries them,
    alphaArray[V,H,W] / depthArray[V,H,W] (depth supervision: ``GaussianTrainer.swift:486-498,949``)."""
    cameras: List[Camera]
    rgbArray: Sequence[np.ndarray]
    alphaArray: Optional[Sequence[np.ndarray]] = None
    depthArray: Optional[Sequence[np.ndarray]] = None

    def getNumCameras(self) -> int:
        return len(self.cameras)

    def getViewPointCamera(self, index: int) -> Camera:
        return self.cameras[index]


class GaussianTrainer:
    # reference defaults (GaussianTrainer.swift:277-300)
    lambda_dssim = 0.2
    lambda_depth = 0.0
    # split and prune parameters (GaussianTrainer.swift:281, 293-300)
    split_and_prune_per_iteration = 100
    gradientThreshold = 0.0002
    minOpacity = 0.005
    maxScale = 0.01
    densifyFromIter = 500
    densifyUntilIter = 15000
    maxGaussians = 1_000_000

    def __init__(self, model: GaussModel, data: TrainData, gaussRender: GaussianRenderer, iterationCount: int,
                 views_per_step: int = 1, parallel: Optional[ViewParallel] = None, seed: Optional[int] = None,
                 reset_optimizer_state: bool = True, outputDirectoryURL: Optional[str] = None,
                 save_snapshot_per_iteration: int = 100, step_log_path: Optional[str] = None):
        self.model, self.data, self.gaussRender = model, data, gaussRender
        self.iterationCount = iterationCount
        self.views_per_step = views_per_step
        self.parallel = parallel or ViewParallel()
        self.reset_optimizer_state = reset_optimizer_state
        self.densify = True
        self.outputDirectoryURL = outputDirectoryURL
        self.save_snapshot_per_iteration = save_snapshot_per_iteration
        self.snapshots: List[str] = []
        self.densify_seed = 0 if seed is None else int(seed)
        self.densify_log: List[Dict[str, int]] = []
        self.forceStop = False
        self.stopped_early = False
        # JSONL step log (one line per loss read-back: iteration, loss, Gaussian count, pairs of the last view, wall ms):
        # the machine-readable counterpart of the reference's "[Profile] iter=..." report lines
        self.step_log_path = step_log_path
        self.delegate: Optional[Callable[[float, int], None]] = None   # pushLoss(loss, iteration)
        self._rng = random.Random(seed)
        ctx = gaussRender.ctx
        ctx.trainer_init({k: torch.from_numpy(np.ascontiguousarray(getattr(model, k))) for k in PARAM_ORDER})
        self._gcams = [_lib.make_camera(c) for c in data.cameras]
        self._targets = [_pin(torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32))) for t in data.rgbArray]
        # effectiveLambdaDepth = data.depthArray != nil ? lambda_depth : 0 (GaussianTrainer.swift:949); depthMask = alpha > 0.5 (:492)
        self._depths = self._masks = None
        if data.depthArray is not None:
            self._depths = [_pin(torch.from_numpy(np.ascontiguousarray(d, dtype=np.float32))) for d in data.depthArray]
            alphas = data.alphaArray if data.alphaArray is not None else [np.ones_like(d) for d in data.depthArray]
            self._masks = [_pin(torch.from_numpy(np.ascontiguousarray(np.asarray(a) > 0.5).astype(np.uint8))) for a in alphas]
        self._grad_block = ctx.trainer_grad_block() if self.parallel.world > 1 else None
        # one process per GPU over NCCL: the step is fused with its collective over NVLink peer memory (dp.peer_step);
        # otherwise (gloo, a single rank) all-reduce + gsb_trainer_apply
        self._peers = self.parallel.world > 1 and self.parallel.enable_peers(ctx)
        # ... and through the NVSwitch (NVLS multimem exchange kernel) where the box has multicast: measured 0.65 ms against
        # 0.81 ms per step at 8 GPUs (profiles/r2/r2e_exchange_8gpu.jsonl)
        self._multicast = self._peers and self.parallel.enable_multicast(ctx)
        self.losses: List[float] = []

    def stopTrain(self):
        self.forceStop = True

    def fetchTrainData(self) -> List[int]:
        """Random view indices for this iteration (``GaussianTrainer.swift:486-498``, batched)."""
        n = self.data.getNumCameras()
        return [self._rng.randrange(n) for _ in range(self.views_per_step)]

    def train_iteration(self, iteration: int, view_indices: Optional[Sequence[int]] = None, want_loss: bool = True,
                        early_stopping_threshold: Optional[float] = None):
        """One iteration of ``startTrain`` (``GaussianTrainer.swift:958-1114``).  Returns the loss (None when not asked for).
        With ``early_stopping_threshold`` set and a loss below it the iteration stops BEFORE the optimiser update, like
        the reference (:1044-1058), and ``self.stopped_early`` is raised."""
        ctx = self.gaussRender.ctx
        views = list(view_indices) if view_indices is not None else self.fetchTrainData()
        B = len(views)
        mine = [views[i] for i in self.parallel.my_views(B)]
        loss = None
        fused = self._peers and self._depths is None and not (want_loss and early_stopping_threshold is not None)
        if fused:
            # the whole step in the library, replicas synchronised by flags in peer memory (no NCCL call on the step); used
            # whenever no early-stopping decision has to be taken between the backward and the update
            loss = self.parallel.fused_step(ctx, [self._gcams[v] for v in mine], [self._targets[v] for v in mine], 1.0 / B, iteration,
                                            self.iterationCount, want_loss=want_loss and bool(mine))
            if want_loss and self.parallel.world > 1:
                loss = self.parallel.all_reduce_scalar_sum(loss or 0.0, ctx.device)
            self._after_update(iteration)
            return loss
        if mine:
            kw = {}
            if self._depths is not None:
                kw = dict(target_depths=[self._depths[v] for v in mine], depth_masks=[self._masks[v] for v in mine],
                          lambda_depth=self.lambda_depth)
            loss = ctx.trainer_accumulate([self._gcams[v] for v in mine], [self._targets[v] for v in mine], zero_grads=True,
                                          grad_scale=1.0 / B, want_loss=want_loss, **kw)
        elif self.parallel.world > 1:
            self._grad_block.zero_()
        if want_loss and self.parallel.world > 1:
            # every rank holds (1/B) x the sum over ITS views: the batch loss is the SUM over ranks (a rank without views
            # contributes 0), and all ranks must take the same early-stopping decision or the next collective deadlocks
            loss = self.parallel.all_reduce_scalar_sum(loss or 0.0, ctx.device)
        if want_loss and early_stopping_threshold is not None and loss is not None and loss < early_stopping_threshold:
            self.stopped_early = True
            return loss
        if self._peers:
            (self.parallel.multicast_step if self._multicast else self.parallel.peer_step)(ctx, iteration, self.iterationCount, reset_state=False)
        else:
            if self.parallel.world > 1:
                self.parallel.all_reduce_sum(self._grad_block)
            ctx.trainer_apply(iteration, self.iterationCount, reset_state=False)
        self._after_update(iteration)
        return loss

    def _after_update(self, iteration: int):
        ctx = self.gaussRender.ctx
        if self.outputDirectoryURL and iteration % self.save_snapshot_per_iteration == 0:
            self.save_snapshot(iteration)                     # GaussianTrainer.swift:1092 (before split_and_prune)
        if iteration % self.split_and_prune_per_iteration == 0:   # :1098-1114, independent of the optimiser policy
            if self.densify:
                self.split_and_prune(iteration)
            if self.reset_optimizer_state:
                # the reference re-creates the Adam state after split_and_prune whether or not N changed (:1104-1109)
                tt = ctx.trainer_tensors()
                for k in PARAM_ORDER:
                    tt["m"][k].zero_(); tt["v"][k].zero_()

    def split_and_prune(self, iteration: int):
        """``GaussianTrainer.split_and_prune`` (``GaussianTrainer.swift:766-908``): iteration guard on the host, the rest in
        ``gsb_trainer_densify``.  The noise is drawn from a counter-based generator keyed by (seed, iteration) so that every
        view-parallel replica produces identical tensors."""
        if not (self.densifyFromIter <= iteration <= self.densifyUntilIter):
            return None
        ctx = self.gaussRender.ctx
        had_peers = self._peers
        if had_peers:
            # Densification swaps (and may free) the trainer slab the other replicas have mapped through CUDA IPC: every
            # replica unmaps first and passes a barrier, then densifies, then the new slabs are exchanged again
            # (gsb_trainer_densify refuses to run while peer mappings are open).
            self.parallel.disable_peers(ctx)
            self._peers = False
        info = ctx.trainer_densify(self.gradientThreshold, self.maxScale, self.minOpacity, self.maxGaussians,
                                   seed=(self.densify_seed << 32) ^ iteration)
        self.densify_log.append(dict(info, iteration=iteration))
        if self.parallel.world > 1:
            self._grad_block = ctx.trainer_grad_block()
            if had_peers:
                self._peers = self.parallel.enable_peers(ctx)
                self._multicast = self._peers and self._multicast and self.parallel.enable_multicast(ctx)
        return info

    def close(self):
        """Unmap the other replicas' slabs (on every rank, with a barrier) before any context is destroyed."""
        if self._peers:
            self.parallel.disable_peers(self.gaussRender.ctx)
            self._peers = False

    def save_snapshot(self, iteration: int):
        """``GaussianTrainer.save_snapshot`` (``GaussianTrainer.swift:909-929``) + the resume sidecar the reference lacks."""
        from .ply import save_snapshot, save_resume
        ctx = self.gaussRender.ctx
        tt = ctx.trainer_tensors()
        host = lambda d: {k: v.cpu().numpy() for k, v in d.items()}
        out = save_snapshot(host(tt["params"]), self.outputDirectoryURL, iteration)
        save_resume(str(out)[:-4] + ".resume.npz", iteration, host(tt["m"]), host(tt["v"]), tt["accum"].cpu().numpy(),
                    ctx.trainer_count()[1])
        self.snapshots.append(str(out))
        return out

    def startTrain(self, earlyStoppingThreshold: float = 1e-4):
        """``startTrain`` (``GaussianTrainer.swift:934-1114``): the loss is read back on iterations 9, 19, ... (FPS window) and
        0, 20, 40, ... (preview image) (:1003-1012) and early stopping is decided before the optimiser update (:1044)."""
        self.stopped_early = False
        for iteration in range(self.iterationCount):
            if self.forceStop:
                break
            report = iteration % 10 == 9 or iteration % 20 == 0
            loss = self.train_iteration(iteration, want_loss=report, early_stopping_threshold=earlyStoppingThreshold)
            if report and loss is not None:
                self.losses.append(loss)
                if self.delegate:
                    self.delegate(loss, iteration)
                if self.step_log_path and self.parallel.rank == 0:
                    import json, time
                    st = self.gaussRender.ctx.stats()
                    with open(self.step_log_path, "a") as f:
                        f.write(json.dumps({"iteration": iteration, "loss": loss, "gaussians": self.gaussRender.ctx.trainer_count()[0],
                                            "pairs_last_view": st["pairs_last_view"], "kernel_launches": st["kernel_launches"],
                                            "t": time.time()}) + "\n")
            if self.stopped_early:
                break
        self.sync_model()
        self.close()

    def sync_model(self):
        """Copy the trained parameters back into the host-side ``GaussModel``."""
        tt = self.gaussRender.ctx.trainer_tensors()
        for k in PARAM_ORDER:
            setattr(self.model, k, tt["params"][k].cpu().numpy().copy())

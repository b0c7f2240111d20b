#!/usr/bin/env python3
"""Correctness check of the data-parallel step variants on 2+ GPUs (run under torchrun, or imported by bench.py in its
untimed set-up: `run_dp_check`).  Same initial model, same views, three steps each through

    nccl       all-reduce of the gradient block + gsb_trainer_apply on every replica            (the reference point)
    peers      barrier + gsb_trainer_apply_peers + barrier (one kernel over NVLink peer memory)
    fused      gsb_trainer_step_peers: flags in peer memory, chunked projection backward overlapping the exchange
    fused_b0   the same with a step in which only rank 0 has a view (B = 0 on the others)
    multicast  fused, NVLS multimem exchange (when the box has multicast)
    densify    peers closed -> gsb_trainer_densify with growth on every replica -> peers re-opened -> fused step

and compares, against `nccl`: the D1 accumulators (no Adam amplification: tight), the parameter deltas (fraction of
elements off by more than 1e-3 of the largest delta; Adam without bias correction turns last-bit gradient noise into
O(lr) steps) - and, for every variant, that all replicas hold bit-identical parameters and accumulators.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def run_dp_check(rank: int, world: int, local_rank: int, n: int = 50001, W: int = 320, H: int = 192, want_multicast: bool = True):
    from gaussiansplattingmlx_b200 import _lib as L
    from gaussiansplattingmlx_b200.context import Context
    from gaussiansplattingmlx_b200.dp import ViewParallel
    from gaussiansplattingmlx_b200.scene import make_cameras, make_gaussians, make_targets

    dev = torch.device("cuda", local_rank)
    params = make_gaussians(n, 7, 3)                    # n not a multiple of 4 x world: ragged last slices
    views = 2 * world
    cams = make_cameras(W, H, views)
    targets = make_targets(W, H, views, 7)
    mine = [v for v in range(views) if v % world == rank]
    gc = [L.make_camera(cams[v]) for v in mine]
    tg = [torch.from_numpy(targets[v]).to(dev) for v in mine]
    gc0 = [L.make_camera(cams[0])] if rank == 0 else []
    tg0 = [torch.from_numpy(targets[0]).to(dev)] if rank == 0 else []
    vp = ViewParallel(rank, world)
    steps = 3

    def snapshot(ctx):
        torch.cuda.synchronize(dev)
        tt = ctx.trainer_tensors()
        out = {k: v.cpu().numpy().copy() for k, v in tt["params"].items()}
        out["_accum"] = tt["accum"].cpu().numpy().copy()
        return out

    def run(mode: str):
        ctx = Context(W, H, device=local_rank)
        ctx.trainer_init({k: torch.from_numpy(v) for k, v in params.items()})
        ok = True
        if mode != "nccl":
            ok = vp.enable_peers(ctx)
        if ok and mode == "multicast":
            ok = vp.enable_multicast(ctx)
        if not ok:
            ctx.close()
            return None
        info = {}
        for it in range(steps):
            b0 = mode == "fused_b0" and it == 1
            cams_it, tg_it, scale = (gc0, tg0, 1.0) if b0 else (gc, tg, 1.0 / views)
            if mode == "nccl":
                gb = ctx.trainer_grad_block()
                if cams_it:
                    ctx.trainer_accumulate(cams_it, tg_it, zero_grads=True, grad_scale=scale, want_loss=False)
                else:
                    gb.zero_()
                dist.all_reduce(gb)
                ctx.trainer_apply(it, 100)
            elif mode == "peers":
                ctx.trainer_accumulate(cams_it, tg_it, zero_grads=True, grad_scale=scale, want_loss=False)
                vp.peer_step(ctx, it, 100)
            else:
                vp.fused_step(ctx, cams_it, tg_it, scale, it, 100)
                if mode == "densify" and it == 0:
                    # clone / split / prune with growth between two fused steps: every replica unmaps, passes a barrier,
                    # densifies identically (counter-based noise), and the new slabs are exchanged again
                    vp.disable_peers(ctx)
                    acc = ctx.trainer_tensors()["accum"]
                    thr = float(torch.quantile(acc, 0.9))
                    info = ctx.trainer_densify(thr, 0.01, 0.005, 10 * n, seed=11)
                    assert info["n"] > n, "the densification of this check must grow the model"
                    assert vp.enable_peers(ctx)
        out = snapshot(ctx)
        if mode != "nccl":
            ctx.trainer_peers_check()
            vp.disable_peers(ctx)
        dist.barrier()
        ctx.close()
        out["_info"] = info
        return out

    def run_nccl_b0_or_densify(kind: str):
        """The NCCL reference of the two special sequences."""
        ctx = Context(W, H, device=local_rank)
        ctx.trainer_init({k: torch.from_numpy(v) for k, v in params.items()})
        for it in range(steps):
            b0 = kind == "fused_b0" and it == 1
            cams_it, tg_it, scale = (gc0, tg0, 1.0) if b0 else (gc, tg, 1.0 / views)
            gb = ctx.trainer_grad_block()
            if cams_it:
                ctx.trainer_accumulate(cams_it, tg_it, zero_grads=True, grad_scale=scale, want_loss=False)
            else:
                gb.zero_()
            dist.all_reduce(gb)
            ctx.trainer_apply(it, 100)
            if kind == "densify" and it == 0:
                acc = ctx.trainer_tensors()["accum"]
                thr = float(torch.quantile(acc, 0.9))
                ctx.trainer_densify(thr, 0.01, 0.005, 10 * n, seed=11)
        out = snapshot(ctx)
        dist.barrier()
        ctx.close()
        return out

    def replicas_identical(res) -> bool:
        same = True
        for k, v in res.items():
            if k == "_info":
                continue
            t = torch.from_numpy(v.view(np.int32).astype(np.int64)).to(dev).sum()
            lo, hi = t.clone(), t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            same &= bool(lo == hi)
        return same

    def compare(res, ref, p0):
        acc_rel = float(np.abs(res["_accum"] - ref["_accum"]).max() / max(np.abs(ref["_accum"]).max(), 1e-12))
        worst = 0.0
        for k in p0:
            if res[k].shape != ref[k].shape:
                return {"accum_rel": acc_rel, "param_off_frac": 1.0}
            base = p0[k].reshape(ref[k].shape) if p0[k].size == ref[k].size else None
            da = ref[k] - (base if base is not None else 0.0)
            db = res[k] - (base if base is not None else 0.0)
            scale = max(np.abs(da).max(), 1e-12)
            worst = max(worst, float((np.abs(da - db) > 1e-3 * scale).mean()))
        return {"accum_rel": acc_rel, "param_off_frac": worst}

    report = {"world": world, "gaussians": n, "steps": steps, "variants": {}}
    ref = run("nccl")
    ok_all = True
    modes = ["peers", "fused", "fused_b0", "densify"] + (["multicast"] if want_multicast else [])
    for mode in modes:
        res = run(mode)
        if res is None:
            report["variants"][mode] = "unavailable"
            continue
        r = ref if mode in ("peers", "fused", "multicast") else run_nccl_b0_or_densify(mode)
        if mode == "densify":
            # the two paths sum the gradients in different orders, so a handful of Gaussians sit on the other side of the
            # densification threshold: the counts agree closely, not exactly; what must hold exactly is replica identity
            n_res, n_ref = int(res["_xyz"].shape[0]), int(r["_xyz"].shape[0])
            cmp = {"n_after": n_res, "n_after_nccl": n_ref, "finite": bool(all(np.isfinite(v).all() for k, v in res.items() if k != "_info")),
                   "replicas_identical": replicas_identical(res)}
            ok = cmp["replicas_identical"] and cmp["finite"] and n_res > n and abs(n_res - n_ref) <= max(8, n_ref // 100)
        else:
            cmp = compare(res, r, params)
            cmp["replicas_identical"] = replicas_identical(res)
            ok = cmp["replicas_identical"] and cmp["accum_rel"] < 1e-4 and cmp["param_off_frac"] < 2e-2
        # every rank must agree on the verdict
        t = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        cmp["ok"] = bool(t.item())
        ok_all &= cmp["ok"]
        report["variants"][mode] = cmp
    report["ok"] = ok_all
    return report


if __name__ == "__main__":
    import json
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rep = run_dp_check(rank, world, local, n=int(os.environ.get("N", 50001)))
    if rank == 0:
        print("DP_CHECK " + json.dumps(rep), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if rep["ok"] else 1)

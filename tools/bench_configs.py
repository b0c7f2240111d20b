#!/usr/bin/env python3
"""Timings of BASELINE.json's other configurations (they are parity-test cases, not bench lines): C2 single-view train
step, C4 densification stress (6 M Gaussians at 3840x2160 with clone/split/prune), C5 forward-only render sweep.
Prints one JSON object per line; run on a B200 (`gpurun -- python tools/bench_configs.py > gpurun_out/configs.jsonl`)."""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gaussiansplattingmlx_b200 import _lib                      # noqa: E402
from gaussiansplattingmlx_b200.context import Context           # noqa: E402
from gaussiansplattingmlx_b200.scene import make_workload, make_gaussians, make_cameras, make_targets  # noqa: E402


def cuda_ms(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_train(name, n=None, views=None, steps=10, max_gaussians=0):
    wl, params, cams, targets = make_workload(name, n_override=n, views_override=views)
    ctx = Context(wl.width, wl.height, sh_degree=wl.sh_degree, max_gaussians=max_gaussians)
    ctx.trainer_init({k: torch.from_numpy(v) for k, v in params.items()})
    gc = [_lib.make_camera(c) for c in cams]
    tg = [torch.from_numpy(t).cuda() for t in targets]
    for it in range(3):
        ctx.train_step(gc, tg, it, 30000, want_loss=False)
    ms = cuda_ms(lambda i: ctx.train_step(gc, tg, 3 + i, 30000, want_loss=False), steps)
    st = ctx.stats()
    out = {"config": name, "gaussians": params["_xyz"].shape[0], "image": [wl.width, wl.height], "views_per_step": len(cams),
           "ms_per_step": ms, "steps_per_s": 1e3 / ms, "pairs_last_view": st["pairs_last_view"]}
    print(json.dumps(out), flush=True)
    return ctx, gc, tg


def bench_c4(cycles=3, iters_per_cycle=5):
    """C4: train `iters_per_cycle` iterations, then split_and_prune, `cycles` times (the reference cadence is every 100
    iterations; the cycle is shortened so that the run stays within seconds — the per-call costs are what is reported)."""
    ctx, gc, tg = bench_train("C4", steps=5, max_gaussians=8_000_000)   # maxGaussians raised to 8 M
    n_hist, dens_ms, step_ms = [ctx.trainer_count()[0]], [], []
    it = 100
    for c in range(cycles):
        step_ms.append(cuda_ms(lambda i: ctx.train_step(gc, tg, it + i, 30000, want_loss=False), iters_per_cycle))
        it += iters_per_cycle
        acc = ctx.trainer_tensors()["accum"]
        thr = float(torch.quantile(acc[:: max(1, acc.numel() // 1_000_000)], 0.97)) / max(ctx.trainer_count()[1], 1)  # densify ~3 %
        torch.cuda.synchronize(); t0 = time.perf_counter()
        info = ctx.trainer_densify(thr, 0.01, 0.005, 8_000_000, seed=c)
        torch.cuda.synchronize(); dens_ms.append((time.perf_counter() - t0) * 1e3)
        n_hist.append(info["n"])
    step_ms.append(cuda_ms(lambda i: ctx.train_step(gc, tg, it + i, 30000, want_loss=False), iters_per_cycle))
    print(json.dumps({"config": "C4-densify", "gaussians": n_hist, "densify_ms": dens_ms, "train_ms_per_step": step_ms,
                      "pairs_last_view": ctx.stats()["pairs_last_view"]}), flush=True)
    ctx.close()


def bench_c5(sizes=(100_000, 300_000, 1_000_000, 3_000_000, 10_000_000), reps=20):
    W, H = 1920, 1080
    cam = _lib.make_camera(make_cameras(W, H, 8)[0])
    for n in sizes:
        params = make_gaussians(n, 5, 3)
        ctx = Context(W, H, max_gaussians=n)
        dp = {k: torch.from_numpy(v).cuda() for k, v in params.items()}
        del params
        for _ in range(3):
            ctx.render_forward(dp, cam, want_outputs=False)
        ms = cuda_ms(lambda i: ctx.render_forward(dp, cam, want_outputs=False), reps)
        evals = ctx.last_contrib_sum()
        print(json.dumps({"config": "C5", "gaussians": n, "image": [W, H], "render_ms": ms, "fps": 1e3 / ms,
                          "pairs": ctx.stats()["pairs_last_view"], "blend_evals": evals}), flush=True)
        ctx.close()
        del dp
        torch.cuda.empty_cache()


def bench_app_default(sizes=(("C2", 300_000, 800, 800, 1), ("C3", 1_000_000, 1920, 1080, 8)), steps=6):
    """The configuration the reference APP runs (UI/TrainView.swift:160-188, Data/ColmapDataLoader.swift:495): SH degree 4
    (K = 25 coefficients) and TILE_SIZE = W/4 x H/4 - 16 tiles per image, every one covered by many 16x16 raster blocks that
    all walk the same (long) tile list - next to the same scene with the benchmark's 16x16 tiles.  Per-stage times from a
    pass with the view pipeline off."""
    for name, n, W, H, views in sizes:
        params = make_gaussians(n, 3 if name == "C3" else 2, 4)
        cams = make_cameras(W, H, views)
        targets = make_targets(W, H, views, 3 if name == "C3" else 2)
        for label, tile in (("app default: tile W/4 x H/4", (W // 4, H // 4)), ("16x16 tiles", (16, 16))):
            ctx = Context(W, H, tile_w=tile[0], tile_h=tile[1], sh_degree=4, max_gaussians=n)
            ctx.trainer_init({k: torch.from_numpy(v) for k, v in params.items()})
            gc = [_lib.make_camera(c) for c in cams]
            tg = [torch.from_numpy(t).cuda() for t in targets]
            for it in range(3):
                ctx.train_step(gc, tg, it, 30000, want_loss=False)
            ms = cuda_ms(lambda i: ctx.train_step(gc, tg, 3 + i, 30000, want_loss=False), steps)
            ctx.set_flags(_lib.GSB_FLAG_NO_OVERLAP)
            ctx.stats_reset(); ctx.enable_stage_timing(True)
            ms_serial = cuda_ms(lambda i: ctx.train_step(gc, tg, 3 + steps + i, 30000, want_loss=False), steps)
            st = ctx.stats()
            ctx.render_forward(ctx.trainer_tensors()["params"], gc[0], want_outputs=False)
            evals = ctx.last_contrib_sum()
            per_view = {k: round(v / (steps * views), 4) for k, v in st["stage_ms"].items() if st["stage_calls"][k] and k != "adam"}
            per_view["adam_per_step"] = round(st["stage_ms"]["adam"] / steps, 4)
            print(json.dumps({"config": f"{name} scene, SH degree 4 (K=25), {label}", "gaussians": n, "image": [W, H], "tile": list(tile),
                              "tiles": ctx.num_tiles, "views_per_step": views, "ms_per_step": ms, "steps_per_s": 1e3 / ms,
                              "ms_per_step_serialized": ms_serial, "pairs_last_view": st["pairs_last_view"], "blend_evals_view0": evals,
                              "stage_ms_per_view": per_view}), flush=True)
            ctx.close()
            del ctx, tg
            torch.cuda.empty_cache()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="C2,C4,C5")
    ap.add_argument("--max-n", type=int, default=10_000_000)
    a = ap.parse_args()
    sel = a.only.split(",")
    if "C2" in sel:
        bench_train("C2")[0].close()
    if "C4" in sel:
        bench_c4()
    if "APP" in sel:
        bench_app_default()
    if "C5" in sel:
        bench_c5(tuple(n for n in (100_000, 300_000, 1_000_000, 3_000_000, 10_000_000) if n <= a.max_n))

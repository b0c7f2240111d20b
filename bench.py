#!/usr/bin/env python3
"""bench.py — fwd+bwd train steps/s on BASELINE.json's headline configuration.

Workload (config C3 of BASELINE.json / SURVEY.md §8d): 1 M synthetic Gaussians, SH degree 3,
1920x1080, a batch of 8 views per step, loss 0.8*L1 + 0.2*(1-SSIM), Adam on all six tensors.
A *step* = for each view: activations → projection → tile binning + sort → raster → loss →
raster backward → projection backward (gradients accumulated), then [DP all-reduce] and Adam.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo (libgsb.so)
    python bench.py --impl reference ...                            # the reference's kernels on host cores

N > 1 is launched by torchrun (one rank per GPU).  Default = STRONG scaling, BASELINE config 3 as written (SURVEY.md
8e): the batch of 8 views is split round-robin over the ranks, every rank renders its views into its own gradient
block, and the data-parallel exchange (gradient sum -> Adam on the owned slice -> parameters into every replica) runs
as kernels over NVLink peer memory synchronised by flags (gsb_trainer_step_peers; GSB_DP selects the other variants).
The same run also measures WEAK scaling (8 views per GPU per step, global batch 8 N views; `value` counts 8-view
steps: all ranks' views / 8 / time) and reports it under the key "weak".  The set-up runs the data-parallel
correctness check of tools/dp_check.py (key "dp_check").  Prints ONE JSON line on rank 0 (library chatter on stdout
is sent to stderr).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "fwd+bwd train steps/s (1M Gaussians, 1080p, batch of 8 views)"
UNIT = "steps/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": float(d["hbm_gbs"]), "sm_max_mhz": float(d.get("sm_max_mhz", 1965.0)), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------
# clocks (nvidia-smi sampled DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own kernels (oracle/_ref) or the C port, on the host cores
# ------------------------------------------------------------------------------------------------
def host_threads() -> int:
    """Make the CPU arm use every host thread and return the count OpenMP will actually use.  torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would time the reference kernels on one core: the variable is overridden
    before the oracle libraries load, and libgomp is told directly as well (it may already be initialised)."""
    import ctypes
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    want = int(os.environ.get("GSB_REF_THREADS", cores))
    os.environ["OMP_NUM_THREADS"] = str(want)
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(want)
        return int(gomp.omp_get_max_threads())
    except OSError:
        return want


def cpu_view_sample(wl, params, cams, targets, view: int):
    """One view's forward + loss + backward through the CPU oracle; returns (seconds, kind, M)."""
    from oracle import pipeline as pl
    from oracle.api import Port, Ref
    o = Ref() if Ref.available() else Port()
    t0 = time.perf_counter()
    fr, lo, bw = pl.loss_and_grads(o, params, cams[view], targets[view], wl.sh_degree, 0.2, wl.tile, wl.tile)
    dt = time.perf_counter() - t0
    return dt, o.kind, int(fr["bins"]["M"]), bw["grads"], float(lo["loss"])


def cpu_adam_sample(params, grads):
    from oracle.api import Port
    from oracle import pipeline as pl
    o = Port()
    p = {k: v.copy() for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in p.items()}
    v_ = {k: np.zeros_like(v) for k, v in p.items()}
    acc = np.zeros(p["_xyz"].shape[0], np.float32)
    lrs = pl.learning_rates(0, 30000)
    t0 = time.perf_counter()
    o.accum_grad_norm(grads["_xyz"], acc)
    for i, k in enumerate(pl.PARAM_ORDER):
        o.adam(p[k], np.ascontiguousarray(grads[k].reshape(p[k].shape)), m[k], v_[k], lrs[i])
    return time.perf_counter() - t0


def cpu_baseline(wl, params, cams, targets, views: int):
    cores = host_threads()
    dt, kind, M, grads, _ = cpu_view_sample(wl, params, cams, targets, 0)
    t_adam = cpu_adam_sample(params, grads)
    step_s = views * dt + t_adam
    return {"value": 1.0 / step_s, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"1 of {views} views (fwd+L1/SSIM loss+bwd, {dt:.2f} s, M={M}) x{views} + Adam ({t_adam:.3f} s), "
                      f"OpenMP on {cores} host threads",
            "extrapolated": f"x{views} from one measured view",
            "seconds_per_view": dt}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from gaussiansplattingmlx_b200.scene import make_workload
    wl, params, cams, targets = make_workload(args.workload, n_override=args.n, views_override=args.views)
    views = len(cams)
    cores = host_threads()
    log(f"[reference arm] OpenMP threads: {cores} (OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')})")
    budget = float(os.environ.get("GSB_REF_BUDGET_S", "240"))
    t_start = time.perf_counter()
    times, kind, grads = [], "port", None
    done_warm = 0
    for i in range(args.warmup + args.steps):
        v = i % views
        dt, kind, M, grads, _ = cpu_view_sample(wl, params, cams, targets, v)
        if i < args.warmup and (time.perf_counter() - t_start) < budget * 0.4:
            done_warm += 1
            continue
        times.append(dt)
        if time.perf_counter() - t_start > budget:
            break
    t_adam = cpu_adam_sample(params, grads)
    t_view = float(np.mean(times))
    step_s = views * t_view + t_adam
    value = 1.0 / step_s
    sample = (f"each step = 1 of {views} views of the workload (fwd+L1/SSIM loss+bwd) through the reference's shipped kernels "
              f"compiled for the host ({kind}); step time = {views} x mean view time ({t_view:.2f} s) + Adam ({t_adam:.3f} s); "
              f"{len(times)} timed samples (requested {args.steps}, wall budget {budget:.0f} s), OpenMP on {cores} threads")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
           "warmup": done_warm, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": workload_config(wl, params, views, world=1),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           # every "step" of this arm is ONE measured view; the step time is views x the mean view time + Adam
           "extrapolated": f"x{views} from {len(times)} single-view samples (one view = {t_view:.2f} s on {cores} threads)",
           "timed_samples": len(times), "seconds_per_view": t_view,
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def workload_config(wl, params, views, world, global_views=None):
    n = params["_xyz"].shape[0]
    global_views = global_views or views
    return {"workload": f"{wl.name}: {n} Gaussians, SH deg {wl.sh_degree}, {wl.width}x{wl.height}, batch of {views} views per step"
                        f"{' per GPU' if global_views != views else ''}, L1+SSIM loss, Adam; seed {wl.seed} (SURVEY.md 8d generator)",
            "gaussians": n, "width": wl.width, "height": wl.height, "views_per_step": views, "global_views_per_step": global_views,
            "tile": wl.tile,
            "parallelism": f"view-parallel dp{world}" if world > 1 else "single GPU",
            "state_policy": "every timed region starts from the freshly initialised model (W warm-up + K timed steps of the same "
                            "trajectory): training on random targets inflates the Gaussians, so later steps are heavier",
            "l2_policy": "inputs larger than L2 (per step: 236 MB params + 236 MB grads + 472 MB Adam state + per-view "
                         "key/record streams > 126 MB L2); no explicit flush"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gsb(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from gaussiansplattingmlx_b200 import _lib
    from gaussiansplattingmlx_b200.context import Context
    from gaussiansplattingmlx_b200.scene import make_cameras, make_targets, make_workload

    # keep stdout to the one JSON line: NCCL prints its version banner there
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl, params, cams8, targets8 = make_workload(args.workload, n_override=args.n, views_override=args.views)
    per_step = len(cams8)                                  # views of one "step" of the metric (8)
    n = params["_xyz"].shape[0]
    base_flags = int(os.environ.get("GSB_FLAGS", "0"))   # debugging: 1 = CUB sort baseline, 2 = no view pipeline
    ctx = Context(wl.width, wl.height, tile_w=wl.tile, tile_h=wl.tile, sh_degree=wl.sh_degree, max_gaussians=n, device=local_rank,
                  flags=base_flags)
    host_params = {k: torch.from_numpy(v) for k, v in params.items()}
    ctx.trainer_init(host_params)
    total_iters = 30000

    # ---- data-parallel step variant (N > 1)
    #   fused_mc  (default) gsb_trainer_step_peers: flags in peer memory, chunked projection backward overlapping the exchange,
    #             NVLS multimem exchange kernel (torch symmetric memory); falls back to `fused` without multicast
    #   fused     the same with plain peer loads / stores
    #   peers / multicast   barrier + one exchange kernel + barrier (two 4-byte NCCL all-reduces per step)
    #   nccl      NCCL all-reduce of the gradient block + Adam on every replica
    dp = {"mode": "single", "check": None}
    vp_dp = None
    if world > 1:
        from gaussiansplattingmlx_b200.dp import ViewParallel
        vp_dp = ViewParallel(rank, world)
        if not args.no_dp_check:
            from tools.dp_check import run_dp_check
            dp["check"] = run_dp_check(rank, world, local_rank, n=20001, W=256, H=160)
            log(f"[rank {rank}] dp_check: {json.dumps(dp['check'])}")
        want = os.environ.get("GSB_DP", "fused_mc")   # the NVLS exchange where the box has multicast, else peer loads / stores
        dp["mode"] = "nccl"
        if want != "nccl" and vp_dp.enable_peers(ctx):
            dp["mode"] = "peers"
            if want in ("fused", "fused_mc"):
                dp["mode"] = "fused"
            if want in ("multicast", "fused_mc") and vp_dp.enable_multicast(ctx):
                dp["mode"] = "fused_mc" if want == "fused_mc" else "multicast"
            tune = [int(os.environ.get(k, "0")) for k in ("GSB_PEER_CHUNKS", "GSB_PEER_BLOCKS", "GSB_MC_BLOCKS")]
            if any(tune):
                ctx.trainer_peers_tune(*tune)
        log(f"[rank {rank}] data-parallel step: {dp['mode']}")

    def reattach():
        if dp["mode"] in ("multicast", "fused_mc"):
            assert vp_dp.enable_multicast(ctx)      # trainer_init moved the parameters back into the slab: re-attach

    loss_slots = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_events = [torch.cuda.Event(), torch.cuda.Event()]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def measure(scaling: str, full: bool):
        """One scaling mode: builds its views and runs the timed regions.  full = also the per-kernel (serialised) pass."""
        weak = scaling == "weak" and world > 1
        if weak:   # ring of 8 N cameras, rank r takes views r, r + N, ...; targets drawn in view order from the same seed
            views = per_step * world
            cams = make_cameras(wl.width, wl.height, views)
            my_views = [v for v in range(views) if v % world == rank]
            all_t = make_targets(wl.width, wl.height, views, wl.seed)
            targets = {v: all_t[v] for v in my_views}
            del all_t
        else:
            views = per_step
            cams = cams8
            my_views = [v for v in range(views) if v % world == rank]
            targets = {v: targets8[v] for v in my_views}
        log(f"[rank {rank}] {scaling} scaling: N={n}, {wl.width}x{wl.height}, views {my_views} of {views}")
        gcams = [_lib.make_camera(cams[v]) for v in my_views]
        host_targets = [torch.from_numpy(targets[v]).pin_memory() for v in my_views]
        dev_targets = [t.to(dev, non_blocking=True) for t in host_targets]
        gscale = 1.0 / views
        grad_block = {"t": ctx.trainer_grad_block() if world > 1 else None}
        e2e_state = {"pending": None, "losses": []}

        def drain_loss():
            if e2e_state["pending"] is not None:
                s_ = e2e_state["pending"]
                loss_events[s_].synchronize()
                e2e_state["losses"].append(float(loss_slots[s_][0]))
                e2e_state["pending"] = None

        # end-to-end mode: pinned host targets go H2D every step and every step's loss comes back D2H into a pinned slot
        # (asynchronously, GSB_FLAG_ASYNC_LOSS); the value of step i is read after step i+1 has been enqueued, so the
        # view pipeline never drains.  Every loss is read inside the timed region.
        def step(it, host: bool, want_loss: bool):
            tg = host_targets if host else dev_targets
            slot = it & 1
            lo = loss_slots[slot] if (want_loss and my_views) else None
            if dp["mode"] in ("fused", "fused_mc"):
                vp_dp.fused_step(ctx, gcams, tg, gscale, it, total_iters, loss_out=lo)
            else:
                if my_views:
                    if lo is not None:
                        ctx.trainer_accumulate(gcams, tg, zero_grads=True, grad_scale=gscale, loss_out=lo)
                    else:
                        ctx.trainer_accumulate(gcams, tg, zero_grads=True, grad_scale=gscale, want_loss=False)
                elif world > 1:
                    grad_block["t"].zero_()
                if dp["mode"] == "multicast":
                    vp_dp.multicast_step(ctx, it, total_iters)
                elif dp["mode"] == "peers":
                    vp_dp.peer_step(ctx, it, total_iters)
                else:
                    if world > 1:
                        dist.all_reduce(grad_block["t"])
                    ctx.trainer_apply(it, total_iters, reset_state=False)
            if lo is not None:
                loss_events[slot].record()
                drain_loss()                       # the PREVIOUS step's loss: its copy finished long ago
                e2e_state["pending"] = slot

        def timed(host: bool, want_loss: bool, k: int, it0: int):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(k):
                step(it0 + i, host, want_loss)
            if want_loss:
                drain_loss()                       # the last step's loss is read inside the timed region too
            if dp["mode"] in ("fused", "fused_mc"):
                ctx.synchronize()                  # the other replicas' last parameter stores have landed (flags), inside the timing
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms

        # Training on synthetic random targets moves the Gaussians (pairs and blend evaluations per view grow step after
        # step), so every timed region starts from the SAME state: parameters re-initialised, optimiser state zeroed, then
        # the W warm-up steps, then the K timed steps - region to region the work is identical.
        def fresh_region(host: bool, want_loss: bool):
            barrier()
            ctx.trainer_init(host_params)
            reattach()
            if world > 1:
                grad_block["t"] = ctx.trainer_grad_block()
            for i in range(args.warmup):
                step(i, host, want_loss)
            if want_loss:
                drain_loss()
            barrier()
            return args.warmup

        # ---- set-up steps (size the intersection buffers, fault in every allocation, let clocks settle)
        ctx.set_flags(base_flags)
        fresh_region(False, False)
        for i in range(4):
            step(args.warmup + i, False, False)
        barrier()
        out = {"views": views, "my_views": my_views}

        if full:
            # blend evaluations per step on this rank (for the raster rooflines), from lastContrib of each view, measured
            # on the state in the middle of a timed region
            fresh_region(False, False)
            for i in range(args.steps // 2):
                step(args.warmup + i, False, False)
            barrier()
            evals = pairs = l1_pairs = 0
            for cam in gcams:
                ctx.render_forward(ctx.trainer_tensors()["params"], cam, want_outputs=False)
                evals += ctx.last_contrib_sum()
                pairs += ctx.stats()["pairs_last_view"]
                l1_pairs += ctx.stats()["sb_pairs_last_view"]
            out.update(evals=evals, pairs=pairs, l1_pairs=l1_pairs)

        # ---- timed region 1: inputs resident in HBM (the product configuration: view pipeline on)
        it = fresh_region(False, False)
        ctx.stats_reset()
        clocks = ClockSampler(local_rank)
        if rank == 0 and full:
            clocks.start()
        out["ms_dev"] = timed(False, False, args.steps, it)
        out["launches"] = ctx.stats()["kernel_launches"]
        if args.debug_overlap and rank == 0 and full:   # how much each stage stretches when the streams share the SMs
            it = fresh_region(False, False)
            ctx.stats_reset(); ctx.enable_stage_timing(True)
            ms_dbg = timed(False, False, args.steps, it)
            sd = ctx.stats(); ctx.enable_stage_timing(False)
            log(f"[overlap on] {ms_dbg / args.steps:.3f} ms/step; per-launch stage ms: " +
                ", ".join(f"{k} {sd['stage_ms'][k] / max(sd['stage_calls'][k], 1):.3f}" for k in sd["stage_ms"] if sd["stage_calls"][k]))

        if full:
            # ---- per-kernel durations: same steps with the view pipeline OFF (every kernel alone on the work stream,
            #      bracketed by CUDA events on that stream), so a kernel's time is not inflated by the kernels of the
            #      next view that overlap it in region 1
            ctx.set_flags(base_flags | _lib.GSB_FLAG_NO_OVERLAP)
            it = fresh_region(False, False)
            ctx.stats_reset()
            ctx.enable_stage_timing(True)
            out["ms_serial"] = timed(False, False, args.steps, it)
            out["stage"] = ctx.stats()
            ctx.enable_stage_timing(False)
            ctx.set_flags(base_flags)

        # ---- timed region 2: end to end through the public API (pinned host targets H2D + loss D2H every step)
        ctx.set_flags(base_flags | _lib.GSB_FLAG_ASYNC_LOSS)
        it = fresh_region(True, True)
        e2e_state["losses"].clear()
        out["ms_e2e"] = timed(True, True, args.steps, it)
        if my_views:
            assert len(e2e_state["losses"]) >= args.steps and all(np.isfinite(l) for l in e2e_state["losses"])
        ctx.set_flags(base_flags)
        out["clk"] = clocks.stop() if (rank == 0 and full) else None
        if dp["mode"] != "nccl" and world > 1:
            ctx.trainer_peers_check()              # no bounded wait of the flag protocol ran out
        return out

    primary = args.scaling
    res = {primary: measure(primary, True)}
    if world > 1 and not args.single_scaling:
        other = "weak" if primary == "strong" else "strong"
        res[other] = measure(other, False)
    if world > 1 and dp["mode"] != "nccl":
        vp_dp.disable_peers(ctx)   # unmap the replicas' slabs on every rank before any context is destroyed

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    K = args.steps
    m = res[primary]
    views, my_views = m["views"], m["my_views"]
    unit_steps = views / per_step                        # 8-view steps processed per global step (N for weak scaling, else 1)
    value = K * unit_steps / (m["ms_dev"] * 1e-3)
    e2e_value = K * unit_steps / (m["ms_e2e"] * 1e-3)
    ms_serial, st = m["ms_serial"], m["stage"]
    img_bytes = wl.width * wl.height * 3 * 4
    peaks = measured_peaks()
    fp32_peak = 148 * 128 * 2 * peaks["sm_max_mhz"] * 1e6 / 1e12   # TFLOP/s, non-tensor FP32 pipe at max clock

    # per-stage averages on rank 0 (CUDA-event pairs recorded on the work stream during the serialised pass)
    nv = max(len(my_views), 1)
    P = wl.width * wl.height
    passes = ctx.tile_list_info()["sort_passes"]               # onesweep passes over the (superblock id, Gaussian) pairs
    L1 = m["l1_pairs"] / nv
    M = m["pairs"] / nv
    E = m["evals"] / nv
    own = 1.0 / world if dp["mode"] != "nccl" else 1.0        # fraction of Adam's state a replica touches per step
    alg = {  # algorithmic bytes / flops per LAUNCH, per view unless noted
        "project_fwd": ("hbm", 284.0 * n), "project_bwd": ("hbm", (516.0 + 236.0) * n), "scan": ("hbm", 8.0 * n),   # bwd: +236 B/G gradient read (accumulated RMW across views)
        "depth_sort": ("hbm", n * (4.0 + 16.0 * 4)),          # 32-bit key + index, 4 onesweep passes (r+w) + histogram read
        "keygen": ("hbm", 16.0 * n + 8.0 * L1),             # perm, offset, rect in; (superblock id, index) out
        "sort": ("hbm", L1 * (4.0 + 16.0 * passes)),         # 8-byte pairs, `passes` onesweep passes on the superblock id
        "tile_lists": ("hbm", 2 * 12.0 * L1 + 4.0 * L1 + 4.0 * M + 16.0 * ctx.num_tiles),   # 2 walks (index + rect), ranges, M list entries out
        "raster_fwd": ("fp32", 27.0 * E), "raster_bwd": ("fp32", 80.0 * E),
        "loss": ("fp32", (225.0 + 170.0) * P * 3), "adam": ("hbm", (28.0 * 59 * n + 12.0 * n) * own),
    }
    per_step_stages = {"adam"}                               # launched once (or once per Gaussian chunk) per STEP, not per view
    kernels = {}
    for name, (bound, work) in alg.items():
        calls = st["stage_calls"].get(name, 0)
        if not calls:
            continue
        ms = st["stage_ms"][name] / (K if name in per_step_stages else K * nv)   # per step / per view (a stage may be several launches)
        if bound == "hbm":
            ach = work / (ms * 1e-3) / 1e9
            kernels[name] = {"bound": "hbm", "ms": ms, "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": ach / peaks["hbm_gbs"], "share_of_step": st["stage_ms"][name] / ms_serial}
        else:
            ach = work / (ms * 1e-3) / 1e12
            kernels[name] = {"bound": "fp32", "ms": ms, "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                             "frac": ach / fp32_peak, "share_of_step": st["stage_ms"][name] / ms_serial}
    if world > 1 and "adam" in kernels:
        kernels["adam"]["note"] = ("data-parallel exchange + Adam on the owned slice; NVLink-bound, the HBM fraction is not its roofline "
                                   "(per GPU and direction it moves 2 (W-1)/W x 236 MB over the links)")
    dominant = max(kernels, key=lambda k: kernels[k]["share_of_step"]) if kernels else None
    roof = None
    if dominant:
        d = kernels[dominant]
        # DRAM bytes per launch of that kernel from the committed ncu --set full capture (tools/ncu_summary.py --traffic)
        traffic, traffic_src = None, None
        tfile = ROOT / "profiles" / "traffic.json"
        if tfile.exists():
            t = json.loads(tfile.read_text()).get("k_" + dominant)
            if t:
                traffic, traffic_src = t["dram_bytes_per_launch"], f"profiles/traffic.json ({t['source']})"
        roof = {"kernel": dominant, "bound": d["bound"], "achieved": d["achieved"], "peak": d["peak"], "unit": d["unit"],
                "frac": d["frac"], "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_src,
                "peak_source": (f"{peaks['source']} MEASURED_PEAKS.json hbm_gbs" if d["bound"] == "hbm" else
                                f"non-tensor FP32 pipe: 148 SM x 128 lanes x 2 flop x sm_max_mhz ({peaks['source']}); "
                                "no measured FP32 figure exists in MEASURED_PEAKS.json; tools/microbench/f32x2_rate.cu "
                                "measured 71.5 TFLOP/s (123 of 128 fma/clk/SM) on this pool (profiles/f32x2_rate.txt)"),
                "peak_measured": 71.54 if d["bound"] == "fp32" else None,
                "frac_of_measured": (d["achieved"] / 71.54) if d["bound"] == "fp32" else None,
                "units_per_launch": {"pairs_M": M, "superblock_pairs_L1": L1, "blend_evals_E": E, "gaussians": n, "pixels": P},
                "timing": "CUDA-event pairs around every launch on the library's work stream, averaged over a second pass of "
                          "the same steps with the view pipeline disabled (kernels back to back); share_of_step is relative to "
                          "that pass (ms_per_step_serialized)"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(wl, params, cams8, targets8, per_step)
        except Exception as ex:  # the oracle is a checker, never a dependency of the GPU arm
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    dp_text = {"fused": "gsb_trainer_step_peers: no NCCL call on the step; replicas meet in flags kept in NVLink peer memory; the last view's "
                        "projection backward runs in Gaussian chunks and the exchange kernel of chunk k (reduce the owned slice over all "
                        "replicas with peer loads + Adam + D1 + peer stores of the parameters into every replica) overlaps chunk k+1",
               "fused_mc": "gsb_trainer_step_peers with the NVLS exchange kernel: multimem.ld_reduce of the owned gradient slice through the "
                           "NVSwitch + Adam + multimem.st of the new parameters; flags in peer memory, chunks overlapping the projection backward",
               "multicast": "one kernel through the NVSwitch (symmetric memory): multimem.ld_reduce of the owned gradient slice + Adam + "
                            "multimem.st of the new parameters, between two 4-byte NCCL barriers",
               "peers": "one kernel over NVLink peer memory: reduce the gradient slices of all replicas + Adam + store the parameters "
                        "into every replica, between two 4-byte NCCL barriers",
               "nccl": "NCCL all-reduce of the 236 MB gradient block + Adam on every replica"}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
           "ms_per_step": m["ms_dev"] / K, "higher_is_better": True, "scaling": primary,
           "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": dict(workload_config(wl, params, per_step, world, views),
                                               **({"dp_step": dp_text[dp["mode"]], "dp_mode": dp["mode"]} if world > 1 else {})),
           "views_per_s": value * per_step,
           "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": m["ms_e2e"] / K,
                   "h2d_bytes_per_step": img_bytes * views, "d2h_bytes_per_step": 4 * min(world, views),
                   "api": "Context.trainer_accumulate(pinned host targets, pinned loss slot) + trainer_apply [N > 1: dp.ViewParallel.fused_step -> "
                          "gsb_trainer_step_peers]; every step's loss is copied D2H asynchronously and read on the host one step later"},
           "gpu_launches": m["launches"], "clocks": m["clk"], "roofline": roof, "roofline_kernels": kernels,
           "ms_per_step_serialized": ms_serial / K,
           "cpu_baseline": cpu}
    if world > 1:
        out["dp_check"] = dp["check"]
        for other, mo in res.items():
            if other == primary:
                continue
            us = mo["views"] / per_step
            out[other] = {"scaling": other, "global_views_per_step": mo["views"], "value": K * us / (mo["ms_dev"] * 1e-3), "unit": UNIT,
                          "ms_per_step": mo["ms_dev"] / K, "e2e_value": K * us / (mo["ms_e2e"] * 1e-3),
                          "note": "same run, same protocol (device-resident `value`, end-to-end `e2e_value`), the other scaling mode: "
                                  "weak = 8 views per GPU per step (global batch 8 N), strong = BASELINE config 3's fixed batch of 8 views"}
    os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gsb", choices=["gsb", "reference"])
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--n", type=int, default=None, help="override the Gaussian count (debugging only)")
    ap.add_argument("--views", type=int, default=None, help="override the views per step (debugging only)")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong (default) = BASELINE config 3 as written, the batch of 8 views split over the GPUs; "
                         "weak = 8 views per GPU per step.  The other mode is measured in the same run and reported under its own key")
    ap.add_argument("--single-scaling", action="store_true", help="N > 1: measure only --scaling")
    ap.add_argument("--no-dp-check", action="store_true", help="N > 1: skip the data-parallel correctness check of the set-up")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--debug-overlap", action="store_true", help="also print per-stage times measured WITH the view pipeline on")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        log(f"--gpus {args.gpus} requires torchrun (WORLD_SIZE={world}); running on 1 GPU")
    run_gsb(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Times the data-parallel exchange alone (no views) on N GPUs with the C3 trainer state (1 M Gaussians, 236 MB gradient
block): every variant of the step is run with B = 0 views on every replica, so a "step" is the gradient reset + the
synchronisation + the exchange kernels + Adam - the part of a train step the collective is responsible for.

    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/dp_exchange_bench.py > gpurun_out/exchange_N.jsonl

One JSON line per variant (rank 0): ms per step (CUDA events, max over ranks) and the implied NVLink rate per GPU and
direction (peer variants move 2 (W-1)/W x block bytes per direction, the NVLS variants (W-1)/W + 1/W... see DESIGN.md)."""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gaussiansplattingmlx_b200.context import Context          # noqa: E402
from gaussiansplattingmlx_b200.dp import ViewParallel           # noqa: E402
from gaussiansplattingmlx_b200.scene import make_gaussians      # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = int(os.environ.get("N", 1_000_000))
    iters = int(os.environ.get("ITERS", 30))
    params = {k: torch.from_numpy(v) for k, v in make_gaussians(n, 3, 3).items()}
    vp = ViewParallel(rank, world)
    ctx = Context(1920, 1080, max_gaussians=n, device=local)
    ctx.trainer_init(params)
    block_bytes = int(ctx.trainer_grad_block().numel()) * 4
    have_peers = vp.enable_peers(ctx)

    def barrier():
        torch.cuda.synchronize(dev); dist.barrier(); torch.cuda.synchronize(dev)

    def timed(fn, name, extra=None):
        for i in range(5):
            fn(i)
        ctx.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(5 + i)
        ctx.synchronize()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        if rank == 0:
            out = {"variant": name, "world": world, "gaussians": n, "block_MB": block_bytes / 1e6, "ms_per_step": ms}
            out.update(extra or {})
            print(json.dumps(out), flush=True)
        return ms

    gb = ctx.trainer_grad_block()

    def nccl_step(it):
        gb.zero_()
        dist.all_reduce(gb)
        ctx.trainer_apply(it, 30000)

    timed(nccl_step, "nccl all-reduce + Adam on every replica")
    timed(lambda it: (gb.zero_(), dist.all_reduce(gb)), "nccl all-reduce alone")
    timed(lambda it: (gb.zero_(), ctx.trainer_apply(it, 30000)), "local: gradient reset + full Adam (no exchange)")
    if have_peers:
        def peers_step(it):
            gb.zero_()
            vp.peer_step(ctx, it, 30000)
        timed(peers_step, "peers: barrier + exchange kernel + barrier")
        for chunks in (1, 2, 4, 8):
            for blocks in (148 * 2, 148 * 4):
                ctx.trainer_peers_tune(chunks, blocks, 0)
                timed(lambda it: vp.fused_step(ctx, [], [], 1.0, it, 30000), f"fused flags, peer loads/stores, {chunks} chunks, {blocks} CTAs",
                      {"chunks": chunks, "ctas": blocks})
        ctx.trainer_peers_tune(4, 0, 0)
        if vp.enable_multicast(ctx):
            gb = ctx.trainer_grad_block()

            def mc_step(it):
                gb.zero_()
                vp.multicast_step(ctx, it, 30000)
            timed(mc_step, "multicast: barrier + NVLS exchange kernel + barrier")
            for chunks in (1, 4):
                for blocks in (148, 148 * 2, 148 * 8):
                    ctx.trainer_peers_tune(chunks, 0, blocks)
                    timed(lambda it: vp.fused_step(ctx, [], [], 1.0, it, 30000), f"fused flags, NVLS multimem, {chunks} chunks, {blocks} CTAs",
                          {"chunks": chunks, "ctas": blocks})
        ctx.trainer_peers_check()
        vp.disable_peers(ctx)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

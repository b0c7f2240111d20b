#!/usr/bin/env python3
"""Work model of the rasterisers from the CPU oracle's forward (no GPU): where the blend evaluations of a view go.

For one view of a workload (default C3, view 0) this prints
  * E = sum of lastContrib (the algorithmic evaluations bench.py's roofline uses) and the overshoot of block-granular
    early exit (a warp / CTA runs until the LAST of its pixels has terminated) for several footprints;
  * how the evaluations of a 16x16 block split into the dense part (every pixel active) and the sparse tail;
  * the makespan of greedy list scheduling of the backward's work items over W one-warp workers, for whole-block items
    (by list length, by true cost) and for 256-Gaussian segments - the reason the backward was cut into segments
    (DESIGN.md section 4): a whole-block item can be longer than the ideal makespan.
TEST / ANALYSIS INFRASTRUCTURE: imports oracle/ (never the product path).
usage: python tools/analysis/raster_work_model.py [C2|C3] > profiles/<round>_work_model.txt"""
import heapq
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from gaussiansplattingmlx_b200.scene import make_workload   # noqa: E402
from oracle import api, pipeline                             # noqa: E402


def makespan(order, workers, cost):
    h = [0.0] * workers
    heapq.heapify(h)
    for t in order:
        heapq.heappush(h, heapq.heappop(h) + cost[t])
    return max(h) / (cost.sum() / workers)


def main(name="C3"):
    wl, params, cams, _ = make_workload(name, views_override=1)
    fr = pipeline.render_forward(api.Port(), params, cams[0], wl.sh_degree)
    H, W = wl.height, wl.width
    last = fr["fwd"]["lastContrib"].reshape(H, W).astype(np.int64)
    counts = np.asarray(fr["bins"]["tileCounts"]).astype(np.int64)
    E = int(last.sum())
    print(f"{name} view 0: {params['_xyz'].shape[0]} Gaussians, {W}x{H}, M = {fr['bins']['M']} pairs, E = {E} evaluations "
          f"({E / (W * H):.1f} per pixel), tile list mean/max {counts.mean():.0f}/{counts.max()}")

    def blocks(bh, bw):
        Hp, Wp = (H + bh - 1) // bh * bh, (W + bw - 1) // bw * bw
        L = np.zeros((Hp, Wp), np.int64)
        L[:H, :W] = last
        return L.reshape(Hp // bh, bh, Wp // bw, bw).transpose(0, 2, 1, 3).reshape(-1, bh * bw)

    print("\novershoot of footprint-granular early exit (footprint evaluations / E):")
    for bh, bw in ((16, 16), (8, 16), (8, 8), (4, 16), (4, 8)):
        B = blocks(bh, bw)
        print(f"  {bh:2d} x {bw:2d}: {B.max(axis=1).sum() * bh * bw / E:.3f}")

    B = blocks(16, 16)
    used, nmin = B.max(axis=1), B.min(axis=1)
    print(f"\n16x16 blocks: {len(used)}, used (max lastContrib) mean {used.mean():.0f}, p50 {np.median(used):.0f}, p95 "
          f"{np.percentile(used, 95):.0f}, max {used.max()}")
    print(f"  dense part (all 256 pixels active): {256 * nmin.sum() / E:.3f} E;  sparse tail: {256 * (used - nmin).sum() / E:.3f} E of "
          f"footprint evaluations for {(E - 256 * nmin.sum()) / E:.3f} E active ones")
    q = np.sort(B, axis=1)
    print("  footprint evaluations if a block ran only to its k-th smallest lastContrib (k of 256): " +
          ", ".join(f"k={k + 1}: {q[:, k].sum() * 256 / E:.3f}" for k in (127, 191, 223, 239, 255)))

    # ---- activity of the (block, Gaussian) iterations: what a lower-cost body for sparsely active iterations can recover
    total_it = int(used.sum())
    print(f"\nbackward / forward iterations by the number of pixels still active (one 16x16 block x one Gaussian; {total_it} in all):")
    for m in (16, 32, 64, 96, 128, 192):
        km = q[:, 256 - m - 1]                       # the (m+1)-th largest lastContrib: from there on <= m pixels are active
        print(f"  <= {m:3d} of 256 active: {(used - np.maximum(km, nmin)).clip(0).sum() / total_it:.3f}")
    masked = int((used - nmin).sum())
    print(f"  masked iterations (at least one pixel inactive): {masked / total_it:.3f}; mean active pixels in them: "
          f"{(B - nmin[:, None]).clip(0).sum() / masked:.1f}")
    # the backward's row pairs: lane = column x half, pair k = rows (2k, 2k+1) of both halves = block rows {2k, 2k+1, 8+2k, 9+2k}
    B3 = B.reshape(-1, 16, 16)
    amax = np.stack([B3[:, [2 * k, 2 * k + 1, 8 + 2 * k, 9 + 2 * k], :].reshape(len(B3), -1).max(1) for k in range(4)], 1)
    am = np.sort(amax, 1)
    frac = []
    for j in range(1, 5):                            # iterations with exactly j of the 4 pairs holding an active pixel
        hi, lo = am[:, 4 - j], (am[:, 4 - j - 1] if j < 4 else np.zeros(len(B3), np.int64))
        frac.append((np.maximum(hi, nmin) - np.maximum(lo, nmin)).clip(0).sum() / masked)
    print("  masked iterations by the number of row pairs (64 pixels each) with an active pixel, 1..4: " +
          ", ".join(f"{f:.3f}" for f in frac) + "  (skipping dead pairs: warp-uniform, but rarely applicable)")
    H8 = blocks(8, 16)
    u8 = H8.max(axis=1)
    q8 = np.sort(H8, axis=1)
    print("  forward, one warp = a 16x8 half (128 pixels): iterations with <= 32 live pixels "
          f"{(u8 - q8[:, 128 - 33]).sum() / u8.sum():.3f}, <= 64: {(u8 - q8[:, 128 - 65]).sum() / u8.sum():.3f}")

    cost = used + 10.0
    print("\nbackward, greedy list scheduling, makespan / ideal:")
    for workers in (148 * 16, 148 * 12):
        print(f"  {workers} one-warp workers, whole blocks by list length: {makespan(np.argsort(-counts, kind='stable'), workers, cost):.3f}, "
              f"by true cost (LPT): {makespan(np.argsort(-used), workers, cost):.3f}")
    for seg in (128, 256, 512):
        items = []
        for u in used:
            u = int(u)
            while u > 0:
                t = min(seg, u)
                items.append(t + 10.0)
                u -= t
        items = np.array(items)
        print(f"  {148 * 16} workers, {seg}-Gaussian segments ({len(items)} items): full segments first {makespan(np.argsort(-items), 148 * 16, items):.3f}, "
              f"work {items.sum() / cost.sum():.3f}x")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "C3")

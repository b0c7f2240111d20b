#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the REFERENCE's own kernels (oracle/_ref/libgsref.so, built by
oracle/build_ref.py from /root/reference/GaussianSplattingMlx/Slang/*_mlx.json).  Run in the
container that has /root/reference; the fixtures are committed, this script documents how they
were made:

    python oracle/build_ref.py && python tests/golden/make_golden.py
"""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from gaussiansplattingmlx_b200.scene import make_cameras, make_gaussians, make_targets  # noqa: E402
from oracle import pipeline as pl  # noqa: E402
from oracle.api import Ref  # noqa: E402

CASES = {
    # name: (N, W, H, seed, degree, view index of a 3-camera ring)
    "c1": (1000, 64, 64, 1, 3, 0),          # BASELINE.json configs[0] (TinyTests-style scene), view 0 of 1
    "deg4_ragged": (400, 72, 40, 31, 4, 1),  # app default SH degree 4, image not a multiple of the tile
}


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    o = Ref()
    for name, (n, W, H, seed, degree, view) in CASES.items():
        params = make_gaussians(n, seed, degree)
        cams = make_cameras(W, H, 1 if name == "c1" else 3)
        cam = cams[view]
        target = make_targets(W, H, 1, seed)[0]
        fr, lo, bw = pl.loss_and_grads(o, params, cam, target, degree)
        b = fr["bins"]
        out = {
            "meta": np.array([n, W, H, seed, degree, view, 1 if name == "c1" else 3], np.int64),
            "M": np.array([b["M"]], np.int64),
            "tilesTouched": b["tilesTouched"], "tileCounts": b["tileCounts"], "tileRanges": b["tileRanges"],
            "sortedKeysHigh": b["sortedKeysHigh"], "sortedKeysLow": b["sortedKeysLow"], "sortedGaussIdx": b["sortedGaussIdx"],
            "means2d": fr["proj"]["means2d"], "depths": fr["proj"]["depths"], "radii": fr["proj"]["radii"],
            "conic": fr["proj"]["conic"], "color": fr["proj"]["color"],
            "render": fr["render"], "depth": fr["depth"], "alpha": fr["alpha"], "lastContrib": fr["fwd"]["lastContrib"],
            "loss": np.array([lo["loss"], lo["l1"], lo["ssim_loss"]], np.float64),
            "ssim_map": lo["ssim"]["ssim"], "cot_render": lo["cot_render"],
            "grad_packed": bw["grad_packed"],
        }
        for k, v in bw["grads"].items():
            out["grad" + k] = v
        path = Path(__file__).resolve().parent / f"{name}.npz"
        np.savez_compressed(path, **out)
        print(name, "M =", b["M"], "loss =", lo["loss"], "->", path.name, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()

"""Seeded synthetic scenes / cameras / targets, shared by tests, bench and oracle.

The reference has no "TinyTests scene" and no dataset in-tree (loaders download at run time), so
the workloads of BASELINE.json are defined here once, deterministically from a seed:

=====  ==========  ===========  =====  ====  =====
name   Gaussians   image        views  seed  K(deg)
=====  ==========  ===========  =====  ====  =====
C1     1 000       64 x 64      1      1     16 (3)
C2     300 000     800 x 800    1      2     16 (3)
C3     1 000 000   1920 x 1080  8      3     16 (3)
C4     6 000 000   3840 x 2160  1      4     16 (3)
C5     1e5..1e7    1920 x 1080  1      5     16 (3)
=====  ==========  ===========  =====  ====  =====

Parameter tensors use the reference's own layouts and names (``GaussianModel.swift:33-55``):
``_xyz[N,3]``, ``_features_dc[N,1,3]``, ``_features_rest[N,K-1,3]``, ``_scales[N,3]`` (log),
``_rotation[N,4]`` (w,x,y,z), ``_opacity[N,1]`` (logit); all f32.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List

import numpy as np

from .camera import Camera

PARAM_NAMES = ("_xyz", "_features_dc", "_features_rest", "_scales", "_rotation", "_opacity")


@dataclass(frozen=True)
class Workload:
    name: str
    n_gaussians: int
    width: int
    height: int
    views: int
    seed: int
    sh_degree: int = 3
    tile: int = 16


WORKLOADS: Dict[str, Workload] = {
    "C1": Workload("C1", 1_000, 64, 64, 1, 1),
    "C2": Workload("C2", 300_000, 800, 800, 1, 2),
    "C3": Workload("C3", 1_000_000, 1920, 1080, 8, 3),
    "C4": Workload("C4", 6_000_000, 3840, 2160, 1, 4),
    "C5": Workload("C5", 1_000_000, 1920, 1080, 1, 5),
}


def make_gaussians(n: int, seed: int, sh_degree: int = 3) -> Dict[str, np.ndarray]:
    """Draw order is part of the definition."""
    rng = np.random.default_rng(seed)
    K = (sh_degree + 1) ** 2
    xyz = rng.uniform(-1.0, 1.0, size=(n, 3)).astype(np.float32)
    mu = np.log(0.25 * (8.0 / n) ** (1.0 / 3.0))
    log_scale = rng.normal(mu, 0.35, size=(n, 3)).astype(np.float32)
    rotation = rng.normal(0.0, 1.0, size=(n, 4)).astype(np.float32)
    opacity = rng.normal(0.0, 1.0, size=(n, 1)).astype(np.float32)
    f_dc = (rng.uniform(-1.0, 1.0, size=(n, 1, 3)) / 0.2820948 * 0.5).astype(np.float32)
    f_rest = rng.normal(0.0, 0.1, size=(n, K - 1, 3)).astype(np.float32)
    return {
        "_xyz": xyz,
        "_features_dc": f_dc,
        "_features_rest": f_rest,
        "_scales": log_scale,
        "_rotation": rotation,
        "_opacity": opacity,
    }


def make_cameras(width: int, height: int, views: int) -> List[Camera]:
    """Ring of OpenCV-convention cameras (+z forward, y down) looking at the origin."""
    cams = []
    focal = width / (2.0 * np.tan(np.deg2rad(25.0)))
    for v in range(views):
        a = 0.3 + 2.0 * np.pi * v / max(views, 1)
        eye = np.array([4.0 * np.sin(a), 0.5, -4.0 * np.cos(a)], dtype=np.float64)
        f = -eye / np.linalg.norm(eye)
        up = np.array([0.0, -1.0, 0.0])
        r = np.cross(up, f)
        r /= np.linalg.norm(r)
        d = np.cross(f, r)
        c2w = np.eye(4, dtype=np.float64)
        c2w[:3, 0] = r
        c2w[:3, 1] = d
        c2w[:3, 2] = f
        c2w[:3, 3] = eye
        cams.append(Camera(width, height, focal, focal, c2w))
    return cams


def make_targets(width: int, height: int, views: int, seed: int) -> List[np.ndarray]:
    """``target_rgb ~ U(0,1)`` f32 [H,W,3]; generator seeded with ``seed + 1``, views drawn in order."""
    rng = np.random.default_rng(seed + 1)
    return [rng.random((height, width, 3), dtype=np.float32) for _ in range(views)]


def make_workload(name_or_wl, n_override: int | None = None, views_override: int | None = None):
    wl = WORKLOADS[name_or_wl] if isinstance(name_or_wl, str) else name_or_wl
    n = n_override or wl.n_gaussians
    views = views_override or wl.views
    params = make_gaussians(n, wl.seed, wl.sh_degree)
    cams = make_cameras(wl.width, wl.height, views)
    targets = make_targets(wl.width, wl.height, views, wl.seed)
    return wl, params, cams, targets

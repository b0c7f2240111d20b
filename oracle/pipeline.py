"""Restatement of the reference's Swift glue on top of an oracle backend (``api.Port`` / ``api.Ref``).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, in order:
* ``GaussianTrainer.swift:634-716``  lossFn: activations → forwardWithCameraParams → L1 / SSIM / total
* ``GaussianRenderer.swift:823-880`` forwardWithCameraParams (K1 → render)
* ``GaussianRenderer.swift:769-821`` render (pack → slice info → K9)
* ``GaussianRenderer.swift:187-226`` raster VJP, ``:605-701`` projection VJP (cotCov2d == 0, radii/rect ignored)
* ``GaussianTrainer.swift:555-625``  SSIM custom function
* ``GaussianTrainer.swift:941-948,1060-1086`` Adam loop, ``GaussianModel.swift:56-65`` learning rates
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from .api import ssim_window

f32 = np.float32
PARAM_ORDER = ("_xyz", "_features_dc", "_features_rest", "_scales", "_rotation", "_opacity")


def learning_rates(current: int, total: int) -> List[float]:
    """``GaussianModel.swift:56-65`` (f32 arithmetic)."""
    return [float(f32(0.00016) * max(f32(1.0) - f32(current) / f32(total), f32(0.01))),
            0.0025, float(f32(0.0025) / f32(20)), 0.005, 0.001, 0.025]


def render_forward(o, params, cam, degree, tileW=16, tileH=16, white_bg=False):
    W, H = cam.imageWidth, cam.imageHeight
    act = o.activate_fwd(params)
    proj = o.project_fwd(act, cam, degree)
    packed = o.pack(proj, act["opacity"])
    bins = o.bin(proj, W, H, tileW, tileH)
    fwd = o.raster_fwd(packed, bins, W, H, tileW, tileH, white_bg)
    return {"act": act, "proj": proj, "packed": packed, "bins": bins, "fwd": fwd,
            "render": fwd["color"].reshape(H, W, 3), "depth": fwd["depth"].reshape(H, W, 1),
            "alpha": fwd["alpha"].reshape(H, W, 1), "visibility_filter": proj["radii"] > 0, "radii": proj["radii"]}


def loss_forward_backward(o, render, target, lambda_dssim=0.2):
    """total = (1-l)*mean|I-T| + l*(1-mean(ssim_map)); returns loss parts and d total / d render."""
    H, W, Cc = render.shape
    n = H * W * Cc
    diff = render - target
    l1 = np.abs(diff).mean(dtype=np.float64)
    _, window = ssim_window(11, 1.5)
    s = o.ssim_fwd(render, target, window)
    ssim_loss = 1.0 - s["ssim"].mean(dtype=np.float64)
    total = (1.0 - lambda_dssim) * l1 + lambda_dssim * ssim_loss
    up = np.full(n, f32(-lambda_dssim / n), dtype=f32)
    g_ssim, _ = o.ssim_bwd(up, render, target, window, s)
    g_l1 = (np.sign(diff) * f32((1.0 - lambda_dssim) / n)).astype(f32)
    return {"loss": float(total), "l1": float(l1), "ssim_loss": float(ssim_loss), "ssim": s,
            "cot_render": (g_l1 + g_ssim).astype(f32)}


def depth_loss_forward_backward(depth, target_depth, depth_mask, lambda_depth):
    """Depth supervision of lossFn: ``GaussianTrainer.swift:492`` (depthMask = alpha > 0.5, passed in here as a bool
    array), ``:693-699`` (sum(|depth[...,0] - trainDepth| * mask) / max(sum(mask), 1e-6)), ``:710-714`` (weight
    lambdaDepth).  Returns the weighted loss term and d total / d depth [H,W,1] (MLX abs: derivative 0 at 0)."""
    m = np.asarray(depth_mask).astype(f32)
    diff = depth[..., 0].astype(f32) - np.asarray(target_depth, f32)
    weight = max(f32(m.sum(dtype=np.float64)), f32(1e-6))
    term = float((np.abs(diff) * m).sum(dtype=np.float64) / float(weight))
    cot = (f32(lambda_depth) * m * np.sign(diff) / weight).astype(f32)
    return {"depth_loss": term, "weighted": float(lambda_depth) * term, "cot_depth": cot[..., None]}


def backward(o, params, cam, degree, fr, cot_render, cot_depth=None, cot_alpha=None, tileW=16, tileH=16,
             white_bg=False):
    W, H = cam.imageWidth, cam.imageHeight
    P = W * H
    cot = {"color": cot_render.reshape(P, 3),
           "depth": np.zeros((P, 1), f32) if cot_depth is None else cot_depth.reshape(P, 1),
           "alpha": np.zeros((P, 1), f32) if cot_alpha is None else cot_alpha.reshape(P, 1)}
    g64 = o.raster_bwd(fr["packed"], fr["bins"], W, H, tileW, tileH, white_bg, cot, fr["fwd"])
    gp = g64.astype(f32)
    n = gp.shape[0]
    cotp = {"means2d": gp[:, 0:2], "conic": gp[:, 2:6], "color": gp[:, 6:9], "depths": gp[:, 10],
            "cov2d": np.zeros((n, 4), f32)}
    gproj = o.project_bwd(fr["act"], cam, degree, cotp)
    g_act = {"means3d": gproj["means3d"], "shs": gproj["shs"], "scales": gproj["scales"],
             "rotations": gproj["rotations"], "opacity": np.ascontiguousarray(gp[:, 9:10])}
    grads = o.activate_bwd(params, g_act)
    return {"grads": grads, "grad_packed": gp, "grad_packed64": g64, "g_act": g_act,
            "grad_camera_center": gproj["cameraCenterPoint"].sum(axis=0, dtype=np.float64)}


def loss_and_grads(o, params, cam, target, degree, lambda_dssim=0.2, tileW=16, tileH=16, white_bg=False,
                   target_depth=None, depth_mask=None, lambda_depth=0.0):
    fr = render_forward(o, params, cam, degree, tileW, tileH, white_bg)
    lo = loss_forward_backward(o, fr["render"], target, lambda_dssim)
    cot_depth = None
    if target_depth is not None:
        dl = depth_loss_forward_backward(fr["depth"], target_depth, depth_mask, lambda_depth)
        lo = dict(lo, loss=lo["loss"] + dl["weighted"], depth_loss=dl["depth_loss"], cot_depth=dl["cot_depth"])
        cot_depth = dl["cot_depth"]
    bw = backward(o, params, cam, degree, fr, lo["cot_render"], cot_depth=cot_depth, tileW=tileW, tileH=tileH, white_bg=white_bg)
    return fr, lo, bw


def train_steps(o, params, cams, targets, degree, iterations, total_iterations, lambda_dssim=0.2, view_order=None):
    """Batch of views per step: L = (1/B) sum_v L_v (B = 1 reproduces the reference loop)."""
    params = {k: np.array(v, dtype=f32, copy=True) for k, v in params.items()}
    m = {k: np.zeros_like(v) for k, v in params.items()}
    v_ = {k: np.zeros_like(v) for k, v in params.items()}
    accum = np.zeros(params["_xyz"].shape[0], f32)
    losses = []
    B = len(cams)
    for it in range(iterations):
        gsum = {k: np.zeros(v.shape, np.float64) for k, v in params.items()}
        loss = 0.0
        for b in range(B):
            _, lo, bw = loss_and_grads(o, params, cams[b], targets[b], degree, lambda_dssim)
            loss += lo["loss"] / B
            for k in gsum:
                gsum[k] += bw["grads"][k].reshape(gsum[k].shape).astype(np.float64) / B
        grads = {k: g.astype(f32) for k, g in gsum.items()}
        o.accum_grad_norm(grads["_xyz"], accum)
        lrs = learning_rates(it, total_iterations)
        for i, k in enumerate(PARAM_ORDER):
            o.adam(params[k], grads[k], m[k], v_[k], lrs[i])
        losses.append(loss)
    return params, m, v_, accum, losses


# densification defaults (Trainer/GaussianTrainer.swift:293-300)
DENSIFY_DEFAULTS = dict(gradientThreshold=0.0002, maxScale=0.01, minOpacity=0.005, densifyFromIter=500, densifyUntilIter=15000,
                        maxGaussians=1_000_000)


def split_and_prune(o, params, accum, denom, iteration, base_noise, **kw):
    """Restatement of ``split_and_prune`` (Trainer/GaussianTrainer.swift:766-908).

    ``base_noise[>=totalOutput,3]`` stands for ``MLXRandom.normal([totalOutput,3])`` (:881), which cannot be
    reproduced outside MLX.  Returns ``(new_params | None, info)``; ``None`` = the model is unchanged (guards
    :767, :777-780, :820-838).  The caller resets the gradient accumulation whenever this function is entered
    past the iteration guard, like the reference does on every exit path.
    """
    cfg = dict(DENSIFY_DEFAULTS); cfg.update(kw)
    info = {"ran": False}
    if not (cfg["densifyFromIter"] <= iteration <= cfg["densifyUntilIter"]):
        return None, info
    n = params["_xyz"].shape[0]
    info["ran"] = True
    if n == 0:
        return None, info
    allow = n < cfg["maxGaussians"]
    actions, counts = o.classify_gaussians(accum, float(denom), params["_scales"], params["_opacity"], cfg["gradientThreshold"],
                                           cfg["maxScale"], cfg["minOpacity"], allow)
    inclusive = np.cumsum(counts, dtype=np.int64).astype(np.int32)
    offsets = inclusive - counts
    total = int(inclusive[-1])
    num_split, num_clone, num_prune = int((actions == 1).sum()), int((actions == 2).sum()), int((actions == 3).sum())
    info.update(keep=n - num_split - num_clone - num_prune, split=num_split, clone=num_clone, prune=num_prune, total=total,
                actions=actions, counts=counts, offsets=offsets)
    if total == 0 or (num_split == 0 and num_clone == 0 and num_prune == 0):
        return None, info
    gather, mode = o.build_densify_output_map(actions, offsets, total)
    info.update(gather=gather, noise_mode=mode)
    new = o.densify_apply(params, gather, mode, np.ascontiguousarray(base_noise[:total], np.float32))
    return new, info

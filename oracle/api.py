"""numpy-level interface over the two CPU oracles.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``Port()``  — oracle/gsb_oracle.c (plain-C restatement, always available).
``Ref()``   — oracle/_ref/libgsref.so (the reference's own shipped kernels compiled for CPU).

Both expose the same stage functions so tests can diff them, and ``pipeline.py`` restates the Swift
glue (packing, slice pipeline order, loss, VJP wiring, Adam loop) once on top of either.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path
from typing import Dict

import numpy as np

HERE = Path(__file__).resolve().parent
PORT_SO = HERE / "libgsoracle.so"
REF_SO = HERE / "_ref" / "libgsref.so"

f32 = np.float32
u32 = np.uint32


def build_port(force: bool = False) -> Path:
    src = HERE / "gsb_oracle.c"
    if force or not PORT_SO.exists() or PORT_SO.stat().st_mtime < src.stat().st_mtime:
        cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
               "-fvisibility=hidden", str(src), "-o", str(PORT_SO), "-lm"]
        subprocess.run(cmd, check=True)
    return PORT_SO


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be contiguous"
    return a.ctypes.data_as(C.c_void_p)


def _f(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=f32)


def tile_bit_count(num_tiles: int) -> int:
    """``GaussianRenderer.swift:246-255`` bitWidthForExclusiveUpperBound."""
    if num_tiles <= 1:
        return 1
    v, bits = num_tiles - 1, 0
    while v > 0:
        bits += 1
        v >>= 1
    return bits


def ssim_window(K: int = 11, sigma: float = 1.5):
    """``LossUtil.swift:47-54`` + ``GaussianTrainer.swift:308-314`` evaluated in f32."""
    center = f32(K) / f32(2.0)
    vals = np.array([np.exp(-np.power(f32(x) - center, f32(2.0), dtype=f32) / (f32(2.0) * np.power(f32(sigma), f32(2.0), dtype=f32)), dtype=f32)
                     for x in range(K)], dtype=f32)
    g = (vals / vals.sum(dtype=f32)).astype(f32)
    w2 = (g.reshape(K, 1) @ g.reshape(1, K)).astype(f32)
    return g, np.ascontiguousarray(w2.reshape(-1))


class _Base:
    kind = "?"

    # ---- stages shared through subclasses -------------------------------------------------
    def pack(self, proj: Dict[str, np.ndarray], opacity: np.ndarray) -> np.ndarray:
        """``GaussianRenderer.swift:85-99`` buildPackedGaussians → [N,11]."""
        n = proj["means2d"].shape[0]
        return np.ascontiguousarray(np.concatenate(
            [proj["means2d"], proj["conic"].reshape(n, 4), proj["color"], opacity.reshape(n, 1),
             proj["depths"].reshape(n, 1)], axis=1), dtype=f32)


class Port(_Base):
    kind = "port"

    def __init__(self):
        self.lib = C.CDLL(str(build_port()))
        self.lib.gso_exclusive_scan.restype = C.c_uint32

    def activate_fwd(self, params):
        n = params["_xyz"].shape[0]
        K = params["_features_rest"].shape[1] + 1
        shs = np.empty((n, K, 3), f32); scales = np.empty((n, 3), f32)
        rot = np.empty((n, 4), f32); op = np.empty((n, 1), f32)
        self.lib.gso_activate_fwd(n, K, _p(_f(params["_features_dc"])), _p(_f(params["_features_rest"])),
                                  _p(_f(params["_scales"])), _p(_f(params["_rotation"])), _p(_f(params["_opacity"])),
                                  _p(shs), _p(scales), _p(rot), _p(op))
        return {"means3d": _f(params["_xyz"]), "shs": shs, "scales": scales, "rotations": rot, "opacity": op}

    def activate_bwd(self, params, g_act):
        n = params["_xyz"].shape[0]
        K = params["_features_rest"].shape[1] + 1
        out = {k: np.zeros_like(_f(params[k])) for k in ("_features_dc", "_features_rest", "_scales", "_rotation", "_opacity")}
        self.lib.gso_activate_bwd(n, K, _p(_f(params["_scales"])), _p(_f(params["_rotation"])), _p(_f(params["_opacity"])),
                                  _p(_f(g_act["shs"])), _p(_f(g_act["scales"])), _p(_f(g_act["rotations"])),
                                  _p(_f(g_act["opacity"])), _p(out["_features_dc"]), _p(out["_features_rest"]),
                                  _p(out["_scales"]), _p(out["_rotation"]), _p(out["_opacity"]))
        out["_xyz"] = _f(g_act["means3d"])
        return out

    def _cam_args(self, cam):
        return (_p(cam.camera_center_f32()), _p(cam.worldViewTransform), _p(cam.projectionMatrix),
                C.c_float(cam.FoVx), C.c_float(cam.FoVy), C.c_float(cam.focalX), C.c_float(cam.focalY),
                C.c_float(cam.imageWidth), C.c_float(cam.imageHeight))

    def project_fwd(self, act, cam, degree):
        n, K = act["shs"].shape[0], act["shs"].shape[1]
        o = {"means2d": np.empty((n, 2), f32), "depths": np.empty((n,), f32), "color": np.empty((n, 3), f32),
             "cov2d": np.empty((n, 2, 2), f32), "conic": np.empty((n, 2, 2), f32), "radii": np.empty((n,), f32),
             "rectMin": np.empty((n, 2), f32), "rectMax": np.empty((n, 2), f32)}
        self.lib.gso_project_fwd(n, degree, K, _p(act["scales"]), _p(act["rotations"]), _p(act["means3d"]),
                                 _p(act["shs"]), *self._cam_args(cam), _p(o["means2d"]), _p(o["depths"]),
                                 _p(o["color"]), _p(o["cov2d"]), _p(o["conic"]), _p(o["radii"]), _p(o["rectMin"]),
                                 _p(o["rectMax"]))
        return o

    def project_bwd(self, act, cam, degree, cot):
        n, K = act["shs"].shape[0], act["shs"].shape[1]
        g = {"scales": np.zeros((n, 3), f32), "rotations": np.zeros((n, 4), f32), "means3d": np.zeros((n, 3), f32),
             "shs": np.zeros((n, K, 3), f32), "cameraCenterPoint": np.zeros((n, 3), f32)}
        self.lib.gso_project_bwd(n, degree, K, _p(act["scales"]), _p(act["rotations"]), _p(act["means3d"]),
                                 _p(act["shs"]), *self._cam_args(cam), _p(_f(cot["depths"])), _p(_f(cot["means2d"])),
                                 _p(_f(cot["cov2d"])), _p(_f(cot["color"])), _p(_f(cot["conic"])), _p(g["scales"]),
                                 _p(g["rotations"]), _p(g["means3d"]), _p(g["shs"]), _p(g["cameraCenterPoint"]))
        return g

    def bin(self, proj, W, H, tileW, tileH):
        """``GaussianRenderer.swift:333-490`` buildGlobalTileSliceInfo, CSR form."""
        n = proj["radii"].shape[0]
        gridW, gridH = (W + tileW - 1) // tileW, (H + tileH - 1) // tileH
        numTiles = gridW * gridH
        touched = np.zeros(n, u32); offsets = np.zeros(n, u32)
        rmin, rmax, radii, depths = _f(proj["rectMin"]), _f(proj["rectMax"]), _f(proj["radii"]), _f(proj["depths"])
        self.lib.gso_count_tiles(n, _p(rmin), _p(rmax), _p(radii), tileW, tileH, W, H, _p(touched))
        M = int(self.lib.gso_exclusive_scan(n, _p(touched), _p(offsets)))
        kh = np.zeros(M, u32); kl = np.zeros(M, u32); gi = np.zeros(M, u32)
        if M:
            self.lib.gso_generate_keys(n, _p(depths), _p(rmin), _p(rmax), _p(radii), _p(offsets), tileW, tileH, W, H,
                                       _p(kh), _p(kl), _p(gi))
        sh = np.zeros(M, u32); sl = np.zeros(M, u32); sv = np.zeros(M, u32)
        tb = tile_bit_count(numTiles)
        if M > 1:
            self.lib.gso_radix_sort_tile_keys(M, tb, _p(kh), _p(kl), _p(gi), _p(sh), _p(sl), _p(sv))
        else:
            sh, sl, sv = kh.copy(), kl.copy(), gi.copy()
        ranges = np.zeros((numTiles, 2), u32); counts = np.zeros(numTiles, u32)
        self.lib.gso_tile_ranges(M, numTiles, _p(sh), _p(ranges), _p(counts))
        return {"tilesTouched": touched, "offsets": offsets, "M": M, "keysHigh": kh, "keysLow": kl, "gaussIdx": gi,
                "sortedKeysHigh": sh, "sortedKeysLow": sl, "sortedGaussIdx": sv, "tileRanges": ranges,
                "tileCounts": counts, "tileBits": tb, "gridW": gridW, "gridH": gridH}

    def raster_fwd(self, packed, bins, W, H, tileW, tileH, white_bg):
        P = W * H
        o = {"color": np.empty((P, 3), f32), "depth": np.empty((P, 1), f32), "alpha": np.empty((P, 1), f32),
             "lastContrib": np.empty((P, 1), u32)}
        self.lib.gso_raster_fwd(W, H, tileW, tileH, int(white_bg), _p(packed), _p(bins["sortedGaussIdx"]),
                                _p(bins["tileRanges"]), _p(o["color"]), _p(o["depth"]), _p(o["alpha"]),
                                _p(o["lastContrib"]))
        return o

    def raster_bwd(self, packed, bins, W, H, tileW, tileH, white_bg, cot, fwd):
        g64 = np.zeros(packed.shape, np.float64)
        self.lib.gso_raster_bwd(W, H, tileW, tileH, int(white_bg), _p(packed), _p(bins["sortedGaussIdx"]),
                                _p(bins["tileRanges"]), _p(_f(cot["color"])), _p(_f(cot["depth"])),
                                _p(_f(cot["alpha"])), _p(fwd["color"]), _p(fwd["depth"]), _p(fwd["alpha"]),
                                _p(fwd["lastContrib"]), _p(g64))
        return g64

    def ssim_fwd(self, img1, img2, window, K=11):
        H, W, Cc = img1.shape
        outs = [np.empty(H * W * Cc, f32) for _ in range(6)]
        self.lib.gso_ssim_fwd(H, W, Cc, K, _p(_f(img1)), _p(_f(img2)), _p(window), *[_p(o) for o in outs])
        return dict(zip(("ssim", "mu1", "mu2", "sigma1", "sigma2", "sigma12"), outs))

    def ssim_bwd(self, grad_out, img1, img2, window, saved, K=11):
        H, W, Cc = img1.shape
        g1 = np.empty(H * W * Cc, f32); g2 = np.empty(H * W * Cc, f32)
        self.lib.gso_ssim_bwd(H, W, Cc, K, _p(_f(grad_out).reshape(-1)), _p(_f(img1)), _p(_f(img2)), _p(window),
                              _p(saved["mu1"]), _p(saved["mu2"]), _p(saved["sigma1"]), _p(saved["sigma2"]),
                              _p(saved["sigma12"]), _p(g1), _p(g2))
        return g1.reshape(H, W, Cc), g2.reshape(H, W, Cc)

    def adam(self, p, g, m, v, lr, b1=0.9, b2=0.999, eps=1e-15):
        self.lib.gso_adam(C.c_long(p.size), _p(p), _p(_f(g)), _p(m), _p(v), C.c_float(lr), C.c_float(b1),
                          C.c_float(b2), C.c_float(eps))

    def accum_grad_norm(self, xyz_grad, accum):
        self.lib.gso_accum_grad_norm(xyz_grad.shape[0], _p(_f(xyz_grad)), _p(accum))

    # ---- densification (D2, D3, gather + noise) ----------------------------------------------
    def classify_gaussians(self, accum, denom, scales_log, opacity_logit, grad_threshold, max_scale, min_opacity, allow_densify):
        n = accum.shape[0]
        actions = np.zeros(n, np.int32); counts = np.zeros(n, np.int32)
        self.lib.gso_classify_gaussians(n, _p(_f(accum)), C.c_float(denom), _p(_f(scales_log)), _p(_f(opacity_logit).reshape(-1)),
                                        C.c_float(grad_threshold), C.c_float(max_scale), C.c_float(min_opacity),
                                        int(bool(allow_densify)), _p(actions), _p(counts))
        return actions, counts

    def build_densify_output_map(self, actions, offsets, total):
        gather = np.zeros(total, np.int32); mode = np.zeros(total, np.int32)
        self.lib.gso_build_densify_output_map(actions.shape[0], _p(np.ascontiguousarray(actions, np.int32)),
                                              _p(np.ascontiguousarray(offsets, np.int32)), _p(gather), _p(mode))
        return gather, mode

    def densify_apply(self, params, gather, noise_mode, base_noise):
        """Phases 4-5 of split_and_prune; returns the new parameter dict."""
        nout = gather.shape[0]
        K = params["_features_rest"].shape[1] + 1
        out = {"_xyz": np.zeros((nout, 3), f32), "_features_dc": np.zeros((nout, 1, 3), f32),
               "_features_rest": np.zeros((nout, K - 1, 3), f32), "_scales": np.zeros((nout, 3), f32),
               "_rotation": np.zeros((nout, 4), f32), "_opacity": np.zeros((nout, 1), f32)}
        self.lib.gso_densify_apply(nout, K, _p(np.ascontiguousarray(gather, np.int32)), _p(np.ascontiguousarray(noise_mode, np.int32)),
                                   _p(_f(base_noise)), _p(_f(params["_xyz"])), _p(_f(params["_features_dc"])),
                                   _p(_f(params["_features_rest"])), _p(_f(params["_scales"])), _p(_f(params["_rotation"])),
                                   _p(_f(params["_opacity"])), _p(out["_xyz"]), _p(out["_features_dc"]), _p(out["_features_rest"]),
                                   _p(out["_scales"]), _p(out["_rotation"]), _p(out["_opacity"]))
        return out


class Ref(_Base):
    """The reference's own kernels (JSON MSL → g++).  Buffer order = each JSON's buffer_parameters."""
    kind = "reference"

    def __init__(self):
        if not REF_SO.exists():
            raise FileNotFoundError(f"{REF_SO} missing - run oracle/build_ref.py where /root/reference exists")
        self.lib = C.CDLL(str(REF_SO))

    @staticmethod
    def available() -> bool:
        return REF_SO.exists()

    def _run(self, name, gx, gy, bufs):
        arr = (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
        getattr(self.lib, "ref_" + name)(C.c_uint(gx), C.c_uint(gy), arr)

    @staticmethod
    def _cam_bufs(cam):
        one = lambda v: np.array([v], f32)
        return [cam.camera_center_f32(), np.ascontiguousarray(cam.worldViewTransform), np.ascontiguousarray(cam.projectionMatrix),
                one(cam.FoVx), one(cam.FoVy), one(cam.focalX), one(cam.focalY), one(cam.imageWidth), one(cam.imageHeight)]

    def project_fwd(self, act, cam, degree):
        n, K = act["shs"].shape[0], act["shs"].shape[1]
        o = {"means2d": np.zeros((n, 2), f32), "depths": np.zeros((n,), f32), "color": np.zeros((n, 3), f32),
             "cov2d": np.zeros((n, 2, 2), f32), "conic": np.zeros((n, 2, 2), f32), "radii": np.zeros((n,), f32),
             "rectMin": np.zeros((n, 2), f32), "rectMax": np.zeros((n, 2), f32)}
        counts = np.array([n, degree, K], u32)
        bufs = [act["scales"], act["rotations"], act["means3d"], act["shs"], *self._cam_bufs(cam), counts,
                o["means2d"], o["depths"], o["color"], o["cov2d"], o["conic"], o["radii"], o["rectMin"], o["rectMax"]]
        self._run("gaussian_projection_screen_fused_forward", n, 1, bufs)
        return o

    def project_bwd(self, act, cam, degree, cot):
        n, K = act["shs"].shape[0], act["shs"].shape[1]
        g = {"scales": np.zeros((n, 3), f32), "rotations": np.zeros((n, 4), f32), "means3d": np.zeros((n, 3), f32),
             "shs": np.zeros((n, K, 3), f32), "cameraCenterPoint": np.zeros((n, 3), f32)}
        counts = np.array([n, degree, K], u32)
        bufs = [act["scales"], act["rotations"], act["means3d"], act["shs"], *self._cam_bufs(cam),
                _f(cot["depths"]), _f(cot["means2d"]), _f(cot["cov2d"]), _f(cot["color"]), _f(cot["conic"]), counts,
                g["scales"], g["rotations"], g["means3d"], g["shs"], g["cameraCenterPoint"]]
        self._run("gaussian_projection_screen_fused_backward", n, 1, bufs)
        return g

    def bin(self, proj, W, H, tileW, tileH):
        n = proj["radii"].shape[0]
        gridW, gridH = (W + tileW - 1) // tileW, (H + tileH - 1) // tileH
        numTiles = gridW * gridH
        rmin, rmax, radii, depths = _f(proj["rectMin"]), _f(proj["rectMax"]), _f(proj["radii"]), _f(proj["depths"])
        cnt = np.array([n, tileW, tileH, W, H], u32)
        touched = np.zeros(n, u32)
        self._run("count_tiles_per_gaussian", n, 1, [rmin, rmax, radii, cnt, touched])
        cumsum = np.cumsum(touched, dtype=np.uint64)
        M = int(cumsum[-1]) if n else 0
        offsets = (cumsum - touched).astype(u32)
        kh = np.zeros(M, u32); kl = np.zeros(M, u32); gi = np.zeros(M, u32)
        if M:
            self._run("generate_keys", n, 1, [depths, rmin, rmax, radii, offsets, cnt, kh, kl, gi])
        tb = tile_bit_count(numTiles)
        sh = np.zeros(M, u32); sl = np.zeros(M, u32); sv = np.zeros(M, u32)
        if M > 1:
            self.lib.ref_stable_sort_tile_keys(_p(kh), _p(kl), _p(gi), C.c_uint(M), C.c_uint(max(tb, 1)), _p(sh), _p(sl), _p(sv))
        else:
            sh, sl, sv = kh.copy(), kl.copy(), gi.copy()
        ranges = np.zeros((numTiles, 2), u32)
        if M:
            self._run("compute_tile_ranges", M, 1, [sh, np.array([M, numTiles], u32), ranges])
        counts = np.zeros(numTiles, u32)
        self._run("compute_tile_counts_from_ranges", numTiles, 1, [ranges, np.array([numTiles], u32), counts])
        maxPairs = int(counts.max()) if numTiles else 0
        packedIdx = np.zeros(max(numTiles * maxPairs, 1), np.int32)
        if maxPairs:
            self._run("build_packed_tile_indices", numTiles * maxPairs, 1, [sv, ranges, np.array([numTiles, maxPairs], u32), packedIdx])
        return {"tilesTouched": touched, "offsets": offsets, "M": M, "keysHigh": kh, "keysLow": kl, "gaussIdx": gi,
                "sortedKeysHigh": sh, "sortedKeysLow": sl, "sortedGaussIdx": sv, "tileRanges": ranges,
                "tileCounts": counts, "tileBits": tb, "gridW": gridW, "gridH": gridH,
                "packedTileIndices": packedIdx, "maxTilePairs": maxPairs}

    def _render_counts(self, bins, W, H, tileW, tileH, white_bg):
        return np.array([W * H, bins["maxTilePairs"], bins["gridW"], tileW, tileH, W, H, int(white_bg)], u32)

    def raster_fwd(self, packed, bins, W, H, tileW, tileH, white_bg):
        P = W * H
        o = {"color": np.zeros((P, 3), f32), "depth": np.zeros((P, 1), f32), "alpha": np.zeros((P, 1), f32),
             "lastContrib": np.zeros((P, 1), u32)}
        rc = self._render_counts(bins, W, H, tileW, tileH, white_bg)
        self._run("gaussian_tile_global_forward", P, 1, [packed, bins["packedTileIndices"], bins["tileCounts"], rc,
                                                         o["color"], o["depth"], o["alpha"], o["lastContrib"]])
        return o

    def raster_bwd(self, packed, bins, W, H, tileW, tileH, white_bg, cot, fwd):
        g64 = np.zeros(packed.shape, np.float64)
        rc = self._render_counts(bins, W, H, tileW, tileH, white_bg)
        self.lib.ref_raster_backward(_p(packed), _p(bins["packedTileIndices"]), _p(bins["tileCounts"]),
                                     _p(_f(cot["color"])), _p(_f(cot["depth"])), _p(_f(cot["alpha"])),
                                     _p(fwd["color"]), _p(fwd["depth"]), _p(rc), _p(fwd["alpha"]),
                                     _p(fwd["lastContrib"]), _p(g64))
        return g64

    def ssim_fwd(self, img1, img2, window, K=11):
        H, W, Cc = img1.shape
        tot = H * W * Cc
        outs = [np.zeros(tot, f32) for _ in range(6)]
        self._run("ssim_forward", tot, 1, [_f(img1).reshape(-1), _f(img2).reshape(-1), window, np.array([H, W, Cc, K], u32), *outs])
        return dict(zip(("ssim", "mu1", "mu2", "sigma1", "sigma2", "sigma12"), outs))

    def ssim_bwd(self, grad_out, img1, img2, window, saved, K=11):
        H, W, Cc = img1.shape
        tot = H * W * Cc
        g1 = np.zeros(tot, f32); g2 = np.zeros(tot, f32)
        self._run("ssim_backward", tot, 1, [_f(grad_out).reshape(-1), _f(img1).reshape(-1), _f(img2).reshape(-1), window,
                                            saved["mu1"], saved["mu2"], saved["sigma1"], saved["sigma2"], saved["sigma12"],
                                            np.array([H, W, Cc, K], u32), g1, g2])
        return g1.reshape(H, W, Cc), g2.reshape(H, W, Cc)

    # Activations / Adam are MLX ops in the reference (no shipped kernel): borrow the port's.
    def _port(self):
        if not hasattr(self, "_port_obj"):
            self._port_obj = Port()
        return self._port_obj

    def activate_fwd(self, params):
        return self._port().activate_fwd(params)

    def activate_bwd(self, params, g_act):
        return self._port().activate_bwd(params, g_act)

    def adam(self, *a, **k):
        return self._port().adam(*a, **k)

    def accum_grad_norm(self, xyz_grad, accum):
        """D1 through the reference's own inline Metal kernel (GaussianTrainer.swift:321-339); in place."""
        n = accum.shape[0]
        out = np.zeros(n, f32)
        self._run("accum_grad_norm", n, 1, [_f(xyz_grad).reshape(-1), accum, np.array([n], np.int32), out])
        accum[:] = out

    def classify_gaussians(self, accum, denom, scales_log, opacity_logit, grad_threshold, max_scale, min_opacity, allow_densify):
        n = accum.shape[0]
        actions = np.zeros(n, np.int32); counts = np.zeros(n, np.int32)
        sl = _f(scales_log)
        self._run("classify_gaussians", n, 1, [
            _f(accum), np.array([n], np.int32), np.full(n, denom, f32), sl, np.array(sl.shape, np.int32), _f(opacity_logit).reshape(-1),
            np.array([grad_threshold], f32), np.array([max_scale], f32), np.array([min_opacity], f32),
            np.array([1 if allow_densify else 0], np.int32), actions, counts])
        return actions, counts

    def build_densify_output_map(self, actions, offsets, total):
        gather = np.zeros(total, np.int32); mode = np.zeros(total, np.int32)
        a = np.ascontiguousarray(actions, np.int32)
        self._run("build_densify_output_map", a.shape[0], 1, [a, np.array([a.shape[0]], np.int32),
                                                              np.ascontiguousarray(offsets, np.int32), gather, mode])
        return gather, mode

    def densify_apply(self, *a, **k):   # MLX gather + arithmetic in the reference (no kernel): the port's restatement
        return self._port().densify_apply(*a, **k)

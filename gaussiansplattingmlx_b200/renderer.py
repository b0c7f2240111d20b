"""``GaussianRenderer`` — host-side mirror of the reference renderer class
(``Trainer/GaussianRenderer.swift:703-963``): same constructor arguments, method names, argument
order and return tuple, with every kernel call replaced by the C ABI of ``libgsb.so``.

* ``forward`` / ``forwardWithCameraParams`` / ``render`` take ACTIVATED tensors exactly like the
  reference and are differentiable through ``torch.autograd`` (two custom Functions standing where
  the reference has its two ``CustomFunction{Forward; VJP}`` pairs, ``GaussianRenderer.swift:150-184``
  and ``:576-602``).
* ``forward_raw`` is the fused training path (raw tensors in, activations fused into projection).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .camera import Camera
from .context import Context, PARAM_NAMES


class _ProjectFn(torch.autograd.Function):
    """projectionScreenFusedCustomFunction (``GaussianRenderer.swift:510-701``)."""

    @staticmethod
    def forward(fn, ctx: Context, cam, scales, rotations, means3d, shs):
        act = {"scales": scales.contiguous(), "rotations": rotations.contiguous(), "means3d": means3d.contiguous(),
               "shs": shs.contiguous()}
        o = ctx.project_fwd(act, cam)
        fn.gsb, fn.cam = ctx, cam
        fn.save_for_backward(act["scales"], act["rotations"], act["means3d"], act["shs"])
        for k in ("radii", "rectMin", "rectMax"):
            fn.mark_non_differentiable(o[k])          # stopGradient, GaussianRenderer.swift:863-865
        return (o["means2d"], o["depths"], o["color"], o["cov2d"], o["conic"], o["radii"], o["rectMin"], o["rectMax"])

    @staticmethod
    def backward(fn, g_means2d, g_depths, g_color, g_cov2d, g_conic, *_):
        scales, rotations, means3d, shs = fn.saved_tensors
        n = means3d.shape[0]
        z = lambda g, *shape: torch.zeros(*shape, device=means3d.device) if g is None else g.contiguous()
        cot = {"depths": z(g_depths, n), "means2d": z(g_means2d, n, 2), "cov2d": z(g_cov2d, n, 2, 2),
               "color": z(g_color, n, 3), "conic": z(g_conic, n, 2, 2)}
        g = fn.gsb.project_bwd({"scales": scales, "rotations": rotations, "means3d": means3d, "shs": shs}, fn.cam, cot)
        return None, None, g["scales"], g["rotations"], g["means3d"], g["shs"]


class _CompositeFn(torch.autograd.Function):
    """renderGlobalTileCompositeCustomOp (``GaussianRenderer.swift:103-226``)."""

    @staticmethod
    def forward(fn, ctx: Context, packed):
        packed = packed.contiguous()
        o = ctx.raster_fwd(packed)
        fn.gsb = ctx
        fn.bin_generation = ctx.bin_generation()     # the reference's closure captures the slice info per call (:119-122)
        fn.save_for_backward(packed, o["color"], o["depth"], o["alpha"], o["lastContrib"])
        fn.mark_non_differentiable(o["lastContrib"])
        return o["color"], o["depth"], o["alpha"], o["lastContrib"]

    @staticmethod
    def backward(fn, g_color, g_depth, g_alpha, _):
        packed, color, depth, alpha, last = fn.saved_tensors
        if fn.gsb.bin_generation() != fn.bin_generation:
            raise RuntimeError("GaussianRenderer: the context's tile lists were rebuilt (another forward / bin call) between this "
                               "render's forward and its backward; render and differentiate one view at a time per renderer")
        P = color.shape[0]
        z = lambda g, *shape: torch.zeros(*shape, device=packed.device) if g is None else g.contiguous()
        cot = {"color": z(g_color, P, 3), "depth": z(g_depth, P, 1), "alpha": z(g_alpha, P, 1)}
        g = fn.gsb.raster_bwd(packed, cot, {"color": color, "depth": depth, "alpha": alpha, "lastContrib": last})
        return None, g


class GaussianRenderer:
    def __init__(self, active_sh_degree: int, W: int, H: int, TILE_SIZE: Tuple[int, int] = (16, 16),
                 whiteBackground: bool = False, useScreenSpaceCustomOp: bool = True, device: int = 0,
                 sh_coeffs: Optional[int] = None, max_gaussians: int = 0):
        # reference signature: init(active_sh_degree:W:H:TILE_SIZE:whiteBackground:useScreenSpaceCustomOp:)
        self.active_sh_degree = active_sh_degree
        self.W, self.H = W, H
        self.TILE_SIZE = (int(TILE_SIZE[0]), int(TILE_SIZE[1]))   # (w, h)
        self.whiteBackground = whiteBackground
        self.ctx = Context(W, H, tile_w=self.TILE_SIZE[0], tile_h=self.TILE_SIZE[1], sh_degree=active_sh_degree,
                           sh_coeffs=sh_coeffs, white_background=whiteBackground, device=device, max_gaussians=max_gaussians)

    # ---- activations (GaussianRenderer.swift:936-963) -----------------------------------------
    @staticmethod
    def get_scales_from(scales): return torch.exp(scales)
    @staticmethod
    def get_rotation_from(rotation): return rotation / (rotation.norm(dim=-1, keepdim=True) + 1e-8)
    @staticmethod
    def get_xyz_from(xyz): return xyz
    @staticmethod
    def get_features_from(features_dc, features_rest): return torch.cat([features_dc, features_rest], dim=1)
    @staticmethod
    def get_opacity_from(opacity): return torch.sigmoid(opacity)

    # ---- reference-shaped differentiable path -------------------------------------------------
    def render(self, means2d, depths, color, conic, opacity, radii, rectMin, rectMax):
        """``render`` (``GaussianRenderer.swift:769-821``): pack → slice info → tile composite."""
        n = means2d.shape[0]
        packed = torch.cat([means2d, conic.reshape(n, 4), color, opacity.reshape(n, 1), depths.reshape(n, 1)], dim=1)
        self.ctx.bin({"rectMin": rectMin.detach(), "rectMax": rectMax.detach(), "radii": radii.detach(),
                      "depths": depths.detach().contiguous()}, read_lists=False)
        color_o, depth_o, alpha_o, _ = _CompositeFn.apply(self.ctx, packed)
        return (color_o.reshape(self.H, self.W, 3), depth_o.reshape(self.H, self.W, 1), alpha_o.reshape(self.H, self.W, 1))

    def forwardWithCameraParams(self, viewMatrix, projMatrix, cameraCenter, fovX, fovY, focalX, focalY, imageWidth, imageHeight,
                                means3d, shs, opacity, scales, rotations):
        """``GaussianRenderer.swift:823-880``.  Camera arguments are host values (numpy / floats)."""
        assert int(imageWidth) == self.W and int(imageHeight) == self.H
        cam = _lib.GsbCamera()
        import numpy as np
        v = np.asarray(viewMatrix, np.float32).reshape(-1); p = np.asarray(projMatrix, np.float32).reshape(-1)
        cc = np.asarray(cameraCenter, np.float32).reshape(-1)
        for i in range(16):
            cam.view[i] = float(v[i]); cam.proj[i] = float(p[i])
        for i in range(3):
            cam.cam_center[i] = float(cc[i])
        cam.fov_x, cam.fov_y, cam.focal_x, cam.focal_y = float(fovX), float(fovY), float(focalX), float(focalY)
        means2d, depths, color, cov2d, conic, radii, rectMin, rectMax = _ProjectFn.apply(self.ctx, cam, scales, rotations, means3d, shs)
        render, depth, alpha = self.render(means2d, depths, color, conic, opacity, radii, rectMin, rectMax)
        return render, depth, alpha, radii > 0, radii

    def forward(self, camera: Camera, means3d, shs, opacity, scales, rotations):
        """``GaussianRenderer.swift:882-934``."""
        return self.forwardWithCameraParams(camera.worldViewTransform, camera.projectionMatrix, camera.camera_center_f32(),
                                            camera.FoVx, camera.FoVy, camera.focalX, camera.focalY, camera.imageWidth,
                                            camera.imageHeight, means3d, shs, opacity, scales, rotations)

    # ---- fused path (training / inference) ----------------------------------------------------
    def forward_raw(self, camera: Camera, params: Dict[str, torch.Tensor]):
        """Raw model tensors in; activations + projection + binning + raster in five kernels.
        Returns the reference tuple (render, depth, alpha, visibility_filter, radii)."""
        return self.ctx.render_forward(params, _lib.make_camera(camera))

    def backward_raw(self, cot_render, cot_depth=None, cot_alpha=None, grads=None, accumulate=False):
        return self.ctx.render_backward(cot_render, cot_depth, cot_alpha, grads, accumulate)

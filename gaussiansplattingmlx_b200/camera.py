"""Camera → view / projection matrices, mirroring the reference's conventions exactly.

Reference: ``Trainer/CameraUtil.swift:5-102`` (``Camera``, ``focal2fov``, ``getProjectionMatrix``)
and ``Trainer/simd+ext.swift:45-55`` (``toMLXArray`` row-major export).

* ``worldViewTransform`` = ``(c2w^-1)^T`` stored row-major, **row-vector convention**:
  ``p_view = [x y z 1] @ V``.
* ``projectionMatrix`` = ``P^T`` stored row-major, with ``P(0,0)=1/tan(fovX/2)``,
  ``P(1,1)=1/tan(fovY/2)``, ``P(2,2)=zf/(zf-zn)``, ``P(3,2)=1``, ``P(2,3)=-zf*zn/(zf-zn)``.
* ``FoV = 2*atan(pixels/(2*focal))`` is evaluated in f32 (it is an MLXArray op in the
  reference), then widened to f64 to build ``P``; matrices are computed in f64 and rounded to
  f32 once.  The principal point is ignored (only ``intrinsic[0][0]``, ``[1][1]`` are used).
* ``cameraCenter`` = translation column of ``c2w``.

Host-side only (numpy); the device never sees anything but the 16+16+7 floats below.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def focal2fov(focal, pixels: float):
    """``CameraUtil.swift:74-80`` — f32 when ``focal`` is f32 (MLXArray path)."""
    focal = np.float32(focal)
    return np.float32(2.0) * np.arctan(np.float32(pixels) / (np.float32(2.0) * focal), dtype=np.float32)


def fov2focal(fov: float, pixels: float) -> float:
    """``CameraUtil.swift:66-71``."""
    return pixels / (2.0 * np.tan(fov / 2.0))


def get_projection_matrix(znear: float, zfar: float, fovX: float, fovY: float) -> np.ndarray:
    """``CameraUtil.swift:82-102`` — returns the mathematical matrix P (f64, P[r, c])."""
    tanHalfY = np.tan(fovY / 2.0)
    tanHalfX = np.tan(fovX / 2.0)
    top = tanHalfY * znear
    bottom = -top
    right = tanHalfX * znear
    left = -right
    P = np.zeros((4, 4), dtype=np.float64)
    P[0, 0] = 2 * znear / (right - left)
    P[1, 1] = 2 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[2, 2] = zfar / (zfar - znear)
    P[3, 2] = 1.0
    P[2, 3] = -znear * zfar / (zfar - znear)
    return P


@dataclass
class Camera:
    """Same fields as the reference ``Camera`` class (``CameraUtil.swift:5-62``)."""

    imageWidth: int
    imageHeight: int
    focalX: np.float32
    focalY: np.float32
    FoVx: np.float32
    FoVy: np.float32
    worldViewTransform: np.ndarray  # [4,4] f32 row-major, row-vector convention
    projectionMatrix: np.ndarray    # [4,4] f32 row-major (= P^T)
    cameraCenter: np.ndarray        # [3] f64

    def __init__(self, width: int, height: int, focalX: float, focalY: float, c2w: np.ndarray,
                 znear: float = 0.1, zfar: float = 100.0):
        c2w = np.asarray(c2w, dtype=np.float64).reshape(4, 4)
        self.imageWidth = int(width)
        self.imageHeight = int(height)
        self.focalX = np.float32(focalX)
        self.focalY = np.float32(focalY)
        self.FoVx = focal2fov(self.focalX, float(width))
        self.FoVy = focal2fov(self.focalY, float(height))
        wvt = np.linalg.inv(c2w).T
        P = get_projection_matrix(znear, zfar, float(self.FoVx), float(self.FoVy))
        self.worldViewTransform = np.ascontiguousarray(wvt, dtype=np.float32)
        self.projectionMatrix = np.ascontiguousarray(P.T, dtype=np.float32)
        inv_wv = np.linalg.inv(wvt).T  # == c2w
        self.cameraCenter = np.array([inv_wv[0, 3], inv_wv[1, 3], inv_wv[2, 3]], dtype=np.float64)

    @classmethod
    def from_intrinsic(cls, width: int, height: int, intrinsic: np.ndarray, c2w: np.ndarray,
                       znear: float = 0.1, zfar: float = 100.0) -> "Camera":
        """First initialiser of the reference class (``CameraUtil.swift:16-39``)."""
        intrinsic = np.asarray(intrinsic)
        return cls(width, height, float(intrinsic[0][0]), float(intrinsic[1][1]), c2w, znear, zfar)

    def camera_center_f32(self) -> np.ndarray:
        """``GaussianTrainer.swift:500-505`` — [1,3] f32."""
        return self.cameraCenter.astype(np.float32).reshape(1, 3)

    def pack(self) -> np.ndarray:
        """The 39-float camera block the C ABI takes (``gsb_camera`` in include/gsb.h):
        view[16], proj[16], cameraCenter[3], fovX, fovY, focalX, focalY.
        ``tan(fov*0.5)`` (``gaussian_projection_screen_shared.slang:200-201``) is evaluated by the
        library's host code with libm ``tanf``, not here (numpy's f32 tan may differ by an ulp)."""
        out = np.zeros(39, dtype=np.float32)
        out[0:16] = self.worldViewTransform.reshape(-1)
        out[16:32] = self.projectionMatrix.reshape(-1)
        out[32:35] = self.cameraCenter.astype(np.float32)
        out[35] = self.FoVx
        out[36] = self.FoVy
        out[37] = self.focalX
        out[38] = self.focalY
        return out

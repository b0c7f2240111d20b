"""``GaussModel`` — the parameter container of the reference (``Trainer/GaussianModel.swift:33-65``):
six raw tensors with the reference's names, shapes and learning-rate schedule.  Host-side only."""
from __future__ import annotations

from typing import Dict, List

import numpy as np

PARAM_ORDER = ("_xyz", "_features_dc", "_features_rest", "_scales", "_rotation", "_opacity")


class GaussModel:
    def __init__(self, sh_degree: int = 3):
        self.max_sh_degree = sh_degree
        K = (sh_degree + 1) ** 2
        self._xyz = np.zeros((0, 3), np.float32)
        self._features_dc = np.zeros((0, 1, 3), np.float32)
        self._features_rest = np.zeros((0, K - 1, 3), np.float32)
        self._scales = np.zeros((0, 3), np.float32)
        self._rotation = np.zeros((0, 4), np.float32)
        self._opacity = np.zeros((0, 1), np.float32)

    @classmethod
    def from_arrays(cls, params: Dict[str, np.ndarray], sh_degree: int = 3) -> "GaussModel":
        m = cls(sh_degree)
        for k in PARAM_ORDER:
            setattr(m, k, np.ascontiguousarray(params[k], dtype=np.float32))
        K = (sh_degree + 1) ** 2
        n = m._xyz.shape[0]
        assert m._features_dc.shape == (n, 1, 3) and m._features_rest.shape == (n, K - 1, 3)
        assert m._scales.shape == (n, 3) and m._rotation.shape == (n, 4) and m._opacity.shape == (n, 1)
        return m

    def getParams(self) -> List[np.ndarray]:
        """``GaussianModel.swift:44-55`` order."""
        return [getattr(self, k) for k in PARAM_ORDER]

    def as_dict(self) -> Dict[str, np.ndarray]:
        return {k: getattr(self, k) for k in PARAM_ORDER}

    @staticmethod
    def getLearningRates(current: int, total: int) -> List[float]:
        """``GaussianModel.swift:56-65`` in f32: xyz decays linearly to 1 % of 1.6e-4."""
        f32 = np.float32
        return [float(f32(0.00016) * max(f32(1.0) - f32(current) / f32(total), f32(0.01))),
                0.0025, float(f32(0.0025) / f32(20)), 0.005, 0.001, 0.025]

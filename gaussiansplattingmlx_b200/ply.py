"""``PlyWriter`` — the reference's only on-disk model format (``Data/PlyWriter.swift:22-146`` writer, ``:149-265``
loader; written every 100 iterations by ``GaussianTrainer.save_snapshot``, ``GaussianTrainer.swift:909-929``).

binary_little_endian 1.0, one ``element vertex``, all properties ``float``, in this order:
``x y z f_dc_0..2 f_rest_0..(3M-1) opacity scale_0..2 rot_0..3`` with the header comment
``features_rest_shape M 3``.  ``f_rest`` is the row-major flattening of ``[M, 3]`` (coefficient-major, channel
fastest — NOT the INRIA channel-major order); opacity is the raw logit, scales are log-scales, rotations raw (w,x,y,z).

The reference has no resume path (no optimiser state, no iteration counter); ``save_resume`` / ``load_resume`` add one
as a sidecar ``.npz`` holding Adam m/v, the gradient-norm accumulation and the iteration.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy as np

from .model import PARAM_ORDER


class PlyError(ValueError):
    pass


def _header(num_points: int, M: int) -> str:
    # line for line Data/PlyWriter.swift:45-67
    h = "ply\nformat binary_little_endian 1.0\n"
    h += f"comment features_rest_shape {M} 3\n"
    h += f"element vertex {num_points}\n"
    for name in ("x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2"):
        h += f"property float {name}\n"
    for i in range(M * 3):
        h += f"property float f_rest_{i}\n"
    for name in ("opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"):
        h += f"property float {name}\n"
    return h + "end_header\n"


class PlyWriter:
    @staticmethod
    def writeGaussianBinary(positions, features_dc, features_rest, opacities, scales, rotations, to) -> None:
        """Same argument order as the MLXArray overload (``PlyWriter.swift:117-146``): positions[N,3],
        features_dc[N,1,3] or [N,3], features_rest[N,M,3], opacities[N] or [N,1], scales[N,3], rotations[N,4]."""
        pos = np.asarray(positions, np.float32).reshape(-1, 3)
        n = pos.shape[0]
        dc = np.asarray(features_dc, np.float32).reshape(-1, 3)
        rest = np.asarray(features_rest, np.float32)
        if rest.ndim != 3 or rest.shape[2] != 3:
            raise PlyError("features_rest must be [N, M, 3]")
        M = rest.shape[1]
        opa = np.asarray(opacities, np.float32).reshape(-1)
        scl = np.asarray(scales, np.float32).reshape(-1, 3)
        rot = np.asarray(rotations, np.float32).reshape(-1, 4)
        if not (dc.shape[0] == rest.shape[0] == opa.shape[0] == scl.shape[0] == rot.shape[0] == n):
            raise PlyError("Attribute array size mismatch")                       # PlyWriter.swift:34-42
        rows = np.concatenate([pos, dc, rest.reshape(n, M * 3), opa[:, None], scl, rot], axis=1).astype("<f4")
        path = Path(to)
        path.parent.mkdir(parents=True, exist_ok=True)
        with open(path, "wb") as f:
            f.write(_header(n, M).encode("ascii"))
            f.write(np.ascontiguousarray(rows).tobytes())

    @staticmethod
    def loadGaussianBinaryPLY(path) -> Dict[str, np.ndarray]:
        """``loadGaussianBinaryPLYAsMLX`` (``PlyWriter.swift:149-265``): properties are looked up BY NAME, so any
        property order is accepted; returns the six tensors with the model's shapes and names."""
        data = Path(path).read_bytes()
        end = data.find(b"end_header\n")
        if end < 0:
            raise PlyError("No end_header")
        end += len(b"end_header\n")
        try:
            header = data[:end].decode("ascii")
        except UnicodeDecodeError as ex:
            raise PlyError("Header parse error") from ex
        num_points, rest_shape, fields = 0, None, []
        for line in header.split("\n"):
            parts = line.split(" ")
            if line.startswith("comment features_rest_shape") and len(parts) >= 4:
                rest_shape = (int(parts[2] or 0), int(parts[3] or 0))
            if len(parts) >= 3 and parts[0] == "element" and parts[1] == "vertex":
                num_points = int(parts[2])
            elif len(parts) == 3 and parts[0] == "property" and parts[1] == "float":
                fields.append(parts[2])
        if rest_shape is None:
            raise PlyError("No features_rest_shape comment")
        M, D = rest_shape
        fpv = len(fields)
        body = np.frombuffer(data, dtype="<f4", count=num_points * fpv, offset=end).reshape(num_points, fpv)
        idx = {name: i for i, name in reversed(list(enumerate(fields)))}   # firstIndex(of:)

        def cols(names):
            try:
                return np.ascontiguousarray(body[:, [idx[nm] for nm in names]], dtype=np.float32)
            except KeyError as ex:
                raise PlyError(f"missing property {ex.args[0]}") from ex
        return {
            "_xyz": cols(["x", "y", "z"]),
            "_features_dc": cols(["f_dc_0", "f_dc_1", "f_dc_2"]).reshape(num_points, 1, 3),
            "_features_rest": cols([f"f_rest_{i}" for i in range(M * D)]).reshape(num_points, M, D),
            "_opacity": cols(["opacity"]).reshape(num_points, 1),
            "_scales": cols(["scale_0", "scale_1", "scale_2"]),
            "_rotation": cols(["rot_0", "rot_1", "rot_2", "rot_3"]),
        }


def save_snapshot(params: Dict[str, np.ndarray], output_dir, iteration: int) -> Path:
    """``GaussianTrainer.save_snapshot`` (``GaussianTrainer.swift:909-929``): ``iteration_<n>.ply``."""
    out = Path(output_dir) / f"iteration_{iteration}.ply"
    PlyWriter.writeGaussianBinary(params["_xyz"], params["_features_dc"], params["_features_rest"], params["_opacity"],
                                  params["_scales"], params["_rotation"], out)
    return out


def save_resume(path, iteration: int, m: Dict[str, np.ndarray], v: Dict[str, np.ndarray], accum: np.ndarray,
                accum_steps: int) -> None:
    arrays = {f"m{k}": np.asarray(m[k], np.float32) for k in PARAM_ORDER}
    arrays.update({f"v{k}": np.asarray(v[k], np.float32) for k in PARAM_ORDER})
    np.savez(path, iteration=np.int64(iteration), accum=np.asarray(accum, np.float32), accum_steps=np.int64(accum_steps), **arrays)


def load_resume(path) -> Tuple[int, Dict[str, np.ndarray], Dict[str, np.ndarray], np.ndarray, int]:
    z = np.load(path)
    m = {k: z[f"m{k}"] for k in PARAM_ORDER}
    v = {k: z[f"v{k}"] for k in PARAM_ORDER}
    return int(z["iteration"]), m, v, z["accum"], int(z["accum_steps"])

"""Dataset loaders → ``TrainData`` + ``PointCloud`` + tile size, and point-cloud initialisation of the model
(SURVEY.md §8 row f4).  Host-side IO only (numpy + Pillow); nothing here runs on the GPU.

Reference:
* ``Data/DataLoaderProtocol.swift:8-13``   ``load(resizeFactor:whiteBackground:) -> (TrainData, PointCloud, TILE_SIZE_H_W)``
* ``Data/ColmapDataLoader.swift:165-498``  COLMAP ``cameras.bin`` / ``images.bin`` / ``points3D.bin`` (little endian)
* ``Data/NerfStudioDataLoader.swift``      ``transforms.json`` (+ per-frame intrinsics), ascii / binary PLY point cloud,
                                         OpenGL → OpenCV camera conversion (rows 1, 2 of w2c negated, ``:352-361``)
* ``Data/BlenderDataLoader.swift``         ``info.json`` with intrinsic / pose / rgb / ``*_depth.png`` / ``*_alpha.png``; the
                                         point cloud is un-projected from the depth maps (``PointCloudUtil.swift:101-138``)
* ``Trainer/PointCloudUtil.swift:139-191`` ``PointCloud`` (``select_channels``, ``randomSample``, ``centering``)
* ``Trainer/GaussianModel.swift:11-31,87-125`` ``distTopK`` and ``create_from_pcd``
* every loader returns ``TILE_SIZE = (W / 4, H / 4)``; the app then calls ``centering`` and samples 16 384 points
  (``UI/TrainView.swift:158-175``).

Deliberate differences (Apple-only APIs have no Linux equivalent): images are decoded and resized with Pillow
(bilinear), not UIKit/CoreGraphics, so resized pixels differ in the last bits; demo datasets are not downloaded
(no network) — point the loader at a directory; zip archives are unpacked with ``zipfile``.
"""
from __future__ import annotations

import json
import struct
import zipfile
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .camera import Camera
from .model import GaussModel

C0 = 0.28209479177387814


def RGB2SH(rgb):
    """``ShUtils.swift``: (rgb - 0.5) / C0."""
    return (np.asarray(rgb, np.float32) - np.float32(0.5)) / np.float32(C0)


def inverse_sigmoid(x):
    x = np.asarray(x, np.float32)
    return np.log(x / (np.float32(1.0) - x))


# ------------------------------------------------------------------------------------------------
# containers
# ------------------------------------------------------------------------------------------------
class LoadedTrainData:
    """``TrainData`` (``GaussianTrainer.swift:14-60``): Hs, Ws, intrinsicArray[B,3|4,3|4], c2wArray[B,4,4] (OpenCV
    convention, mathematical layout M[r, c]), rgbArray[B,H,W,3], alphaArray[B,H,W], depthArray | None."""

    def __init__(self, Hs, Ws, intrinsicArray, c2wArray, rgbArray, alphaArray, depthArray=None):
        self.Hs = np.asarray(Hs, np.float32)
        self.Ws = np.asarray(Ws, np.float32)
        self.intrinsicArray = np.asarray(intrinsicArray, np.float32)
        self.c2wArray = np.asarray(c2wArray, np.float32)
        self.rgbArray = rgbArray
        self.alphaArray = alphaArray
        self.depthArray = depthArray

    def getNumCameras(self) -> int:
        return int(self.Hs.shape[0])

    def getViewPointCamera(self, index: int) -> Camera:
        """``Camera(width:height:intrinsic:c2w:)`` — only intrinsic[0][0] and [1][1] are used (``CameraUtil.swift:26-27``)."""
        K = self.intrinsicArray[index]
        return Camera(int(self.Ws[index]), int(self.Hs[index]), float(K[0, 0]), float(K[1, 1]), self.c2wArray[index].astype(np.float64))

    @property
    def cameras(self) -> List[Camera]:
        return [self.getViewPointCamera(i) for i in range(self.getNumCameras())]

    def getCameraParams(self):
        return self.Hs, self.Ws, self.intrinsicArray, self.c2wArray


class PointCloud:
    """``PointCloudUtil.swift:139-191``: coords[N,3] + named channels in [0, 1]."""
    COLORS = {"R", "G", "B", "A"}

    def __init__(self, coords, channels: Dict[str, np.ndarray]):
        self.coords = np.asarray(coords, np.float32)
        self.channels = {k: np.asarray(v, np.float32) for k, v in channels.items()}

    def select_channels(self, channel_names: Sequence[str]) -> np.ndarray:
        cols = []
        for name in channel_names:
            d = self.channels[name]
            cols.append(np.round(d * np.float32(255.0)) if name in self.COLORS else d)
        return np.stack(cols, axis=-1)

    def randomSample(self, numPoints: int, rng: Optional[np.random.Generator] = None) -> "PointCloud":
        n = self.coords.shape[0]
        if n <= numPoints:
            return self
        rng = rng or np.random.default_rng()
        pick = rng.permutation(n)[:numPoints]
        return PointCloud(self.coords[pick], {k: v[pick] for k, v in self.channels.items()})

    def centering(self, data: LoadedTrainData, outlierSigma: float = 3.0) -> None:
        """Moves the cloud's mean to the origin (cameras follow), then drops points beyond ``outlierSigma`` standard
        deviations on any axis (``:171-190``; MLX ``std`` is the population standard deviation)."""
        center = self.coords.mean(axis=0, dtype=np.float32)
        data.c2wArray = data.c2wArray.copy()
        data.c2wArray[:, :3, 3] -= center
        self.coords = self.coords - center
        std = self.coords.std(axis=0, dtype=np.float32)
        s = np.float32(outlierSigma)
        keep = np.all((self.coords > -s * std) & (self.coords < s * std), axis=1)
        self.coords = self.coords[keep]
        self.channels = {k: v[keep] for k, v in self.channels.items()}


# ------------------------------------------------------------------------------------------------
# image IO (Pillow stands in for UIKit)
# ------------------------------------------------------------------------------------------------
def _load_rgba(path: Path, scale: float) -> np.ndarray:
    from PIL import Image
    img = Image.open(path).convert("RGBA")
    if scale != 1.0:
        img = img.resize((int(img.width * scale), int(img.height * scale)), Image.BILINEAR)
    a = np.asarray(img, np.float32) / np.float32(255.0)
    # CoreGraphics draws into a premultiplied-alpha context (ColmapDataLoader.swift:136-148)
    a[..., :3] *= a[..., 3:4]
    return a


def _load_gray(path: Path, scale: float) -> np.ndarray:
    from PIL import Image
    img = Image.open(path).convert("L")
    if scale != 1.0:
        img = img.resize((int(img.width * scale), int(img.height * scale)), Image.BILINEAR)
    return np.asarray(img, np.float32) / np.float32(255.0)


def _compose(rgbs: np.ndarray, alphas: np.ndarray, whiteBackground: bool) -> np.ndarray:
    if whiteBackground:
        return alphas[..., None] * rgbs + (np.float32(1.0) - alphas)[..., None]
    return rgbs


def _tile_size(data: LoadedTrainData) -> Tuple[int, int]:
    """TILE_SIZE_H_W(w: W / 4, h: H / 4) — returned as (w, h)."""
    return int(data.Ws[0]) // 4, int(data.Hs[0]) // 4


def _maybe_unzip(path: Path) -> Path:
    if path.is_file() and path.suffix == ".zip":
        out = path.with_suffix("")
        if not out.exists():
            with zipfile.ZipFile(path) as z:
                z.extractall(out)
        return out
    return path


# ------------------------------------------------------------------------------------------------
# COLMAP
# ------------------------------------------------------------------------------------------------
_COLMAP_MODEL_PARAMS = {0: 3, 1: 4, 2: 4, 3: 8}   # SimplePinhole, Pinhole, SimpleRadial, OpenCV


def _quat_to_rotmat(q) -> np.ndarray:
    w, x, y, z = (float(v) for v in q)
    n = np.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]], np.float64)


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.o = data, 0

    def take(self, fmt: str):
        size = struct.calcsize(fmt)
        if self.o + size > len(self.d):
            raise ValueError("Not enough data")          # ColmapDataLoader.swift:375-381
        v = struct.unpack_from(fmt, self.d, self.o)
        self.o += size
        return v if len(v) > 1 else v[0]

    def cstring(self) -> str:
        end = self.d.find(b"\0", self.o)
        end = len(self.d) if end < 0 else end
        s = self.d[self.o:end].decode("utf-8")
        self.o = end + 1
        return s


def colmap_read_cameras(path: Path) -> Dict[int, Dict[str, float]]:
    r = _Reader(Path(path).read_bytes())
    cams = {}
    for _ in range(r.take("<Q")):
        cam_id, model = r.take("<Ii")
        width, height = r.take("<QQ")
        model = model if model in _COLMAP_MODEL_PARAMS else 1      # unknown models are read as Pinhole (:196)
        p = [r.take("<d") for _ in range(_COLMAP_MODEL_PARAMS[model])]
        if model in (0, 2):
            fx, fy, cx, cy = p[0], p[0], p[1], p[2]
        else:
            fx, fy, cx, cy = p[0], p[1], p[2], p[3]
        cams[cam_id] = {"width": width, "height": height, "fx": fx, "fy": fy, "cx": cx, "cy": cy, "model": model}
    return cams


def colmap_read_images(path: Path) -> List[Dict]:
    """Returns name, camera id and the camera-to-world pose (R^T, -R^T t) of every image, in file order."""
    r = _Reader(Path(path).read_bytes())
    out = []
    for _ in range(r.take("<Q")):
        r.take("<I")
        q = r.take("<dddd")
        t = np.array(r.take("<ddd"), np.float64)
        cam_id = r.take("<I")
        name = r.cstring()
        n2d = r.take("<Q")
        r.o += n2d * 24
        Rinv = _quat_to_rotmat(q).T
        pose = np.eye(4, dtype=np.float64)
        pose[:3, :3] = Rinv
        pose[:3, 3] = -(Rinv @ t)
        out.append({"name": name, "camera_id": cam_id, "c2w": pose})
    return out


def colmap_read_points3d(path: Path) -> Tuple[np.ndarray, np.ndarray]:
    r = _Reader(Path(path).read_bytes())
    n = r.take("<Q")
    pts = np.zeros((n, 3), np.float64)
    cols = np.zeros((n, 3), np.uint8)
    for i in range(n):
        r.take("<Q")
        pts[i] = r.take("<ddd")
        cols[i] = r.take("<BBB")
        r.take("<d")
        track = r.take("<Q")
        r.o += track * 8
    return pts, cols


class ColmapDataLoader:
    """``<root>/colmap/sparse/0/{cameras,images,points3D}.bin`` + ``<root>/images/`` (``:499-517``); ``root`` may be a zip."""

    def __init__(self, root, bin_subdir: str = "colmap/sparse/0", image_subdir: str = "images"):
        self.root = _maybe_unzip(Path(root))
        self.bin_root = self.root / bin_subdir
        self.image_root = self.root / image_subdir

    def getOriginalImageSize(self) -> Tuple[int, int]:
        cams = colmap_read_cameras(self.bin_root / "cameras.bin")
        first = colmap_read_images(self.bin_root / "images.bin")[0]
        c = cams[first["camera_id"]]
        return int(c["width"]), int(c["height"])

    def load(self, resizeFactor: float = 1.0, whiteBackground: bool = False):
        if not (self.bin_root / "cameras.bin").exists() or not (self.bin_root / "images.bin").exists():
            raise FileNotFoundError("Colmap files missing")
        cams = colmap_read_cameras(self.bin_root / "cameras.bin")
        images = colmap_read_images(self.bin_root / "images.bin")
        K = []
        for im in images:
            c = cams[im["camera_id"]]
            k = np.array([[c["fx"], 0, c["cx"]], [0, c["fy"], c["cy"]], [0, 0, 1]], np.float64)
            if resizeFactor != 1.0:
                k[:2, :3] *= resizeFactor
            K.append(k)
        rgba = [_load_rgba(self.image_root / im["name"], resizeFactor) for im in images]
        rgbs = np.stack([a[..., :3] for a in rgba])
        alphas = np.stack([a[..., 3] for a in rgba])
        data = LoadedTrainData(Hs=[a.shape[0] for a in rgba], Ws=[a.shape[1] for a in rgba], intrinsicArray=np.stack(K),
                               c2wArray=np.stack([im["c2w"] for im in images]), rgbArray=_compose(rgbs, alphas, whiteBackground),
                               alphaArray=alphas)
        pts, cols = colmap_read_points3d(self.bin_root / "points3D.bin")
        ch = cols.astype(np.float32) / np.float32(255.0)
        pc = PointCloud(pts.astype(np.float32), {"R": ch[:, 0], "G": ch[:, 1], "B": ch[:, 2]})
        return data, pc, _tile_size(data)


# ------------------------------------------------------------------------------------------------
# NerfStudio
# ------------------------------------------------------------------------------------------------
def parse_point_ply(path: Path) -> Tuple[np.ndarray, np.ndarray]:
    """``NerfStudioDataLoader.parsePLY`` (``:111-211``): x y z (float) + r g b (uchar), ascii or binary (15-byte vertices)."""
    data = Path(path).read_bytes()
    end = data.find(b"end_header\n")
    if end < 0:
        raise ValueError("No end_header")
    end += len(b"end_header\n")
    header = data[:end].decode("ascii")
    vline = next((l for l in header.split("\n") if l.startswith("element vertex")), None)
    if vline is None:
        raise ValueError("No vertex count")
    n = int(vline.split(" ")[-1])
    if "format ascii" in header:
        rows = [l.split() for l in data[end:].decode("ascii").splitlines()[:n]]
        rows = [r for r in rows if len(r) >= 6]
        xyz = np.array([[float(v) for v in r[:3]] for r in rows], np.float32).reshape(-1, 3)
        rgb = np.array([[int(v) for v in r[3:6]] for r in rows], np.uint8).reshape(-1, 3)
        return xyz, rgb
    if end + n * 15 > len(data):
        raise ValueError("File too small")
    rec = np.frombuffer(data, dtype=np.dtype([("xyz", "<f4", 3), ("rgb", "u1", 3)]), count=n, offset=end)
    return rec["xyz"].astype(np.float32), rec["rgb"].copy()


def opengl_to_opencv_c2w(c2w: np.ndarray) -> np.ndarray:
    """``:352-361``: invert, negate rows 1 and 2 of the world-to-camera matrix, invert back."""
    w2c = np.linalg.inv(np.asarray(c2w, np.float64))
    w2c[1:3, :] *= -1.0
    return np.linalg.inv(w2c)


class NerfStudioDataLoader:
    def __init__(self, root):
        self.root = _maybe_unzip(Path(root))

    def _meta(self):
        return json.loads((self.root / "transforms.json").read_text())

    @staticmethod
    def _intrinsic(d) -> Optional[np.ndarray]:
        if all(d.get(k) is not None for k in ("fl_x", "fl_y", "cx", "cy")):
            return np.array([[d["fl_x"], 0, d["cx"]], [0, d["fl_y"], d["cy"]], [0, 0, 1]], np.float64)
        return None

    def getOriginalImageSize(self) -> Tuple[int, int]:
        from PIL import Image
        meta = self._meta()
        with Image.open(self.root / meta["frames"][0]["file_path"]) as im:
            return im.width, im.height

    def load(self, resizeFactor: float = 1.0, whiteBackground: bool = False):
        meta = self._meta()
        xyz, rgb = parse_point_ply(self.root / meta["ply_file_path"])
        ch = rgb.astype(np.float32) / np.float32(255.0)
        pc = PointCloud(xyz, {"R": ch[:, 0], "G": ch[:, 1], "B": ch[:, 2]})
        K, c2w, rgba = [], [], []
        for fr in meta["frames"]:
            k = self._intrinsic(fr)
            k = k if k is not None else self._intrinsic(meta)
            if k is None:
                raise ValueError("Failed to load intrinsic matrix")
            if resizeFactor != 1.0:
                k[:2, :3] *= resizeFactor
            K.append(k)
            rgba.append(_load_rgba(self.root / fr["file_path"], resizeFactor))
            c2w.append(opengl_to_opencv_c2w(np.asarray(fr["transform_matrix"], np.float32)))
        rgbs = np.stack([a[..., :3] for a in rgba])
        alphas = np.stack([a[..., 3] for a in rgba])
        data = LoadedTrainData(Hs=[a.shape[0] for a in rgba], Ws=[a.shape[1] for a in rgba], intrinsicArray=np.stack(K),
                               c2wArray=np.stack(c2w), rgbArray=_compose(rgbs, alphas, whiteBackground)[..., :3], alphaArray=alphas)
        return data, pc, _tile_size(data)


# ------------------------------------------------------------------------------------------------
# Blender demo format (info.json + depth / alpha maps)
# ------------------------------------------------------------------------------------------------
def get_rays_from_images(H: int, W: int, intrinsics: np.ndarray, c2w: np.ndarray, renderStride: int = 1):
    """``getRaysFromImages`` (``PointCloudUtil.swift:50-99``): pixel (u, v, 1) at integer coordinates."""
    u = np.arange(0, W, renderStride, dtype=np.float32)
    v = np.arange(0, H, renderStride, dtype=np.float32)
    ug, vg = np.meshgrid(u, v, indexing="xy")
    pix = np.stack([ug.reshape(-1), vg.reshape(-1), np.ones(ug.size, np.float32)], axis=0)   # [3, HW]
    invK = np.linalg.inv(intrinsics[:, :3, :3].astype(np.float64))
    rays_d = np.einsum("bij,bjk,kn->bni", c2w[:, :3, :3].astype(np.float64), invK, pix.astype(np.float64)).astype(np.float32)
    rays_o = np.repeat(c2w[:, None, :3, 3].astype(np.float32), rays_d.shape[1], axis=1)
    return rays_o, rays_d


def point_cloud_from_train_data(data: LoadedTrainData) -> PointCloud:
    """``getPointCloudsFromTrainData`` (``:101-138``): un-project every pixel with alpha == 1 along its ray by its depth."""
    if data.depthArray is None:
        raise ValueError("unexpected nil depth")
    H, W = int(data.Hs[0]), int(data.Ws[0])
    rays_o, rays_d = get_rays_from_images(H, W, data.intrinsicArray, data.c2wArray)
    B = rays_o.shape[0]
    pts = rays_o + rays_d * data.depthArray.reshape(B, -1, 1)
    mask = data.alphaArray.reshape(-1) == 1.0
    rgb = np.asarray(data.rgbArray, np.float32).reshape(-1, 3)[mask]
    return PointCloud(pts.reshape(-1, 3)[mask], {"R": rgb[:, 0], "G": rgb[:, 1], "B": rgb[:, 2], "A": data.alphaArray.reshape(-1)[mask]})


class BlenderDemoDataLoader:
    def __init__(self, root):
        self.root = _maybe_unzip(Path(root))

    def _info(self):
        return json.loads((self.root / "info.json").read_text())

    def getOriginalImageSize(self) -> Tuple[int, int]:
        hw = self._info()["images"][0]["HW"]
        return int(hw[1]), int(hw[0])

    def load(self, resizeFactor: float = 1.0, whiteBackground: bool = False):
        info = self._info()
        images = info["images"]
        max_depth = float(images[0].get("max_depth", 1.0)) if images else 1.0
        K, c2w, rgbs, alphas, depths = [], [], [], [], []
        for im in images:
            path = self.root / im["rgb"]
            base = path.stem.split("_")[0]
            rgb = _load_rgba(path, resizeFactor)[..., :3]      # drawn premultiplied like the COLMAP loader (:157-196)
            depths.append(_load_gray(path.parent / f"{base}_depth.png", resizeFactor) * np.float32(max_depth))
            alphas.append(_load_gray(path.parent / f"{base}_alpha.png", resizeFactor))
            rgbs.append(rgb)
            k = np.eye(4, dtype=np.float64)
            k[:3, :3] = np.asarray(im["intrinsic"], np.float64)[:3, :3]
            if resizeFactor != 1.0:
                k[:2, :3] *= resizeFactor
            K.append(k)
            c2w.append(opengl_to_opencv_c2w(np.asarray(im["pose"], np.float64)))
        rgbs, alphas = np.stack(rgbs), np.stack(alphas)
        data = LoadedTrainData(Hs=[r.shape[0] for r in rgbs], Ws=[r.shape[1] for r in rgbs], intrinsicArray=np.stack(K),
                               c2wArray=np.stack(c2w), rgbArray=_compose(rgbs, alphas, whiteBackground)[..., :3], alphaArray=alphas,
                               depthArray=np.stack(depths))
        pc = point_cloud_from_train_data(data)
        return data, pc, _tile_size(data)


# ------------------------------------------------------------------------------------------------
# model initialisation from a point cloud
# ------------------------------------------------------------------------------------------------
def distTopK(X: np.ndarray, k: int) -> np.ndarray:
    """``GaussianModel.swift:11-31``: mean squared distance to the k nearest points (the point itself included), computed
    for chunks of 256 rows — with the reference's loop bounds ``stride(from: 0, to: N / 256 + 1, by: 256)``, i.e. only
    rows ``i .. i + 256`` for ``i = 0, 256, ... < N / 256 + 1``; every other row keeps 0 (then clamped to 1e-7 by the
    caller).  Reproduced as is."""
    X = np.asarray(X, np.float32)
    n = X.shape[0]
    out = np.zeros(n, np.float32)
    chunk = 1 << 8
    for i in range(0, n // chunk + 1, chunk):
        a = X[i:i + chunk]
        if a.shape[0] == 0:
            continue
        d2 = ((a[:, None, :] - X[None, :, :]) ** 2).sum(-1)
        kk = min(k, n)
        out[i:i + chunk] = np.sort(d2, axis=1)[:, :kk].mean(axis=1)
    return out


def create_from_pcd(pcd: PointCloud, sh_degree: int = 3) -> GaussModel:
    """``GaussModel.create_from_pcd`` (``GaussianModel.swift:87-125``)."""
    points = pcd.coords
    colors = pcd.select_channels(["R", "G", "B"]) / np.float32(255.0)
    n = points.shape[0]
    K = (sh_degree + 1) ** 2
    features = np.zeros((n, 3, K), np.float32)
    features[:, :3, 0] = RGB2SH(colors)
    dist2 = np.maximum(distTopK(points, 3), np.float32(1e-7))
    scales = np.repeat(np.log(np.sqrt(dist2)).reshape(n, 1), 3, axis=1).astype(np.float32)
    rots = np.zeros((n, 4), np.float32)
    rots[:, 0] = 1.0
    opac = inverse_sigmoid(np.float32(0.1) * np.ones((n, 1), np.float32))
    params = {"_xyz": points.astype(np.float32), "_features_dc": np.ascontiguousarray(features[:, :, 0:1].transpose(0, 2, 1)),
              "_features_rest": np.ascontiguousarray(features[:, :, 1:].transpose(0, 2, 1)), "_scales": scales, "_rotation": rots,
              "_opacity": opac.astype(np.float32)}
    return GaussModel.from_arrays(params, sh_degree)


def createModel(sh_degree: int, pointCloud: PointCloud, sampleCount: int, rng: Optional[np.random.Generator] = None) -> GaussModel:
    """``GaussianTrainer.createModel(sh_degree:pointCloud:sampleCount:)`` (``GaussianTrainer.swift:1130``)."""
    return create_from_pcd(pointCloud.randomSample(sampleCount, rng), sh_degree)

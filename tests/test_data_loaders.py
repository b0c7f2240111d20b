"""Dataset loaders and point-cloud initialisation on fixtures written by the tests themselves
(the reference downloads its demo data at run time; there is no network here).  Reference: Data/ColmapDataLoader.swift,
Data/NerfStudioDataLoader.swift, Data/BlenderDataLoader.swift, Trainer/PointCloudUtil.swift, Trainer/GaussianModel.swift."""
import json
import struct

import numpy as np
import pytest
from PIL import Image

from gaussiansplattingmlx_b200 import data as D
from gaussiansplattingmlx_b200.scene import make_cameras


def _rot_to_quat(R):
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    x = (R[2, 1] - R[1, 2]) / (4 * w); y = (R[0, 2] - R[2, 0]) / (4 * w); z = (R[1, 0] - R[0, 1]) / (4 * w)
    return np.array([w, x, y, z])


def _write_colmap(root, c2ws, W, H, fx, fy, pts, cols, model=1):
    b = root / "colmap" / "sparse" / "0"
    b.mkdir(parents=True)
    (root / "images").mkdir()
    params = {0: [fx, W / 2, H / 2], 1: [fx, fy, W / 2, H / 2], 2: [fx, W / 2, H / 2, 0.01], 3: [fx, fy, W / 2, H / 2, 0.01, 0.0, 0.0, 0.0]}[model]
    cam = struct.pack("<Q", 1) + struct.pack("<IiQQ", 7, model, W, H) + struct.pack(f"<{len(params)}d", *params)
    (b / "cameras.bin").write_bytes(cam)
    img = struct.pack("<Q", len(c2ws))
    rng = np.random.default_rng(0)
    for i, c2w in enumerate(c2ws):
        w2c = np.linalg.inv(c2w)
        q, t = _rot_to_quat(w2c[:3, :3]), w2c[:3, 3]
        name = f"im_{i}.png"
        img += struct.pack("<I", i + 1) + struct.pack("<4d", *q) + struct.pack("<3d", *t) + struct.pack("<I", 7) + name.encode() + b"\0"
        img += struct.pack("<Q", 2) + struct.pack("<ddQ", 1.0, 2.0, 5) * 2                 # two 2-D points to skip
        a = (rng.random((H, W, 4)) * 255).astype(np.uint8); a[..., 3] = 255; a[0, 0, 3] = 128
        Image.fromarray(a, "RGBA").save(root / "images" / name)
    (b / "images.bin").write_bytes(img)
    p3 = struct.pack("<Q", len(pts))
    for i, (p, c) in enumerate(zip(pts, cols)):
        p3 += struct.pack("<Q", i) + struct.pack("<3d", *p) + struct.pack("<3B", *c) + struct.pack("<d", 0.5) + struct.pack("<Q", 1) + struct.pack("<II", 1, 0)
    (b / "points3D.bin").write_bytes(p3)


def _cams(W, H, n):
    cams = make_cameras(W, H, n)
    return [np.linalg.inv(c.worldViewTransform.astype(np.float64).T) for c in cams], float(cams[0].focalX)


@pytest.mark.parametrize("model", [0, 1, 2, 3])
def test_colmap_loader_round_trip(tmp_path, model):
    W, H = 32, 24
    c2ws, f = _cams(W, H, 3)
    rng = np.random.default_rng(1)
    pts = rng.normal(size=(50, 3)); cols = rng.integers(0, 256, (50, 3))
    _write_colmap(tmp_path, c2ws, W, H, f, f * 1.1, pts, cols, model)
    ld = D.ColmapDataLoader(tmp_path)
    assert ld.getOriginalImageSize() == (W, H)
    data, pc, tile = ld.load(resizeFactor=1.0, whiteBackground=False)
    assert tile == (W // 4, H // 4) and data.getNumCameras() == 3
    assert np.allclose(data.c2wArray, np.stack(c2ws), atol=1e-5)
    assert data.rgbArray.shape == (3, H, W, 3) and data.alphaArray.shape == (3, H, W)
    fy = f if model in (0, 2) else f * 1.1
    assert np.allclose(data.intrinsicArray[0], [[f, 0, W / 2], [0, fy, H / 2], [0, 0, 1]], rtol=1e-6)
    assert np.allclose(pc.coords, pts, atol=1e-6) and np.array_equal(pc.select_channels(["R", "G", "B"]), cols.astype(np.float32))
    cam = data.getViewPointCamera(1)
    assert cam.imageWidth == W and abs(float(cam.focalX) - f) < 1e-3
    # premultiplied decode: the half-transparent pixel is darker than its stored colour
    raw = np.asarray(Image.open(tmp_path / "images" / "im_0.png"), np.float32) / 255
    assert np.allclose(data.rgbArray[0, 0, 0], raw[0, 0, :3] * raw[0, 0, 3], atol=1e-6)
    # white background composite and resize (intrinsics scale with the image)
    d2, _, tile2 = ld.load(resizeFactor=0.5, whiteBackground=True)
    assert d2.rgbArray.shape == (3, H // 2, W // 2, 3) and tile2 == (W // 8, H // 8)
    assert np.allclose(d2.intrinsicArray[0][:2], data.intrinsicArray[0][:2] * 0.5)
    assert float(d2.rgbArray.min()) >= 0.0 and float(d2.rgbArray.max()) <= 1.0 + 1e-6


def test_colmap_errors(tmp_path):
    with pytest.raises(FileNotFoundError, match="Colmap files missing"):
        D.ColmapDataLoader(tmp_path).load()
    b = tmp_path / "colmap" / "sparse" / "0"; b.mkdir(parents=True)
    (b / "cameras.bin").write_bytes(struct.pack("<Q", 1) + b"\x01\x02")
    (b / "images.bin").write_bytes(struct.pack("<Q", 0))
    with pytest.raises(ValueError, match="Not enough data"):
        D.ColmapDataLoader(tmp_path).load()


@pytest.mark.parametrize("binary", [False, True])
def test_nerfstudio_loader(tmp_path, binary):
    W, H = 20, 16
    c2ws_cv, f = _cams(W, H, 2)
    # store OpenGL-convention matrices: c2w_gl = c2w_cv with the y and z camera axes flipped
    flip = np.diag([1.0, -1.0, -1.0, 1.0])
    frames = []
    for i, c in enumerate(c2ws_cv):
        Image.fromarray((np.random.default_rng(i).random((H, W, 3)) * 255).astype(np.uint8)).save(tmp_path / f"f{i}.png")
        fr = {"file_path": f"f{i}.png", "transform_matrix": (c @ flip).tolist()}
        if i == 1:
            fr.update(fl_x=f * 2, fl_y=f * 2, cx=1.0, cy=2.0)               # per-frame intrinsics win over the global ones
        frames.append(fr)
    (tmp_path / "transforms.json").write_text(json.dumps({"frames": frames, "ply_file_path": "pts.ply", "fl_x": f, "fl_y": f,
                                                          "cx": W / 2, "cy": H / 2, "w": W, "h": H}))
    pts = np.random.default_rng(3).normal(size=(7, 3)).astype(np.float32)
    cols = np.random.default_rng(4).integers(0, 256, (7, 3)).astype(np.uint8)
    hdr = f"ply\nformat {'binary_little_endian' if binary else 'ascii'} 1.0\nelement vertex 7\nproperty float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n"
    if binary:
        body = b"".join(struct.pack("<3f3B", *p, *c) for p, c in zip(pts, cols))
    else:
        body = "".join(f"{p[0]:.9g} {p[1]:.9g} {p[2]:.9g} {c[0]} {c[1]} {c[2]}\n" for p, c in zip(pts, cols)).encode()
    (tmp_path / "pts.ply").write_bytes(hdr.encode() + body)
    data, pc, tile = D.NerfStudioDataLoader(tmp_path).load()
    assert np.allclose(data.c2wArray, np.stack(c2ws_cv), atol=1e-5), "OpenGL -> OpenCV conversion"
    assert np.allclose(pc.coords, pts, atol=1e-6) and np.array_equal(pc.select_channels(["R", "G", "B"]), cols.astype(np.float32))
    assert abs(data.intrinsicArray[0][0, 0] - f) < 1e-3 and abs(data.intrinsicArray[1][0, 0] - 2 * f) < 1e-3
    assert tile == (W // 4, H // 4) and np.all(data.alphaArray == 1.0)


def test_blender_loader_unprojects_depth(tmp_path):
    W, H = 16, 12
    c2ws_cv, f = _cams(W, H, 2)
    flip = np.diag([1.0, -1.0, -1.0, 1.0])
    K = [[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]]
    images = []
    for i, c in enumerate(c2ws_cv):
        Image.fromarray(np.full((H, W, 3), 200, np.uint8)).save(tmp_path / f"{i:03d}_rgb.png")
        depth = np.full((H, W), 128, np.uint8)                        # 128/255 * max_depth
        alpha = np.zeros((H, W), np.uint8); alpha[2:6, 3:9] = 255
        Image.fromarray(depth, "L").save(tmp_path / f"{i:03d}_depth.png")
        Image.fromarray(alpha, "L").save(tmp_path / f"{i:03d}_alpha.png")
        images.append({"intrinsic": K, "pose": (c @ flip).tolist(), "rgb": f"{i:03d}_rgb.png", "depth": f"{i:03d}_depth.png",
                       "alpha": f"{i:03d}_alpha.png", "max_depth": 8.0, "HW": [H, W]})
    (tmp_path / "info.json").write_text(json.dumps({"backend": "x", "light_mode": "y", "fast_mode": False, "format_version": 1,
                                                    "channels": [], "scale": 1.0, "images": images, "bbox": [[0, 0, 0], [1, 1, 1]]}))
    ld = D.BlenderDemoDataLoader(tmp_path)
    assert ld.getOriginalImageSize() == (W, H)
    data, pc, tile = ld.load()
    assert data.depthArray.shape == (2, H, W) and abs(float(data.depthArray[0, 0, 0]) - 128 / 255 * 8.0) < 1e-5
    assert pc.coords.shape[0] == 2 * 4 * 6, "one point per pixel with alpha == 1"
    # every point lies at depth z = 128/255*8 in front of its camera
    cam0 = np.linalg.inv(c2ws_cv[0])
    z = (cam0[:3, :3] @ pc.coords[:24].T.astype(np.float64) + cam0[:3, 3:4])[2]
    assert np.allclose(z, 128 / 255 * 8.0, atol=1e-4)
    assert np.allclose(pc.select_channels(["R"]), 200.0)


def test_point_cloud_centering_and_sampling():
    rng = np.random.default_rng(5)
    pts = rng.normal(size=(2000, 3)).astype(np.float32) + np.float32([5, -3, 2])
    pts[0] = [500, 0, 0]                                              # an outlier
    ch = {k: rng.random(2000).astype(np.float32) for k in "RGB"}
    pc = D.PointCloud(pts.copy(), ch)
    data = D.LoadedTrainData([8], [8], np.eye(3)[None], np.eye(4)[None], np.zeros((1, 8, 8, 3), np.float32), np.ones((1, 8, 8), np.float32))
    center = pts.mean(axis=0)
    pc.centering(data)
    assert np.allclose(data.c2wArray[0, :3, 3], -center, atol=1e-4)
    assert pc.coords.shape[0] < 2000 and np.abs(pc.coords).max() < 200 and pc.channels["R"].shape[0] == pc.coords.shape[0]
    s = pc.randomSample(100, np.random.default_rng(1))
    assert s.coords.shape == (100, 3) and s.channels["G"].shape == (100,)
    assert pc.randomSample(10 ** 6) is pc


def test_create_from_pcd_matches_reference_semantics():
    rng = np.random.default_rng(6)
    n = 700
    pc = D.PointCloud(rng.normal(size=(n, 3)).astype(np.float32), {k: rng.random(n).astype(np.float32) for k in "RGB"})
    m = D.create_from_pcd(pc, sh_degree=3)
    assert m._xyz.shape == (n, 3) and m._features_dc.shape == (n, 1, 3) and m._features_rest.shape == (n, 15, 3)
    col = np.round(np.stack([pc.channels[k] for k in "RGB"], -1) * 255) / 255
    assert np.allclose(m._features_dc[:, 0], (col - 0.5) / 0.28209479177387814, atol=1e-6)
    assert not m._features_rest.any() and np.array_equal(m._rotation, np.tile(np.float32([1, 0, 0, 0]), (n, 1)))
    assert np.allclose(m._opacity, np.log(0.1 / 0.9), atol=1e-6)
    # distTopK quirk: with the reference's loop bounds only the first 256 rows get a k-NN scale, the rest the 1e-7 floor
    x = pc.coords
    d2 = ((x[:256, None] - x[None]) ** 2).sum(-1)
    want = np.sort(d2, axis=1)[:, :3].mean(axis=1)
    assert np.allclose(m._scales[:256, 0], np.log(np.sqrt(np.maximum(want, 1e-7))), atol=1e-5)
    assert np.allclose(m._scales[256:], np.log(np.sqrt(np.float32(1e-7))))
    assert np.array_equal(m._scales[:, 0], m._scales[:, 1]) and np.array_equal(m._scales[:, 0], m._scales[:, 2])
    m2 = D.createModel(3, pc, 128, np.random.default_rng(0))
    assert m2._xyz.shape == (128, 3)


@pytest.mark.gpu
def test_colmap_dataset_to_training_end_to_end(tmp_path):
    """The app's path (UI/TrainView.swift:141-205): load → centering → createModel → GaussianRenderer(TILE_SIZE W/4 x H/4)
    → GaussianTrainer.startTrain, on a COLMAP fixture."""
    import torch
    from gaussiansplattingmlx_b200.renderer import GaussianRenderer
    from gaussiansplattingmlx_b200.trainer import GaussianTrainer
    W, H = 64, 48
    c2ws, f = _cams(W, H, 4)
    rng = np.random.default_rng(2)
    pts = rng.normal(scale=0.5, size=(600, 3)); cols = rng.integers(0, 256, (600, 3))
    _write_colmap(tmp_path, c2ws, W, H, f, f, pts, cols)
    data, pc, tile = D.ColmapDataLoader(tmp_path).load(resizeFactor=1.0, whiteBackground=False)
    pc.centering(data)
    model = D.createModel(3, pc, 512, np.random.default_rng(0))
    r = GaussianRenderer(active_sh_degree=3, W=W, H=H, TILE_SIZE=tile, whiteBackground=False)
    tr = GaussianTrainer(model, data, r, iterationCount=30, seed=1)
    before = model._xyz.copy()
    tr.startTrain(earlyStoppingThreshold=-1.0)
    assert len(tr.losses) == 5 and all(np.isfinite(l) for l in tr.losses)
    assert tr.losses[-1] < tr.losses[0], "30 Adam iterations on 4 views must lower the loss"
    assert model._xyz.shape == before.shape and not np.array_equal(model._xyz, before)

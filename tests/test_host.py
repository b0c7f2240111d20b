"""Host-side logic: camera conventions, seeded workloads, learning-rate schedule, view sharding."""
import numpy as np
import pytest

from gaussiansplattingmlx_b200.camera import Camera, focal2fov, get_projection_matrix
from gaussiansplattingmlx_b200.dp import ViewParallel, grad_scale, shard_views
from gaussiansplattingmlx_b200.model import GaussModel
from gaussiansplattingmlx_b200.scene import WORKLOADS, make_cameras, make_gaussians, make_targets, make_workload


def test_camera_row_vector_convention():
    cam = make_cameras(64, 48, 5)[2]
    c2w = np.linalg.inv(cam.worldViewTransform.astype(np.float64).T)
    # the camera centre maps to the view-space origin:  [c 1] @ V = [0 0 0 1]
    pv = np.append(cam.cameraCenter, 1.0) @ cam.worldViewTransform.astype(np.float64)
    assert np.abs(pv - np.array([0, 0, 0, 1.0])).max() < 1e-5
    # the origin (looked at) lies on +z at distance |eye|
    o = np.array([0, 0, 0, 1.0]) @ cam.worldViewTransform.astype(np.float64)
    assert abs(o[0]) < 1e-5 and abs(o[1]) < 1e-5 and abs(o[2] - np.linalg.norm(cam.cameraCenter)) < 1e-5
    assert np.abs(c2w[:3, 3] - cam.cameraCenter).max() < 1e-5


def test_projection_matrix_entries():
    cam = make_cameras(64, 48, 1)[0]
    P = cam.projectionMatrix.astype(np.float64).T          # stored as P^T row-major
    assert abs(P[0, 0] - 1.0 / np.tan(float(cam.FoVx) / 2)) < 1e-5
    assert abs(P[1, 1] - 1.0 / np.tan(float(cam.FoVy) / 2)) < 1e-5
    assert abs(P[2, 2] - 100.0 / 99.9) < 1e-6 and P[3, 2] == 1.0 and abs(P[2, 3] + 10.0 / 99.9) < 1e-6
    assert focal2fov(np.float32(32.0), 64.0).dtype == np.float32
    assert abs(float(focal2fov(32.0, 64.0)) - np.pi / 2) < 1e-6
    assert cam.pack().shape == (39,) and cam.pack().dtype == np.float32


def test_workloads_are_the_baseline_configs():
    assert (WORKLOADS["C1"].n_gaussians, WORKLOADS["C1"].width) == (1000, 64)
    assert (WORKLOADS["C2"].n_gaussians, WORKLOADS["C2"].width, WORKLOADS["C2"].height) == (300_000, 800, 800)
    c3 = WORKLOADS["C3"]
    assert (c3.n_gaussians, c3.width, c3.height, c3.views, c3.sh_degree) == (1_000_000, 1920, 1080, 8, 3)


def test_scene_generator_is_deterministic_and_shaped():
    a = make_gaussians(50, 7, 3); b = make_gaussians(50, 7, 3)
    for k in a:
        assert a[k].dtype == np.float32 and np.array_equal(a[k], b[k])
    assert a["_features_rest"].shape == (50, 15, 3) and a["_features_dc"].shape == (50, 1, 3)
    assert make_gaussians(5, 7, 4)["_features_rest"].shape == (5, 24, 3)
    t = make_targets(8, 4, 2, 3)
    assert t[0].shape == (4, 8, 3) and not np.array_equal(t[0], t[1])
    wl, p, cams, tg = make_workload("C1")
    assert len(cams) == 1 and tg[0].shape == (64, 64, 3) and p["_xyz"].shape == (1000, 3)


def test_learning_rates():
    lr0 = GaussModel.getLearningRates(0, 1000)
    assert abs(lr0[0] - 1.6e-4) < 1e-10 and lr0[1:] == [0.0025, float(np.float32(0.0025) / np.float32(20)), 0.005, 0.001, 0.025]
    assert abs(GaussModel.getLearningRates(500, 1000)[0] - 0.8e-4) < 1e-9
    assert abs(GaussModel.getLearningRates(999_999, 1_000_000)[0] - 1.6e-6) < 1e-9     # floor at 1 %
    from oracle.pipeline import learning_rates
    assert learning_rates(123, 1000) == GaussModel.getLearningRates(123, 1000)


def test_model_container():
    p = make_gaussians(10, 1, 3)
    m = GaussModel.from_arrays(p, 3)
    assert [x.shape for x in m.getParams()] == [(10, 3), (10, 1, 3), (10, 15, 3), (10, 3), (10, 4), (10, 1)]


def test_shard_views_partitions():
    for world in (1, 2, 4, 8):
        seen = sorted(v for r in range(world) for v in shard_views(8, r, world))
        assert seen == list(range(8))
        assert all(len(shard_views(8, r, world)) == 8 // world for r in range(world))
    assert shard_views(3, 3, 4) == []
    assert grad_scale(8) == 0.125
    with pytest.raises(ValueError):
        shard_views(8, 2, 2)
    with pytest.raises(ValueError):
        grad_scale(0)
    assert ViewParallel().all_reduce_max(3.0) == 3.0

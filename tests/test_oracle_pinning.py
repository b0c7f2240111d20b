"""Pins the CPU oracle (oracle/gsb_oracle.c, the "port") before anything is checked against it:

1. the reference's own known-answer unit tests that touch the hot path
   (GaussianSplattingMlxTests/ShUtilsTests.swift:15-151 — SH polynomial on UN-normalised directions;
   GaussianSplattingMlxTests.swift:73-130 — build_rotation / build_scaling_rotation);
2. the committed golden fixtures tests/golden/*.npz, which are outputs of the reference's own shipped
   kernels (oracle/_ref) on seeded scenes (tests/golden/make_golden.py);
3. when oracle/_ref/libgsref.so is present (container with /root/reference), a live port-vs-reference diff.
"""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from gaussiansplattingmlx_b200.camera import Camera
from gaussiansplattingmlx_b200.scene import make_cameras, make_gaussians, make_targets
from oracle import pipeline as pl

GOLDEN = Path(__file__).resolve().parent / "golden"

# Trainer/ShUtils.swift:4-32
C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]
C4 = [2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892, 0.10578554691520431,
      -0.6690465435572892, 0.47308734787878004, -1.7701307697799304, 0.6258357354491761]


def expected_sh(deg, sh, d):
    """The hand-expanded polynomial of ShUtilsTests.swift, in float64."""
    x, y, z = d
    r = C0 * sh[0]
    if deg > 0:
        r += -C1 * y * sh[1] + C1 * z * sh[2] - C1 * x * sh[3]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    if deg > 1:
        r += C2[0] * xy * sh[4] + C2[1] * yz * sh[5] + C2[2] * (2 * zz - xx - yy) * sh[6] + C2[3] * xz * sh[7] + C2[4] * (xx - yy) * sh[8]
    if deg > 2:
        r += (C3[0] * y * (3 * xx - yy) * sh[9] + C3[1] * xy * z * sh[10] + C3[2] * y * (4 * zz - xx - yy) * sh[11]
              + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[12] + C3[4] * x * (4 * zz - xx - yy) * sh[13]
              + C3[5] * z * (xx - yy) * sh[14] + C3[6] * x * (xx - 3 * yy) * sh[15])
    if deg > 3:
        r += (C4[0] * xy * (xx - yy) * sh[16] + C4[1] * yz * (3 * xx - yy) * sh[17] + C4[2] * xy * (7 * zz - 1) * sh[18]
              + C4[3] * yz * (7 * zz - 3) * sh[19] + C4[4] * (zz * (35 * zz - 30) + 3) * sh[20] + C4[5] * xz * (7 * zz - 3) * sh[21]
              + C4[6] * (xx - yy) * (7 * zz - 1) * sh[22] + C4[7] * xz * (xx - 3 * yy) * sh[23]
              + C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy)) * sh[24])
    return r


# the exact vectors of ShUtilsTests.swift:15-151
SH_CASES = [
    (0, [1.0], (0.0, 0.0, 1.0)),
    (0, [-0.5], (0.1, 0.2, 0.3)),
    (1, [1.0, 0.2, 0.3, 0.4], (1.0, 2.0, 3.0)),
    (2, [0.5, 0.2, -0.1, 0.1, 1.0, -1.0, 2.0, 0.5, -2.0], (0.5, -1.0, 2.0)),
    (3, [0.1 * (i + 1) for i in range(16)], (0.1, -0.2, 0.3)),
    (4, [0.1 * (i + 1) for i in range(25)], (-0.3, 0.2, 0.7)),
]


@pytest.mark.parametrize("deg,sh,d", SH_CASES)
def test_sh_known_answers_basis(port, deg, sh, d):
    basis = np.zeros(25, np.float32)
    port.lib.gso_sh_basis(C.c_float(d[0]), C.c_float(d[1]), C.c_float(d[2]), deg, basis.ctypes.data_as(C.c_void_p))
    got = float(np.dot(basis[:len(sh)].astype(np.float64), np.array(sh, np.float64)))
    assert abs(got - expected_sh(deg, sh, d)) < 2e-6


def _color_through_k1(o, deg, sh, d):
    """Push the same vector through K1: colour = max(evalSh + 0.5, 0) with dir = mean - cameraCenter."""
    K = (deg + 1) ** 2
    cam = make_cameras(32, 32, 1)[0]
    cc = cam.cameraCenter.astype(np.float32).astype(np.float64)
    act = {"means3d": (cc + np.array(d)).astype(np.float32).reshape(1, 3), "scales": np.full((1, 3), 0.01, np.float32),
           "rotations": np.array([[1, 0, 0, 0]], np.float32), "shs": np.zeros((1, K, 3), np.float32)}
    act["shs"][0, :, 1] = np.array(sh, np.float32)
    # direction actually seen by the kernel (f32 subtraction)
    d_seen = (act["means3d"][0] - cam.camera_center_f32()[0]).astype(np.float64)
    proj = o.project_fwd(act, cam, deg)
    return float(proj["color"][0, 1]), max(expected_sh(deg, sh, d_seen) + 0.5, 0.0), float(proj["color"][0, 0])


@pytest.mark.parametrize("deg,sh,d", SH_CASES)
def test_sh_known_answers_through_projection(port, deg, sh, d):
    got, want, other = _color_through_k1(port, deg, sh, d)
    assert abs(got - want) < 5e-6 and abs(other - 0.5) < 1e-7


@pytest.mark.parametrize("deg,sh,d", SH_CASES)
def test_sh_known_answers_reference_kernels(ref, deg, sh, d):
    got, want, _ = _color_through_k1(ref, deg, sh, d)
    assert abs(got - want) < 5e-6


def _scaling_rotation(port, s, q):
    L = np.zeros(9, np.float32); cov = np.zeros(9, np.float32)
    port.lib.gso_build_scaling_rotation(np.array(s, np.float32).ctypes.data_as(C.c_void_p),
                                        np.array(q, np.float32).ctypes.data_as(C.c_void_p),
                                        L.ctypes.data_as(C.c_void_p), cov.ctypes.data_as(C.c_void_p))
    return L.reshape(3, 3), cov.reshape(3, 3)


def test_build_rotation_known_answers(port):
    # GaussianSplattingMlxTests.swift:73-108: quaternion order is (w, x, y, z)
    L, _ = _scaling_rotation(port, (1, 1, 1), (1, 0, 0, 0))
    assert np.allclose(L, np.eye(3), atol=1e-5)
    L, _ = _scaling_rotation(port, (1, 1, 1), (0, 1, 0, 0))      # 180 deg about X
    assert np.allclose(L, np.diag([1.0, -1.0, -1.0]), atol=1e-5)


def test_build_scaling_rotation_known_answer(port):
    # GaussianSplattingMlxTests.swift:110-130
    L, cov = _scaling_rotation(port, (2, 2, 2), (1, 0, 0, 0))
    assert np.allclose(L, 2 * np.eye(3), atol=1e-5)
    assert np.allclose(cov, 4 * np.eye(3), atol=1e-5)


def test_ssim_identical_images_is_one(port):
    # TrainTests.swift:55 prints ssim(identical) — the value the metric defines is exactly 1
    from oracle.api import ssim_window
    img = np.random.default_rng(0).random((20, 24, 3), dtype=np.float32)
    _, w = ssim_window()
    s = port.ssim_fwd(img, img, w)
    assert np.abs(s["ssim"] - 1.0).max() < 1e-5
    g, w2 = ssim_window()
    assert abs(float(g.sum()) - 1.0) < 1e-6 and g[5] == g[6] and g[0] != g[10]    # centre 5.5 → asymmetric


# ------------------------------------------------------------------------------------------------
# golden fixtures produced by the reference kernels
# ------------------------------------------------------------------------------------------------
def load_case(name):
    g = np.load(GOLDEN / f"{name}.npz")
    n, W, H, seed, degree, view, ncams = (int(x) for x in g["meta"])
    params = make_gaussians(n, seed, degree)
    cam = make_cameras(W, H, ncams)[view]
    target = make_targets(W, H, 1, seed)[0]
    return g, params, cam, target, degree


@pytest.mark.parametrize("name", ["c1", "deg4_ragged"])
def test_port_reproduces_reference_golden(port, name):
    g, params, cam, target, degree = load_case(name)
    fr, lo, bw = pl.loss_and_grads(port, params, cam, target, degree)
    b = fr["bins"]
    assert b["M"] == int(g["M"][0])
    for k in ("tilesTouched", "tileCounts", "sortedKeysHigh", "sortedKeysLow", "sortedGaussIdx"):
        assert np.array_equal(b[k], g[k]), k                         # bit-exact integer work
    for k in ("means2d", "depths", "radii", "conic", "color"):
        assert np.array_equal(fr["proj"][k].view(np.uint32), g[k].view(np.uint32)), k
    assert np.array_equal(fr["render"].view(np.uint32), g["render"].view(np.uint32))
    assert np.array_equal(fr["fwd"]["lastContrib"], g["lastContrib"])
    assert abs(lo["loss"] - float(g["loss"][0])) < 1e-7
    assert np.abs(lo["ssim"]["ssim"] - g["ssim_map"]).max() < 1e-6
    for k in params:
        a, r = bw["grads"][k], g["grad" + k]
        assert np.abs(a - r).max() / max(np.abs(r).max(), 1e-8) < 1e-5, k


def test_c1_matches_survey_workload_numbers():
    g = np.load(GOLDEN / "c1.npz")
    assert int(g["M"][0]) == 3359                     # SURVEY.md 8d-workload, C1 seed 1
    assert int(g["tileCounts"].max()) == 491


@pytest.mark.parametrize("seed,n,W,H,degree", [(41, 300, 48, 32, 3), (42, 500, 40, 56, 4), (43, 200, 33, 17, 1)])
def test_port_vs_reference_live(port, ref, seed, n, W, H, degree):
    params = make_gaussians(n, seed, degree)
    cam = make_cameras(W, H, 4)[seed % 4]
    target = make_targets(W, H, 1, seed)[0]
    fp, lp, bp = pl.loss_and_grads(port, params, cam, target, degree)
    fr, lr, br = pl.loss_and_grads(ref, params, cam, target, degree)
    for k in ("sortedKeysHigh", "sortedKeysLow", "sortedGaussIdx", "tileCounts"):
        assert np.array_equal(fp["bins"][k], fr["bins"][k]), k
    assert np.array_equal(fp["render"].view(np.uint32), fr["render"].view(np.uint32))
    assert abs(lp["loss"] - lr["loss"]) < 1e-7
    for k in params:
        a, r = bp["grads"][k], br["grads"][k]
        assert np.abs(a - r).max() / max(np.abs(r).max(), 1e-8) < 1e-5, k


def test_oracle_gradients_match_finite_differences(port):
    """Independent check that 'parity with the reference' means 'correct gradients': central
    differences of the oracle loss in the smooth parameters of a tiny scene."""
    n, W, H = 12, 24, 24
    params = make_gaussians(n, 5, 1)
    params["_scales"] += 1.0          # big, overlapping splats: smooth loss
    cam = make_cameras(W, H, 1)[0]
    target = make_targets(W, H, 1, 5)[0]
    _, lo, bw = pl.loss_and_grads(port, params, cam, target, 1)
    rng = np.random.default_rng(1)
    for key in ("_features_dc", "_opacity", "_scales"):
        for _ in range(3):
            idx = tuple(rng.integers(0, s) for s in params[key].shape)
            eps = 2e-3
            p1 = {k: v.copy() for k, v in params.items()}; p1[key][idx] += eps
            p2 = {k: v.copy() for k, v in params.items()}; p2[key][idx] -= eps
            l1 = pl.loss_and_grads(port, p1, cam, target, 1)[1]["loss"]
            l2 = pl.loss_and_grads(port, p2, cam, target, 1)[1]["loss"]
            fd = (l1 - l2) / (2 * eps)
            an = float(bw["grads"][key][idx])
            assert abs(fd - an) < 5e-2 * max(abs(an), abs(fd)) + 2e-6, (key, idx, fd, an)


def test_activation_exp_convention_is_within_one_ulp(port):
    """The activation exponential is pinned by convention (oracle gso_expf == csrc/common.cuh gsb_expf, op for op): it
    must stay within 1 ulp of the exact value over the range of log-scales / opacity logits, saturate and flush like
    expf, and pass NaN through."""
    import ctypes as C
    x = np.concatenate([np.linspace(-104.5, 89.0, 1_000_001), np.random.default_rng(0).normal(0, 3, 500_000),
                        [np.inf, -np.inf, 0.0, -0.0, 88.9, -200.0]]).astype(np.float32)
    y = np.empty_like(x)
    port.lib.gso_expf_array(len(x), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p))
    exact = np.exp(x.astype(np.float64))
    with np.errstate(over="ignore"):
        r32 = exact.astype(np.float32)
    normal = np.isfinite(r32) & (r32 > 1.2e-38)
    ulp = np.abs(y[normal].astype(np.float64) - exact[normal]) / np.spacing(r32[normal]).astype(np.float64)
    assert ulp.max() < 1.0
    assert np.all(y[np.isposinf(r32)] == np.inf) and np.all(y[x < -104.0] == 0.0)
    tiny = np.isfinite(r32) & ~normal
    assert np.abs(y[tiny].astype(np.float64) - exact[tiny]).max() <= 1.5e-45      # one subnormal step
    nan = np.array([np.nan], np.float32); out = np.empty(1, np.float32)
    port.lib.gso_expf_array(1, nan.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert np.isnan(out[0])

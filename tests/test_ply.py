"""PLY snapshots: the reference's format, Data/PlyWriter.swift:22-146 (writer), :149-265 (loader)."""
import numpy as np
import pytest

from gaussiansplattingmlx_b200.ply import PlyWriter, PlyError, save_snapshot, save_resume, load_resume
from gaussiansplattingmlx_b200.scene import make_gaussians


def test_header_and_record_layout_byte_for_byte(tmp_path):
    p = make_gaussians(3, 1, 1)                                  # K = 4 -> M = 3 rest coefficients
    f = tmp_path / "a" / "b" / "iteration_0.ply"                # parent directories are created (PlyWriter.swift:105-110)
    PlyWriter.writeGaussianBinary(p["_xyz"], p["_features_dc"], p["_features_rest"], p["_opacity"], p["_scales"], p["_rotation"], f)
    raw = f.read_bytes()
    expect = ("ply\nformat binary_little_endian 1.0\ncomment features_rest_shape 3 3\nelement vertex 3\n"
              "property float x\nproperty float y\nproperty float z\n"
              "property float f_dc_0\nproperty float f_dc_1\nproperty float f_dc_2\n"
              + "".join(f"property float f_rest_{i}\n" for i in range(9))
              + "property float opacity\nproperty float scale_0\nproperty float scale_1\nproperty float scale_2\n"
              "property float rot_0\nproperty float rot_1\nproperty float rot_2\nproperty float rot_3\nend_header\n").encode("ascii")
    assert raw.startswith(expect)
    body = np.frombuffer(raw[len(expect):], "<f4").reshape(3, 3 + 3 + 9 + 1 + 3 + 4)
    assert np.array_equal(body[:, 0:3], p["_xyz"]) and np.array_equal(body[:, 3:6], p["_features_dc"].reshape(3, 3))
    # f_rest_i: row-major [M,3] (coefficient-major, channel fastest)
    assert np.array_equal(body[:, 6:15], p["_features_rest"].reshape(3, 9))
    assert body[1, 6 + 2 * 3 + 1] == p["_features_rest"][1, 2, 1]
    assert np.array_equal(body[:, 15], p["_opacity"][:, 0]) and np.array_equal(body[:, 16:19], p["_scales"])
    assert np.array_equal(body[:, 19:23], p["_rotation"])


@pytest.mark.parametrize("n,deg", [(0, 3), (1, 0), (257, 3), (50, 4)])
def test_round_trip_is_bit_exact(tmp_path, n, deg):
    p = make_gaussians(max(n, 1), 7, deg)
    p = {k: v[:n] for k, v in p.items()}
    f = save_snapshot(p, tmp_path, 300)
    assert f.name == "iteration_300.ply"
    q = PlyWriter.loadGaussianBinaryPLY(f)
    for k in p:
        assert q[k].shape == p[k].shape and np.array_equal(q[k].view(np.uint32), p[k].view(np.uint32)), k


def test_loader_finds_properties_by_name_and_reports_errors(tmp_path):
    p = make_gaussians(4, 2, 0)
    f = tmp_path / "x.ply"
    # a file with the properties in a different order (the loader indexes by name, PlyWriter.swift:191)
    hdr = ("ply\nformat binary_little_endian 1.0\ncomment features_rest_shape 0 3\nelement vertex 4\n"
           "property float opacity\nproperty float x\nproperty float y\nproperty float z\n"
           "property float rot_0\nproperty float rot_1\nproperty float rot_2\nproperty float rot_3\n"
           "property float scale_0\nproperty float scale_1\nproperty float scale_2\n"
           "property float f_dc_0\nproperty float f_dc_1\nproperty float f_dc_2\nend_header\n")
    rows = np.concatenate([p["_opacity"], p["_xyz"], p["_rotation"], p["_scales"], p["_features_dc"].reshape(4, 3)], axis=1).astype("<f4")
    f.write_bytes(hdr.encode() + rows.tobytes())
    q = PlyWriter.loadGaussianBinaryPLY(f)
    for k in ("_xyz", "_opacity", "_rotation", "_scales", "_features_dc"):
        assert np.array_equal(q[k], p[k]), k
    assert q["_features_rest"].shape == (4, 0, 3)
    (tmp_path / "bad1.ply").write_bytes(b"ply\nformat binary_little_endian 1.0\n")
    with pytest.raises(PlyError, match="No end_header"):
        PlyWriter.loadGaussianBinaryPLY(tmp_path / "bad1.ply")
    (tmp_path / "bad2.ply").write_bytes(b"ply\nformat binary_little_endian 1.0\nelement vertex 0\nend_header\n")
    with pytest.raises(PlyError, match="features_rest_shape"):
        PlyWriter.loadGaussianBinaryPLY(tmp_path / "bad2.ply")
    with pytest.raises(PlyError, match="size mismatch"):
        PlyWriter.writeGaussianBinary(p["_xyz"], p["_features_dc"][:2], p["_features_rest"], p["_opacity"], p["_scales"],
                                      p["_rotation"], tmp_path / "bad3.ply")


def test_resume_sidecar_round_trip(tmp_path):
    p = make_gaussians(10, 3, 3)
    m = {k: v * 0.5 for k, v in p.items()}
    v = {k: v * v for k, v in p.items()}
    acc = np.arange(10, dtype=np.float32)
    save_resume(tmp_path / "r.npz", 1234, m, v, acc, 34)
    it, m2, v2, acc2, steps = load_resume(tmp_path / "r.npz")
    assert it == 1234 and steps == 34 and np.array_equal(acc, acc2)
    for k in p:
        assert np.array_equal(m[k], m2[k]) and np.array_equal(v[k], v2[k])

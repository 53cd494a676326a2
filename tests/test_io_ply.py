"""CPU test of the Inria-compatible .ply import / export (format of gs_toolkit/scripts/exporter.py:83-148)."""
import numpy as np
import pytest


def test_ply_round_trip_and_layout(tmp_path):
    from rasterizer.io_ply import load_gaussians_ply, save_gaussians_ply

    g = np.random.default_rng(0)
    n, k = 257, 16
    p = dict(means=g.normal(size=(n, 3)), features_dc=g.normal(size=(n, 3)), features_rest=g.normal(size=(n, k - 1, 3)),
             opacities=g.normal(size=(n, 1)), scales=g.normal(size=(n, 3)), quats=g.normal(size=(n, 4)))
    p = {a: b.astype(np.float32) for a, b in p.items()}
    path = str(tmp_path / "gaussians.ply")
    save_gaussians_ply(path, **p)
    raw = open(path, "rb").read()
    header, body = raw.split(b"end_header\n", 1)
    lines = header.decode().splitlines()
    assert lines[0] == "ply" and lines[1] == "format binary_little_endian 1.0" and lines[2] == f"element vertex {n}"
    names = [ln.split()[2] for ln in lines[3:]]
    assert names[:6] == ["x", "y", "z", "nx", "ny", "nz"] and names[6:9] == ["f_dc_0", "f_dc_1", "f_dc_2"]
    assert names[9] == "f_rest_0" and names[9 + 45] == "opacity" and names[-4:] == ["rot_0", "rot_1", "rot_2", "rot_3"]
    assert len(body) == n * len(names) * 4 and len(names) == 6 + 3 + 45 + 1 + 3 + 4
    rows = np.frombuffer(body, "<f4").reshape(n, len(names))
    # f_rest is channel-major: f_rest_j = features_rest[:, j % 15, j // 15]  (transpose(1, 2).flatten(1))
    assert np.array_equal(rows[:, 9 + 0], p["features_rest"][:, 0, 0])
    assert np.array_equal(rows[:, 9 + 1], p["features_rest"][:, 1, 0])
    assert np.array_equal(rows[:, 9 + 15], p["features_rest"][:, 0, 1])
    assert np.all(rows[:, 3:6] == 0)
    back = load_gaussians_ply(path)
    for a in p:
        assert back[a].shape == p[a].shape and np.array_equal(back[a], p[a]), a


def test_ply_rejects_unsupported(tmp_path):
    from rasterizer.io_ply import load_gaussians_ply

    bad = tmp_path / "ascii.ply"
    bad.write_text("ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nend_header\n0.0\n")
    with pytest.raises(ValueError, match="binary_little_endian"):
        load_gaussians_ply(str(bad))
    nope = tmp_path / "nope.ply"
    nope.write_text("hello\n")
    with pytest.raises(ValueError, match="not a PLY"):
        load_gaussians_ply(str(nope))

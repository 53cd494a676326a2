"""Generates tests/golden/io_orient_ref.npz with the REFERENCE's own `auto_orient_and_center_poses`
(gs_toolkit/cameras/camera_utils.py:552-660) for every orientation x centring method on two seeded pose sets
(an inward-looking ring and a forward-facing array).  Build container only.
    python tests/golden/gen_golden_orient.py"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden_densify  # noqa: F401,E402  (installs the stub finder for gs_toolkit's absent third-party imports)

from gs_toolkit.cameras.camera_utils import auto_orient_and_center_poses  # noqa: E402


def look_at(pos, target, up0):
    fwd = (target - pos) / np.linalg.norm(target - pos)
    right = np.cross(fwd, up0)
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, up, -fwd, pos
    return c2w


def pose_sets():
    g = np.random.default_rng(11)
    ring = []
    for i in range(17):
        a = 2 * math.pi * i / 17
        pos = np.array([4 * math.cos(a) + 0.5, 4 * math.sin(a) - 0.3, 1.2 + 0.4 * math.sin(2 * a)])
        ring.append(look_at(pos, np.array([0.5, -0.3, 0.2]) + 0.05 * g.normal(size=3), np.array([0.05 * g.normal(), 0.05 * g.normal(), 1.0])))
    array = []
    for i in range(12):
        pos = np.array([0.4 * (i % 4) - 0.6, 0.3 * (i // 4) - 0.3, 0.02 * g.normal()]) + np.array([2.0, 1.0, 0.5])
        array.append(look_at(pos, pos + np.array([0.1 * g.normal(), 1.0, 0.1 * g.normal()]), np.array([0.0, 0.02 * g.normal(), 1.0])))
    return {"ring": np.stack(ring).astype(np.float32), "array": np.stack(array).astype(np.float32)}


def main():
    out = {}
    for name, poses in pose_sets().items():
        out[f"{name}_poses"] = poses
        for method in ("pca", "up", "vertical", "none"):
            for center in ("poses", "focus", "none"):
                o, t = auto_orient_and_center_poses(torch.from_numpy(poses.copy()), method=method, center_method=center)
                out[f"{name}_{method}_{center}_oriented"] = o.numpy()
                out[f"{name}_{method}_{center}_transform"] = t.numpy()
    np.savez_compressed(os.path.join(HERE, "io_orient_ref.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()

"""Inria-compatible Gaussian .ply import / export — the on-disk format the reference's
`gs-export gaussian-splat` writes (gs_toolkit/scripts/exporter.py:83-148): one `vertex` element with float32
properties x y z nx ny nz f_dc_* f_rest_* opacity scale_* rot_*, binary little endian (plyfile's default), RAW
parameters (log-scales, logit opacities, unnormalised wxyz quaternions), `f_rest` stored channel-major
(features_rest.transpose(1, 2).flatten(1)).  Needed to load real trained scenes into the operators; pure host code
(numpy), no third-party ply library."""
from __future__ import annotations

from typing import Dict

import numpy as np


def _attributes(n_dc: int, n_rest: int, n_scale: int, n_rot: int):
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)]
    names += [f"f_rest_{i}" for i in range(n_rest)]
    names += ["opacity"]
    names += [f"scale_{i}" for i in range(n_scale)]
    names += [f"rot_{i}" for i in range(n_rot)]
    return names


def save_gaussians_ply(path: str, means, features_dc, features_rest, opacities, scales, quats) -> None:
    """Arrays (numpy or torch, any device): means [N,3], features_dc [N,3], features_rest [N,K-1,3], opacities [N,1],
    scales [N,3], quats [N,4] — the model's raw parameter tensors, written exactly like exporter.py:100-128."""
    def np32(a):
        a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
        return np.ascontiguousarray(a, dtype=np.float32)

    xyz, f_dc, opac, scale, rot = np32(means), np32(features_dc), np32(opacities), np32(scales), np32(quats)
    rest = np32(features_rest)
    n = xyz.shape[0]
    f_dc = f_dc.reshape(n, -1)
    f_rest = np.ascontiguousarray(np.transpose(rest.reshape(n, -1, 3), (0, 2, 1))).reshape(n, -1)  # channel-major
    opac = opac.reshape(n, 1)
    cols = np.concatenate([xyz, np.zeros_like(xyz), f_dc, f_rest, opac, scale, rot], axis=1).astype("<f4")
    names = _attributes(f_dc.shape[1], f_rest.shape[1], scale.shape[1], rot.shape[1])
    assert cols.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {n}\n"
    header += "".join(f"property float {nm}\n" for nm in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(cols.tobytes())


def load_gaussians_ply(path: str) -> Dict[str, np.ndarray]:
    """Inverse of save_gaussians_ply: dict of float32 arrays with the model's parameter names and shapes."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, n, props = None, None, []
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii").split()
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                if tok[1] != "vertex" and n is None:
                    raise ValueError(f"{path}: expected the vertex element first")
                if tok[1] == "vertex":
                    n = int(tok[2])
            elif tok[0] == "property":
                if tok[1] not in ("float", "float32"):
                    raise ValueError(f"{path}: only float32 properties are supported (got {tok[1]})")
                props.append(tok[2])
            elif tok[0] == "end_header":
                break
        if fmt != "binary_little_endian":
            raise ValueError(f"{path}: only binary_little_endian PLY is supported (got {fmt})")
        data = np.frombuffer(f.read(n * len(props) * 4), dtype="<f4").reshape(n, len(props))
    col = {nm: i for i, nm in enumerate(props)}

    def block(prefix):
        idx = sorted((int(nm[len(prefix):]), i) for nm, i in col.items() if nm.startswith(prefix) and nm[len(prefix):].isdigit())
        return np.ascontiguousarray(data[:, [i for _, i in idx]])

    f_rest = block("f_rest_")
    k_rest = f_rest.shape[1] // 3
    return dict(
        means=np.ascontiguousarray(data[:, [col["x"], col["y"], col["z"]]]),
        features_dc=block("f_dc_"),
        features_rest=np.ascontiguousarray(np.transpose(f_rest.reshape(n, 3, k_rest), (0, 2, 1))),
        opacities=np.ascontiguousarray(data[:, [col["opacity"]]]),
        scales=block("scale_"),
        quats=block("rot_"),
    )

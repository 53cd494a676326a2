"""Seeded synthetic scenes for the benchmark, the smoke test and the parity tests.

Pure host code (torch CPU generator -> numpy); nothing here touches the GPU or the oracle.  The camera
conventions are those the reference models feed to the operator: `viewmat` is world->camera with the
camera looking down +z, `projmat = projection_matrix(znear, zfar, fovx, fovy) @ viewmat`
(reference gs_toolkit/utils/comms.py:103-123, gs_toolkit/models/vanilla_gs.py:722-771).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

SH_C0 = 0.28209479177387814

# the configurations BASELINE.json names (N, W, H, s_min, s_max, frustum margin)
CONFIGS = {
    "cfg1": dict(num_points=10_000, img_width=256, img_height=256, s_min=0.01, s_max=0.1, margin=1.0),
    "cfg2": dict(num_points=1_000_000, img_width=1920, img_height=1080, s_min=0.002, s_max=0.02, margin=1.1),
    "cfg4": dict(num_points=5_000_000, img_width=3840, img_height=2160, s_min=0.0015, s_max=0.015, margin=1.1),
    # one object-centric 800x800 view of a cfg3-like scene (BASELINE configs[2] trains on such views): 85 % of the
    # Gaussians sit in a blob around the look-at point, so a few hundred of the 2500 tiles hold most of the pairs
    "cfg3view": dict(num_points=200_000, img_width=800, img_height=800, s_min=0.006, s_max=0.06, margin=1.1, clustered=True),
}


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> np.ndarray:
    """OpenGL-style perspective matrix with the reference's sign conventions (utils/comms.py:103-123)."""
    t = znear * math.tan(0.5 * fovy)
    b = -t
    r = znear * math.tan(0.5 * fovx)
    l = -r  # noqa: E741
    n, f = znear, zfar
    return np.array(
        [
            [2 * n / (r - l), 0.0, (r + l) / (r - l), 0.0],
            [0.0, 2 * n / (t - b), (t + b) / (t - b), 0.0],
            [0.0, 0.0, (f + n) / (f - n), -1.0 * f * n / (f - n)],
            [0.0, 0.0, 1.0, 0.0],
        ],
        dtype=np.float32,
    )


def look_at_viewmat(yaw_deg: float = 0.0, pitch_deg: float = 0.0, centre=(0.0, 0.0, 6.0),
                    shift=(0.0, 0.0, 0.0)) -> np.ndarray:
    """World->camera matrix of a camera orbiting `centre` (yaw about y, pitch about x) at the radius the
    default camera (origin, looking down +z) has.  yaw = pitch = 0, shift = 0 gives the identity."""
    cy, sy = math.cos(math.radians(yaw_deg)), math.sin(math.radians(yaw_deg))
    cp, sp = math.cos(math.radians(pitch_deg)), math.sin(math.radians(pitch_deg))
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=np.float64)
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]], dtype=np.float64)
    R_c2w = Ry @ Rx
    c = np.asarray(centre, dtype=np.float64)
    cam_pos = c + R_c2w @ (np.zeros(3) - c) + np.asarray(shift, dtype=np.float64)
    V = np.eye(4, dtype=np.float64)
    V[:3, :3] = R_c2w.T
    V[:3, 3] = -R_c2w.T @ cam_pos
    return V.astype(np.float32)


def make_scene(num_points: int, img_width: int, img_height: int, s_min: float, s_max: float,
               margin: float = 1.1, seed: int = 0, sh_degree: int = 3, degrees_to_use: Optional[int] = None,
               block_width: int = 16, viewmat: Optional[np.ndarray] = None, opacity_clip: Optional[float] = None,
               channels: int = 3, clustered: bool = False) -> Dict[str, object]:
    """The seeded generator of SURVEY §8(d).  Returns numpy arrays (float32 / int) plus python scalars.
    `clustered`: non-uniform, object-centric scene — 85 % of the Gaussians are drawn from an isotropic normal blob
    (sigma 0.7) around (0, 0, 6), the rest fill the frustum as usual (tile lists then differ by orders of magnitude)."""
    g = torch.Generator().manual_seed(seed)
    W, H, N = img_width, img_height, num_points
    fovx = math.radians(60.0)
    fx = fy = 0.5 * W / math.tan(0.5 * fovx)
    fovy = 2.0 * math.atan(0.5 * H / fy)
    cx, cy = W / 2.0, H / 2.0

    def U(*shape):
        return torch.rand(*shape, generator=g, dtype=torch.float32)

    def Nrm(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32)

    z = 2.0 + 8.0 * U(N)
    x = (2.0 * U(N) - 1.0) * margin * (0.5 * W / fx) * z
    y = (2.0 * U(N) - 1.0) * margin * (0.5 * H / fy) * z
    means3d = torch.stack([x, y, z], dim=-1)
    if clustered:
        n_obj = (N * 85) // 100
        blob = 0.7 * Nrm(n_obj, 3) + torch.tensor([0.0, 0.0, 6.0])
        means3d = torch.cat([blob, means3d[n_obj:]], dim=0)
    scales = torch.exp(math.log(s_min) + (math.log(s_max) - math.log(s_min)) * U(N, 3))
    quats = Nrm(N, 4)
    quats = quats / quats.norm(dim=-1, keepdim=True)
    K = (sh_degree + 1) ** 2
    dc = (U(N, 1, 3) - 0.5) / SH_C0
    rest = 0.05 * Nrm(N, K - 1, 3)
    sh_coeffs = torch.cat([dc, rest], dim=1).contiguous()
    opac = torch.sigmoid(1.5 * Nrm(N))
    if opacity_clip is not None:
        opac = opac.clamp(max=opacity_clip)
    background = U(3)
    v_out_img = (2.0 * U(H, W, 3) - 1.0) * 1e-3
    v_out_alpha = (2.0 * U(H, W) - 1.0) * 1e-3
    extra = {}
    if channels != 3:
        extra["nd_colors"] = U(N, channels).numpy()
        extra["nd_background"] = U(channels).numpy()
        extra["nd_v_out_img"] = ((2.0 * U(H, W, channels) - 1.0) * 1e-3).numpy()

    if viewmat is None:
        viewmat = np.eye(4, dtype=np.float32)
    viewmat = np.ascontiguousarray(viewmat, dtype=np.float32)
    projmat = (projection_matrix(0.001, 1000.0, fovx, fovy).astype(np.float64) @ viewmat.astype(np.float64)).astype(np.float32)
    cam_pos = (-viewmat[:3, :3].T.astype(np.float64) @ viewmat[:3, 3].astype(np.float64)).astype(np.float32)
    scene = dict(
        means3d=means3d.numpy(), scales=scales.numpy(), quats=quats.numpy(), sh_coeffs=sh_coeffs.numpy(),
        opacities=opac.numpy(), background=background.numpy(), v_out_img=v_out_img.numpy(),
        v_out_alpha=v_out_alpha.numpy(), viewmat=viewmat, projmat=np.ascontiguousarray(projmat), cam_pos=cam_pos,
        fx=float(fx), fy=float(fy), cx=float(cx), cy=float(cy), img_width=W, img_height=H,
        block_width=block_width, glob_scale=1.0, clip_thresh=0.01, sh_degree=sh_degree,
        degrees_to_use=sh_degree if degrees_to_use is None else degrees_to_use, seed=seed,
    )
    scene.update(extra)
    return scene


def make_config_scene(name: str, seed: int = 0, **overrides) -> Dict[str, object]:
    kw = dict(CONFIGS[name])
    kw.update(overrides)
    return make_scene(seed=seed, **kw)


def scene_to_torch(scene: Dict[str, object], device) -> Dict[str, object]:
    out = {}
    for k, v in scene.items():
        out[k] = torch.from_numpy(v).to(device) if isinstance(v, np.ndarray) else v
    return out

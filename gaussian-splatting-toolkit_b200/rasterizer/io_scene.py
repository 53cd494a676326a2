"""On-disk formats either side of the hot path — pure host code (numpy / torch CPU):

  * checkpoints  `step-%09d.ckpt` exactly as `Trainer.save_checkpoint` writes them (gs_toolkit/engine/trainer.py:443-476):
        {"step", "pipeline": {"_model.gauss_params.<name>": tensor, ...}, "optimizers": {group: Adam.state_dict()},
         "schedulers": {group: LambdaLR.state_dict()}, "scalers": GradScaler.state_dict()}
    and as `Trainer._load_checkpoint` / `Pipeline.load_pipeline` / `GaussianSplattingModel.load_state_dict` read them
    (trainer.py:404-441, pipelines/base_pipeline.py:355-367, models/vanilla_gs.py:236-258: "module." prefix stripped,
    pre-`gauss_params` names remapped, parameters resized to the stored Gaussian count);
  * `transforms.json`  the reference's own dataset format (data/dataparsers/gs_toolkit_dataparser.py:77-457): frames
    sorted by file name, intrinsics global or per frame, train / eval split by fraction (dataparsers_utils.py:10-32) or
    by explicit `<split>_filenames`, poses oriented ("up") and centred ("poses") as camera_utils.py:552-668 does by
    default, optional auto-scale; the file names are returned and `load_image` / `load_depth_image` decode them the
    way the reference's dataset does (data/datasets/base_dataset.py:48-118; needs Pillow, like the reference);
  * `camera_to_view_proj`  camera-to-world (OpenGL axes) -> the (viewmat, projmat, fovs) the operators take, i.e. the
    per-view prologue of `get_outputs` (vanilla_gs.py:722-741, utils/comms.py:103-123).
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .synthetic import projection_matrix

GROUPS = ("means", "scales", "quats", "features_dc", "features_rest", "opacities")
_PREFIX = "_model.gauss_params."


# ------------------------------------------------------------------------------------------------- images
def load_image_uint8(filename: str, scale_factor: float = 1.0) -> np.ndarray:
    """InputDataset.get_numpy_image (base_dataset.py:48-66): uint8 [H,W,3 or 4]; bilinear resize by `scale_factor`
    (new size truncated to int, as the reference does), grey images repeated to three channels."""
    from PIL import Image  # the reference's own dependency; imported here so that the package loads without it

    pil_image = Image.open(filename)
    if scale_factor != 1.0:
        width, height = pil_image.size
        pil_image = pil_image.resize((int(width * scale_factor), int(height * scale_factor)), resample=Image.BILINEAR)
    image = np.array(pil_image, dtype="uint8")
    if image.ndim == 2:
        image = image[:, :, None].repeat(3, axis=2)
    if image.ndim != 3 or image.shape[2] not in (3, 4):
        raise ValueError(f"Image shape of {image.shape} is in correct.")
    return image


def load_image(filename: str, scale_factor: float = 1.0, alpha_color: Optional[torch.Tensor] = None) -> torch.Tensor:
    """InputDataset.get_image (base_dataset.py:68-87): float32 [H,W,3 or 4] in [0,1]; an RGBA image is composited over
    `alpha_color` (rgb * a + alpha_color * (1 - a)) when one is given, otherwise its four channels are returned."""
    image = torch.from_numpy(load_image_uint8(filename, scale_factor).astype("float32") / 255.0)
    if alpha_color is not None and image.shape[-1] == 4:
        image = image[:, :, :3] * image[:, :, -1:] + torch.as_tensor(alpha_color, dtype=torch.float32) * (1.0 - image[:, :, -1:])
    return image


def load_depth_image(filename: str, scale_factor: float = 1.0, mono_depth: bool = False) -> torch.Tensor:
    """InputDataset.get_numpy_depth_image / get_depth_image (base_dataset.py:89-131): `.png` / `.jpg` / `.jpeg` through
    Pillow (bilinear resize by `scale_factor`) or `.npy`; monocular depth maps are stored as 8-bit (/ 255), metric depth
    in millimetres (/ 1000)."""
    ext = os.path.splitext(filename)[1]
    if ext in (".png", ".jpg", ".jpeg"):
        from PIL import Image

        pil_image = Image.open(filename)
        if scale_factor != 1.0:
            width, height = pil_image.size
            pil_image = pil_image.resize((int(width * scale_factor), int(height * scale_factor)), resample=Image.BILINEAR)
        depth = np.array(pil_image)
    elif ext == ".npy":
        depth = np.load(filename)
    else:
        raise ValueError(f"Depth file format {ext} not supported.")
    return torch.from_numpy(depth.astype("float32") / (255.0 if mono_depth else 1000.0))


# ------------------------------------------------------------------------------------------------- checkpoints
def checkpoint_path(checkpoint_dir: str, step: int) -> str:
    return os.path.join(checkpoint_dir, f"step-{step:09d}.ckpt")


def save_checkpoint(checkpoint_dir: str, step: int, params: Dict[str, torch.Tensor], optimizers=None,
                    schedulers: Optional[Dict[str, dict]] = None, save_only_latest_checkpoint: bool = True,
                    extra_pipeline_state: Optional[Dict[str, torch.Tensor]] = None) -> str:
    """trainer.py:443-476.  `optimizers` is a rasterizer.optim.GaussianOptimizers (or anything with a
    `state_dict()` returning {group: Adam.state_dict()}), or None.  `schedulers` defaults to the optimizers' own
    `scheduler_state_dict()` ({group: LambdaLR.state_dict()}, trainer.py:470) so that a resumed run continues its
    learning-rate decay."""
    os.makedirs(checkpoint_dir, exist_ok=True)
    path = checkpoint_path(checkpoint_dir, step)
    pipeline = {_PREFIX + k: v.detach().cpu() for k, v in params.items()}
    if extra_pipeline_state:
        pipeline.update({k: v.detach().cpu() for k, v in extra_pipeline_state.items()})
    opt_sd = {}
    if optimizers is not None:
        for name, sd in optimizers.state_dict().items():
            opt_sd[name] = {"state": {i: {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in st.items()}
                                      for i, st in sd["state"].items()}, "param_groups": sd["param_groups"]}
    if schedulers is None and optimizers is not None and hasattr(optimizers, "scheduler_state_dict"):
        schedulers = optimizers.scheduler_state_dict()
    torch.save({"step": step, "pipeline": pipeline, "optimizers": opt_sd, "schedulers": schedulers or {},
                "scalers": {}}, path)
    if save_only_latest_checkpoint:
        for f in os.listdir(checkpoint_dir):
            if os.path.join(checkpoint_dir, f) != path:
                os.unlink(os.path.join(checkpoint_dir, f))
    return path


def latest_checkpoint_step(load_dir: str) -> int:
    """trainer.py:410-416 (specific to the `step-<n>.ckpt` name format)."""
    return sorted(int(x[x.find("-") + 1: x.find(".")]) for x in os.listdir(load_dir))[-1]


def load_checkpoint(load_dir_or_file: str, load_step: Optional[int] = None, device="cpu") -> Dict[str, object]:
    """Returns {"step", "params": {name: tensor}, "optimizers": {group: Adam.state_dict()}, "schedulers", "extra"}.
    Accepts checkpoints written by the reference trainer (incl. DDP "module." prefixes and the pre-gauss_params
    parameter names) and by `save_checkpoint`."""
    path = load_dir_or_file
    if os.path.isdir(path):
        step = latest_checkpoint_step(path) if load_step is None else load_step
        path = checkpoint_path(path, step)
    assert os.path.exists(path), f"Checkpoint {path} does not exist"
    loaded = torch.load(path, map_location="cpu", weights_only=False)
    state = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in loaded["pipeline"].items()}
    params, extra = {}, {}
    for k, v in state.items():
        if k.startswith(_PREFIX):
            params[k[len(_PREFIX):]] = v
        elif k.startswith("_model.") and k[len("_model."):] in GROUPS:   # old checkpoints (vanilla_gs.py:239-251)
            params.setdefault(k[len("_model."):], v)
        else:
            extra[k] = v
    missing = [g for g in GROUPS if g not in params]
    if missing:
        raise RuntimeError(f"checkpoint {path} holds no Gaussian parameters {missing}")
    n = params["means"].shape[0]
    for k, v in params.items():
        if v.shape[0] != n:
            raise RuntimeError(f"checkpoint {path}: {k} has {v.shape[0]} rows, means has {n}")
    params = {k: v.to(device=device, dtype=torch.float32).contiguous() for k, v in params.items()}
    return {"step": int(loaded["step"]), "params": params, "optimizers": loaded.get("optimizers", {}),
            "schedulers": loaded.get("schedulers", {}), "extra": extra, "path": path}


# ------------------------------------------------------------------------------------------------- cameras
def camera_to_view_proj(camera_to_world: np.ndarray, fx: float, fy: float, width: int, height: int,
                        znear: float = 0.001, zfar: float = 1000.0) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """vanilla_gs.py:722-741: flip y and z of the camera axes (OpenGL -> the operator's +z-forward convention),
    invert analytically, projmat = projection_matrix(0.001, 1000, fovx, fovy) @ viewmat.
    Returns (viewmat [4,4], projmat [4,4] full projection, camera position [3]), float32."""
    c2w = np.asarray(camera_to_world, dtype=np.float64)
    R = c2w[:3, :3] @ np.diag([1.0, -1.0, -1.0])
    T = c2w[:3, 3:4]
    V = np.eye(4)
    V[:3, :3] = R.T
    V[:3, 3:4] = -R.T @ T
    fovx = 2 * math.atan(width / (2 * fx))
    fovy = 2 * math.atan(height / (2 * fy))
    P = projection_matrix(znear, zfar, fovx, fovy).astype(np.float64)
    return V.astype(np.float32), (P @ V).astype(np.float32), c2w[:3, 3].astype(np.float32)


def _rotation_matrix(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """camera_utils.py:462-493 (rotation taking a to b); float32 arithmetic like the reference."""
    a = (a / np.linalg.norm(a)).astype(np.float32)
    b = (b / np.linalg.norm(b)).astype(np.float32)
    v = np.cross(a, b)
    c = np.dot(a, b)
    if c < -1 + 1e-8:
        raise ValueError("average up vector is exactly opposite to +z (the reference perturbs it randomly)")
    s = np.linalg.norm(v)
    K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float32)
    return (np.eye(3, dtype=np.float32) + K + (K @ K) * ((1 - c) / (s ** 2 + 1e-8))).astype(np.float32)


def focus_of_attention(poses: np.ndarray, initial_focus: np.ndarray) -> np.ndarray:
    """camera_utils.py:496-549: the point closest to the optical axes of the cameras that have it in front of them
    (iterated until the set of such cameras stops changing).  poses [n,4,4] / [n,3,4] camera-to-world, float32."""
    dirs = -poses[:, :3, 2:3]
    origins = poses[:, :3, 3:4]
    focus = np.asarray(initial_focus, dtype=np.float32)
    active = np.sum(dirs[..., 0] * (focus - origins[..., 0]), axis=-1) > 0
    done = False
    while int(active.sum()) > 1 and not done:
        dirs, origins = dirs[active], origins[active]
        m = np.eye(3, dtype=np.float32) - dirs * np.transpose(dirs, (0, 2, 1))
        mt_m = np.transpose(m, (0, 2, 1)) @ m
        focus = (np.linalg.inv(mt_m.mean(0)) @ (mt_m @ origins).mean(0)[:, 0]).astype(np.float32)
        active = np.sum(dirs[..., 0] * (focus - origins[..., 0]), axis=-1) > 0
        if active.all():
            done = True
    return focus


def auto_orient_and_center_poses(poses: np.ndarray, method: str = "up", center_method: str = "poses"):
    """camera_utils.py:552-660: orientation "pca" | "up" | "vertical" | "none", centring "poses" | "focus" | "none".
    poses: [n,4,4] or [n,3,4] camera-to-world.  Returns (poses [n,3,4], transform [3,4]); float32 arithmetic like the
    reference."""
    poses = np.asarray(poses, dtype=np.float32)
    if poses.shape[-2] == 3:
        poses = np.concatenate([poses, np.tile(np.array([[[0, 0, 0, 1]]], np.float32), (poses.shape[0], 1, 1))], axis=1)
    origins = poses[:, :3, 3]
    mean_origin = origins.mean(axis=0)
    translation_diff = origins - mean_origin
    if center_method == "poses":
        translation = mean_origin
    elif center_method == "focus":
        translation = focus_of_attention(poses, mean_origin)
    elif center_method == "none":
        translation = np.zeros_like(mean_origin)
    else:
        raise ValueError(f"Unknown value for center_method: {center_method}")
    if method == "pca":
        # eigenvectors are defined up to sign and the reference uses them as LAPACK returns them: same routine
        # (torch.linalg.eigh) so that the same signs come out
        _, eigvec_t = torch.linalg.eigh(torch.from_numpy(np.ascontiguousarray(translation_diff.T @ translation_diff)))
        eigvec = np.ascontiguousarray(eigvec_t.numpy()[:, ::-1])  # largest principal direction first (torch.flip, :603)
        if np.linalg.det(eigvec) < 0:
            eigvec[:, 2] = -eigvec[:, 2]
        # NOTE the reference multiplies by `eigvec` itself, not its transpose (:608) — reproduced
        transform = np.concatenate([eigvec, eigvec @ -translation[:, None]], axis=-1)
        oriented = transform @ poses
        if oriented.mean(axis=0)[2, 1] < 0:
            oriented[:, 1:3] = -1 * oriented[:, 1:3]
    elif method in ("up", "vertical"):
        up = poses[:, :3, 1].mean(axis=0)
        up = up / np.linalg.norm(up)
        if method == "vertical":
            # the 3-D direction that projects most vertically in all cameras: total least squares ||X u|| -> min
            # over the cameras' x axes (:616-647)
            _, S_t, Vh_t = torch.linalg.svd(torch.from_numpy(np.ascontiguousarray(poses[:, :3, 0])), full_matrices=False)
            S, Vh = S_t.numpy(), Vh_t.numpy()
            if S[1] > 0.17 * math.sqrt(poses.shape[0]):
                up_vertical = Vh[2, :]
                up = up_vertical if np.dot(up_vertical, up) > 0 else -up_vertical
            else:  # degenerate configuration: project "up" on the plane orthogonal to the first singular vector
                up = up - Vh[0, :] * np.dot(up, Vh[0, :])
                up = up / np.linalg.norm(up)
        rotation = _rotation_matrix(up, np.array([0, 0, 1], np.float32))
        transform = np.concatenate([rotation, rotation @ -translation[:, None]], axis=-1)
        oriented = transform @ poses
    elif method == "none":
        transform = np.eye(4, dtype=np.float32)
        transform[:3, 3] = -translation
        transform = transform[:3, :]
        oriented = transform @ poses
    else:
        raise ValueError(f"Unknown value for method: {method}")
    return oriented.astype(np.float32), transform.astype(np.float32)


def train_eval_split_fraction(num_images: int, train_split_fraction: float):
    """dataparsers_utils.py:10-32."""
    num_train = math.ceil(num_images * train_split_fraction)
    i_all = np.arange(num_images)
    i_train = np.linspace(0, num_images - 1, num_train, dtype=int)
    i_eval = np.setdiff1d(i_all, i_train)
    assert len(i_eval) == num_images - num_train
    return i_train, i_eval


def load_transforms(data: str, split: str = "train", train_split_fraction: float = 0.9,
                    orientation_method: str = "up", center_method: str = "poses", auto_scale_poses: bool = False,
                    scale_factor: float = 1.0, downscale_factor: int = 1) -> Dict[str, object]:
    """gs_toolkit_dataparser.py:77-457 for perspective cameras.  `data` is a directory holding transforms.json or the
    json file itself.  Returns {"image_filenames": [...], "camera_to_worlds": [n,3,4], "fx","fy","cx","cy": [n] float32,
    "height","width": [n] int32, "transform": [3,4], "scale_factor": float, "indices": [n]} for the requested split."""
    assert os.path.exists(data), f"Data directory {data} does not exist."
    if data.endswith(".json"):
        meta_path, data_dir = data, os.path.dirname(data)
    else:
        meta_path, data_dir = os.path.join(data, "transforms.json"), data
    with open(meta_path, "r", encoding="UTF-8") as f:
        meta = json.load(f)
    if "applied_scale" in meta:
        scale_factor = meta["applied_scale"]

    def fname(fp: str) -> str:
        if downscale_factor > 1:
            return os.path.join(data_dir, f"images_{downscale_factor}", os.path.basename(fp))
        return os.path.join(data_dir, fp)

    frames = meta["frames"]
    frames = [frames[i] for i in np.argsort([fname(fr["file_path"]) for fr in frames])]
    names: List[str] = [fname(fr["file_path"]) for fr in frames]
    n = len(frames)

    def intrinsic(key: str, cast):
        if key in meta:
            return np.full(n, cast(meta[key]))
        for fr in frames:
            assert key in fr, f"{key} not specified in frame"
        return np.array([cast(fr[key]) for fr in frames])

    fx, fy, cx, cy = (intrinsic(k, float).astype(np.float32) for k in ("fl_x", "fl_y", "cx", "cy"))
    height, width = (intrinsic(k, int).astype(np.int32) for k in ("h", "w"))
    poses = np.array([fr["transform_matrix"] for fr in frames], dtype=np.float32)

    has_split_files_spec = any(f"{s}_filenames" in meta for s in ("train", "val", "test"))
    if f"{split}_filenames" in meta:
        wanted = set(fname(x) for x in meta[f"{split}_filenames"])
        unmatched = wanted.difference(names)
        if unmatched:
            raise RuntimeError(f"Some filenames for split {split} were not found: {unmatched}.")
        indices = np.array([i for i, p in enumerate(names) if p in wanted], dtype=np.int32)
    elif has_split_files_spec:
        raise RuntimeError(f"The dataset's list of filenames for split {split} is missing.")
    else:
        i_train, i_eval = train_eval_split_fraction(n, train_split_fraction)
        if split == "train":
            indices = i_train
        elif split in ("val", "test"):
            indices = i_eval
        else:
            raise ValueError(f"Unknown dataparser split {split}")

    method = meta.get("orientation_override", orientation_method)
    poses, transform = auto_orient_and_center_poses(poses, method=method, center_method=center_method)
    s = 1.0
    if auto_scale_poses:
        s /= float(np.max(np.abs(poses[:, :3, 3])))
    s *= scale_factor
    poses[:, :3, 3] *= s
    idx = np.asarray(indices, dtype=np.int64)
    d = float(downscale_factor)
    return {"image_filenames": [names[i] for i in idx], "camera_to_worlds": poses[idx, :3, :4],
            "fx": fx[idx] / d, "fy": fy[idx] / d, "cx": cx[idx] / d, "cy": cy[idx] / d,
            "height": (height[idx] // downscale_factor).astype(np.int32), "width": (width[idx] // downscale_factor).astype(np.int32),
            "transform": transform, "scale_factor": s, "indices": idx}


def load_views(data: str, split: str = "train", image_scale_factor: float = 1.0, alpha_color=None, load_images: bool = True,
               **dataparser_kwargs) -> List[Dict[str, object]]:
    """A whole dataset split ready for the operators: `load_transforms` + per view the (viewmat, projmat, camera position)
    of `camera_to_view_proj`, the intrinsics rescaled like `Cameras.rescale_output_resolution` does for a dataset scale
    factor (cameras.py: fx, fy, cx, cy multiplied, height / width truncated to int), and — unless `load_images=False` —
    the decoded image of `load_image` ([H,W,3] float32 in [0,1] when `alpha_color` is given or the file has no alpha).
    This is what `InputDataset.get_data` + the model's per-view prologue hand to `get_outputs` / `get_loss_dict`."""
    t = load_transforms(data, split=split, **dataparser_kwargs)
    views = []
    for i, name in enumerate(t["image_filenames"]):
        s = float(image_scale_factor)
        fx, fy, cx, cy = (float(t[k][i]) * s for k in ("fx", "fy", "cx", "cy"))
        height, width = int(int(t["height"][i]) * s), int(int(t["width"][i]) * s)
        viewmat, projmat, cam_pos = camera_to_view_proj(t["camera_to_worlds"][i], fx, fy, width, height)
        v = {"image_filename": name, "index": int(t["indices"][i]), "fx": fx, "fy": fy, "cx": cx, "cy": cy,
             "height": height, "width": width, "viewmat": viewmat, "projmat": projmat, "cam_pos": cam_pos}
        if load_images:
            v["image"] = load_image(name, image_scale_factor, alpha_color)
        views.append(v)
    return views

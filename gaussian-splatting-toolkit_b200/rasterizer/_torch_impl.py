"""Pure-torch helpers the reference models import from `rasterizer._torch_impl`
(gs_toolkit/models/vanilla_gs.py:13,552 uses quat_to_rotmat).  Only the small, model-facing helpers live
here; the reference's slow per-pixel PyTorch renderer (rasterizer/_torch_impl.py:280-470) is test
infrastructure of the reference and is not part of this package."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor


def normalized_quat_to_rotmat(quat: Tensor) -> Tensor:
    """(w,x,y,z) unit quaternions [...,4] -> rotation matrices [...,3,3] (rasterizer/_torch_impl.py:116-133)."""
    assert quat.shape[-1] == 4, quat.shape
    w, x, y, z = torch.unbind(quat, dim=-1)
    rows = [
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y),
    ]
    return torch.stack(rows, dim=-1).reshape(quat.shape[:-1] + (3, 3))


def quat_to_rotmat(quat: Tensor) -> Tensor:
    """Normalise, then convert (rasterizer/_torch_impl.py:136-138)."""
    assert quat.shape[-1] == 4, quat.shape
    return normalized_quat_to_rotmat(F.normalize(quat, dim=-1))


def scale_rot_to_cov3d(scale: Tensor, glob_scale: float, quat: Tensor) -> Tensor:
    """Sigma = (R S)(R S)^T for unit quaternions (rasterizer/_torch_impl.py:141-148)."""
    assert scale.shape[-1] == 3 and quat.shape[-1] == 4 and scale.shape[:-1] == quat.shape[:-1]
    M = normalized_quat_to_rotmat(quat) * glob_scale * scale[..., None, :]
    return M @ M.transpose(-1, -2)

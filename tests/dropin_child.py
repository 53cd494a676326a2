"""TEST INFRASTRUCTURE — child process of tests/test_gpu_dropin_reference_callers.py.

Imports the REFERENCE's own, unmodified Python package `rasterizer` (and `gs_toolkit`) from baseline/_ref/ and runs
  A. the reference wrappers project_gaussians -> spherical_harmonics -> rasterize_gaussians (+ backward), exactly the
     calls of gs_toolkit/models/vanilla_gs.py:765-837;
  B. the reference model class' own `GaussianSplattingModel.get_outputs` (vanilla_gs.py:672-855) in training mode with
     depth output, followed by a backward pass through rgb / depth;
with the native module `rasterizer.csrc` being either the reference's own CUDA extension (mode "ref") or this
repository's `rasterizer/csrc.so` = the eleven bindings over libgsr_b200.so (mode "ours", injected into sys.modules
BEFORE the reference's lazy `from rasterizer import csrc as _C`, rasterizer/cuda/_backend.py:61-63).
Nothing of this repository's Python package is imported here.

    python dropin_child.py <mode> <scene.npz> <out.npz>
"""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import time
import types
from unittest.mock import MagicMock

mode, scene_path, out_path = sys.argv[1:4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
sys.path = [p for p in sys.path if "gaussian-splatting-toolkit_b200" not in p]
sys.path.insert(0, REF)

import numpy as np
import torch

if mode == "ours":
    so = os.path.join(ROOT, "gaussian-splatting-toolkit_b200", "rasterizer", "csrc.so")
    spec = importlib.util.spec_from_file_location("rasterizer.csrc", so)
    ours = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ours)
    sys.modules["rasterizer.csrc"] = ours

import rasterizer  # the reference's package

assert os.path.realpath(rasterizer.__file__).startswith(os.path.realpath(REF)), rasterizer.__file__
from rasterizer.cuda._backend import _C  # noqa: E402

native = os.path.realpath(_C.__file__)
if mode == "ours":
    assert native.endswith("gaussian-splatting-toolkit_b200/rasterizer/csrc.so"), native
    assert hasattr(_C, "gsr_version")
else:
    assert native.startswith(os.path.realpath(REF)), native
from rasterizer.project_gaussians import project_gaussians  # noqa: E402
from rasterizer.rasterize import rasterize_gaussians  # noqa: E402
from rasterizer.sh import spherical_harmonics  # noqa: E402

z = np.load(scene_path)
dev = torch.device("cuda")
t = lambda k: torch.from_numpy(z[k]).to(dev)
H, W, bw = int(z["img_height"]), int(z["img_width"]), int(z["block_width"])
out = {"native": np.array(native)}


# ---- A. the reference wrappers
def wrappers_view():
    means, scales, quats = (t(k).requires_grad_(True) for k in ("means3d", "scales", "quats"))
    coeffs = t("sh_coeffs").requires_grad_(True)
    opac = t("opacities").reshape(-1, 1).requires_grad_(True)
    xys, depths, radii, conics, comp, nth, cov3d = project_gaussians(
        means, scales, 1, quats, t("viewmat"), t("projmat"), float(z["fx"]), float(z["fy"]), float(z["cx"]), float(z["cy"]),
        H, W, bw)
    xys.retain_grad()
    viewdirs = means.detach() - t("cam_pos")[None, :]
    rgbs = torch.clamp(spherical_harmonics(int(z["degrees_to_use"]), viewdirs, coeffs) + 0.5, min=0.0)
    img, alpha = rasterize_gaussians(xys, depths, radii, conics, nth, rgbs, opac, H, W, bw, background=t("background"),
                                     return_alpha=True)
    ((img * t("v_out_img")).sum() + (alpha * t("v_out_alpha")).sum()).backward()
    return dict(img=img, alpha=alpha, radii=radii, num_tiles_hit=nth, xys=xys, v_xy=xys.grad, v_mean3d=means.grad,
                v_scale=scales.grad, v_quat=quats.grad, v_coeffs=coeffs.grad, v_opacity=opac.grad)


r = wrappers_view()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    wrappers_view()
torch.cuda.synchronize()
out["A_wall_ms_per_view"] = np.array((time.perf_counter() - t0) * 100.0)
for k, v in r.items():
    out["A_" + k] = v.detach().cpu().numpy()


# ---- B. the reference model class
class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """absent third-party modules of gs_toolkit (viewer, metrics, mesh tools ...) -> inert stubs, for the import only"""
    ROOTS = ("viser", "pytorch_msssim", "torchmetrics", "open3d", "plyfile", "comet_ml", "wandb", "tyro", "cv2", "mediapy",
             "splines", "nerfacc", "tensorboard", "xatlas", "trimesh", "pymeshlab", "imageio", "PIL", "matplotlib", "skimage")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


sys.meta_path.append(_StubFinder())
from gs_toolkit.cameras.cameras import Cameras, CameraType  # noqa: E402
from gs_toolkit.models.vanilla_gs import GaussianSplattingModel, GaussianSplattingModelConfig  # noqa: E402

o = np.clip(z["opacities"].astype(np.float64), 1e-6, 1 - 1e-6)
params = {
    "means": t("means3d"), "scales": torch.log(t("scales")), "quats": t("quats"),
    "features_dc": t("sh_coeffs")[:, 0, :].contiguous(), "features_rest": t("sh_coeffs")[:, 1:, :].contiguous(),
    "opacities": torch.from_numpy(np.log(o / (1 - o)).astype(np.float32)).to(dev)[:, None],
}
m = object.__new__(GaussianSplattingModel)
torch.nn.Module.__init__(m)
m.config = GaussianSplattingModelConfig(sh_degree=int(z["sh_degree"]), background_color="black",
                                        output_depth_during_training=True, num_downscales=0)
m.device_indicator_param = torch.nn.Parameter(torch.empty(0, device=dev))
m.gauss_params = torch.nn.ParameterDict({k: torch.nn.Parameter(v.clone()) for k, v in params.items()})
m.step = 30000  # all SH bands on, full resolution
m.crop_box = None
m.train()
# camera: the scene's world-to-camera matrix (gsplat convention) back to a nerfstudio-style camera_to_world
V = z["viewmat"].astype(np.float64)
R_inv, T_inv = V[:3, :3], V[:3, 3:4]
R = R_inv.T @ np.diag([1.0, -1.0, -1.0])  # undo the y/z flip of vanilla_gs.py:731-735
T = -R_inv.T @ T_inv
c2w = torch.from_numpy(np.concatenate([R, T], axis=1).astype(np.float32))[None]
cam = Cameras(camera_to_worlds=c2w, fx=float(z["fx"]), fy=float(z["fy"]), cx=float(z["cx"]), cy=float(z["cy"]),
              width=W, height=H, camera_type=CameraType.PERSPECTIVE).to(dev)
outs = m.get_outputs(cam)
w_rgb, w_d = t("v_out_img"), (t("v_out_alpha") * 0.1)[..., None]
((outs["rgb"] * w_rgb).sum() + (outs["depth"] * w_d).sum()).backward()
out["B_rgb"] = outs["rgb"].detach().cpu().numpy()
out["B_depth"] = outs["depth"].detach().cpu().numpy()
out["B_v_xy"] = m.xys.grad.detach().cpu().numpy()
for k, p in m.gauss_params.items():
    out["B_grad_" + k] = p.grad.detach().cpu().numpy()
np.savez(out_path, **out)
print(f"[dropin child {mode}] native module: {native}; wrappers view {float(out['A_wall_ms_per_view']):.3f} ms wall")

"""Generates tests/golden/densify_*.npz by running the REFERENCE model's own methods — `after_train`,
`refinement_after` (with `split_gaussians`, `dup_gaussians`, `cull_gaussians`, `dup_in_all_optim`,
`remove_from_all_optim`) of gs_toolkit/models/vanilla_gs.py — and real `torch.optim.Adam(eps=1e-15)` steps
(engine/optimizers.py) on small seeded Gaussian sets, on CPU.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/gen_golden_densify.py
`gs_toolkit` cannot normally be imported here (viser, pytorch_msssim, torchmetrics, ... are absent and there is no
network): the missing third-party modules are replaced by inert stubs FOR THE IMPORT ONLY; none of the methods
exercised touches them, and the reference sources are used where they lie, unmodified.  The model object is created
without running `populate_modules` (which needs a dataset); the attributes the methods read are set by hand.
Inputs, the standard-normal draws of split_gaussians, and the reference's outputs are stored together.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types
from types import SimpleNamespace
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, "/root/reference/gs_toolkit/gs_components")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
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


sys.meta_path.append(_StubFinder())  # last: only consulted for modules that are really missing
from gs_toolkit.models.vanilla_gs import GaussianSplattingModel, GaussianSplattingModelConfig  # noqa: E402

GROUPS = ("means", "scales", "quats", "features_dc", "features_rest", "opacities")
LRS = {"means": 1.6e-4, "features_dc": 0.0025, "features_rest": 0.0025 / 20, "opacities": 0.05, "scales": 0.005, "quats": 0.001}
H, W = 540, 960


def make_params(n, gen):
    r = lambda *s: torch.rand(*s, generator=gen)
    g = lambda *s: torch.randn(*s, generator=gen)
    # log-scales spread across both size thresholds (densify 0.01, cull 0.5)
    scales = (r(n, 1) * (np.log(0.8) - np.log(0.001)) + np.log(0.001)) + torch.log(0.5 + 0.5 * r(n, 3))
    return {"means": g(n, 3) * 2.0, "scales": scales, "quats": g(n, 4), "features_dc": r(n, 3),
            "features_rest": g(n, 3, 3) * 0.05, "opacities": g(n, 1) * 2.0}


def make_model(params, cfg_over, step, num_train_data):
    m = object.__new__(GaussianSplattingModel)
    torch.nn.Module.__init__(m)
    m.config = GaussianSplattingModelConfig(**cfg_over)
    m.device_indicator_param = torch.nn.Parameter(torch.empty(0))
    m.gauss_params = torch.nn.ParameterDict({k: torch.nn.Parameter(v.clone()) for k, v in params.items()})
    m.step, m.num_train_data = step, num_train_data
    m.xys_grad_norm = m.vis_counts = m.max_2Dsize = None
    m.last_size = (H, W)
    return m


def run_case(name, n, step, seed, cfg_over=None, adam_steps=2, views=4, num_train_data=200):
    gen = torch.Generator().manual_seed(seed)
    params = make_params(n, gen)
    model = make_model(params, cfg_over or {}, step, num_train_data)
    opts = SimpleNamespace(optimizers={k: torch.optim.Adam([model.gauss_params[k]], lr=LRS[k], eps=1e-15) for k in GROUPS})
    out = {"meta_step": step, "meta_num_train_data": num_train_data, "meta_H": H, "meta_W": W,
           "meta_cfg": np.array(repr(sorted((cfg_over or {}).items())))}
    for k in GROUPS:
        out["in_" + k] = params[k].numpy().copy()
    # --- Adam steps with random gradients through the real torch.optim.Adam
    for s in range(adam_steps):
        for k in GROUPS:
            g = torch.randn(model.gauss_params[k].shape, generator=gen) * (10.0 ** float(torch.randint(-6, -1, (1,), generator=gen)))
            g[torch.rand(n, generator=gen) < 0.3] = 0.0   # invisible Gaussians get exact zeros
            model.gauss_params[k].grad = g
            out[f"adam{s}_grad_{k}"] = g.numpy().copy()
        for k in GROUPS:
            opts.optimizers[k].step()
    for k in GROUPS:   # state after the last step (the input of the refinement)
        st = opts.optimizers[k].state[model.gauss_params[k]]
        out[f"adam_p_{k}"] = model.gauss_params[k].detach().numpy().copy()
        out[f"adam_m_{k}"] = st["exp_avg"].numpy().copy()
        out[f"adam_v_{k}"] = st["exp_avg_sq"].numpy().copy()
    # --- running statistics over a few views (the reference's after_train)
    for v in range(views):
        radii = torch.randint(-2, 140, (n,), generator=gen, dtype=torch.int32)
        radii[torch.rand(n, generator=gen) < 0.25] = 0
        xg = torch.randn(n, 2, generator=gen) * 4e-7 * torch.exp(torch.randn(n, 1, generator=gen) * 1.5)
        xg[radii <= 0] = 0.0
        model.radii = radii
        model.xys = torch.zeros(n, 2, requires_grad=True)
        model.xys.grad = xg
        model.after_train(step)
        out[f"view{v}_radii"], out[f"view{v}_xys_grad"] = radii.numpy().copy(), xg.numpy().copy()
    out["meta_views"] = views
    out["meta_adam_steps"] = adam_steps
    for k in ("xys_grad_norm", "vis_counts", "max_2Dsize"):
        if getattr(model, k) is not None:   # after_train returns early once step >= stop_split_at (:347-348)
            out["stats_" + k] = getattr(model, k).numpy().copy()
    # --- refinement; the draws of split_gaussians are reproduced by re-seeding torch's global CPU generator
    torch.manual_seed(seed + 1000)
    model.refinement_after(opts, step)
    n_after = model.num_points
    for k in GROUPS:
        out["ref_" + k] = model.gauss_params[k].detach().numpy().copy()
        st = opts.optimizers[k].state[opts.optimizers[k].param_groups[0]["params"][0]]
        out["ref_m_" + k], out["ref_v_" + k] = st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()
    out["meta_n_after"] = n_after
    return out, model


def samples_for(out, n_split, n_samples, seed):
    """torch.randn((samps * n_splits, 3)) right after torch.manual_seed(seed + 1000) — what split_gaussians drew."""
    torch.manual_seed(seed + 1000)
    return torch.randn((n_samples * n_split, 3)).numpy()


def main():
    cases = [
        ("a_densify_cullbig_screen", 600, 3500, 1, {}),
        ("b_densify_early", 600, 1000, 2, {}),
        ("c_densify_late_noscreen", 600, 6500, 3, {"n_split_samples": 3}),
        ("d_cull_only", 600, 12000, 4, {}),
        ("e_opacity_reset", 400, 3100, 5, {}),
        ("f_warmup_noop", 200, 400, 6, {}),
    ]
    for name, n, step, seed, cfg_over in cases:
        out, model = run_case(name, n, step, seed, cfg_over)
        # number of splits = what the oracle restatement finds from the same inputs (checked against the reference's
        # final count below); store exactly the draws the reference consumed
        cfg = {f: getattr(model.config, f) for f in ("warmup_length", "refine_every", "cull_alpha_thresh", "cull_scale_thresh",
                                                      "continue_cull_post_densification", "reset_alpha_every", "densify_grad_thresh",
                                                      "densify_size_thresh", "n_split_samples", "cull_screen_size", "split_screen_size",
                                                      "stop_screen_size_at", "stop_split_at")}
        p = {k: torch.from_numpy(out[f"adam_p_{k}"]) for k in GROUPS}
        stats = {k: torch.from_numpy(out["stats_" + k]) for k in ("xys_grad_norm", "vis_counts", "max_2Dsize") if "stats_" + k in out}
        reset_interval = cfg["reset_alpha_every"] * cfg["refine_every"]
        do_dens = step > cfg["warmup_length"] and step < cfg["stop_split_at"] and step % reset_interval > 200 + cfg["refine_every"]
        n_split = 0
        if do_dens:
            avg = (stats["xys_grad_norm"] / stats["vis_counts"]) * 0.5 * max(H, W)
            splits = p["scales"].exp().max(dim=-1).values > cfg["densify_size_thresh"]
            if step < cfg["stop_screen_size_at"]:
                splits |= stats["max_2Dsize"] > cfg["split_screen_size"]
            splits &= avg > cfg["densify_grad_thresh"]
            n_split = int(splits.sum())
        out["samples"] = samples_for(out, n_split, cfg["n_split_samples"], seed)
        for k, v in cfg.items():
            out["cfg_" + k] = v
        path = os.path.join(HERE, f"densify_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: N {n} -> {out['meta_n_after']}  (n_split={n_split})  {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()

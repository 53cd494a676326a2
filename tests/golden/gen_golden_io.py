"""Generates tests/golden/io_*.{npz,json,ckpt} with the REFERENCE's own code (build container only):
  * io_transforms.json + io_transforms_ref.npz — a small synthetic `transforms.json` and what the reference dataparser
    (gs_toolkit/data/dataparsers/gs_toolkit_dataparser.py) makes of it for the train and the val split;
  * io_cameras_ref.npz — viewmat / projmat the reference model builds from those cameras (the literal prologue of
    GaussianSplattingModel.get_outputs, models/vanilla_gs.py:722-741, with utils/comms.py projection_matrix);
  * io_ref_step-000000123.ckpt — a checkpoint dict assembled the way Trainer.save_checkpoint does (engine/trainer.py:
    443-476) from a reference GaussianSplattingModel inside a pipeline-like module and real torch.optim.Adam states.
Missing third-party modules are stubbed for the import only (see gen_golden_densify.py)."""
import json
import math
import os
import sys
from pathlib import Path

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden_densify as gd  # noqa: E402  (installs the stub finder, imports the reference model)

from gs_toolkit.data.dataparsers.gs_toolkit_dataparser import GSToolkitDataParserConfig  # noqa: E402
from gs_toolkit.utils.comms import projection_matrix  # noqa: E402


def make_transforms(path, n=11):
    g = np.random.default_rng(3)
    frames = []
    for i in range(n):
        ang = 2 * math.pi * i / n
        pos = np.array([3 * math.cos(ang), 3 * math.sin(ang), 1.0 + 0.3 * math.sin(3 * ang)])
        fwd = -pos / np.linalg.norm(pos)               # camera looks at the origin along -z (OpenGL)
        up0 = np.array([0.1 * g.normal(), 0.1 * g.normal(), 1.0])
        right = np.cross(fwd, up0); right /= np.linalg.norm(right)
        up = np.cross(right, fwd)
        c2w = np.eye(4)
        c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, up, -fwd, pos
        frames.append({"file_path": f"images/frame_{(7 * i) % n:05d}.png", "transform_matrix": c2w.tolist(),
                       "fl_x": 500.0 + i, "fl_y": 505.0 + i})
    meta = {"w": 640, "h": 480, "cx": 320.5, "cy": 239.5, "k1": 0.0, "k2": 0.0, "p1": 0.0, "p2": 0.0, "frames": frames}
    json.dump(meta, open(path, "w"), indent=1)


def main():
    tj = os.path.join(HERE, "io_transforms.json")
    make_transforms(tj)
    out = {}
    for split in ("train", "val"):
        cfg = GSToolkitDataParserConfig(data=Path(tj), downscale_factor=1)
        parser = cfg.setup()
        dpo = parser._generate_dataparser_outputs(split)
        cams = dpo.cameras
        out[f"{split}_c2w"] = cams.camera_to_worlds.numpy()
        for k in ("fx", "fy", "cx", "cy", "height", "width"):
            out[f"{split}_{k}"] = getattr(cams, k).numpy().reshape(-1)
        out[f"{split}_files"] = np.array([os.path.basename(str(f)) for f in dpo.image_filenames])
        out[f"{split}_transform"] = dpo.dataparser_transform.numpy()
        out[f"{split}_scale"] = dpo.dataparser_scale
        if split == "train":
            vm, pm = [], []
            for i in range(cams.camera_to_worlds.shape[0]):
                camera = cams[i:i + 1]
                # --- literal lines of vanilla_gs.py:722-741
                R = camera.camera_to_worlds[0, :3, :3]
                T = camera.camera_to_worlds[0, :3, 3:4]
                R_edit = torch.diag(torch.tensor([1, -1, -1], dtype=R.dtype))
                R = R @ R_edit
                R_inv = R.T
                T_inv = -R_inv @ T
                viewmat = torch.eye(4, dtype=R.dtype)
                viewmat[:3, :3] = R_inv
                viewmat[:3, 3:4] = T_inv
                fovx = 2 * math.atan(camera.width / (2 * camera.fx))
                fovy = 2 * math.atan(camera.height / (2 * camera.fy))
                projmat = projection_matrix(0.001, 1000, fovx, fovy)
                vm.append(viewmat.numpy())
                pm.append((projmat.squeeze() @ viewmat.squeeze()).numpy())
            np.savez_compressed(os.path.join(HERE, "io_cameras_ref.npz"), viewmat=np.stack(vm), projmat=np.stack(pm))
    np.savez_compressed(os.path.join(HERE, "io_transforms_ref.npz"), **out)

    # --- checkpoint
    gen = torch.Generator().manual_seed(0)
    params = gd.make_params(40, gen)
    model = gd.make_model(params, {}, 123, 10)
    pipe = torch.nn.Module()
    pipe._model = model
    opts = {k: torch.optim.Adam([model.gauss_params[k]], lr=gd.LRS[k], eps=1e-15) for k in gd.GROUPS}
    for _ in range(2):
        for k in gd.GROUPS:
            model.gauss_params[k].grad = torch.randn(model.gauss_params[k].shape, generator=gen) * 1e-3
            opts[k].step()
    ck = {"step": 123, "pipeline": pipe.state_dict(), "optimizers": {k: v.state_dict() for k, v in opts.items()},
          "schedulers": {}, "scalers": torch.amp.GradScaler("cpu", enabled=False).state_dict()}
    torch.save(ck, os.path.join(HERE, "io_ref_step-000000123.ckpt"))
    print({k: tuple(v.shape) for k, v in ck["pipeline"].items()})
    print("train files", out["train_files"], "val files", out["val_files"])


def gen_scheduler():
    """io_scheduler_ref.npz: learning rates of the reference's ExponentialDecayScheduler (engine/schedulers.py:94-135)
    driving a real torch.optim.Adam through LambdaLR, for the `means` configuration of configs/method_configs.py:98-105 and
    for a warm-up variant."""
    from gs_toolkit.engine.schedulers import ExponentialDecaySchedulerConfig

    out = {}
    for tag, lr_init, kw in (("means", 1.6e-4, dict(lr_final=1.6e-6, max_steps=30000)),
                             ("warm_cos", 1e-3, dict(lr_final=1e-4, max_steps=500, warmup_steps=100)),
                             ("warm_lin", 1e-3, dict(lr_final=None, max_steps=400, warmup_steps=50, ramp="linear"))):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adam([p], lr=lr_init, eps=1e-15)
        sched = ExponentialDecaySchedulerConfig(**kw).setup().get_scheduler(optimizer=opt, lr_init=lr_init)
        lrs = []
        for _ in range(600):
            opt.step()
            sched.step()
            lrs.append(sched.get_last_lr()[0])
        out[tag] = np.array(lrs, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "io_scheduler_ref.npz"), **out)


if __name__ == "__main__":
    main()
    gen_scheduler()

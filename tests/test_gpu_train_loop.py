"""cfg3-style integration test (BASELINE configs[2], restated because gs_toolkit cannot be imported here — viser,
comet_ml, pytorch_msssim, torchmetrics are absent): a short training run of the reference model's per-iteration maths
(gs_toolkit/models/vanilla_gs.py:759-947, engine/trainer.py:478-498, configs/method_configs.py:98-125) on a synthetic
multi-view scene, three times from the same initial state:
  (ref)   the unmodified reference CUDA extension behind reference-style autograd wrappers + torch glue,
  (ours)  this package's drop-in operators + the same torch glue,
  (fused) this package's fused operator,
  (fused+loss) the fused operator and the fused L1+SSIM loss kernel.
Loss = 0.8 L1 + 0.2 (1 - SSIM) (SSIM restated from pytorch_msssim defaults: 11x11 Gaussian window, sigma 1.5, valid
padding), six Adam groups with the reference learning rates, no densification (f2 is out of scope).  Checks: the loss
goes down, the three runs reach the same PSNR, and records iterations/s (gpurun_out/train_loop.json)."""
import json
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, W, BW = 240, 320, 16
LRS = {"means": 1.6e-4, "features_dc": 0.0025, "features_rest": 0.0025 / 20, "opacities": 0.05, "scales": 0.005, "quats": 0.001}


def _ssim(a, b):
    """pytorch_msssim.ssim(a, b, data_range=1.0, size_average=True) for [1,3,H,W] tensors."""
    g = torch.arange(11, dtype=torch.float32, device=a.device) - 5
    g = torch.exp(-(g ** 2) / (2 * 1.5 ** 2))
    g = (g / g.sum())
    win_h, win_v = g.view(1, 1, 1, 11).repeat(3, 1, 1, 1), g.view(1, 1, 11, 1).repeat(3, 1, 1, 1)

    def blur(x):
        return F.conv2d(F.conv2d(x, win_v, groups=3), win_h, groups=3)

    c1, c2 = 0.01 ** 2, 0.03 ** 2
    mu1, mu2 = blur(a), blur(b)
    s11, s22, s12 = blur(a * a) - mu1 * mu1, blur(b * b) - mu2 * mu2, blur(a * b) - mu1 * mu2
    cs = (2 * s12 + c2) / (s11 + s22 + c2)
    return (((2 * mu1 * mu2 + c1) / (mu1 * mu1 + mu2 * mu2 + c1)) * cs).mean()


def _cameras(scene):
    from rasterizer.synthetic import look_at_viewmat, projection_matrix

    fovx = math.radians(60.0)
    fovy = 2.0 * math.atan(0.5 * H / scene["fy"])
    P = projection_matrix(0.001, 1000.0, fovx, fovy).astype(np.float64)
    cams = []
    for k in range(24):
        V = look_at_viewmat(yaw_deg=-24.0 + 2.0 * k, pitch_deg=6.0 * math.sin(k))
        PM = (P @ V.astype(np.float64)).astype(np.float32)
        cam_pos = (-V[:3, :3].T.astype(np.float64) @ V[:3, 3].astype(np.float64)).astype(np.float32)
        cams.append(tuple(torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (V, PM, cam_pos)))
    return cams


def test_short_training_run_three_backends():
    import rasterizer
    from oracle import oracle as orc
    from oracle.build_ref import load_ref
    from rasterizer.fused import render_gaussians
    from rasterizer.losses import l1_ssim_loss
    from rasterizer.sh import spherical_harmonics
    from rasterizer.synthetic import make_scene
    from ref_autograd import make_ops

    teacher = make_scene(40_000, W, H, 0.02, 0.12, margin=1.0, seed=100)
    student = make_scene(40_000, W, H, 0.02, 0.12, margin=1.0, seed=101)
    cams = _cameras(teacher)
    fx, fy, cx, cy = teacher["fx"], teacher["fy"], teacher["cx"], teacher["cy"]
    bg = torch.zeros(3, device="cuda")

    def params_of(scene, grad):
        raw = orc.raw_parameters(scene)
        d = {"means": scene["means3d"], "scales": raw["scales_raw"], "quats": raw["quats_raw"], "features_dc": raw["features_dc"],
             "features_rest": raw["features_rest"], "opacities": raw["opacities_raw"]}
        return {k: torch.from_numpy(np.ascontiguousarray(v)).cuda().requires_grad_(grad) for k, v in d.items()}

    def render_fused(p, cam):
        V, PM, _ = cam
        rgb, _, _ = render_gaussians(p["means"], p["scales"], p["quats"], p["features_dc"], p["features_rest"], p["opacities"], V, PM,
                                     fx, fy, cx, cy, H, W, 3, background=bg, render_depth=False)
        return rgb

    def render_glue(ops):
        sh_fn, proj_fn, rast_fn = ops

        def fn(p, cam):
            V, PM, cam_pos = cam
            scales, quats = torch.exp(p["scales"]), p["quats"] / p["quats"].norm(dim=-1, keepdim=True)
            coeffs = torch.cat((p["features_dc"][:, None, :], p["features_rest"]), dim=1)
            xys, depths, radii, conics, comp, nth, cov3d = proj_fn(p["means"], scales, 1.0, quats, V, PM, fx, fy, cx, cy, H, W, BW, 0.01)
            rgbs = torch.clamp(sh_fn(3, (p["means"].detach() - cam_pos[None]).contiguous(), coeffs) + 0.5, min=0.0)
            return rast_fn(xys, depths, radii, conics, nth, rgbs, torch.sigmoid(p["opacities"]))

        return fn

    ours_ops = (spherical_harmonics, rasterizer.project_gaussians,
                lambda xys, d, r, c, n_, col, op_: rasterizer.rasterize_gaussians(xys, d, r, c, n_, col, op_, H, W, BW, background=bg))
    backends = {"ours": render_glue(ours_ops), "fused": render_fused, "fused+loss": render_fused}
    ref_ext = load_ref()
    if ref_ext is not None:
        r_sh, r_proj, r_rast = make_ops(ref_ext)
        backends["ref"] = render_glue((r_sh, r_proj, lambda xys, d, r, c, n_, col, op_: r_rast(xys, d, r, c, n_, col, op_, H, W, BW, bg)[0]))

    with torch.no_grad():
        tp = params_of(teacher, False)
        gts = [torch.clamp(render_fused(tp, cam), max=1.0) for cam in cams]

    iters, warm = 200, 10
    report = {"workload": f"{len(cams)} views {W}x{H}, 40k Gaussians, {iters} Adam iterations, L1 + 0.2 (1 - SSIM), no densification"}
    for name, render in backends.items():
        p = params_of(student, True)
        opt = torch.optim.Adam([{"params": [p[k]], "lr": lr, "name": k} for k, lr in LRS.items()], eps=1e-15)
        losses = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(iters + warm):
            if it == warm:
                e0.record()
            cam, gt = cams[it % len(cams)], gts[it % len(cams)]
            opt.zero_grad(set_to_none=True)
            pred = torch.clamp(render(p, cam), max=1.0)
            if name == "fused+loss":   # the fused photometric loss kernel (f3) instead of the torch formulation
                loss = l1_ssim_loss(pred, gt, 0.2)
            else:
                l1 = (gt - pred).abs().mean()
                sim = 1 - _ssim(gt.permute(2, 0, 1)[None], pred.permute(2, 0, 1)[None])
                loss = 0.8 * l1 + 0.2 * sim
            loss.backward()
            opt.step()
            if it % 20 == 0 or it == iters + warm - 1:
                losses.append(float(loss))
        e1.record()
        torch.cuda.synchronize()
        with torch.no_grad():
            mse = torch.stack([((torch.clamp(render(p, cam), max=1.0) - gt) ** 2).mean() for cam, gt in zip(cams, gts)]).mean()
        psnr = float(-10 * torch.log10(mse))
        report[name] = {"iters_per_s": iters / (e0.elapsed_time(e1) * 1e-3), "loss_first": losses[0], "loss_last": losses[-1], "psnr": psnr}
        print(f"[train loop] {name:5s}: {report[name]}")
        assert losses[-1] < 0.85 * losses[0], f"{name}: loss did not decrease ({losses[0]:.4f} -> {losses[-1]:.4f})"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(report, open(os.path.join(ROOT, "gpurun_out", "train_loop.json"), "w"), indent=1)
    assert abs(report["ours"]["psnr"] - report["fused"]["psnr"]) < 0.3
    assert abs(report["fused+loss"]["psnr"] - report["fused"]["psnr"]) < 0.3
    if "ref" in report:
        assert abs(report["ours"]["psnr"] - report["ref"]["psnr"]) < 0.3
        assert abs(report["fused"]["psnr"] - report["ref"]["psnr"]) < 0.3

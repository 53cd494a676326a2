"""Generates tests/golden/refcuda_*.npz ON THE GPU BOX by running the compiled, UNMODIFIED reference CUDA
extension (oracle/_ref/rasterizer_ref_cuda.so, built from /root/reference by oracle/build_ref.py with the
reference's packaged flags) on small seeded scenes:
    gpurun -- python tests/golden/gen_golden_ref_cuda.py      # writes gpurun_out/golden/*.npz
then copy gpurun_out/golden/*.npz to tests/golden/ and commit.  Inputs and reference outputs are stored
together; forward AND backward (all six parameter gradients) are recorded."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "gaussian-splatting-toolkit_b200"), os.path.dirname(HERE)):
    sys.path.insert(0, p)


def main():
    from oracle.build_ref import load_ref
    from pipelines import run_view_bindings
    from rasterizer.synthetic import look_at_viewmat, make_scene, scene_to_torch

    ref = load_ref()
    assert ref is not None, "oracle/_ref/rasterizer_ref_cuda.so missing"
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    cases = {
        # ragged image, clipped / culled Gaussians (margin > 1), SH degree 3
        "a_2k_144x100": make_scene(2000, 144, 100, 0.02, 0.25, margin=1.2, seed=21),
        # rotated camera, block width 8, 2 of 3 SH degrees, many opaque Gaussians (early termination)
        "b_1500_96x96_bw8": make_scene(1500, 96, 96, 0.05, 0.4, margin=0.9, seed=22, block_width=8, degrees_to_use=2,
                                       viewmat=look_at_viewmat(yaw_deg=15.0, pitch_deg=-8.0, shift=(0.1, -0.05, 0.3))),
    }
    for name, scene in cases.items():
        s = scene_to_torch(scene, "cuda")
        out = run_view_bindings(ref, s, backward=True, sort_impl="torch")
        torch.cuda.synchronize()
        blob = {("in_" + k): v for k, v in scene.items()}
        for k, v in out.items():
            blob["ref_" + k] = v.cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
        np.savez_compressed(os.path.join(out_dir, f"refcuda_{name}.npz"), **blob)
        print(name, "M =", out["num_intersects"], "visible =", int((out["radii"] > 0).sum()))
    gen_model_step(ref, out_dir)


def gen_model_step(ref, out_dir):
    """Golden vectors for the model-level view: RAW parameters -> rgb / depth / alpha and the gradients of
    the six raw parameter tensors, computed by the reference extension behind reference-style autograd wrappers
    (tests/ref_autograd.py) + the torch glue of gs_toolkit/models/vanilla_gs.py:759-855."""
    from oracle import oracle as orc
    from rasterizer.synthetic import look_at_viewmat, make_scene, scene_to_torch
    from ref_autograd import make_ops

    sh_fn, proj_fn, rast_fn = make_ops(ref)
    scene = make_scene(1800, 128, 96, 0.03, 0.25, margin=1.1, seed=31,
                       viewmat=look_at_viewmat(yaw_deg=-9.0, pitch_deg=4.0, shift=(0.0, 0.1, 0.1)))
    raw = orc.raw_parameters(scene)
    s = scene_to_torch(scene, "cuda")
    H, W, bw = 96, 128, 16
    p = {k: torch.from_numpy(v).cuda().requires_grad_(True) for k, v in raw.items()}
    means = s["means3d"].clone().requires_grad_(True)
    scales, quats = torch.exp(p["scales_raw"]), p["quats_raw"] / p["quats_raw"].norm(dim=-1, keepdim=True)
    coeffs = torch.cat((p["features_dc"][:, None, :], p["features_rest"]), dim=1)
    xys, depths, radii, conics, comp, nth, cov3d = proj_fn(means, scales, 1.0, quats, s["viewmat"], s["projmat"], s["fx"],
                                                           s["fy"], s["cx"], s["cy"], H, W, bw, 0.01)
    rgbs = torch.clamp(sh_fn(3, (means.detach() - s["cam_pos"][None]).contiguous(), coeffs) + 0.5, min=0.0)
    opac = torch.sigmoid(p["opacities_raw"])
    rgb, alpha = rast_fn(xys, depths, radii, conics, nth, rgbs, opac, H, W, bw, s["background"])
    depth, _ = rast_fn(xys, depths, radii, conics, nth, depths[:, None].repeat(1, 3), opac, H, W, bw,
                       torch.zeros(3, device="cuda"))
    g = torch.Generator().manual_seed(7)
    v_rgb = ((torch.rand(H, W, 3, generator=g) - 0.5) * 2e-3).cuda()
    v_depth = ((torch.rand(H, W, generator=g) - 0.5) * 2e-4).cuda()
    v_alpha = ((torch.rand(H, W, generator=g) - 0.5) * 2e-3).cuda()
    torch.autograd.backward([rgb, depth[..., 0], alpha], [v_rgb, v_depth, v_alpha])
    blob = {("in_" + k): v for k, v in scene.items()}
    blob.update({("raw_" + k): v for k, v in raw.items()})
    blob.update(up_v_rgb=v_rgb.cpu().numpy(), up_v_depth=v_depth.cpu().numpy(), up_v_alpha=v_alpha.cpu().numpy(),
                ref_rgb=rgb.detach().cpu().numpy(), ref_depth=depth[..., 0].detach().cpu().numpy(),
                ref_alpha=alpha.detach().cpu().numpy(), ref_radii=radii.cpu().numpy(),
                ref_v_means3d=means.grad.cpu().numpy(), ref_v_scales_raw=p["scales_raw"].grad.cpu().numpy(),
                ref_v_quats_raw=p["quats_raw"].grad.cpu().numpy(), ref_v_opacities_raw=p["opacities_raw"].grad.cpu().numpy(),
                ref_v_features_dc=p["features_dc"].grad.cpu().numpy(), ref_v_features_rest=p["features_rest"].grad.cpu().numpy())
    np.savez_compressed(os.path.join(out_dir, "refcuda_modelstep_1800_128x96.npz"), **blob)
    print("modelstep fixture written")


if __name__ == "__main__":
    main()

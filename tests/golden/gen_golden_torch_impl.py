"""Generates tests/golden/torch_impl_*.npz by running the REFERENCE's pure-PyTorch restatement
(/root/reference/gs_toolkit/gs_components/rasterizer/_torch_impl.py) on small seeded scenes, on CPU.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/gen_golden_torch_impl.py
The reference file is loaded by path (it is not copied); inputs and the reference's outputs are stored
together so that the fixtures do not depend on torch's RNG stream.

Known defects of _torch_impl.py are avoided by construction: every Gaussian is visible
(the key-emission loop `break`s at the first culled one, _torch_impl.py:351-352), and `final_idx` is not
recorded (it stores the loop variable, not the last contributor, _torch_impl.py:457-467).
"""
import importlib.util
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting-toolkit_b200"))
REF = "/root/reference/gs_toolkit/gs_components/rasterizer/_torch_impl.py"


def load_ref():
    spec = importlib.util.spec_from_file_location("ref_torch_impl", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def run_case(ti, name, scene, prefix="torch_impl_"):
    s = scene
    t = {k: torch.from_numpy(v) if isinstance(v, np.ndarray) else v for k, v in s.items()}
    H, W, bw = s["img_height"], s["img_width"], s["block_width"]
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    N = s["means3d"].shape[0]
    viewdirs = t["means3d"] - t["cam_pos"][None]
    viewdirs_n = viewdirs / viewdirs.norm(dim=-1, keepdim=True)
    # the torch restatement evaluates every stored basis: hand it only the (degrees_to_use+1)^2 active ones
    k_use = (s["degrees_to_use"] + 1) ** 2
    rgb_sh = ti.compute_sh_color(viewdirs_n, t["sh_coeffs"][:, :k_use, :])
    colors = torch.clamp(rgb_sh + 0.5, min=0.0)
    (cov3d, cov2d, xys, depths, radii, conics, comp, nth, mask) = ti.project_gaussians_forward(
        t["means3d"], t["scales"], s["glob_scale"], t["quats"], t["viewmat"], t["projmat"],
        (s["fx"], s["fy"], s["cx"], s["cy"]), (W, H), bw, s["clip_thresh"])
    assert bool(mask.all()), f"{name}: scene must be all-visible ({int((~mask).sum())} culled)"
    cum = torch.cumsum(nth, dim=0, dtype=torch.int32)
    M = int(cum[-1])
    isect, gids = ti.map_gaussian_to_intersects(N, xys, depths, radii, cum, tb, bw)
    ks, order = torch.sort(isect, stable=True)
    vs = torch.gather(gids, 0, order)
    bins = ti.get_tile_bin_edges(M, ks, tb)
    t0 = time.time()
    img, fT, _ = ti.rasterize_forward(tb, (bw, bw, 1), (W, H, 1), vs, bins, xys, conics, colors,
                                      t["opacities"], t["background"])
    print(f"{name}: N={N} M={M} {W}x{H} bw={bw}: reference rasterize_forward {time.time()-t0:.1f}s")
    out = {("in_" + k): v for k, v in s.items()}
    out.update(
        ref_rgb_sh=rgb_sh.numpy(), ref_colors=colors.numpy(), ref_cov3d=cov3d.numpy(),
        ref_cov2d=cov2d.numpy(), ref_xys=xys.numpy(), ref_depths=depths.numpy(),
        ref_radii=radii.numpy(), ref_conics=conics.numpy(), ref_compensation=comp.numpy(),
        ref_num_tiles_hit=nth.numpy().astype(np.int32), ref_cum_tiles_hit=cum.numpy(),
        ref_isect_ids=isect.numpy(), ref_gaussian_ids=gids.numpy(), ref_isect_ids_sorted=ks.numpy(),
        ref_gaussian_ids_sorted=vs.numpy(), ref_tile_bins=bins.numpy(), ref_out_img=img.numpy(),
        ref_final_Ts=fT.numpy(),
    )
    np.savez_compressed(os.path.join(HERE, f"{prefix}{name}.npz"), **out)


def main():
    from rasterizer.synthetic import look_at_viewmat, make_scene

    ti = load_ref()
    torch.set_num_threads(os.cpu_count() or 1)
    # (a) identity camera, ragged image (not a multiple of 16), SH degree 3
    run_case(ti, "a_identity_48x40", make_scene(90, 48, 40, 0.05, 0.4, margin=0.7, seed=11))
    # (b) rotated + translated camera, block_width 8, SH degree 2 used out of 3
    run_case(ti, "b_rotated_40x32_bw8",
             make_scene(70, 40, 32, 0.05, 0.3, margin=0.45, seed=12, block_width=8, degrees_to_use=2,
                        viewmat=look_at_viewmat(yaw_deg=12.0, pitch_deg=-7.0, shift=(0.1, -0.05, 0.2))))
    # (c), (d): CPU-side pins of the oracle only (`oracle_pin_` prefix: tests/test_oracle_golden.py); the GPU suite keeps
    # its two torch_impl fixtures
    # (c) tall ragged image, block_width 4, SH degree 1 stored and used (every tile must hold a Gaussian: the reference
    #     restatement reads an unbound `idx` for pixels of an empty tile, _torch_impl.py:467)
    run_case(ti, "c_tall_24x54_bw4_deg1", make_scene(150, 24, 54, 0.2, 0.6, margin=1.0, seed=13, block_width=4, sh_degree=1),
             prefix="oracle_pin_torch_impl_")
    # (d) large, mostly opaque Gaussians (early termination, T <= 1e-4), SH degree 0
    run_case(ti, "d_opaque_64x48_deg0", make_scene(120, 64, 48, 0.3, 0.9, margin=0.6, seed=14, sh_degree=0),
             prefix="oracle_pin_torch_impl_")


if __name__ == "__main__":
    main()

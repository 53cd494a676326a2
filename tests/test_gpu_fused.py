"""GPU parity of the FUSED render operator (rasterizer.fused.render_gaussians) against
(i) the oracle restatement of what the reference model computes per view from its raw parameters and
(ii) the separate operators of this package driven exactly like the reference model drives them."""
import numpy as np
import pytest
import torch

from parity import assert_float_parity, to_np

pytestmark = pytest.mark.gpu


def _cases():
    from rasterizer.synthetic import look_at_viewmat, make_scene

    return {
        "8k_256x192_deg3": lambda: make_scene(8000, 256, 192, 0.02, 0.15, margin=1.1, seed=61),
        "3k_120x90_bw8_deg2of3_rot": lambda: make_scene(3000, 120, 90, 0.03, 0.3, margin=1.0, seed=62, block_width=8,
                                                         degrees_to_use=2,
                                                         viewmat=look_at_viewmat(yaw_deg=14.0, pitch_deg=-6.0, shift=(0.1, 0.0, 0.2))),
        "1k_64x64_deg1": lambda: make_scene(1000, 64, 64, 0.05, 0.3, margin=0.9, seed=63, sh_degree=1),
        "2k_96x64_deg4_opaque": lambda: make_scene(2000, 96, 64, 0.08, 0.5, margin=0.9, seed=64, sh_degree=4),
    }


def _to_cuda(d):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in d.items()}


@pytest.mark.parametrize("name", list(_cases().keys()))
def test_fused_render_vs_oracle(oracle, name):
    from rasterizer.fused import RenderAux, render_gaussians
    from rasterizer.synthetic import scene_to_torch

    scene = _cases()[name]()
    raw = oracle.raw_parameters(scene)
    H, W = scene["img_height"], scene["img_width"]
    g = np.random.default_rng(5)
    v_rgb = ((g.random((H, W, 3)) - 0.5) * 2e-3).astype(np.float32)
    v_depth = ((g.random((H, W)) - 0.5) * 2e-4).astype(np.float32)
    v_alpha = ((g.random((H, W)) - 0.5) * 2e-3).astype(np.float32)
    ref = oracle.render_fused_reference(scene, raw, v_rgb, v_depth, v_alpha)

    s = scene_to_torch(scene, "cuda")
    p = {k: v.requires_grad_(True) for k, v in _to_cuda(raw).items()}
    means = s["means3d"].clone().requires_grad_(True)
    aux = RenderAux()
    rgb, depth, alpha = render_gaussians(
        means, p["scales_raw"], p["quats_raw"], p["features_dc"], p["features_rest"], p["opacities_raw"], s["viewmat"],
        s["projmat"], s["fx"], s["fy"], s["cx"], s["cy"], H, W, scene["degrees_to_use"], background=s["background"],
        block_width=scene["block_width"], render_depth=True, aux=aux)
    assert rgb.shape == (H, W, 3) and depth.shape == (H, W, 1) and alpha.shape == (H, W, 1)
    clean = ref["ambiguous"] == 0
    assert_float_parity(rgb, ref["rgb"], "rgb", mask=np.broadcast_to(clean[..., None], ref["rgb"].shape), max_frac_bad=1e-5)
    assert_float_parity(depth[..., 0], ref["depth"], "depth", mask=clean, max_frac_bad=1e-5)
    assert_float_parity(alpha[..., 0], ref["alpha"], "alpha", mask=clean, max_frac_bad=1e-5, atol=1e-6)
    assert int((to_np(aux.radii) != ref["radii"]).sum()) <= max(1, int(2e-4 * ref["radii"].size))
    torch.autograd.backward([rgb, depth, alpha], [torch.from_numpy(v_rgb).cuda(), torch.from_numpy(v_depth).cuda()[..., None],
                                                  torch.from_numpy(v_alpha).cuda()[..., None]])
    got = dict(v_means3d=means.grad, v_scales_raw=p["scales_raw"].grad, v_quats_raw=p["quats_raw"].grad,
               v_opacities_raw=p["opacities_raw"].grad, v_features_dc=p["features_dc"].grad,
               v_features_rest=p["features_rest"].grad, v_xy=aux.xys_grad)
    for k, v in got.items():
        assert_float_parity(to_np(v).reshape(ref[k].shape), ref[k], k, max_norm_rel=5e-5, max_frac_bad=5e-4)


@pytest.mark.parametrize("mode", ["classic", "antialiased"])
def test_fused_render_equals_separate_operators(mode):
    """Same image / depth / alpha and the same raw-parameter gradients as the three operators + torch glue the way
    gs_toolkit/models/vanilla_gs.py:759-855 chains them, in both rasterize modes (antialiased: opacities multiplied by
    the EWA compensation factor, :813-816, whose gradient flows back through project_gaussians)."""
    import rasterizer
    from oracle import oracle as orc
    from rasterizer.fused import render_gaussians
    from rasterizer.sh import spherical_harmonics
    from rasterizer.synthetic import make_scene, scene_to_torch

    scene = make_scene(20_000, 320, 240, 0.01, 0.1, margin=1.1, seed=66)
    raw = _to_cuda(orc.raw_parameters(scene))
    s = scene_to_torch(scene, "cuda")
    H, W, bw = 240, 320, 16

    def leaves():
        return [s["means3d"].clone().requires_grad_(True)] + [raw[k].clone().requires_grad_(True) for k in
                                                               ("scales_raw", "quats_raw", "features_dc", "features_rest", "opacities_raw")]

    # (a) the model's way
    means, sc, q, dc, rest, op = a_leaves = leaves()
    scales, quats = torch.exp(sc), q / q.norm(dim=-1, keepdim=True)
    colors_all = torch.cat((dc[:, None, :], rest), dim=1)
    xys, depths, radii, conics, comp, nth, cov3d = rasterizer.project_gaussians(
        means, scales, 1.0, quats, s["viewmat"], s["projmat"], s["fx"], s["fy"], s["cx"], s["cy"], H, W, bw)
    viewdirs = means.detach() - s["cam_pos"][None]
    rgbs = torch.clamp(spherical_harmonics(3, viewdirs, colors_all) + 0.5, min=0.0)
    opac = torch.sigmoid(op) * comp[:, None] if mode == "antialiased" else torch.sigmoid(op)
    rgb_a, alpha_a = rasterizer.rasterize_gaussians(xys, depths, radii, conics, nth, rgbs, opac, H, W, bw,
                                                    background=s["background"], return_alpha=True)
    depth_a = rasterizer.rasterize_gaussians(xys, depths, radii, conics, nth, depths[:, None].repeat(1, 3), opac, H, W, bw,
                                             background=torch.zeros(3, device="cuda"))[..., 0:1]
    # (b) fused
    b_leaves = leaves()
    rgb_b, depth_b, alpha_b = render_gaussians(b_leaves[0], b_leaves[1], b_leaves[2], b_leaves[3], b_leaves[4], b_leaves[5],
                                               s["viewmat"], s["projmat"], s["fx"], s["fy"], s["cx"], s["cy"], H, W, 3,
                                               background=s["background"], rasterize_mode=mode)
    assert_float_parity(rgb_b, rgb_a, "rgb", max_frac_bad=2e-5)
    assert_float_parity(depth_b, depth_a, "depth", max_frac_bad=2e-5)
    assert_float_parity(alpha_b[..., 0], alpha_a, "alpha", max_frac_bad=2e-5, atol=1e-6)
    g = torch.Generator(device="cuda").manual_seed(3)
    w_rgb = (torch.rand(H, W, 3, device="cuda", generator=g) - 0.5) * 2e-3
    w_d = (torch.rand(H, W, 1, device="cuda", generator=g) - 0.5) * 2e-4
    w_a = (torch.rand(H, W, device="cuda", generator=g) - 0.5) * 2e-3
    torch.autograd.backward([rgb_a, depth_a, alpha_a], [w_rgb, w_d, w_a])
    torch.autograd.backward([rgb_b, depth_b, alpha_b], [w_rgb, w_d, w_a[..., None]])
    for name, la, lb in zip(("means3d", "scales", "quats", "features_dc", "features_rest", "opacities"), a_leaves, b_leaves):
        assert_float_parity(lb.grad, la.grad, "grad " + name, max_norm_rel=1e-4, max_frac_bad=2e-3)


def test_fused_render_empty_and_no_depth():
    from oracle import oracle as orc
    from rasterizer.fused import render_gaussians
    from rasterizer.synthetic import make_scene, scene_to_torch

    scene = make_scene(500, 48, 32, 0.05, 0.2, seed=67)
    raw = _to_cuda(orc.raw_parameters(scene))
    s = scene_to_torch(scene, "cuda")
    args = (raw["scales_raw"], raw["quats_raw"], raw["features_dc"], raw["features_rest"], raw["opacities_raw"],
            s["viewmat"], s["projmat"], s["fx"], s["fy"], s["cx"], s["cy"], 32, 48, 3)
    rgb, depth, alpha = render_gaussians(s["means3d"], *args, background=s["background"], render_depth=False)
    assert depth is None and rgb.shape == (32, 48, 3)
    rgb2, depth2, alpha2 = render_gaussians(s["means3d"], *args, background=s["background"], render_depth=True)
    assert torch.equal(rgb, rgb2) and torch.equal(alpha, alpha2)
    with pytest.raises(ValueError, match="Unknown rasterize_mode"):
        render_gaussians(s["means3d"], *args, rasterize_mode="fancy")
    behind = s["means3d"].clone()
    behind[:, 2] *= -1
    behind.requires_grad_(True)
    rgb3, depth3, alpha3 = render_gaussians(behind, *args, background=s["background"])
    assert torch.allclose(rgb3, s["background"].expand(32, 48, 3)) and float(alpha3.abs().sum()) == 0.0
    (rgb3.sum() + depth3.sum()).backward()
    assert float(behind.grad.abs().sum()) == 0.0


def test_fused_render_vs_reference_cuda_golden():
    """Fused operator vs what the unmodified reference extension + torch autograd glue produced on a B200 for the
    model-level view (committed fixture tests/golden/refcuda_modelstep_*.npz)."""
    import glob
    import os

    from rasterizer.fused import render_gaussians

    paths = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refcuda_modelstep_*.npz")))
    assert paths, "model-step golden fixture missing"
    for path in paths:
        z = np.load(path)
        s = {k[3:]: (z[k] if z[k].ndim else z[k].item()) for k in z.files if k.startswith("in_")}
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        p = {k[4:]: d(z[k]).requires_grad_(True) for k in z.files if k.startswith("raw_")}
        means = d(s["means3d"]).requires_grad_(True)
        H, W = s["img_height"], s["img_width"]
        rgb, depth, alpha = render_gaussians(means, p["scales_raw"], p["quats_raw"], p["features_dc"], p["features_rest"],
                                             p["opacities_raw"], d(s["viewmat"]), d(s["projmat"]), s["fx"], s["fy"], s["cx"],
                                             s["cy"], H, W, s["degrees_to_use"], background=d(s["background"]),
                                             block_width=s["block_width"])
        assert_float_parity(rgb, z["ref_rgb"], "rgb", max_frac_bad=2e-4)
        assert_float_parity(depth[..., 0], z["ref_depth"], "depth", max_frac_bad=2e-4)
        assert_float_parity(alpha[..., 0], z["ref_alpha"], "alpha", max_frac_bad=2e-4, atol=1e-6)
        torch.autograd.backward([rgb, depth, alpha], [d(z["up_v_rgb"]), d(z["up_v_depth"])[..., None], d(z["up_v_alpha"])[..., None]])
        got = dict(v_means3d=means.grad, v_scales_raw=p["scales_raw"].grad, v_quats_raw=p["quats_raw"].grad,
                   v_opacities_raw=p["opacities_raw"].grad, v_features_dc=p["features_dc"].grad,
                   v_features_rest=p["features_rest"].grad)
        for k, v in got.items():
            assert_float_parity(to_np(v).reshape(z["ref_" + k].shape), z["ref_" + k], k, max_norm_rel=5e-5, max_frac_bad=5e-4)


def test_fused_render_refuses_sh_degree_zero_models():
    """sh_degree = 0 models colour with sigmoid(features_dc) (vanilla_gs.py:808), which the fused operator does not
    implement: it must say so instead of rendering SH-band-0 + 0.5 colours."""
    from rasterizer.fused import render_gaussians

    n = 16
    z = lambda *s: torch.zeros(*s, device="cuda")
    with pytest.raises(ValueError, match="sh_degree = 0"):
        render_gaussians(z(n, 3), z(n, 3), torch.ones(n, 4, device="cuda"), z(n, 3), z(n, 0, 3), z(n, 1),
                         torch.eye(4, device="cuda"), torch.eye(4, device="cuda"), 50.0, 50.0, 16.0, 16.0, 32, 32, 0)

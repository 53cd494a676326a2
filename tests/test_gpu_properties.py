"""Size-independent properties of the hot path checked at the BASELINE sizes (cfg2: 1 M Gaussians @1080p; cfg4:
5 M Gaussians @4K), where the CPU oracle is too slow to be the checker:
  * binning: every tile's list is depth-sorted (ties in index order), bins partition [0, M), kept pairs lie inside
    the reference bounding boxes;
  * compositing: sum of weights == 1 - T (render ones on a black background == alpha), linearity in the colours,
    bitwise determinism of the forward;
  * adjoint: linearity in the upstream gradients, <J v, w> == <v, J^T w> dot-product test through the whole
    operator chain (forward-mode side by finite differences in FP32 on a colour perturbation, which is exact
    because the image is linear in the colours);
  * side by side with the live reference extension when it travelled to the box."""
import numpy as np
import pytest
import torch

from parity import assert_float_parity, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg2():
    from rasterizer.synthetic import make_config_scene, scene_to_torch

    scene = make_config_scene("cfg2")
    return scene, scene_to_torch(scene, "cuda")


def _project(s):
    from rasterizer import cuda as C

    N = s["means3d"].shape[0]
    return C.project_gaussians_forward(N, s["means3d"], s["scales"], s["glob_scale"], s["quats"], s["viewmat"],
                                       s["projmat"], s["fx"], s["fy"], s["cx"], s["cy"], s["img_height"],
                                       s["img_width"], s["block_width"], s["clip_thresh"])


def _blend(s, proj, bins, colors, background):
    from rasterizer import cuda as C

    cov3d, xys, depths, radii, conics, comp, nth = proj
    m, ids, tb_ = bins
    H, W, bw = s["img_height"], s["img_width"], s["block_width"]
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    return C.rasterize_forward(tb, (bw, bw, 1), (W, H, 1), ids, tb_, xys, conics, colors,
                               s["opacities"].reshape(-1, 1).contiguous(), background)


def test_cfg2_binning_invariants(cfg2):
    from rasterizer import cuda as C

    scene, s = cfg2
    proj = _project(s)
    cov3d, xys, depths, radii, conics, comp, nth = proj
    H, W, bw = s["img_height"], s["img_width"], s["block_width"]
    m, ids, bins = C.bin_gaussians_fast(xys, depths, radii, conics, s["opacities"].contiguous(), H, W, bw)
    m_ref = int(nth.sum())
    print(f"[cfg2] visible={int((radii > 0).sum())} M_bbox={m_ref} M_kept={m}")
    assert 0 < m <= m_ref
    # bins partition [0, m): sorted by start, contiguous, empty tiles are (0,0)
    b = bins.long()
    nz = b[:, 1] > b[:, 0]
    starts, ends = b[nz, 0], b[nz, 1]
    assert int(starts[0]) == 0 and int(ends[-1]) == m
    assert bool((starts[1:] == ends[:-1]).all())
    assert bool((b[~nz] == 0).all())
    # per-tile lists are sorted by (depth, Gaussian index)
    tile_of = torch.repeat_interleave(torch.arange(b.shape[0], device="cuda"), (b[:, 1] - b[:, 0]))
    d = depths[ids.long()]
    same_tile = tile_of[1:] == tile_of[:-1]
    ok = (d[1:] > d[:-1]) | ((d[1:] == d[:-1]) & (ids[1:] > ids[:-1]))
    assert bool((ok | ~same_tile).all())
    # every kept pair lies inside the reference bounding box of its Gaussian (helpers.cuh:11-34)
    tiles_x = (W + bw - 1) // bw
    tx, ty = (tile_of % tiles_x).float(), (tile_of // tiles_x).float()
    c, r = xys[ids.long()] / bw, radii[ids.long()].float() / bw
    x0, x1 = (c[:, 0] - r).int().clamp(0, tiles_x).float(), (c[:, 0] + r + 1).int().clamp(0, tiles_x).float()
    tiles_y = (H + bw - 1) // bw
    y0, y1 = (c[:, 1] - r).int().clamp(0, tiles_y).float(), (c[:, 1] + r + 1).int().clamp(0, tiles_y).float()
    assert bool(((tx >= x0) & (tx < x1) & (ty >= y0) & (ty < y1)).all())
    # no (tile, Gaussian) pair twice
    key = tile_of * (1 << 32) + ids.long()
    assert int(torch.unique(key).numel()) == m


def test_cfg2_compositing_properties(cfg2):
    from rasterizer import cuda as C

    scene, s = cfg2
    proj = _project(s)
    cov3d, xys, depths, radii, conics, comp, nth = proj
    H, W, bw = s["img_height"], s["img_width"], s["block_width"]
    bins = C.bin_gaussians_fast(xys, depths, radii, conics, s["opacities"].contiguous(), H, W, bw)
    N = xys.shape[0]
    g = torch.Generator(device="cuda").manual_seed(1)
    c1 = torch.rand(N, 3, device="cuda", generator=g)
    c2 = torch.rand(N, 3, device="cuda", generator=g)
    black = torch.zeros(3, device="cuda")
    i1, T1, idx1 = _blend(s, proj, bins, c1, black)
    i1b, T1b, idx1b = _blend(s, proj, bins, c1, black)
    assert torch.equal(i1, i1b) and torch.equal(T1, T1b) and torch.equal(idx1, idx1b)   # deterministic forward
    i2, _, _ = _blend(s, proj, bins, c2, black)
    i12, _, _ = _blend(s, proj, bins, c1 + c2, black)
    assert_float_parity(i12, i1 + i2, "linearity in colours", atol=2e-6)
    ones, Tn, _ = _blend(s, proj, bins, torch.ones(N, 3, device="cuda"), black)
    assert_float_parity(ones[..., 0], 1 - Tn, "sum of weights == 1 - T", atol=5e-6)
    # background enters as T * bg
    bg = s["background"]
    ibg, Tbg, _ = _blend(s, proj, bins, c1, bg)
    assert_float_parity(ibg, i1 + Tbg[..., None] * bg, "background term", atol=2e-6)
    assert float(Tn.min()) >= 0.0 and float(Tn.max()) <= 1.0
    # reference bounding-box binning renders the very same image (bitwise)
    cum = torch.cumsum(nth, 0, dtype=torch.int32)
    M = int(cum[-1])
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    isect, gids = C.map_gaussian_to_intersects(N, M, xys, depths, radii, cum, tb, bw)
    ks, vs = C.sort_intersects(isect, gids, tb[0] * tb[1])
    rb = C.get_tile_bin_edges(M, ks, tb)
    i1r, T1r, _ = _blend(s, proj, (M, vs, rb), c1, black)
    assert torch.equal(i1r, i1) and torch.equal(T1r, T1)


def test_cfg2_adjoint_properties(cfg2):
    """Linearity of the blend adjoint in the upstream gradients and the dot-product identity
    <d img / d colours [dc], w> == <dc, v_colours(w)> (exact: the image is linear in the colours)."""
    from rasterizer import cuda as C
    from rasterizer.synthetic import make_config_scene, scene_to_torch

    # opacity <= 0.98: above 0.99 the reference's backward clamps alpha at 0.99 while its forward clamps at 0.999
    # (SURVEY 4, quirk 1), so the adjoint it defines is deliberately NOT the exact transpose there
    s = scene_to_torch(make_config_scene("cfg2", opacity_clip=0.98), "cuda")
    proj = _project(s)
    cov3d, xys, depths, radii, conics, comp, nth = proj
    H, W, bw = s["img_height"], s["img_width"], s["block_width"]
    m, ids, tbins = C.bin_gaussians_fast(xys, depths, radii, conics, s["opacities"].contiguous(), H, W, bw)
    N = xys.shape[0]
    g = torch.Generator(device="cuda").manual_seed(2)
    colors = torch.rand(N, 3, device="cuda", generator=g)
    opac = s["opacities"].reshape(-1, 1).contiguous()
    bg = s["background"]
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    img, fT, fi = C.rasterize_forward(tb, (bw, bw, 1), (W, H, 1), ids, tbins, xys, conics, colors, opac, bg)
    w1 = (torch.rand(H, W, 3, device="cuda", generator=g) - 0.5) * 1e-3
    w2 = (torch.rand(H, W, 3, device="cuda", generator=g) - 0.5) * 1e-3
    a1 = (torch.rand(H, W, device="cuda", generator=g) - 0.5) * 1e-3
    zero_a = torch.zeros(H, W, device="cuda")
    bwd = lambda w, a: C.rasterize_backward(H, W, bw, ids, tbins, xys, conics, colors, opac, bg, fT, fi, w, a)
    g1, g2, g12 = bwd(w1, a1), bwd(w2, zero_a), bwd(w1 + w2, a1)
    for name, x1, x2, x12 in zip(("v_xy", "v_conic", "v_colors", "v_opacity"), g1, g2, g12):
        assert_float_parity(x12, x1 + x2, "adjoint linearity " + name, max_norm_rel=2e-5, max_frac_bad=5e-3)
    # coherent (all-positive) probes so that the two inner products are sums of same-signed terms
    dc = torch.rand(N, 3, device="cuda", generator=g)
    wp = torch.rand(H, W, 3, device="cuda", generator=g) * 1e-3
    gp = bwd(wp, zero_a)
    img2, _, _ = C.rasterize_forward(tb, (bw, bw, 1), (W, H, 1), ids, tbins, xys, conics, colors + dc, opac, bg)
    lhs = ((img2 - img).double() * wp.double()).sum()
    rhs = (dc.double() * gp[2].double()).sum()
    rel = abs(float(lhs - rhs)) / max(abs(float(rhs)), 1e-30)
    print(f"[dot-product test] <J dc, w> = {float(lhs):.8e}  <dc, J^T w> = {float(rhs):.8e}  rel = {rel:.2e}")
    assert rel < 1e-4


def _side_by_side_with_reference_extension(s, tag):
    """Full view (forward + all adjoints) through libgsr_b200 (product binning path) and through the UNMODIFIED
    reference CUDA extension with the reference orchestration, same inputs.  Integer outputs: every mismatch is LISTED with
    the values that produced it and at most 2e-6 of the Gaussians may differ (observed on B200: 1 of 1 000 000 at cfg2, 5 of
    5 000 000 at cfg4, each a radius off by one: `radius = ceil(3 sqrt(lambda_max))` (helpers.cuh:55-58) of a cov2d whose
    entries agree with the reference's to 6-7 digits — the 3x3 products of the EWA projection are associated differently
    here than in the reference's glm expressions — lands on the other side of an integer.  The radius only sizes the
    conservative bounding box; with the exact tile culling the image does not depend on it); the image within 1e-4
    on all but 1e-4 of its elements (threshold flips, see parity.py); the seven gradient tensors: see the loop below."""
    from oracle.build_ref import load_ref

    ref_ext = load_ref()
    if ref_ext is None:
        pytest.skip("oracle/_ref/rasterizer_ref_cuda.so not present")
    from pipelines import run_view_bindings
    from rasterizer import cuda as C

    ref = run_view_bindings(ref_ext, s, sort_impl="torch")
    ours = run_view_bindings(C, s, sort_impl="gsr", binning="fast")
    for k in ("radii", "num_tiles_hit"):
        diff = (ours[k] != ref[k]).nonzero().flatten()
        print(f"[{tag} vs reference ext] {k}: {diff.numel()} mismatches of {ours[k].numel()}")
        for g in diff[:8].tolist():  # the offending Gaussians, if any
            print(f"    gaussian {g}: ours {int(ours[k][g])} ref {int(ref[k][g])} xys ours {ours['xys'][g].tolist()} "
                  f"ref {ref['xys'][g].tolist()} conic ours {ours['conics'][g].tolist()} ref {ref['conics'][g].tolist()}")
        assert diff.numel() <= 2e-6 * ours[k].numel(), f"{k}: {diff.numel()} integer outputs differ from the reference extension"
        if diff.numel():
            assert int((ours[k][diff] - ref[k][diff]).abs().max()) <= (1 if k == "radii" else 16)
    bad = ((ours["out_img"] - ref["out_img"]).abs() > 1e-4 * ref["out_img"].abs() + 1e-5).float().mean()
    print(f"[{tag} vs reference ext] image elements outside 1e-4: {float(bad):.2e}; M ours {ours['num_intersects']} "
          f"(exact tile culling) vs reference {ref['num_intersects']}")
    assert float(bad) < 1e-4
    dT = (ours["final_Ts"] - ref["final_Ts"]).abs()
    assert float((dT > 1e-4 * ref["final_Ts"].abs() + 1e-6).float().mean()) < 1e-4
    for k in ("v_coeffs", "v_mean3d", "v_scale", "v_quat", "v_opacity", "v_xy", "v_conic"):
        # whole-tensor norm within the north star's 1e-4 (at cfg4 ONE flipped pixel of v_conic carries 1.0e-4 of a
        # 15 M-element tensor), the non-outlier part within 5e-5, outliers <= 5e-4 of the elements (observed <= 3.2e-5)
        assert_float_parity(to_np(ours[k]).reshape(to_np(ref[k]).shape), ref[k], f"{tag} {k}", max_norm_rel=2e-4,
                            max_norm_rel_trim=5e-5, max_frac_bad=5e-4)


def test_cfg2_vs_live_reference_extension(cfg2):
    scene, s = cfg2
    _side_by_side_with_reference_extension(s, "cfg2")


def test_cfg4_fwd_bwd_vs_live_reference_extension():
    """BASELINE configs[3] (5 M Gaussians, 3840x2160): forward AND backward, all seven gradient tensors, against the
    live reference extension (rasterize.py:185-247 / backward.cu:133-303 at the tile-occupancy-stress size)."""
    from rasterizer.synthetic import make_config_scene, scene_to_torch

    s = scene_to_torch(make_config_scene("cfg4"), "cuda")
    _side_by_side_with_reference_extension(s, "cfg4")
    del s
    torch.cuda.empty_cache()


def test_cfg4_5m_4k_depth_alpha_outputs():
    """BASELINE configs[3]: 5 M Gaussians, 3840x2160, rgb + depth + alpha.  Depth is rendered the way the models do
    (colours = depth repeated, models/vanilla_gs.py:839-855); checks the size-independent properties + records
    occupancy statistics."""
    import rasterizer
    from rasterizer.synthetic import make_config_scene, scene_to_torch

    scene = make_config_scene("cfg4")
    s = scene_to_torch(scene, "cuda")
    H, W, bw = s["img_height"], s["img_width"], s["block_width"]
    xys, depths, radii, conics, comp, nth, cov3d = rasterizer.project_gaussians(
        s["means3d"], s["scales"], 1.0, s["quats"], s["viewmat"], s["projmat"], s["fx"], s["fy"], s["cx"], s["cy"], H, W, bw)
    from rasterizer.sh import spherical_harmonics

    viewdirs = s["means3d"] - s["cam_pos"][None]
    rgbs = torch.clamp(spherical_harmonics(3, viewdirs, s["sh_coeffs"]) + 0.5, min=0.0)
    opac = s["opacities"].reshape(-1, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rgb, alpha = rasterizer.rasterize_gaussians(xys, depths, radii, conics, nth, rgbs, opac, H, W, bw,
                                                background=torch.zeros(3, device="cuda"), return_alpha=True)
    depth_im = rasterizer.rasterize_gaussians(xys, depths, radii, conics, nth, depths[:, None].repeat(1, 3), opac, H, W,
                                              bw, background=torch.zeros(3, device="cuda"))[..., 0:1]
    e1.record()
    torch.cuda.synchronize()
    print(f"[cfg4] visible={int((radii > 0).sum())} M_bbox={int(nth.sum())} rgb+alpha+depth forward {e0.elapsed_time(e1):.2f} ms")
    assert rgb.shape == (H, W, 3) and alpha.shape == (H, W) and depth_im.shape == (H, W, 1)
    assert bool(torch.isfinite(rgb).all()) and float(alpha.min()) >= 0 and float(alpha.max()) <= 1
    # expected depth lies between the nearest and farthest visible Gaussian wherever something was hit
    hit = alpha > 0.5
    dn = (depth_im[..., 0] / alpha.clamp(min=1e-6))[hit]
    vis = radii > 0
    assert float(dn.min()) >= float(depths[vis].min()) * 0.999 and float(dn.max()) <= float(depths[vis].max()) * 1.001
    ones = rasterizer.rasterize_gaussians(xys, depths, radii, conics, nth, torch.ones_like(rgbs), opac, H, W, bw,
                                          background=torch.zeros(3, device="cuda"))
    assert_float_parity(ones[..., 0], alpha, "sum of weights == alpha", atol=5e-6)

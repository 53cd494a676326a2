"""The hand-written binning (per-tile counting + per-tile sorts, device-side M), the hand-written radix sort / scan of
the public utilities, and the asynchronous binning mode."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _projected(scene):
    from rasterizer import cuda as C
    from rasterizer.synthetic import scene_to_torch

    s = scene_to_torch(scene, "cuda")
    N = s["means3d"].shape[0]
    out = C.project_gaussians_forward(N, s["means3d"], s["scales"], 1.0, s["quats"], s["viewmat"], s["projmat"], s["fx"],
                                      s["fy"], s["cx"], s["cy"], s["img_height"], s["img_width"], s["block_width"], 0.01)
    return s, out


@pytest.mark.parametrize("case", ["60k_500x300", "tiny_300_40x24_bw4", "cfg2", "onepass_2k_64x64"])
def test_device_binning_matches_synchronous_binning(case):
    from rasterizer import cuda as C
    from rasterizer.synthetic import make_config_scene, make_scene

    scene = {"60k_500x300": lambda: make_scene(60_000, 500, 300, 0.004, 0.05, margin=1.1, seed=31),
             "tiny_300_40x24_bw4": lambda: make_scene(300, 40, 24, 0.05, 0.2, seed=1, block_width=4),
             "cfg2": lambda: make_config_scene("cfg2"),
             "onepass_2k_64x64": lambda: make_scene(2000, 64, 64, 0.02, 0.2, seed=5)}[case]()  # 16 tiles: one radix pass
    s, (cov3d, xys, depths, radii, conics, comp, nth) = _projected(scene)
    H, W, bw = s["img_height"], s["img_width"], s["block_width"]
    op = s["opacities"].contiguous()
    m, ids, bins = C.bin_gaussians_fast(xys, depths, radii, conics, op, H, W, bw)
    pin = torch.zeros(4, dtype=torch.int32).pin_memory()
    ids_d, bins_d, meta = C.bin_gaussians_device(xys, depths, radii, conics, op, H, W, bw, int(1.3 * m) + 7, meta_pinned=pin)
    torch.cuda.synchronize()
    assert pin.tolist() == [m, 0, m, 0] and meta.tolist() == [m, 0, m, 0]
    assert torch.equal(ids_d[:m], ids) and torch.equal(bins_d, bins)
    # exactly full buffers are not an overflow
    ids_e, bins_e, meta_e = C.bin_gaussians_device(xys, depths, radii, conics, op, H, W, bw, m)
    assert meta_e.tolist() == [m, 0, m, 0] and torch.equal(ids_e[:m], ids) and torch.equal(bins_e, bins)
    # overflow: flagged; tile_bins are clipped to the capacity, the tiles that still fit completely keep their exact lists
    cap = max(1, m // 2)
    ids_o, bins_o, meta_o = C.bin_gaussians_device(xys, depths, radii, conics, op, H, W, bw, cap)
    assert meta_o.tolist() == [m, 1, cap, 0]
    full, part = bins.long(), bins_o.long()
    assert int((part[:, 1] - part[:, 0]).sum()) == cap and int(part.max()) <= cap
    fits = (full[:, 1] <= cap) & (full[:, 1] > full[:, 0])
    assert bool(fits.any()) and torch.equal(part[fits], full[fits])
    last = int(full[fits][:, 1].max())
    assert torch.equal(ids_o[:last], ids[:last])


def test_own_radix_sort_64bit_keys_random_and_ties():
    """gsr_sort_intersects (hand-written LSD radix sort, 6 passes over 45 key bits) == torch.sort(stable)."""
    from rasterizer import cuda as C

    g = torch.Generator(device="cuda").manual_seed(1)
    for m, tiles in ((1, 3), (31, 2), (4096, 100), (4097, 8160), (1_000_003, 32400)):
        tile = torch.randint(0, tiles, (m,), generator=g, device="cuda", dtype=torch.int64)
        depth = torch.randint(0, 2**31 - 1, (m,), generator=g, device="cuda", dtype=torch.int64)
        depth[::3] = int(depth[0])  # many exact ties
        keys = (tile << 32) | depth
        vals = torch.arange(m, device="cuda", dtype=torch.int32)
        ks, vs = C.sort_intersects(keys, vals, tiles)
        ks_t, order = torch.sort(keys, stable=True)
        assert torch.equal(ks, ks_t) and torch.equal(vs.long(), order), (m, tiles)


def test_own_scan_matches_torch_cumsum():
    import rasterizer

    g = torch.Generator(device="cuda").manual_seed(2)
    for n in (1, 255, 4096, 4097, 1_000_001):
        x = torch.randint(0, 50, (n,), generator=g, device="cuda", dtype=torch.int32)
        total, cum = rasterizer.compute_cumulative_intersects(x)
        ref = torch.cumsum(x, 0, dtype=torch.int32)
        assert torch.equal(cum, ref) and total == int(ref[-1])


def test_async_binning_mode_matches_sync_and_reports_overflow():
    import rasterizer
    from rasterizer import binning
    from rasterizer.synthetic import make_scene

    scene = make_scene(30_000, 320, 200, 0.01, 0.08, margin=1.1, seed=11)
    s, (cov3d, xys, depths, radii, conics, comp, nth) = _projected(scene)
    H, W, bw = s["img_height"], s["img_width"], s["block_width"]
    colors = torch.rand(xys.shape[0], 3, device="cuda")
    opac = s["opacities"].reshape(-1, 1).contiguous()

    def render(o):
        return rasterizer.rasterize_gaussians(xys.clone(), depths, radii, conics, nth, colors, o, H, W, bw,
                                              background=s["background"], return_alpha=True)

    binning.reset()
    rasterizer.set_binning_mode("sync")
    img_s, a_s = render(opac)
    rasterizer.set_binning_mode("async")
    try:
        img_1, _ = render(opac)  # first call of the signature: synchronous, learns the capacity
        img_2, a_2 = render(opac)  # asynchronous
        binning.check()
        assert torch.equal(img_1, img_s) and torch.equal(img_2, img_s) and torch.equal(a_2, a_s)
        # a scene with far more pairs than the learned capacity (all opacities -> 1: nothing is culled): overflow
        sig = next(iter(binning._signatures.values()))
        sig.capacity = max(1, sig.max_seen // 4)
        render(opac)
        with pytest.raises(rasterizer.BinningOverflow):
            binning.check()
        assert sig.capacity > sig.max_seen  # raised: the repeated call fits
        img_3, _ = render(opac)
        binning.check()
        assert torch.equal(img_3, img_s)
    finally:
        rasterizer.set_binning_mode("sync")
        binning.reset()


def test_graph_capture_of_public_api_view_matches_eager():
    """rasterizer.graphs.capture_step: forward + backward of a view through the public autograd API, captured once and
    replayed with NEW camera / upstream-gradient contents in the static input tensors, equals the eager result; the
    pair-buffer overflow of a replay is reported by graphs.check()."""
    import rasterizer
    from rasterizer import binning, graphs
    from rasterizer.sh import spherical_harmonics
    from rasterizer.synthetic import look_at_viewmat, make_scene, scene_to_torch

    H, W, bw = 200, 320, 16
    scenes = [scene_to_torch(make_scene(30_000, W, H, 0.01, 0.08, margin=1.1, seed=11, viewmat=look_at_viewmat(yaw_deg=y)), "cuda")
              for y in (0.0, 7.0)]
    s0 = scenes[0]
    params = {k: s0[k].clone().requires_grad_(True) for k in ("means3d", "scales", "quats", "sh_coeffs")}
    opac = s0["opacities"].reshape(-1, 1).clone().requires_grad_(True)
    viewmat, projmat, cam = s0["viewmat"].clone(), s0["projmat"].clone(), s0["cam_pos"].clone()
    v_img, v_alpha = s0["v_out_img"].clone(), s0["v_out_alpha"].clone()
    leaves = list(params.values()) + [opac]

    def view():
        for p in leaves:
            p.grad = None
        xys, depths, radii, conics, comp, nth, cov3d = rasterizer.project_gaussians(
            params["means3d"], params["scales"], 1.0, params["quats"], viewmat, projmat, s0["fx"], s0["fy"], s0["cx"], s0["cy"],
            H, W, bw, 0.01)
        rgbs = torch.clamp(spherical_harmonics(3, params["means3d"].detach() - cam[None], params["sh_coeffs"]) + 0.5, min=0.0)
        img, alpha = rasterizer.rasterize_gaussians(xys, depths, radii, conics, nth, rgbs, opac, H, W, bw,
                                                    background=s0["background"], return_alpha=True)
        torch.autograd.backward([img, alpha], [v_img, v_alpha])
        return img, alpha, [p.grad for p in leaves]

    def load(s, scale):
        viewmat.copy_(s["viewmat"]); projmat.copy_(s["projmat"]); cam.copy_(s["cam_pos"])
        v_img.copy_(s["v_out_img"] * scale); v_alpha.copy_(s["v_out_alpha"] * scale)

    binning.reset()
    rasterizer.set_binning_mode("sync")
    eager = []
    # autograd ties a leaf's AccumulateGrad node to the stream on which the leaf is first used: the eager reference runs on
    # the stream the capture will use, otherwise the captured backward would synchronise with the default stream
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for s, scale in zip(scenes, (1.0, -2.0)):
            load(s, scale)
            img, alpha, grads = view()
            eager.append((img.clone(), alpha.clone(), [g.clone() for g in grads]))
    torch.cuda.synchronize()
    try:
        load(scenes[0], 1.0)
        step = graphs.capture_step(view, stream=side)
        for (s, scale), want in zip(zip(scenes, (1.0, -2.0)), eager):
            load(s, scale)
            img, alpha, grads = step.replay()
            torch.cuda.synchronize()
            graphs.check()
            assert torch.equal(img, want[0]) and torch.equal(alpha, want[1])
            for g, w in zip(grads, want[2]):
                err = float((g - w).norm() / w.norm())
                assert err < 2e-6, err           # atomics order only
        # a replay that needs more pairs than the capacity baked into the graph is flagged
        for slot, key, cap in binning._capture_slots:
            assert int(slot[1]) == 0 and 0 < int(slot[0]) <= cap
        with torch.no_grad():
            saved = opac.detach().clone()
            params["scales"].mul_(4.0)           # 16x the footprint: far more (Gaussian, tile) pairs
        step.replay()
        with pytest.raises(rasterizer.BinningOverflow):
            graphs.check()
        with torch.no_grad():
            params["scales"].div_(4.0)
            opac.copy_(saved)
    finally:
        rasterizer.set_binning_mode("sync")
        binning.reset()
        binning._capture_slots.clear()


def test_long_tiles_take_the_global_radix_path_and_match_key_sort():
    """More than 4096 pairs in one tile (the in-shared-memory sort's limit): the slow path must give the same lists as
    the reference orchestration (64-bit key sort)."""
    from rasterizer import cuda as C
    from rasterizer.synthetic import make_scene

    scene = make_scene(40_000, 64, 48, 0.05, 0.3, margin=0.6, seed=3)  # 12 tiles, ~10 k Gaussians each
    s, (cov3d, xys, depths, radii, conics, comp, nth) = _projected(scene)
    H, W, bw = s["img_height"], s["img_width"], s["block_width"]
    op = s["opacities"].contiguous()
    m, ids, bins = C.bin_gaussians_fast(xys, depths, radii, conics, op, H, W, bw)
    assert int((bins[:, 1] - bins[:, 0]).max()) > 4096
    tiles = C.count_tiles_tight(xys, radii, conics, op, H, W, bw)
    cum = torch.cumsum(tiles, 0, dtype=torch.int32)
    assert int(cum[-1]) == m
    isect, gids = C.map_gaussian_to_intersects_tight(xys.shape[0], m, xys, depths, radii, conics, op, cum, H, W, bw)
    ks, order = torch.sort(isect, stable=True)
    assert torch.equal(ids.long(), gids.long()[order])

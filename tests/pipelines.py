"""TEST INFRASTRUCTURE: one 'view'  through three back-ends with identical orchestration:
  * run_view_bindings(C, ...)  — raw native bindings `C` with the 11 reference names.  `C` is either
    `rasterizer.cuda` (ours, -> libgsr_b200.so C ABI) or the compiled reference extension (oracle/_ref).
    The orchestration between bindings follows the reference wrappers (rasterizer/rasterize.py:92-183,
    utils.py:106-182, sh.py:70-97, project_gaussians.py:83-232) and uses the same ATen ops the reference
    uses (torch.cumsum / torch.sort / torch.gather) so that both back-ends see identical inputs.
  * run_view_public(...)       — our public autograd API (rasterizer.project_gaussians etc.), i.e. what
    gs_toolkit/models/vanilla_gs.py:765-855 calls.
All return dicts of torch tensors with the oracle's key names.
"""
from __future__ import annotations

import torch


def _tile_bounds(W, H, bw):
    return ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)


def run_view_bindings(C, s, backward=True, sort_impl="torch", binning="reference"):
    """s: scene dict of CUDA tensors (rasterizer.synthetic.scene_to_torch).
    binning="tight" (ours only) lists only the (Gaussian, tile) pairs that can reach alpha >= 1/255."""
    H, W, bw = s["img_height"], s["img_width"], s["block_width"]
    N = s["means3d"].shape[0]
    tb = _tile_bounds(W, H, bw)
    degree = {1: 0, 4: 1, 9: 2, 16: 3, 25: 4}[s["sh_coeffs"].shape[1]]
    viewdirs = (s["means3d"] - s["cam_pos"][None, :]).contiguous()
    rgb_sh = C.compute_sh_forward(N, degree, s["degrees_to_use"], viewdirs, s["sh_coeffs"])
    colors = torch.clamp(rgb_sh + 0.5, min=0.0)
    cov3d, xys, depths, radii, conics, comp, nth = C.project_gaussians_forward(
        N, s["means3d"], s["scales"], s["glob_scale"], s["quats"], s["viewmat"], s["projmat"], s["fx"], s["fy"],
        s["cx"], s["cy"], H, W, bw, s["clip_thresh"])
    opac = s["opacities"].reshape(-1, 1).contiguous()
    if binning == "tight":
        tiles = C.count_tiles_tight(xys, radii, conics, opac.reshape(-1), H, W, bw)
        cum = torch.cumsum(tiles, dim=0, dtype=torch.int32)
    else:
        cum = torch.cumsum(nth, dim=0, dtype=torch.int32)
    M = int(cum[-1].item())
    out = dict(rgb_sh=rgb_sh, colors=colors, cov3d=cov3d, xys=xys, depths=depths, radii=radii, conics=conics,
               compensation=comp, num_tiles_hit=nth, cum_tiles_hit=cum, num_intersects=M)
    if binning == "fast":  # the product path of rasterize_gaussians: two-level sort + exact tile culling
        M, vs, bins = C.bin_gaussians_fast(xys, depths, radii, conics, opac.reshape(-1), H, W, bw)
        out["num_intersects"] = M
    if M < 1:
        return out
    if binning == "fast":
        isect = gids = ks = None
    elif binning == "tight":
        isect, gids = C.map_gaussian_to_intersects_tight(N, M, xys, depths, radii, conics, opac.reshape(-1), cum, H, W, bw)
    else:
        isect, gids = C.map_gaussian_to_intersects(N, M, xys, depths, radii, cum, tb, bw)
    if binning == "fast":
        pass
    elif sort_impl == "torch":
        ks, order = torch.sort(isect)
        vs = torch.gather(gids, 0, order)
    else:
        ks, vs = C.sort_intersects(isect, gids, tb[0] * tb[1])
    if binning != "fast":
        bins = C.get_tile_bin_edges(M, ks, tb)
    img, fT, fi = C.rasterize_forward(tb, (bw, bw, 1), (W, H, 1), vs, bins, xys, conics, colors, opac,
                                      s["background"])
    out.update(isect_ids=isect, gaussian_ids=gids, isect_ids_sorted=ks, gaussian_ids_sorted=vs, tile_bins=bins,
               out_img=img, final_Ts=fT, final_idx=fi, out_alpha=1 - fT)
    if not backward:
        return out
    v_xy, v_conic, v_colors, v_opacity = C.rasterize_backward(H, W, bw, vs, bins, xys, conics, colors, opac,
                                                              s["background"], fT, fi, s["v_out_img"],
                                                              s["v_out_alpha"])
    v_rgb_sh = torch.where(rgb_sh + 0.5 > 0, v_colors, torch.zeros_like(v_colors)).contiguous()
    v_coeffs = C.compute_sh_backward(N, degree, s["degrees_to_use"], viewdirs, v_rgb_sh)
    zeros_n = torch.zeros(N, device=xys.device)
    v_cov2d, v_cov3d, v_mean, v_scale, v_quat = C.project_gaussians_backward(
        N, s["means3d"], s["scales"], s["glob_scale"], s["quats"], s["viewmat"], s["projmat"], s["fx"], s["fy"],
        s["cx"], s["cy"], H, W, cov3d, radii, conics, comp, v_xy, zeros_n, v_conic, zeros_n)
    out.update(v_xy=v_xy, v_conic=v_conic, v_colors=v_colors, v_opacity=v_opacity, v_coeffs=v_coeffs,
               v_mean3d=v_mean, v_scale=v_scale, v_quat=v_quat, v_cov2d=v_cov2d, v_cov3d=v_cov3d)
    return out


def run_view_public(s, backward=True):
    """The same view through the public autograd API, the way gs_toolkit/models/vanilla_gs.py drives it."""
    import rasterizer
    from rasterizer.sh import spherical_harmonics

    H, W, bw = s["img_height"], s["img_width"], s["block_width"]
    means = s["means3d"].clone().requires_grad_(True)
    scales = s["scales"].clone().requires_grad_(True)
    quats = s["quats"].clone().requires_grad_(True)
    coeffs = s["sh_coeffs"].clone().requires_grad_(True)
    opac = s["opacities"].reshape(-1, 1).clone().requires_grad_(True)
    xys, depths, radii, conics, comp, nth, cov3d = rasterizer.project_gaussians(
        means, scales, s["glob_scale"], quats, s["viewmat"], s["projmat"], s["fx"], s["fy"], s["cx"], s["cy"], H, W,
        bw, s["clip_thresh"])
    xys.retain_grad()  # models/vanilla_gs.py:797-798
    viewdirs = means.detach() - s["cam_pos"][None, :]
    rgbs = spherical_harmonics(s["degrees_to_use"], viewdirs, coeffs)
    rgbs = torch.clamp(rgbs + 0.5, min=0.0)
    img, alpha = rasterizer.rasterize_gaussians(xys, depths, radii, conics, nth, rgbs, opac, H, W, bw,
                                                background=s["background"], return_alpha=True)
    out = dict(out_img=img, out_alpha=alpha, xys=xys, depths=depths, radii=radii, conics=conics, compensation=comp,
               num_tiles_hit=nth, cov3d=cov3d, colors=rgbs)
    if backward:
        loss = (img * s["v_out_img"]).sum() + (alpha * s["v_out_alpha"]).sum()
        loss.backward()
        out.update(v_mean3d=means.grad, v_scale=scales.grad, v_quat=quats.grad, v_coeffs=coeffs.grad,
                   v_opacity=opac.grad, v_xy=xys.grad)
    return out

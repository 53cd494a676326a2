"""TEST INFRASTRUCTURE — numpy front-end of the CPU oracle (oracle/gsr_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import
this module; the product package (gaussian-splatting-toolkit_b200/) never does.

Each function mirrors one reference operator; see gsr_oracle.c for the file:line each one restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgsr_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
        os.path.join(_HERE, "gsr_oracle.c")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_cumsum.restype = C.c_int64
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(n))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def num_sh_bases(degree: int) -> int:
    return int(lib().orc_num_sh_bases(C.c_int(degree)))


def deg_from_sh(num_bases: int) -> int:
    return {1: 0, 4: 1, 9: 2, 16: 3, 25: 4}[num_bases]


# ---------------------------------------------------------------------------------------------------
def sh_forward(degrees_to_use, viewdirs, coeffs):
    viewdirs, coeffs = _f32(viewdirs), _f32(coeffs)
    n, K = coeffs.shape[0], coeffs.shape[1]
    colors = np.empty((n, 3), np.float32)
    lib().orc_sh_forward(C.c_int(n), C.c_int(deg_from_sh(K)), C.c_int(degrees_to_use), _p(viewdirs),
                         _p(coeffs), _p(colors))
    return colors


def sh_backward(degree, degrees_to_use, viewdirs, v_colors):
    viewdirs, v_colors = _f32(viewdirs), _f32(v_colors)
    n = v_colors.shape[0]
    K = num_sh_bases(degree)
    v_coeffs = np.empty((n, K, 3), np.float32)
    lib().orc_sh_backward(C.c_int(n), C.c_int(degree), C.c_int(degrees_to_use), _p(viewdirs),
                          _p(v_colors), _p(v_coeffs))
    return v_coeffs


def compute_cov2d_bounds(cov2d):
    cov2d = _f32(cov2d)
    n = cov2d.shape[0]
    conics = np.empty((n, 3), np.float32)
    radii = np.empty((n, 1), np.float32)
    lib().orc_compute_cov2d_bounds(C.c_int(n), _p(cov2d), _p(conics), _p(radii))
    return conics, radii


def project_forward(means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_height,
                    img_width, block_width, clip_thresh=0.01):
    """-> (cov3d, xys, depths, radii, conics, compensation, num_tiles_hit) — the binding's order."""
    means3d, scales, quats = _f32(means3d), _f32(scales), _f32(quats)
    viewmat, projmat = _f32(viewmat), _f32(projmat)
    n = means3d.shape[0]
    cov3d = np.empty((n, 6), np.float32)
    xys = np.empty((n, 2), np.float32)
    depths = np.empty((n,), np.float32)
    radii = np.empty((n,), np.int32)
    conics = np.empty((n, 3), np.float32)
    comp = np.empty((n,), np.float32)
    nth = np.empty((n,), np.int32)
    lib().orc_project_forward(C.c_int(n), _p(means3d), _p(scales), C.c_float(glob_scale), _p(quats),
                              _p(viewmat), _p(projmat), C.c_float(fx), C.c_float(fy), C.c_float(cx),
                              C.c_float(cy), C.c_int(img_height), C.c_int(img_width),
                              C.c_int(block_width), C.c_float(clip_thresh), _p(cov3d), _p(xys),
                              _p(depths), _p(radii), _p(conics), _p(comp), _p(nth))
    return cov3d, xys, depths, radii, conics, comp, nth


def project_backward(means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_height,
                     img_width, cov3d, radii, conics, compensation, v_xy, v_depth, v_conic,
                     v_compensation):
    """-> (v_cov2d, v_cov3d, v_mean3d, v_scale, v_quat) — the binding's order."""
    means3d, scales, quats = _f32(means3d), _f32(scales), _f32(quats)
    viewmat, projmat = _f32(viewmat), _f32(projmat)
    cov3d, radii, conics, compensation = _f32(cov3d), _i32(radii), _f32(conics), _f32(compensation)
    v_xy, v_depth, v_conic, v_compensation = _f32(v_xy), _f32(v_depth), _f32(v_conic), _f32(v_compensation)
    n = means3d.shape[0]
    v_cov2d = np.empty((n, 3), np.float32)
    v_cov3d = np.empty((n, 6), np.float32)
    v_mean = np.empty((n, 3), np.float32)
    v_scale = np.empty((n, 3), np.float32)
    v_quat = np.empty((n, 4), np.float32)
    lib().orc_project_backward(C.c_int(n), _p(means3d), _p(scales), C.c_float(glob_scale), _p(quats),
                               _p(viewmat), _p(projmat), C.c_float(fx), C.c_float(fy), C.c_float(cx),
                               C.c_float(cy), C.c_int(img_height), C.c_int(img_width), _p(cov3d),
                               _p(radii), _p(conics), _p(compensation), _p(v_xy), _p(v_depth),
                               _p(v_conic), _p(v_compensation), _p(v_cov2d), _p(v_cov3d), _p(v_mean),
                               _p(v_scale), _p(v_quat))
    return v_cov2d, v_cov3d, v_mean, v_scale, v_quat


# ---------------------------------------------------------------------------------------------------
def compute_cumulative_intersects(num_tiles_hit):
    nth = _i32(num_tiles_hit)
    cum = np.empty_like(nth)
    m = lib().orc_cumsum(C.c_int(nth.shape[0]), _p(nth), _p(cum))
    return int(m), cum


def map_gaussian_to_intersects(num_points, num_intersects, xys, depths, radii, cum_tiles_hit,
                               tile_bounds, block_width):
    xys, depths, radii, cum = _f32(xys), _f32(depths), _i32(radii), _i32(cum_tiles_hit)
    isect = np.zeros((num_intersects,), np.int64)
    gids = np.zeros((num_intersects,), np.int32)
    lib().orc_map_intersects(C.c_int(num_points), _p(xys), _p(depths), _p(radii), _p(cum),
                             C.c_int(tile_bounds[0]), C.c_int(tile_bounds[1]), C.c_int(block_width),
                             _p(isect), _p(gids))
    return isect, gids


def sort_intersects(isect_ids, gaussian_ids):
    isect_ids = np.ascontiguousarray(isect_ids, np.int64)
    gaussian_ids = _i32(gaussian_ids)
    m = isect_ids.shape[0]
    ks = np.empty_like(isect_ids)
    vs = np.empty_like(gaussian_ids)
    lib().orc_sort_intersects(C.c_int64(m), _p(isect_ids), _p(gaussian_ids), _p(ks), _p(vs))
    return ks, vs


def get_tile_bin_edges(num_intersects, isect_ids_sorted, tile_bounds):
    ks = np.ascontiguousarray(isect_ids_sorted, np.int64)
    nt = tile_bounds[0] * tile_bounds[1]
    bins = np.empty((nt, 2), np.int32)
    lib().orc_tile_bin_edges(C.c_int64(num_intersects), _p(ks), C.c_int(nt), _p(bins))
    return bins


def bin_and_sort_gaussians(num_points, num_intersects, xys, depths, radii, cum_tiles_hit, tile_bounds,
                           block_width):
    isect, gids = map_gaussian_to_intersects(num_points, num_intersects, xys, depths, radii,
                                             cum_tiles_hit, tile_bounds, block_width)
    ks, vs = sort_intersects(isect, gids)
    bins = get_tile_bin_edges(num_intersects, ks, tile_bounds)
    return isect, gids, ks, vs, bins


# ---------------------------------------------------------------------------------------------------
def rasterize_forward(img_height, img_width, block_width, gaussian_ids_sorted, tile_bins, xys, conics,
                      colors, opacities, background, nd_numerics=None, want_ambiguous=False, tile_rows=(0, -1)):
    """-> (out_img [H,W,C], final_Ts [H,W], final_idx [H,W][, ambiguous [H,W] u8]).

    nd_numerics: None -> like the reference dispatch (binary16 accumulators iff C != 3).
    """
    gids, bins = _i32(gaussian_ids_sorted), _i32(tile_bins)
    xys, conics, colors = _f32(xys), _f32(conics), _f32(colors)
    opac, bg = _f32(opacities).reshape(-1), _f32(background)
    ch = colors.shape[1]
    half = (ch != 3) if nd_numerics is None else bool(nd_numerics)
    out = np.zeros((img_height, img_width, ch), np.float32)
    fT = np.zeros((img_height, img_width), np.float32)
    fi = np.zeros((img_height, img_width), np.int32)
    amb = np.zeros((img_height, img_width), np.uint8)
    lib().orc_rasterize_forward(C.c_int(img_height), C.c_int(img_width), C.c_int(block_width),
                                C.c_int(ch), C.c_int(int(half)), _p(gids), _p(bins), _p(xys),
                                _p(conics), _p(colors), _p(opac), _p(bg), _p(out), _p(fT), _p(fi),
                                _p(amb), C.c_int(tile_rows[0]), C.c_int(tile_rows[1]))
    if want_ambiguous:
        return out, fT, fi, amb
    return out, fT, fi


def rasterize_backward(img_height, img_width, block_width, gaussian_ids_sorted, tile_bins, xys, conics,
                       colors, opacities, background, final_Ts, final_idx, v_output, v_output_alpha,
                       nd_numerics=None, dtype=np.float64, tile_rows=(0, -1)):
    """-> (v_xy [N,2], v_conic [N,3], v_colors [N,C], v_opacity [N,1]) accumulated in binary64."""
    gids, bins = _i32(gaussian_ids_sorted), _i32(tile_bins)
    xys, conics, colors = _f32(xys), _f32(conics), _f32(colors)
    opac, bg = _f32(opacities).reshape(-1), _f32(background)
    fT, fi = _f32(final_Ts), _i32(final_idx)
    vo, voa = _f32(v_output), _f32(v_output_alpha)
    n, ch = xys.shape[0], colors.shape[1]
    nd = (ch != 3) if nd_numerics is None else bool(nd_numerics)
    v_xy = np.empty((n, 2), np.float64)
    v_conic = np.empty((n, 3), np.float64)
    v_col = np.empty((n, ch), np.float64)
    v_op = np.empty((n, 1), np.float64)
    lib().orc_rasterize_backward(C.c_int(img_height), C.c_int(img_width), C.c_int(block_width),
                                 C.c_int(ch), C.c_int(int(nd)), C.c_int(n), _p(gids), _p(bins),
                                 _p(xys), _p(conics), _p(colors), _p(opac), _p(bg), _p(fT), _p(fi),
                                 _p(vo), _p(voa), _p(v_xy), _p(v_conic), _p(v_col), _p(v_op),
                                 C.c_int(tile_rows[0]), C.c_int(tile_rows[1]))
    return tuple(a.astype(dtype) for a in (v_xy, v_conic, v_col, v_op))


# ---------------------------------------------------------------------------------------------------
def render_view(scene, v_out_img=None, v_out_alpha=None, backward=True, tile_rows=(0, -1)):
    """One whole 'view' of SURVEY §8(d) on the CPU: SH fwd -> project fwd -> bin/sort -> blend fwd
    [-> blend bwd -> SH bwd -> project bwd].  `scene` is the dict made by rasterizer.synthetic.make_scene
    (numpy arrays).  Returns a dict of every intermediate and output.  The model-side glue between the
    operators (viewdirs, clamp(rgb+0.5, 0), its derivative) follows models/vanilla_gs.py:799-807."""
    s = scene
    H, W, bw = s["img_height"], s["img_width"], s["block_width"]
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    n = s["means3d"].shape[0]
    out = {}
    viewdirs = s["means3d"] - s["cam_pos"][None, :]
    rgb_sh = sh_forward(s["degrees_to_use"], viewdirs, s["sh_coeffs"])
    colors = np.maximum(rgb_sh + 0.5, 0.0).astype(np.float32)
    cov3d, xys, depths, radii, conics, comp, nth = project_forward(
        s["means3d"], s["scales"], s["glob_scale"], s["quats"], s["viewmat"], s["projmat"], s["fx"],
        s["fy"], s["cx"], s["cy"], H, W, bw, s["clip_thresh"])
    m, cum = compute_cumulative_intersects(nth)
    isect, gids, ks, vs, bins = bin_and_sort_gaussians(n, m, xys, depths, radii, cum, tb, bw)
    img, fT, fi, amb = rasterize_forward(H, W, bw, vs, bins, xys, conics, colors, s["opacities"],
                                         s["background"], want_ambiguous=True, tile_rows=tile_rows)
    out.update(colors=colors, rgb_sh=rgb_sh, cov3d=cov3d, xys=xys, depths=depths, radii=radii,
               conics=conics, compensation=comp, num_tiles_hit=nth, num_intersects=m, cum_tiles_hit=cum,
               isect_ids=isect, gaussian_ids=gids, isect_ids_sorted=ks, gaussian_ids_sorted=vs,
               tile_bins=bins, out_img=img, final_Ts=fT, final_idx=fi, out_alpha=1.0 - fT,
               ambiguous=amb)
    if not backward:
        return out
    v_xy, v_conic, v_colors, v_opacity = rasterize_backward(
        H, W, bw, vs, bins, xys, conics, colors, s["opacities"], s["background"], fT, fi, v_out_img,
        v_out_alpha, dtype=np.float32, tile_rows=tile_rows)
    v_rgb_sh = np.where(rgb_sh + 0.5 > 0.0, v_colors, 0.0).astype(np.float32)
    v_coeffs = sh_backward(deg_from_sh(s["sh_coeffs"].shape[1]), s["degrees_to_use"], viewdirs, v_rgb_sh)
    zeros_n = np.zeros((n,), np.float32)
    v_cov2d, v_cov3d, v_mean, v_scale, v_quat = project_backward(
        s["means3d"], s["scales"], s["glob_scale"], s["quats"], s["viewmat"], s["projmat"], s["fx"],
        s["fy"], s["cx"], s["cy"], H, W, cov3d, radii, conics, comp, v_xy, zeros_n, v_conic, zeros_n)
    out.update(v_xy=v_xy, v_conic=v_conic, v_colors=v_colors, v_opacity=v_opacity, v_coeffs=v_coeffs,
               v_mean3d=v_mean, v_scale=v_scale, v_quat=v_quat, v_cov2d=v_cov2d, v_cov3d=v_cov3d)
    return out


# ---------------------------------------------------------------------------------------------------
def raw_parameters(scene, seed=0):
    """The model's RAW parameters (log-scales, unnormalised quaternions, logit opacities, features_dc / features_rest,
    gs_toolkit/models/vanilla_gs.py:150-175) that activate to the tensors of a synthetic scene."""
    rng = np.random.default_rng(seed)
    n = scene["means3d"].shape[0]
    o = np.clip(scene["opacities"].astype(np.float64), 1e-6, 1 - 1e-6)
    return dict(
        scales_raw=np.log(scene["scales"]).astype(np.float32),
        quats_raw=(scene["quats"] * rng.uniform(0.5, 2.0, size=(n, 1))).astype(np.float32),
        opacities_raw=np.log(o / (1 - o)).astype(np.float32).reshape(n, 1),
        features_dc=np.ascontiguousarray(scene["sh_coeffs"][:, 0, :]),
        features_rest=np.ascontiguousarray(scene["sh_coeffs"][:, 1:, :]),
    )


def render_fused_reference(scene, raw, v_rgb=None, v_depth=None, v_alpha=None):
    """What the reference model computes per view from its raw parameters (models/vanilla_gs.py:759-855: activations,
    SH colour + clamp, projection, colour rasterization with alpha, depth rasterization), restated with the oracle's
    operators and numpy glue, and the gradients of the six raw parameter tensors for upstream gradients of
    (rgb [H,W,3], depth [H,W], alpha [H,W]).  Depth is composited as a 4th FP32 channel — per pixel the same
    arithmetic as the model's second rasterize_gaussians call."""
    s = scene
    H, W, bw = s["img_height"], s["img_width"], s["block_width"]
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    n = s["means3d"].shape[0]
    scales = np.exp(raw["scales_raw"]).astype(np.float32)
    qnorm = np.linalg.norm(raw["quats_raw"].astype(np.float64), axis=1, keepdims=True)
    qn = (raw["quats_raw"] / qnorm).astype(np.float32)
    opac = (1.0 / (1.0 + np.exp(-raw["opacities_raw"].astype(np.float64)))).astype(np.float32).reshape(-1)
    coeffs = np.concatenate([raw["features_dc"][:, None, :], raw["features_rest"]], axis=1).astype(np.float32)
    viewdirs = s["means3d"] - s["cam_pos"][None, :]
    rgb_sh = sh_forward(s["degrees_to_use"], viewdirs, coeffs)
    colors = np.maximum(rgb_sh + 0.5, 0.0).astype(np.float32)
    cov3d, xys, depths, radii, conics, comp, nth = project_forward(
        s["means3d"], scales, s["glob_scale"], qn, s["viewmat"], s["projmat"], s["fx"], s["fy"], s["cx"], s["cy"], H, W,
        bw, s["clip_thresh"])
    m, cum = compute_cumulative_intersects(nth)
    _, _, _, vs, bins = bin_and_sort_gaussians(n, m, xys, depths, radii, cum, tb, bw)
    colors4 = np.concatenate([colors, depths[:, None]], axis=1).astype(np.float32)
    bg4 = np.concatenate([s["background"], np.zeros(1, np.float32)]).astype(np.float32)
    img4, fT, fi, amb = rasterize_forward(H, W, bw, vs, bins, xys, conics, colors4, opac, bg4, nd_numerics=False,
                                          want_ambiguous=True)
    out = dict(rgb=img4[..., :3], depth=img4[..., 3], alpha=1.0 - fT, ambiguous=amb, radii=radii, xys=xys,
               num_intersects=m)
    if v_rgb is None:
        return out
    v_out4 = np.concatenate([v_rgb, v_depth[..., None]], axis=-1).astype(np.float32)
    v_xy, v_conic, v_colors4, v_opacity = rasterize_backward(
        H, W, bw, vs, bins, xys, conics, colors4, opac, bg4, fT, fi, v_out4, v_alpha, nd_numerics=False, dtype=np.float32)
    v_rgb_sh = np.where(rgb_sh + 0.5 > 0.0, v_colors4[:, :3], 0.0).astype(np.float32)
    v_coeffs = sh_backward(deg_from_sh(coeffs.shape[1]), s["degrees_to_use"], viewdirs, v_rgb_sh)
    zeros_n = np.zeros((n,), np.float32)
    _, _, v_mean, v_scale, v_quat = project_backward(
        s["means3d"], scales, s["glob_scale"], qn, s["viewmat"], s["projmat"], s["fx"], s["fy"], s["cx"], s["cy"], H, W,
        cov3d, radii, conics, comp, v_xy, np.ascontiguousarray(v_colors4[:, 3]), v_conic, zeros_n)
    # chain rules of the activations (autograd of torch.exp / quats / quats.norm / torch.sigmoid in the model)
    vq = v_quat.astype(np.float64)
    qh = qn.astype(np.float64)
    v_quats_raw = (vq - qh * np.sum(qh * vq, axis=1, keepdims=True)) / qnorm
    o = opac.astype(np.float64).reshape(-1, 1)
    out.update(v_means3d=v_mean, v_scales_raw=(v_scale * scales).astype(np.float32),
               v_quats_raw=v_quats_raw.astype(np.float32),
               v_opacities_raw=(v_opacity.astype(np.float64) * o * (1 - o)).astype(np.float32),
               v_features_dc=np.ascontiguousarray(v_coeffs[:, 0, :]), v_features_rest=np.ascontiguousarray(v_coeffs[:, 1:, :]),
               v_xy=v_xy)
    return out

"""`rasterizer.cuda` — the native-binding surface of the reference package, backed by libgsr_b200.so.

The reference module (rasterizer/cuda/__init__.py:4-26) exposes lazy trampolines to the 11 pybind
functions of csrc/ext.cpp:6-17.  The same 11 names with the same positional arguments and return
tuples are provided here; each allocates its outputs with torch (like bindings.cu does with
torch::zeros / torch::empty, but uninitialised: the kernels write every element) and calls the C ABI of
include/gsr_b200.h with raw device pointers on torch's current stream.

Argument checking mirrors bindings.h:10-15 (CHECK_CUDA / CHECK_CONTIGUOUS -> RuntimeError) and the
AT_ERROR shape checks of bindings.cu:65-68,86-93,420-426,490-496.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from .. import _lib

_P = _lib.C.c_void_p


def _ptr(t: Tensor):
    return _P(t.data_ptr())


def _check_input(x: Tensor, name: str, dtype=None) -> None:
    if not isinstance(x, Tensor) or not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not x.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and x.dtype != dtype:
        raise RuntimeError(f"{name}: expected scalar type {dtype} but found {x.dtype}")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
if _raw_stream is None:  # older torch: fall back to the public accessor
    def _raw_stream(dev):  # noqa: E306
        return torch.cuda.current_stream(dev).cuda_stream


class _Guard:
    """DEVICE_GUARD(tensor) of bindings.h:16-17 + the stream to launch on (the reference uses the legacy
    default stream; we use torch's current stream of the tensor's device)."""

    __slots__ = ("dev", "prev", "stream")

    def __init__(self, t: Tensor):
        self.dev = t.device.index if t.device.index is not None else torch.cuda.current_device()

    def __enter__(self):
        self.prev = torch.cuda.current_device()
        if self.prev != self.dev:
            torch.cuda.set_device(self.dev)
        # the raw handle of torch's current stream (torch.cuda.current_stream() builds a Stream object per call: ~6 us,
        # seven times per view on the public path)
        self.stream = _P(_raw_stream(self.dev))
        return self.stream

    def __exit__(self, *exc):
        if self.prev != self.dev:
            torch.cuda.set_device(self.prev)
        return False


def num_sh_bases(degree: int) -> int:
    return {0: 1, 1: 4, 2: 9, 3: 16}.get(degree, 25)


# ---------------------------------------------------------------------------------------------------
def compute_sh_forward(num_points: int, degree: int, degrees_to_use: int, viewdirs: Tensor, coeffs: Tensor) -> Tensor:
    _check_input(viewdirs, "viewdirs", torch.float32)
    _check_input(coeffs, "coeffs", torch.float32)
    nb = num_sh_bases(degree)
    if coeffs.dim() != 3 or coeffs.size(0) != num_points or coeffs.size(1) != nb or coeffs.size(2) != 3:
        raise RuntimeError("coeffs must have dimensions (N, D, 3)")
    colors = torch.empty((num_points, 3), dtype=torch.float32, device=coeffs.device)
    with _Guard(coeffs) as st:
        _lib.check(_lib.load().gsr_compute_sh_forward(num_points, degree, degrees_to_use, _ptr(viewdirs),
                                                      _ptr(coeffs), _ptr(colors), st), "compute_sh_forward")
    return colors


def _out_or_empty(out, shape, dev, name):
    """Optional caller-provided output (e.g. a segment of view_parallel.GradientBucket): must be a contiguous
    float32 CUDA tensor with the right number of elements."""
    if out is None:
        return torch.empty(shape, dtype=torch.float32, device=dev)
    _check_input(out, name, torch.float32)
    n = 1
    for d in shape:
        n *= d
    if out.numel() != n or out.device != dev:
        raise RuntimeError(f"{name}: expected {n} float32 elements on {dev}")
    return out.view(shape)


def compute_sh_backward(num_points: int, degree: int, degrees_to_use: int, viewdirs: Tensor, v_colors: Tensor, *,
                        out: Tensor = None) -> Tensor:
    _check_input(viewdirs, "viewdirs", torch.float32)
    _check_input(v_colors, "v_colors", torch.float32)
    if viewdirs.dim() != 2 or viewdirs.size(0) != num_points or viewdirs.size(1) != 3:
        raise RuntimeError("viewdirs must have dimensions (N, 3)")
    if v_colors.dim() != 2 or v_colors.size(0) != num_points or v_colors.size(1) != 3:
        raise RuntimeError("v_colors must have dimensions (N, 3)")
    nb = num_sh_bases(degree)
    v_coeffs = _out_or_empty(out, (num_points, nb, 3), v_colors.device, "out")
    with _Guard(v_colors) as st:
        _lib.check(_lib.load().gsr_compute_sh_backward(num_points, degree, degrees_to_use, _ptr(viewdirs),
                                                       _ptr(v_colors), _ptr(v_coeffs), st), "compute_sh_backward")
    return v_coeffs


def compute_sh_backward_multiview(degree: int, degrees_to_use: int, means3d: Tensor, cam_positions,
                                  v_colors_views, *, out: Tensor = None) -> Tensor:
    """v_coeffs [N,K,3] = sum over views v of Y(means3d - cam_positions[v]) (x) v_colors_views[v]  (each [N,3]).
    `cam_positions`: CUDA tensor [V,3] or a sequence of V CUDA tensors of 3 floats; `v_colors_views`: sequence of V
    CUDA tensors (or one [V,N,3] tensor).  Per-view tensors may alias peer-GPU memory (symmetric memory)."""
    import ctypes

    _check_input(means3d, "means3d", torch.float32)
    views = list(v_colors_views.unbind(0)) if isinstance(v_colors_views, Tensor) else list(v_colors_views)
    cams = list(cam_positions.reshape(-1, 3).unbind(0)) if isinstance(cam_positions, Tensor) else list(cam_positions)
    n, V = means3d.size(0), len(views)
    if len(cams) != V:
        raise RuntimeError("cam_positions and v_colors_views must have the same number of views")
    for i, c in enumerate(cams):
        if not c.is_cuda or c.dtype != torch.float32 or c.numel() != 3 or not c.is_contiguous():
            raise RuntimeError(f"cam_positions[{i}] must be 3 contiguous float32 values on the GPU")
    for i, v in enumerate(views):
        _check_input(v, f"v_colors_views[{i}]", torch.float32)
        if v.numel() != 3 * n:
            raise RuntimeError("every view's colour gradient must have N*3 elements")
    nb = num_sh_bases(degree)
    v_coeffs = _out_or_empty(out, (n, nb, 3), means3d.device, "out")
    ptrs = (ctypes.c_void_p * V)(*[v.data_ptr() for v in views])
    cam_ptrs = (ctypes.c_void_p * V)(*[c.data_ptr() for c in cams])
    with _Guard(means3d) as st:
        _lib.check(_lib.load().gsr_compute_sh_backward_multiview_ptrs(n, degree, degrees_to_use, V, _ptr(means3d),
                                                                      cam_ptrs, ptrs,
                                                                 _ptr(v_coeffs), st), "compute_sh_backward_multiview")
    return v_coeffs


def peer_all_reduce_phase(phase: str, rank: int, bufs, num_floats: int) -> None:
    """One kernel of the two-kernel all-reduce over NVLink peer memory (gsr_peer_reduce_scatter / gsr_peer_all_gather).
    `bufs`: one CUDA float32 tensor per rank — rank w's buffer as mapped into this process (symmetric memory) — each with at
    least `num_floats` elements.  The caller synchronises the ranks (device-side barriers) around the two phases."""
    import ctypes

    bufs = list(bufs)
    for i, b in enumerate(bufs):
        _check_input(b, f"bufs[{i}]", torch.float32)
        if b.numel() < num_floats:
            raise RuntimeError("peer_all_reduce_phase: buffer smaller than num_floats")
    fn = {"reduce_scatter": "gsr_peer_reduce_scatter", "all_gather": "gsr_peer_all_gather"}[phase]
    ptrs = (ctypes.c_void_p * len(bufs))(*[b.data_ptr() for b in bufs])
    with _Guard(bufs[rank]) as st:
        _lib.check(getattr(_lib.load(), fn)(len(bufs), int(rank), ptrs, int(num_floats), st), fn)


def peer_push(dsts, src: Tensor, src_stride: int, n_per_dst: int, n_total: int = 0) -> None:
    """gsr_peer_push: dsts[w][0 .. n_w) <- src[w * src_stride ..) for every rank w (remote stores into symmetric memory).
    src_stride = 0 broadcasts `n_per_dst` floats of `src` to every destination."""
    import ctypes

    dsts = list(dsts)
    _check_input(src, "src", torch.float32)
    for i, d in enumerate(dsts):
        _check_input(d, f"dsts[{i}]", torch.float32)
    ptrs = (ctypes.c_void_p * len(dsts))(*[d.data_ptr() for d in dsts])
    with _Guard(src) as st:
        _lib.check(_lib.load().gsr_peer_push(len(dsts), ptrs, _ptr(src), int(src_stride), int(n_per_dst), int(n_total), st),
                   "peer_push")


def peer_reduce_broadcast(dsts, slots: Tensor, slot_stride: int, num_floats: int) -> None:
    """gsr_peer_reduce_broadcast: every dsts[w][0 .. num_floats) <- sum over the len(dsts) local slots of `slots`."""
    import ctypes

    dsts = list(dsts)
    _check_input(slots, "slots", torch.float32)
    ptrs = (ctypes.c_void_p * len(dsts))(*[d.data_ptr() for d in dsts])
    with _Guard(slots) as st:
        _lib.check(_lib.load().gsr_peer_reduce_broadcast(len(dsts), ptrs, _ptr(slots), int(slot_stride), int(num_floats), st),
                   "peer_reduce_broadcast")


def compute_cov2d_bounds(num_pts: int, covs2d: Tensor) -> Tuple[Tensor, Tensor]:
    _check_input(covs2d, "covs2d", torch.float32)
    conics = torch.empty((num_pts, covs2d.size(1)), dtype=torch.float32, device=covs2d.device)
    radii = torch.empty((num_pts, 1), dtype=torch.float32, device=covs2d.device)
    with _Guard(covs2d) as st:
        _lib.check(_lib.load().gsr_compute_cov2d_bounds(num_pts, _ptr(covs2d), _ptr(conics), _ptr(radii), st),
                   "compute_cov2d_bounds")
    return conics, radii


# ---------------------------------------------------------------------------------------------------
def project_gaussians_forward(num_points: int, means3d: Tensor, scales: Tensor, glob_scale: float, quats: Tensor,
                              viewmat: Tensor, projmat: Tensor, fx: float, fy: float, cx: float, cy: float,
                              img_height: int, img_width: int, block_width: int, clip_thresh: float):
    for t, n in ((means3d, "means3d"), (scales, "scales"), (quats, "quats"), (viewmat, "viewmat"),
                 (projmat, "projmat")):
        _check_input(t, n, torch.float32)
    if viewmat.numel() < 12 or projmat.numel() < 16:
        raise RuntimeError("viewmat needs >= 12 and projmat 16 elements")
    dev = means3d.device
    f32 = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    cov3d = torch.empty((num_points, 6), **f32)
    xys = torch.empty((num_points, 2), **f32)
    depths = torch.empty((num_points,), **f32)
    radii = torch.empty((num_points,), **i32)
    conics = torch.empty((num_points, 3), **f32)
    compensation = torch.empty((num_points,), **f32)
    num_tiles_hit = torch.empty((num_points,), **i32)
    with _Guard(means3d) as st:
        _lib.check(_lib.load().gsr_project_gaussians_forward(
            num_points, _ptr(means3d), _ptr(scales), float(glob_scale), _ptr(quats), _ptr(viewmat), _ptr(projmat),
            float(fx), float(fy), float(cx), float(cy), int(img_height), int(img_width), int(block_width),
            float(clip_thresh), _ptr(cov3d), _ptr(xys), _ptr(depths), _ptr(radii), _ptr(conics),
            _ptr(compensation), _ptr(num_tiles_hit), st), "project_gaussians_forward")
    return cov3d, xys, depths, radii, conics, compensation, num_tiles_hit


def project_gaussians_backward(num_points: int, means3d: Tensor, scales: Tensor, glob_scale: float, quats: Tensor,
                               viewmat: Tensor, projmat: Tensor, fx: float, fy: float, cx: float, cy: float,
                               img_height: int, img_width: int, cov3d: Tensor, radii: Tensor, conics: Tensor,
                               compensation: Tensor, v_xy: Tensor, v_depth: Tensor, v_conic: Tensor,
                               v_compensation: Tensor, *, out_mean3d: Tensor = None, out_scale: Tensor = None,
                               out_quat: Tensor = None, need_cov_grads: bool = True):
    """`need_cov_grads=False` skips v_cov2d / v_cov3d (the reference binding returns them, its Python wrapper discards
    them, rasterizer/project_gaussians.py:178,203-232): the first two return values are then None."""
    for t, n in ((means3d, "means3d"), (scales, "scales"), (quats, "quats"), (viewmat, "viewmat"),
                 (projmat, "projmat"), (cov3d, "cov3d"), (conics, "conics"), (compensation, "compensation"),
                 (v_xy, "v_xy"), (v_depth, "v_depth"), (v_conic, "v_conic"), (v_compensation, "v_compensation")):
        _check_input(t, n, torch.float32)
    _check_input(radii, "radii", torch.int32)
    dev = means3d.device
    f32 = dict(dtype=torch.float32, device=dev)
    v_cov2d = torch.empty((num_points, 3), **f32) if need_cov_grads else None
    v_cov3d = torch.empty((num_points, 6), **f32) if need_cov_grads else None
    v_mean3d = _out_or_empty(out_mean3d, (num_points, 3), dev, "out_mean3d")
    v_scale = _out_or_empty(out_scale, (num_points, 3), dev, "out_scale")
    v_quat = _out_or_empty(out_quat, (num_points, 4), dev, "out_quat")
    with _Guard(means3d) as st:
        _lib.check(_lib.load().gsr_project_gaussians_backward(
            num_points, _ptr(means3d), _ptr(scales), float(glob_scale), _ptr(quats), _ptr(viewmat), _ptr(projmat),
            float(fx), float(fy), float(cx), float(cy), int(img_height), int(img_width), _ptr(cov3d), _ptr(radii),
            _ptr(conics), _ptr(compensation), _ptr(v_xy), _ptr(v_depth), _ptr(v_conic), _ptr(v_compensation),
            _ptr(v_cov2d) if need_cov_grads else None, _ptr(v_cov3d) if need_cov_grads else None, _ptr(v_mean3d),
            _ptr(v_scale), _ptr(v_quat), st), "project_gaussians_backward")
    return v_cov2d, v_cov3d, v_mean3d, v_scale, v_quat


# ---------------------------------------------------------------------------------------------------
def cumsum_tiles_hit(num_tiles_hit: Tensor, total_pinned: Tensor = None) -> Tensor:
    """int32 inclusive scan (replaces torch.cumsum at rasterizer/utils.py:123).  If `total_pinned` (a pinned
    CPU int32 tensor with one element) is given, the total is copied into it asynchronously."""
    _check_input(num_tiles_hit, "num_tiles_hit", torch.int32)
    n = num_tiles_hit.numel()
    lib = _lib.load()
    cum = torch.empty_like(num_tiles_hit)
    ws_bytes = lib.gsr_cumsum_workspace_bytes(n)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=num_tiles_hit.device)
    with _Guard(num_tiles_hit) as st:
        _lib.check(lib.gsr_cumsum_tiles_hit(n, _ptr(num_tiles_hit), _ptr(cum),
                                            _P(total_pinned.data_ptr()) if total_pinned is not None else None,
                                            _ptr(ws), ws_bytes, st), "cumsum_tiles_hit")
    return cum


def map_gaussian_to_intersects(num_points: int, num_intersects: int, xys: Tensor, depths: Tensor, radii: Tensor,
                               cum_tiles_hit: Tensor, tile_bounds: Tuple[int, int, int], block_width: int):
    _check_input(xys, "xys", torch.float32)
    _check_input(depths, "depths", torch.float32)
    _check_input(radii, "radii", torch.int32)
    _check_input(cum_tiles_hit, "cum_tiles_hit", torch.int32)
    gaussian_ids = torch.zeros((num_intersects,), dtype=torch.int32, device=xys.device)
    isect_ids = torch.zeros((num_intersects,), dtype=torch.int64, device=xys.device)
    with _Guard(xys) as st:
        _lib.check(_lib.load().gsr_map_gaussian_to_intersects(
            num_points, num_intersects, _ptr(xys), _ptr(depths), _ptr(radii), _ptr(cum_tiles_hit),
            int(tile_bounds[0]), int(tile_bounds[1]), int(block_width), _ptr(isect_ids), _ptr(gaussian_ids), st),
            "map_gaussian_to_intersects")
    return isect_ids, gaussian_ids


def count_tiles_tight(xys: Tensor, radii: Tensor, conics: Tensor, opacities: Tensor, img_height: int, img_width: int,
                      block_width: int) -> Tensor:
    """Per-Gaussian number of bounding-box tiles in which some pixel can reach alpha >= 1/255 (exact, conservative
    culling; see include/gsr_b200.h).  Always <= the reference's num_tiles_hit."""
    _check_input(xys, "xys", torch.float32)
    _check_input(radii, "radii", torch.int32)
    _check_input(conics, "conics", torch.float32)
    _check_input(opacities, "opacities", torch.float32)
    n = xys.size(0)
    out = torch.empty((n,), dtype=torch.int32, device=xys.device)
    with _Guard(xys) as st:
        _lib.check(_lib.load().gsr_count_tiles_tight(n, _ptr(xys), _ptr(radii), _ptr(conics), _ptr(opacities),
                                                     int(img_height), int(img_width), int(block_width), _ptr(out), st),
                   "count_tiles_tight")
    return out


def map_gaussian_to_intersects_tight(num_points: int, num_intersects: int, xys: Tensor, depths: Tensor, radii: Tensor,
                                     conics: Tensor, opacities: Tensor, cum_tiles: Tensor, img_height: int,
                                     img_width: int, block_width: int):
    _check_input(xys, "xys", torch.float32)
    _check_input(depths, "depths", torch.float32)
    _check_input(radii, "radii", torch.int32)
    _check_input(conics, "conics", torch.float32)
    _check_input(opacities, "opacities", torch.float32)
    _check_input(cum_tiles, "cum_tiles", torch.int32)
    gaussian_ids = torch.empty((num_intersects,), dtype=torch.int32, device=xys.device)
    isect_ids = torch.empty((num_intersects,), dtype=torch.int64, device=xys.device)
    with _Guard(xys) as st:
        _lib.check(_lib.load().gsr_map_gaussian_to_intersects_tight(
            num_points, num_intersects, _ptr(xys), _ptr(depths), _ptr(radii), _ptr(conics), _ptr(opacities),
            _ptr(cum_tiles), int(img_height), int(img_width), int(block_width), _ptr(isect_ids), _ptr(gaussian_ids),
            st), "map_gaussian_to_intersects_tight")
    return isect_ids, gaussian_ids


_pinned_words = {}


def _pinned_word(dev_index: int, stream_ptr: int) -> Tensor:
    key = (dev_index, stream_ptr)
    w = _pinned_words.get(key)
    if w is None:
        w = _pinned_words[key] = torch.zeros(4, dtype=torch.int32).pin_memory()
    return w


def bin_gaussians_fast(xys: Tensor, depths: Tensor, radii: Tensor, conics: Tensor, opacities: Tensor,
                       img_height: int, img_width: int, block_width: int):
    """Internal binning of rasterize_gaussians, synchronous form: (num_intersects, gaussian_ids_sorted [M] i32,
    tile_bins [T,2] i32) with exactly-sized outputs.  Per-tile counting + per-tile sorts with exact tile culling
    (include/gsr_b200.h, gsr_bin_count / gsr_bin_fill_sort); the per-tile order is the reference's.  One stream
    synchronisation to learn M (the reference has the same one: `.item()` at rasterizer/utils.py:124)."""
    _check_input(xys, "xys", torch.float32)
    _check_input(depths, "depths", torch.float32)
    _check_input(radii, "radii", torch.int32)
    _check_input(conics, "conics", torch.float32)
    _check_input(opacities, "opacities", torch.float32)
    lib = _lib.load()
    n, dev = xys.size(0), xys.device
    H, W, bw = int(img_height), int(img_width), int(block_width)
    tiles_x, tiles_y = (W + bw - 1) // bw, (H + bw - 1) // bw
    ws_bytes = lib.gsr_bin_count_workspace_bytes(n, H, W, bw)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    tile_bins = torch.empty((tiles_x * tiles_y, 2), dtype=torch.int32, device=dev)
    meta = torch.empty((4,), dtype=torch.int32, device=dev)
    with _Guard(xys) as st:
        pin = _pinned_word(dev.index, st.value or 0)
        _lib.check(lib.gsr_bin_count(n, _ptr(xys), _ptr(radii), _ptr(conics), _ptr(opacities), H, W, bw, _ptr(tile_bins),
                                     _ptr(meta), _P(pin.data_ptr()), _ptr(ws), ws_bytes, st), "bin_count")
        torch.cuda.current_stream(dev).synchronize()
        m = int(pin[0].item())
        if m < 1:
            return 0, None, None
        ids_sorted = torch.empty((m,), dtype=torch.int32, device=dev)
        ws2_bytes = lib.gsr_bin_fill_workspace_bytes(m)
        ws2 = torch.empty((ws2_bytes,), dtype=torch.uint8, device=dev)
        _lib.check(lib.gsr_bin_fill_sort(n, m, _ptr(xys), _ptr(depths), _ptr(radii), _ptr(conics), _ptr(opacities), H, W, bw,
                                         _ptr(tile_bins), _ptr(ws), _ptr(ids_sorted), _ptr(ws2), ws2_bytes, st),
                   "bin_fill_sort")
    return m, ids_sorted, tile_bins


def bin_gaussians_device(xys: Tensor, depths: Tensor, radii: Tensor, conics: Tensor, opacities: Tensor,
                         img_height: int, img_width: int, block_width: int, capacity: int,
                         meta_pinned: Optional[Tensor] = None):
    """Asynchronous internal binning (include/gsr_b200.h, gsr_bin_gaussians_device): the pair count M stays on the
    device, nothing synchronises.  Returns (gaussian_ids_sorted [capacity] i32, tile_bins [T,2] i32, meta int32[4] on
    the device = {M, overflow, min(M, capacity), 0}).  `meta_pinned` (pinned host int32[4]) receives a copy on the
    current stream; the caller inspects it later (rasterizer.binning)."""
    _check_input(xys, "xys", torch.float32)
    _check_input(depths, "depths", torch.float32)
    _check_input(radii, "radii", torch.int32)
    _check_input(conics, "conics", torch.float32)
    _check_input(opacities, "opacities", torch.float32)
    lib = _lib.load()
    n, dev = xys.size(0), xys.device
    capacity = int(capacity)
    tiles_x = (img_width + block_width - 1) // block_width
    tiles_y = (img_height + block_width - 1) // block_width
    ids_sorted = torch.empty((max(capacity, 1),), dtype=torch.int32, device=dev)
    tile_bins = torch.empty((tiles_x * tiles_y, 2), dtype=torch.int32, device=dev)
    meta = torch.empty((4,), dtype=torch.int32, device=dev)
    ws_bytes = lib.gsr_bin_device_workspace_bytes(n, capacity, int(img_height), int(img_width), int(block_width))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    with _Guard(xys) as st:
        _lib.check(lib.gsr_bin_gaussians_device(
            n, _ptr(xys), _ptr(depths), _ptr(radii), _ptr(conics), _ptr(opacities), int(img_height), int(img_width),
            int(block_width), capacity, _ptr(ids_sorted), _ptr(tile_bins), _ptr(meta),
            _P(meta_pinned.data_ptr()) if meta_pinned is not None else None, _ptr(ws), ws_bytes, st), "bin_gaussians_device")
    return ids_sorted, tile_bins, meta


def sort_intersects(isect_ids: Tensor, gaussian_ids: Tensor, num_tiles: int) -> Tuple[Tensor, Tensor]:
    """Stable radix sort of (key, Gaussian id) pairs (replaces torch.sort + torch.gather, utils.py:179-180)."""
    _check_input(isect_ids, "isect_ids", torch.int64)
    _check_input(gaussian_ids, "gaussian_ids", torch.int32)
    m = isect_ids.numel()
    lib = _lib.load()
    keys = torch.empty_like(isect_ids)
    vals = torch.empty_like(gaussian_ids)
    ws_bytes = lib.gsr_sort_workspace_bytes(m)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=isect_ids.device)
    with _Guard(isect_ids) as st:
        _lib.check(lib.gsr_sort_intersects(m, int(num_tiles), _ptr(isect_ids), _ptr(gaussian_ids), _ptr(keys),
                                           _ptr(vals), _ptr(ws), ws_bytes, st), "sort_intersects")
    return keys, vals


def get_tile_bin_edges(num_intersects: int, isect_ids_sorted: Tensor, tile_bounds: Tuple[int, int, int]) -> Tensor:
    _check_input(isect_ids_sorted, "isect_ids_sorted", torch.int64)
    num_tiles = int(tile_bounds[0]) * int(tile_bounds[1])
    tile_bins = torch.empty((num_tiles, 2), dtype=torch.int32, device=isect_ids_sorted.device)
    with _Guard(isect_ids_sorted) as st:
        _lib.check(_lib.load().gsr_get_tile_bin_edges(num_intersects, _ptr(isect_ids_sorted), num_tiles,
                                                      _ptr(tile_bins), st), "get_tile_bin_edges")
    return tile_bins


# ---------------------------------------------------------------------------------------------------
def _check_raster_inputs(gaussian_ids_sorted, tile_bins, xys, conics, colors, opacities, background):
    _check_input(gaussian_ids_sorted, "gaussian_ids_sorted", torch.int32)
    _check_input(tile_bins, "tile_bins", torch.int32)
    _check_input(xys, "xys", torch.float32)
    _check_input(conics, "conics", torch.float32)
    _check_input(colors, "colors", torch.float32)
    _check_input(opacities, "opacities", torch.float32)
    _check_input(background, "background", torch.float32)


def _rasterize_forward(nd: bool, tile_bounds, block, img_size, gaussian_ids_sorted, tile_bins, xys, conics, colors,
                       opacities, background):
    _check_raster_inputs(gaussian_ids_sorted, tile_bins, xys, conics, colors, opacities, background)
    channels = colors.size(1)
    img_width, img_height = int(img_size[0]), int(img_size[1])
    block_width = int(block[0])
    dev = xys.device
    out_img = torch.empty((img_height, img_width, channels), dtype=torch.float32, device=dev)
    final_Ts = torch.empty((img_height, img_width), dtype=torch.float32, device=dev)
    final_idx = torch.empty((img_height, img_width), dtype=torch.int32, device=dev)
    lib = _lib.load()
    with _Guard(xys) as st:
        if nd:
            rc = lib.gsr_nd_rasterize_forward(img_height, img_width, block_width, channels, xys.size(0),
                                              _ptr(gaussian_ids_sorted), _ptr(tile_bins), _ptr(xys), _ptr(conics),
                                              _ptr(colors), _ptr(opacities), _ptr(background), _ptr(out_img),
                                              _ptr(final_Ts), _ptr(final_idx), st)
        else:
            if channels != 3:
                raise RuntimeError("rasterize_forward: colors must have 3 channels (use nd_rasterize_forward)")
            rc = lib.gsr_rasterize_forward(img_height, img_width, block_width, xys.size(0),
                                           _ptr(gaussian_ids_sorted), _ptr(tile_bins), _ptr(xys), _ptr(conics),
                                           _ptr(colors), _ptr(opacities), _ptr(background), _ptr(out_img),
                                           _ptr(final_Ts), _ptr(final_idx), st)
        _lib.check(rc, "nd_rasterize_forward" if nd else "rasterize_forward")
    return out_img, final_Ts, final_idx


def rasterize_forward(tile_bounds, block, img_size, gaussian_ids_sorted, tile_bins, xys, conics, colors, opacities,
                      background):
    return _rasterize_forward(False, tile_bounds, block, img_size, gaussian_ids_sorted, tile_bins, xys, conics,
                              colors, opacities, background)


def nd_rasterize_forward(tile_bounds, block, img_size, gaussian_ids_sorted, tile_bins, xys, conics, colors,
                         opacities, background):
    return _rasterize_forward(True, tile_bounds, block, img_size, gaussian_ids_sorted, tile_bins, xys, conics,
                              colors, opacities, background)


def _rasterize_backward(nd: bool, img_height, img_width, block_width, gaussians_ids_sorted, tile_bins, xys, conics,
                        colors, opacities, background, final_Ts, final_idx, v_output, v_output_alpha,
                        out_opacity=None):
    _check_input(xys, "xys", torch.float32)
    _check_input(colors, "colors", torch.float32)
    if xys.dim() != 2 or xys.size(1) != 2:
        raise RuntimeError("xys must have dimensions (num_points, 2)")
    if colors.dim() != 2 or (not nd and colors.size(1) != 3):
        raise RuntimeError("colors must have 2 dimensions")
    # the reference calls .contiguous() on everything else (bindings.cu:510-526)
    gaussians_ids_sorted = gaussians_ids_sorted.contiguous()
    tile_bins, conics, opacities = tile_bins.contiguous(), conics.contiguous(), opacities.contiguous()
    background, final_Ts, final_idx = background.contiguous(), final_Ts.contiguous(), final_idx.contiguous()
    v_output, v_output_alpha = v_output.contiguous(), v_output_alpha.contiguous()
    if v_output.dtype != torch.float32 or v_output_alpha.dtype != torch.float32:
        raise RuntimeError("v_output / v_output_alpha: expected scalar type Float")
    num_points, channels = xys.size(0), colors.size(1)
    dev = xys.device
    v_xy = torch.empty((num_points, 2), dtype=torch.float32, device=dev)
    v_conic = torch.empty((num_points, 3), dtype=torch.float32, device=dev)
    v_colors = torch.empty((num_points, channels), dtype=torch.float32, device=dev)
    v_opacity = _out_or_empty(out_opacity, (num_points, 1), dev, "out_opacity")
    lib = _lib.load()
    with _Guard(xys) as st:
        args = (_ptr(gaussians_ids_sorted), _ptr(tile_bins), _ptr(xys), _ptr(conics), _ptr(colors), _ptr(opacities),
                _ptr(background), _ptr(final_Ts), _ptr(final_idx), _ptr(v_output), _ptr(v_output_alpha), _ptr(v_xy),
                _ptr(v_conic), _ptr(v_colors), _ptr(v_opacity), st)
        if nd:
            rc = lib.gsr_nd_rasterize_backward(int(img_height), int(img_width), int(block_width), channels,
                                               num_points, *args)
        else:
            rc = lib.gsr_rasterize_backward(int(img_height), int(img_width), int(block_width), num_points, *args)
        _lib.check(rc, "nd_rasterize_backward" if nd else "rasterize_backward")
    return v_xy, v_conic, v_colors, v_opacity


def rasterize_backward(img_height, img_width, block_width, gaussians_ids_sorted, tile_bins, xys, conics, colors,
                       opacities, background, final_Ts, final_idx, v_output, v_output_alpha, *, out_opacity=None):
    return _rasterize_backward(False, img_height, img_width, block_width, gaussians_ids_sorted, tile_bins, xys,
                               conics, colors, opacities, background, final_Ts, final_idx, v_output, v_output_alpha,
                               out_opacity=out_opacity)


def nd_rasterize_backward(img_height, img_width, block_width, gaussians_ids_sorted, tile_bins, xys, conics, colors,
                          opacities, background, final_Ts, final_idx, v_output, v_output_alpha):
    return _rasterize_backward(True, img_height, img_width, block_width, gaussians_ids_sorted, tile_bins, xys,
                               conics, colors, opacities, background, final_Ts, final_idx, v_output, v_output_alpha)


# ---------------------------------------------------------------------------------------------------
# fused render operator; see include/gsr_b200.h "FUSED render operator"
def fused_preprocess_forward(means3d: Tensor, scales_raw: Tensor, quats_raw: Tensor, opacities_raw: Tensor,
                             features_dc: Tensor, features_rest: Tensor, viewmat: Tensor, projmat: Tensor,
                             glob_scale: float, fx: float, fy: float, cx: float, cy: float, img_height: int,
                             img_width: int, block_width: int, degrees_to_use: int, clip_thresh: float = 0.01,
                             antialiased: bool = False):
    """-> (records [3,N,4], xys [N,2], depths [N], radii [N] i32, conics [N,3], opacities [N], clamp_mask [N] i32,
    compensation [N] | None).  `antialiased`: opacities = sigmoid(raw) * compensation (vanilla_gs.py:813-816)."""
    for t, nm in ((means3d, "means3d"), (scales_raw, "scales"), (quats_raw, "quats"), (opacities_raw, "opacities"),
                  (features_dc, "features_dc"), (features_rest, "features_rest"), (viewmat, "viewmat"),
                  (projmat, "projmat")):
        _check_input(t, nm, torch.float32)
    n, dev = means3d.size(0), means3d.device
    k_rest = features_rest.size(1) if features_rest.dim() == 3 else 0
    sh_degree = {0: 0, 3: 1, 8: 2, 15: 3, 24: 4}.get(k_rest)
    if sh_degree is None or features_dc.numel() != 3 * n or features_rest.numel() != 3 * n * k_rest:
        raise RuntimeError("features_dc must be [N,3] and features_rest [N,(d+1)^2-1,3]")
    f32 = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    rec = torch.empty((3, n, 4), **f32)
    xys, depths = torch.empty((n, 2), **f32), torch.empty((n,), **f32)
    radii, conics = torch.empty((n,), **i32), torch.empty((n, 3), **f32)
    opac, mask = torch.empty((n,), **f32), torch.empty((n,), **i32)
    comp = torch.empty((n,), **f32) if antialiased else None
    with _Guard(means3d) as st:
        _lib.check(_lib.load().gsr_fused_preprocess_forward(
            n, sh_degree, int(degrees_to_use), _ptr(means3d), _ptr(scales_raw), _ptr(quats_raw), _ptr(opacities_raw),
            _ptr(features_dc), _ptr(features_rest), _ptr(viewmat), _ptr(projmat), float(glob_scale), float(fx), float(fy),
            float(cx), float(cy), int(img_height), int(img_width), int(block_width), float(clip_thresh), _ptr(rec),
            _ptr(xys), _ptr(depths), _ptr(radii), _ptr(conics), _ptr(opac), _ptr(mask),
            _ptr(comp) if comp is not None else None, st), "fused_preprocess_forward")
    return rec, xys, depths, radii, conics, opac, mask, comp


def fused_preprocess_backward(means3d, scales_raw, quats_raw, opacities_raw, k_rest: int, degrees_to_use: int, viewmat,
                              projmat, glob_scale, fx, fy, img_height, img_width, radii, conics, clamp_mask, grad_rec,
                              v_xys_extra=None, out=None, compensation=None):
    """-> (v_means3d, v_scales_raw, v_quats_raw, v_opacities_raw [N,1], v_features_dc [N,3], v_features_rest [N,k_rest,3]).
    `out`: optional dict of preallocated outputs with those names (e.g. views of a GradientBucket)."""
    _check_input(grad_rec, "grad_rec", torch.float32)
    n, dev = means3d.size(0), means3d.device
    sh_degree = {0: 0, 3: 1, 8: 2, 15: 3, 24: 4}[k_rest]
    out = out or {}
    v_means = _out_or_empty(out.get("v_means3d"), (n, 3), dev, "v_means3d")
    v_scales = _out_or_empty(out.get("v_scales"), (n, 3), dev, "v_scales")
    v_quats = _out_or_empty(out.get("v_quats"), (n, 4), dev, "v_quats")
    v_opac = _out_or_empty(out.get("v_opacities"), (n, 1), dev, "v_opacities")
    v_dc = _out_or_empty(out.get("v_features_dc"), (n, 3), dev, "v_features_dc")
    v_rest = _out_or_empty(out.get("v_features_rest"), (n, k_rest, 3), dev, "v_features_rest")
    with _Guard(means3d) as st:
        _lib.check(_lib.load().gsr_fused_preprocess_backward(
            n, sh_degree, int(degrees_to_use), _ptr(means3d), _ptr(scales_raw), _ptr(quats_raw), _ptr(opacities_raw),
            _ptr(viewmat), _ptr(projmat), float(glob_scale), float(fx), float(fy), int(img_height), int(img_width),
            _ptr(radii), _ptr(conics), _ptr(clamp_mask), _ptr(compensation) if compensation is not None else None,
            _ptr(grad_rec), _ptr(v_xys_extra) if v_xys_extra is not None else None, _ptr(v_means), _ptr(v_scales), _ptr(v_quats),
            _ptr(v_opac), _ptr(v_dc), _ptr(v_rest), st), "fused_preprocess_backward")
    return v_means, v_scales, v_quats, v_opac, v_dc, v_rest


def blend_packed_forward(img_height: int, img_width: int, block_width: int, gaussian_ids_sorted: Tensor,
                         tile_bins: Tensor, records: Tensor, background: Tensor, with_depth: bool):
    """-> (out_img [H,W,3], out_depth [H,W] | None, final_Ts [H,W], final_idx [H,W] i32)"""
    _check_input(records, "records", torch.float32)
    _check_input(background, "background", torch.float32)
    n, dev = records.size(1), records.device
    out_img = torch.empty((img_height, img_width, 3), dtype=torch.float32, device=dev)
    out_depth = torch.empty((img_height, img_width), dtype=torch.float32, device=dev) if with_depth else None
    final_Ts = torch.empty((img_height, img_width), dtype=torch.float32, device=dev)
    final_idx = torch.empty((img_height, img_width), dtype=torch.int32, device=dev)
    with _Guard(records) as st:
        _lib.check(_lib.load().gsr_blend_packed_forward(
            int(img_height), int(img_width), int(block_width), n, _ptr(gaussian_ids_sorted), _ptr(tile_bins),
            _ptr(records), _ptr(background), _ptr(out_img), _ptr(out_depth) if with_depth else None, _ptr(final_Ts),
            _ptr(final_idx), st), "blend_packed_forward")
    return out_img, out_depth, final_Ts, final_idx


def blend_packed_backward(img_height: int, img_width: int, block_width: int, gaussian_ids_sorted: Tensor,
                          tile_bins: Tensor, records: Tensor, background: Tensor, final_Ts: Tensor, final_idx: Tensor,
                          v_output: Tensor, v_output_depth, v_output_alpha: Tensor) -> Tensor:
    """-> grad_rec [N,12] = {v_x, v_y, v_opacity, v_depth | v_a, v_b, v_c, - | v_r, v_g, v_b, -}"""
    for t, nm in ((v_output, "v_output"), (v_output_alpha, "v_output_alpha")):
        _check_input(t, nm, torch.float32)
    if v_output_depth is not None:
        _check_input(v_output_depth, "v_output_depth", torch.float32)
    n, dev = records.size(1), records.device
    grad_rec = torch.empty((n, 12), dtype=torch.float32, device=dev)
    with _Guard(records) as st:
        _lib.check(_lib.load().gsr_blend_packed_backward(
            int(img_height), int(img_width), int(block_width), n, _ptr(gaussian_ids_sorted), _ptr(tile_bins),
            _ptr(records), _ptr(background), _ptr(final_Ts), _ptr(final_idx), _ptr(v_output),
            _ptr(v_output_depth) if v_output_depth is not None else None, _ptr(v_output_alpha), _ptr(grad_rec), st),
            "blend_packed_backward")
    return grad_rec

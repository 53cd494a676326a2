"""CPU check that the two culling stages are CONSERVATIVE with respect to the reference's per-pixel alpha test.

The product's helpers — exact tile culling of the binning (csrc/tile_cull.cuh) and the warp-block masks of the 16x16
blend kernels (block_mask_16, csrc/blend_common.cuh) — are __host__ __device__; tests/host_cull/host_cull.cu wraps them
for the host (built here with nvcc, no GPU needed).  For seeded Gaussians, including long thin slanted ones, every
(Gaussian, tile) pair and every (Gaussian, 8x4 pixel block) pair in which SOME pixel passes the reference's test

    sigma = 0.5 (a dx^2 + c dy^2) + b dx dy >= 0   and   alpha = min(0.999, o exp(-sigma)) >= 1/255     (forward.cu:355-363)

evaluated in FP32 in the reference's expression order, must be kept.  The same run checks that the binning's two passes
agree: walking the cached 64-bit tile mask (fill pass, division-free `walk_tile_mask`) visits exactly the tiles `cull_tiles`
counted (count pass), in the same order.  (On the GPU the same property is checked end to end
by test_tight_binning_is_exact / test_block_masks_match_per_warp_tests: bitwise image, superset of pairs.)"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_cull", "host_cull.cu")
CSRC = os.path.join(ROOT, "gaussian-splatting-toolkit_b200", "csrc")
f32 = np.float32


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("host_cull") / "libhost_cull.so")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "--expt-relaxed-constexpr",
                    "-shared", "-Xcompiler", "-fPIC", "-I", CSRC, "-o", out, SRC], check=True)
    L = ctypes.CDLL(out)
    fl, it = ctypes.c_float, ctypes.c_int
    L.host_block_mask_16.argtypes = [fl] * 8
    L.host_block_mask_16.restype = ctypes.c_uint
    L.host_kept_tiles.argtypes = [fl, fl, it, fl, fl, fl, fl, it, it, it, ctypes.POINTER(it), it,
                                  ctypes.POINTER(ctypes.c_ulonglong)]
    L.host_kept_tiles.restype = it
    L.host_walk_mask.argtypes = [ctypes.c_ulonglong, it, it, it, it, ctypes.POINTER(it), it]
    L.host_walk_mask.restype = it
    L.host_tile_bbox.argtypes = [fl, fl, it, it, it, it, ctypes.POINTER(it)]
    return L


def _gaussians(n, seed, extent):
    """2-D Gaussians the way the projection hands them over: centre, conic = inverse of (cov + 0.3 I), 3-sigma radius,
    opacity.  A third are needles (axis ratio up to 300) at arbitrary angles, sizes from sub-pixel to several tiles."""
    g = np.random.default_rng(seed)
    ctr = g.uniform(-24.0, extent + 24.0, (n, 2))
    big = np.exp(g.uniform(np.log(0.3), np.log(60.0), n))
    ratio = np.where(g.random(n) < 0.33, np.exp(g.uniform(0, np.log(300.0), n)), np.exp(g.uniform(0, np.log(4.0), n)))
    small = big / ratio
    th = g.uniform(0, np.pi, n)
    cs, sn = np.cos(th), np.sin(th)
    cxx = (cs * big) ** 2 + (sn * small) ** 2 + 0.3
    cyy = (sn * big) ** 2 + (cs * small) ** 2 + 0.3
    cxy = cs * sn * (big ** 2 - small ** 2)
    cxx, cxy, cyy = cxx.astype(f32), cxy.astype(f32), cyy.astype(f32)
    det = cxx * cyy - cxy * cxy
    a, b, c = (cyy / det).astype(f32), (-cxy / det).astype(f32), (cxx / det).astype(f32)
    mid = f32(0.5) * (cxx + cyy)
    lam = mid + np.sqrt(np.maximum(f32(0.1), mid * mid - det))
    radius = np.ceil(f32(3.0) * np.sqrt(lam)).astype(np.int32)
    opac = np.where(g.random(n) < 0.15, g.uniform(0.003, 0.02, n), g.uniform(0.02, 1.0, n)).astype(f32)
    return ctr.astype(f32), a, b, c, radius, opac


def _contributes(ctr, a, b, c, opac, px, py):
    """the reference's per-pixel test in FP32, in its expression order (px, py: pixel coordinate grids)"""
    dx, dy = f32(ctr[0]) - px.astype(f32), f32(ctr[1]) - py.astype(f32)
    sigma = f32(0.5) * (a * dx * dx + c * dy * dy) + b * dx * dy
    alpha = np.minimum(f32(0.999), opac * np.exp(-sigma, dtype=f32))
    return (sigma >= 0) & (alpha >= f32(1.0 / 255.0))


def test_block_masks_are_conservative(lib):
    ctr, a, b, c, radius, opac = _gaussians(6000, 5, 16.0)
    tile_x0, tile_y0 = 32.0, 48.0
    py, px = np.meshgrid(np.arange(16) + tile_y0, np.arange(16) + tile_x0, indexing="ij")
    kept = needed = 0
    for i in range(len(opac)):
        x, y = ctr[i, 0] + tile_x0, ctr[i, 1] + tile_y0
        m = lib.host_block_mask_16(x, y, a[i], b[i], c[i], opac[i], tile_x0, tile_y0)
        hit = _contributes((x, y), a[i], b[i], c[i], opac[i], px, py)
        for w in range(8):  # warp w: block column w & 1 (8 px wide), block row w >> 1 (4 px high)
            blk = hit[4 * (w >> 1):4 * (w >> 1) + 4, 8 * (w & 1):8 * (w & 1) + 8]
            if blk.any():
                needed += 1
                assert (m >> w) & 1, (i, w, float(x), float(y), float(a[i]), float(b[i]), float(c[i]), float(opac[i]))
            kept += (m >> w) & 1
    assert needed > 5000  # the scene exercises the test
    print(f"[block masks] {needed} (Gaussian, block) pairs contribute, {kept} kept ({kept / needed:.3f}x)")
    assert kept <= 1.35 * needed  # and the masks still cull: few false positives


def test_degenerate_records_keep_every_block(lib):
    nan = float("nan")
    assert lib.host_block_mask_16(40.0, 56.0, nan, 0.0, 1.0, 0.5, 32.0, 48.0) == 0xFF  # NaN conic
    assert lib.host_block_mask_16(40.0, 56.0, 1.0, 0.0, 1.0, nan, 32.0, 48.0) == 0xFF  # NaN opacity
    assert lib.host_block_mask_16(40.0, 56.0, 1.0, 5.0, 1.0, 0.9, 32.0, 48.0) == 0xFF  # indefinite conic: never culled
    assert lib.host_block_mask_16(40.0, 56.0, 1.0, 0.0, 1.0, 0.003, 32.0, 48.0) == 0    # opacity < 1/255: never visible


def test_tile_culling_is_conservative(lib):
    bw, tiles_x, tiles_y = 16, 10, 8
    W, H = bw * tiles_x, bw * tiles_y
    ctr, a, b, c, radius, opac = _gaussians(2500, 9, float(min(W, H)))
    ctr = ctr * f32(W / min(W, H))
    py, px = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    buf = (ctypes.c_int * (tiles_x * tiles_y))()
    buf2 = (ctypes.c_int * (tiles_x * tiles_y))()
    box = (ctypes.c_int * 4)()
    kept = needed = in_box = walked = big_boxes = 0
    for i in range(len(opac)):
        x, y, r = ctr[i, 0], ctr[i, 1], int(radius[i])
        mask = ctypes.c_ulonglong(0)
        n = lib.host_kept_tiles(x, y, r, a[i], b[i], c[i], opac[i], tiles_x, tiles_y, bw, buf, len(buf), ctypes.byref(mask))
        assert n >= 0  # the returned count equals the number of visits
        order = list(buf[:n])
        tiles = set(order)
        assert len(tiles) == n
        lib.host_tile_bbox(x, y, r, tiles_x, tiles_y, bw, box)
        x0, y0, x1, y1 = box
        # count pass and fill pass must agree: walking the cached mask visits exactly the tiles the count pass counted, in
        # the same order (a disagreement would write keys outside the tile's segment); boxes of > 64 tiles cache nothing
        if (x1 - x0) * (y1 - y0) <= 64:
            m = lib.host_walk_mask(mask.value, x0, y0, x1 - x0, tiles_x, buf2, len(buf2))
            assert m == n == bin(mask.value).count("1") and list(buf2[:m]) == order, (i, hex(mask.value))
            walked += 1
        else:
            assert mask.value == 0
            big_boxes += 1
        in_box += max(0, x1 - x0) * max(0, y1 - y0)
        hit = _contributes((x, y), a[i], b[i], c[i], opac[i], px, py)
        per_tile = hit.reshape(tiles_y, bw, tiles_x, bw).any(axis=(1, 3))
        for ty in range(y0, y1):  # the reference only ever visits the tiles of its bounding box
            for tx in range(x0, x1):
                if per_tile[ty, tx]:
                    needed += 1
                    assert ty * tiles_x + tx in tiles, (i, tx, ty, float(x), float(y), r, float(a[i]), float(b[i]), float(c[i]))
        assert all(x0 <= t % tiles_x < x1 and y0 <= t // tiles_x < y1 for t in tiles)
        kept += n
    assert needed > 3000
    print(f"[tile culling] bounding boxes {in_box} pairs, kept {kept}, contributing {needed}")
    assert kept <= 1.25 * needed and kept < in_box
    assert walked > 1500 and big_boxes > 20  # both the cached-mask and the re-evaluation path were exercised

"""CPU-only analysis for DESIGN.md §8: how many blend-loop iterations would a warp need if its two 16-lane halves each
owned a 4x4 pixel block with its own survivor list (iterations per batch = max of the two list lengths) instead of one
8x4 block (iterations = size of the union)?  Uses the CPU oracle's projection + binning of the cfg2-style scene and
evaluates alpha >= 1/255 exactly on a sample of tiles (early termination ignored).
usage: python tools/sim_halfwarp_units.py [num_gaussians] [width] [height] [tiles_sampled]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gaussian-splatting-toolkit_b200")):
    sys.path.insert(0, p)
from oracle import oracle as orc  # noqa: E402  (analysis tool, not the product)
from rasterizer.synthetic import make_scene  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
    n_tiles = int(sys.argv[4]) if len(sys.argv) > 4 else 150
    orc.build()
    scene = make_scene(n, W, H, 0.002, 0.02, margin=1.1, seed=0)
    out = orc.render_view(scene, backward=False)
    xys, conics, opac = out["xys"], out["conics"], np.asarray(scene["opacities"]).reshape(-1)
    ids, bins = out["gaussian_ids_sorted"], out["tile_bins"]
    tiles_x = (W + 15) // 16
    rng = np.random.default_rng(0)
    sample = rng.choice(bins.shape[0], size=min(n_tiles, bins.shape[0]), replace=False)
    it_union = it_half = it_quarter = visits32 = lanes_valid = pairs = 0
    for t in sample:
        lo, hi = bins[t]
        if hi <= lo:
            continue
        g = ids[lo:hi]
        ty, tx = divmod(int(t), tiles_x)
        px = (tx * 16 + np.arange(16))[None, :, None].astype(np.float32)
        py = (ty * 16 + np.arange(16))[:, None, None].astype(np.float32)
        dx, dy = xys[g, 0][None, None, :] - px, xys[g, 1][None, None, :] - py
        sigma = 0.5 * (conics[g, 0] * dx * dx + conics[g, 2] * dy * dy) + conics[g, 1] * dx * dy
        alpha = np.minimum(0.999, opac[g] * np.exp(-sigma))
        hit = (sigma >= 0) & (alpha >= 1.0 / 255.0) & (px < W) & (py < H)          # [16,16,G]
        pairs += g.size
        for b0 in range(0, g.size, 256):                                               # batches of 256 records
            hb = hit[:, :, b0:b0 + 256]
            for wy in range(4):                                                        # warp = 8 wide x 4 high
                for wx in range(2):
                    blk = hb[4 * wy:4 * wy + 4, 8 * wx:8 * wx + 8]
                    any32 = blk.any(axis=(0, 1))
                    a, b = blk[:, :4].any(axis=(0, 1)), blk[:, 4:].any(axis=(0, 1))
                    q = [blk[2 * i:2 * i + 2, 4 * j:4 * j + 4].any(axis=(0, 1)).sum() for i in range(2) for j in range(2)]
                    it_union += int(any32.sum())
                    it_half += int(max(a.sum(), b.sum()))
                    it_quarter += int(max(q))
                    visits32 += int(any32.sum())
                    lanes_valid += int(blk[:, :, any32].sum())
    print(f"{len(sample)} tiles, {pairs} (tile, Gaussian) pairs; contributing (8x4 warp, Gaussian) visits per pair: {visits32 / pairs:.2f}")
    print(f"valid lanes per visit: {lanes_valid / max(1, visits32):.1f} of 32")
    print(f"loop iterations, 8x4 warp units       : {it_union}")
    print(f"loop iterations, two 4x4 half-warps   : {it_half}  ({it_half / it_union:.2f} x)")
    print(f"loop iterations, four 4x2 quarter-warps: {it_quarter}  ({it_quarter / it_union:.2f} x)")


if __name__ == "__main__":
    main()

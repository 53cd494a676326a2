"""CPU-only analysis (not product code): cost model of blend-adjoint designs on the cfg2 scene, with the per-pixel early
termination (final_idx) taken from the CPU oracle.  For a sample of tiles it counts, per design,
  * loop iterations (warp-instructions scale with it), * (unit, Gaussian) flushes (= atomic groups of 9), * valid lanes.
Designs: pixel-parallel 8x4 warps (current kernel); Gaussian-parallel RING of 32 lanes over 8x4 pixels (+31 fill/drain
per tile-warp); two 16-lane rings over 4x4 pixels (+15); 32-lane ring over 16x4 pixels, 2 pixels per lane-step (+31).
usage: python tools/sim_ring_units.py [tiles_sampled]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gaussian-splatting-toolkit_b200")):
    sys.path.insert(0, p)
from oracle import oracle as orc  # noqa: E402
from rasterizer.synthetic import make_config_scene  # noqa: E402


def main():
    n_tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    orc.build()
    scene = make_config_scene("cfg2")
    W, H = scene["img_width"], scene["img_height"]
    out = orc.render_view(scene, backward=False)
    xys, conics, opac = out["xys"], out["conics"], np.asarray(scene["opacities"]).reshape(-1)
    ids, bins, fidx = out["gaussian_ids_sorted"], out["tile_bins"], out["final_idx"]
    tiles_x = (W + 15) // 16
    rng = np.random.default_rng(0)
    sample = rng.choice(bins.shape[0], size=min(n_tiles, bins.shape[0]), replace=False)
    tot = dict(pairs=0, cur_visits=0, cur_lanes=0, ring32_it=0, ring32_fl=0, ring16_it=0, ring16_fl=0, ring64_it=0,
               ring64_fl=0, ring8_it=0, ring8_fl=0, valid_pairs=0)
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
        alpha = np.minimum(0.99, opac[g] * np.exp(-sigma))
        inside = (px < W) & (py < H)
        f = np.zeros((16, 16), np.int64)
        y0, x0 = ty * 16, tx * 16
        sub = fidx[y0:y0 + 16, x0:x0 + 16]
        f[:sub.shape[0], :sub.shape[1]] = sub
        k = lo + np.arange(g.size)[None, None, :]
        hit = (sigma >= 0) & (alpha >= 1.0 / 255.0) & inside & (k <= f[:, :, None])  # the adjoint's `valid`
        tot["pairs"] += g.size
        tot["valid_pairs"] += int(hit.sum())

        def units(h, w):  # per-unit survivor counts for units of h x w pixels
            return hit.reshape(16 // h, h, 16 // w, w, -1).any(axis=(1, 3))  # [uy, ux, G]

        u84 = units(4, 8)
        n84 = u84.sum(-1)
        tot["cur_visits"] += int(n84.sum())
        tot["cur_lanes"] += int(hit.sum())
        tot["ring32_it"] += int((n84 + 31 * (n84 > 0)).sum())
        tot["ring32_fl"] += int(n84.sum())
        u44 = units(4, 4).sum(-1)  # [4, 4]
        pair = u44.reshape(4, 2, 2).max(-1)  # two 4x4 halves of one 8x4 warp in lock-step
        tot["ring16_it"] += int((pair + 15 * (pair > 0)).sum())
        tot["ring16_fl"] += int(u44.sum())
        u42 = units(2, 4).sum(-1)  # [8, 4] units of 4 wide x 2 high
        quad = u42.reshape(4, 2, 2, 2).max(axis=(1, 3))
        tot["ring8_it"] += int((quad + 7 * (quad > 0)).sum())
        tot["ring8_fl"] += int(u42.sum())
        u416 = units(4, 16).sum(-1)  # [4, 1]: 16 wide x 4 high, 2 pixels per lane-step
        tot["ring64_it"] += int((u416 + 31 * (u416 > 0)).sum())
        tot["ring64_fl"] += int(u416.sum())
    P = tot["pairs"]
    print(f"{len(sample)} tiles, {P} (tile, Gaussian) pairs, valid (pixel, Gaussian) pairs per tile pair: {tot['valid_pairs'] / P:.1f}")
    print(f"current 8x4 pixel-parallel : visits/pair {tot['cur_visits'] / P:.2f}  valid lanes/visit {tot['cur_lanes'] / tot['cur_visits']:.1f}")
    for name, ev in (("ring32", 1), ("ring16", 1), ("ring8", 1), ("ring64", 2)):
        it, fl = tot[name + "_it"], tot[name + "_fl"]
        print(f"{name:7s}: iterations/pair {it / P:.2f} (x{it / tot['cur_visits']:.2f} of current visits), evals/iteration {ev}x32, "
              f"flushes/pair {fl / P:.2f} (x{fl / tot['cur_visits']:.2f})")


if __name__ == "__main__":
    main()

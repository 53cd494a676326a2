"""CPU-only: how much does tile-list imbalance cost a plain grid launch?  

Projects + bins one view with the CPU oracle, then list-schedules the tiles (cost = list length + a fixed per-tile
overhead) over 148 SMs x 5 resident CTAs in grid order and longest-first, and compares with the balanced time.
usage: python tools/sim_tile_schedule.py [cfg3|cfg2-small]"""
import heapq
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "gaussian-splatting-toolkit_b200")):
    sys.path.insert(0, p)
from oracle import oracle as orc  # noqa: E402  (test infrastructure; this is an analysis tool, not the product)
from rasterizer.synthetic import look_at_viewmat, make_scene, projection_matrix  # noqa: E402


def cfg3_view():
    import test_gpu_train_densify as t

    W = H = 800
    g = t._gaussians(200_000, 100, 0.004, 0.03, 0.55)
    scene = make_scene(200_000, W, H, 0.01, 0.03, seed=0)
    fov = math.radians(50.0)
    fx = 0.5 * W / math.tan(0.5 * fov)
    V = look_at_viewmat(yaw_deg=30, pitch_deg=-15, centre=(0, 0, 4.0))
    P = projection_matrix(0.001, 1000.0, fov, fov).astype(np.float64)
    scene.update(means3d=g["means"].numpy(), scales=np.exp(g["scales"].numpy()),
                 quats=(g["quats"] / g["quats"].norm(dim=-1, keepdim=True)).numpy(),
                 opacities=torch.sigmoid(g["opacities"]).numpy().reshape(-1), viewmat=V,
                 projmat=(P @ V.astype(np.float64)).astype(np.float32), fx=fx, fy=fx, cx=W / 2, cy=H / 2,
                 cam_pos=(-V[:3, :3].T @ V[:3, 3]).astype(np.float32))
    return scene


def makespan(L, order, slots=148 * 5, overhead=300):
    h = [0] * slots
    heapq.heapify(h)
    for i in order:
        heapq.heappush(h, heapq.heappop(h) + int(L[i]) + overhead)
    return max(h)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    orc.build()
    scene = cfg3_view() if which == "cfg3" else make_scene(250_000, 960, 540, 0.004, 0.04, seed=0)
    tb = orc.render_view(scene, backward=False)["tile_bins"]
    L = (tb[:, 1] - tb[:, 0]).astype(np.int64)
    ideal = (L.sum() + 300 * L.size) / (148 * 5)
    nat, lpt = makespan(L, range(L.size)), makespan(L, np.argsort(-L))
    print(f"{which}: {L.size} tiles, M = {L.sum()}, mean {L.mean():.0f} / max {L.max()} per tile")
    print(f"makespan / balanced: grid order {nat / ideal:.2f}, longest-first {lpt / ideal:.2f} (longest tile alone: {(L.max() + 300) / ideal:.2f})")


if __name__ == "__main__":
    main()

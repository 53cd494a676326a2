"""Generates tests/golden/refcuda_*.npz ON THE GPU BOX by running the compiled, UNMODIFIED reference CUDA
extension (oracle/_ref/rasterizer_ref_cuda.so, built from /root/reference by oracle/build_ref.py with the
reference's packaged flags) on small seeded scenes:
    gpurun -- python tests/golden/gen_golden_ref_cuda.py      # writes gpurun_out/golden/*.npz
then copy gpurun_out/golden/*.npz to tests/golden/ and commit.  Inputs and reference outputs are stored
together; forward AND backward (all six parameter gradients) are recorded."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "gaussian-splatting-toolkit_b200"), os.path.dirname(HERE)):
    sys.path.insert(0, p)


def main():
    from oracle.build_ref import load_ref
    from pipelines import run_view_bindings
    from rasterizer.synthetic import look_at_viewmat, make_scene, scene_to_torch

    ref = load_ref()
    assert ref is not None, "oracle/_ref/rasterizer_ref_cuda.so missing"
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    cases = {
        # ragged image, clipped / culled Gaussians (margin > 1), SH degree 3
        "a_2k_144x100": make_scene(2000, 144, 100, 0.02, 0.25, margin=1.2, seed=21),
        # rotated camera, block width 8, 2 of 3 SH degrees, many opaque Gaussians (early termination)
        "b_1500_96x96_bw8": make_scene(1500, 96, 96, 0.05, 0.4, margin=0.9, seed=22, block_width=8, degrees_to_use=2,
                                       viewmat=look_at_viewmat(yaw_deg=15.0, pitch_deg=-8.0, shift=(0.1, -0.05, 0.3))),
    }
    for name, scene in cases.items():
        s = scene_to_torch(scene, "cuda")
        out = run_view_bindings(ref, s, backward=True, sort_impl="torch")
        torch.cuda.synchronize()
        blob = {("in_" + k): v for k, v in scene.items()}
        for k, v in out.items():
            blob["ref_" + k] = v.cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
        np.savez_compressed(os.path.join(out_dir, f"refcuda_{name}.npz"), **blob)
        print(name, "M =", out["num_intersects"], "visible =", int((out["radii"] > 0).sum()))


if __name__ == "__main__":
    main()

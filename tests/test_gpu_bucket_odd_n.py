"""the flat view-parallel gradient bucket must work for ANY Gaussian count (after densification N is odd as
often as not): the backward kernels write straight into its segments, and the projection adjoint requires 16-byte aligned
output rows."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1001, 33_149])
def test_backward_kernels_write_into_bucket_segments_for_odd_n(n):
    from pipelines import run_view_bindings
    from rasterizer import cuda as C
    from rasterizer.synthetic import make_scene, scene_to_torch
    from rasterizer.view_parallel import GradientBucket

    s = scene_to_torch(make_scene(n, 160, 120, 0.02, 0.15, margin=1.0, seed=n), "cuda")
    ref = run_view_bindings(C, s, sort_impl="gsr", binning="fast")
    bk = GradientBucket(n, s["sh_coeffs"].shape[1], device="cuda")
    H, W, bw = s["img_height"], s["img_width"], s["block_width"]
    opac = s["opacities"].reshape(-1, 1).contiguous()
    v_xy, v_conic, v_colors, v_opacity = C.rasterize_backward(
        H, W, bw, ref["gaussian_ids_sorted"], ref["tile_bins"], ref["xys"], ref["conics"], ref["colors"], opac, s["background"],
        ref["final_Ts"], ref["final_idx"], s["v_out_img"], s["v_out_alpha"], out_opacity=bk["v_opacity"])
    zeros_n = torch.zeros(n, device="cuda")
    C.project_gaussians_backward(n, s["means3d"], s["scales"], 1.0, s["quats"], s["viewmat"], s["projmat"], s["fx"], s["fy"],
                                 s["cx"], s["cy"], H, W, ref["cov3d"], ref["radii"], ref["conics"], ref["compensation"], v_xy,
                                 zeros_n, v_conic, zeros_n, out_mean3d=bk["v_mean3d"], out_scale=bk["v_scale"],
                                 out_quat=bk["v_quat"], need_cov_grads=False)
    torch.cuda.synchronize()
    for name in ("v_mean3d", "v_scale", "v_quat"):
        a, b = bk[name], ref[name]
        assert float((a - b).norm() / b.norm()) < 5e-6, name   # same kernels, atomics order only
    assert float((bk["v_opacity"] - ref["v_opacity"].reshape(-1, 1)).norm() / ref["v_opacity"].norm()) < 5e-6

"""Pins the CPU oracle (oracle/gsr_oracle.c) against golden vectors produced by the reference itself:
  * tests/golden/torch_impl_*.npz — the reference's pure-PyTorch restatement (rasterizer/_torch_impl.py),
    forward outputs, generated in the build container by tests/golden/gen_golden_torch_impl.py;
  * tests/golden/refcuda_*.npz — the compiled, unmodified reference CUDA extension run on a B200
    (tests/golden/gen_golden_ref_cuda.py), forward AND backward outputs.
CPU only."""
import glob
import os

import numpy as np
import pytest

from parity import assert_float_parity, assert_int_equal

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _scene_from_npz(z):
    return {k[3:]: (z[k] if z[k].ndim else z[k].item()) for k in z.files if k.startswith("in_")}


TORCH_IMPL = sorted(glob.glob(os.path.join(GOLD, "torch_impl_*.npz")) + glob.glob(os.path.join(GOLD, "oracle_pin_torch_impl_*.npz")))
REFCUDA = sorted(p for p in glob.glob(os.path.join(GOLD, "refcuda_*.npz")) if "modelstep" not in p)
MODELSTEP = sorted(glob.glob(os.path.join(GOLD, "refcuda_modelstep_*.npz")))


def test_golden_fixtures_present():
    assert len(TORCH_IMPL) >= 2, "torch_impl golden fixtures missing"


@pytest.mark.parametrize("path", TORCH_IMPL, ids=[os.path.basename(p) for p in TORCH_IMPL])
def test_oracle_vs_reference_torch_impl(oracle, path):
    z = np.load(path)
    s = _scene_from_npz(z)
    out = oracle.render_view(s, backward=False)
    # float outputs of SH / projection / blend
    for k in ("rgb_sh", "colors", "cov3d", "xys", "depths", "compensation"):
        assert_float_parity(out[k], z["ref_" + k].reshape(out[k].shape), k)
    # image / transmittance: pixels where a threshold decision (alpha vs 1/255, T vs 1e-4) sits within rounding of its
    # threshold are flagged by the oracle and excluded (two correct FP32 evaluations may decide differently there)
    clean = out["ambiguous"] == 0
    assert clean.mean() > 0.98
    assert_float_parity(out["out_img"], z["ref_out_img"], "out_img", mask=np.broadcast_to(clean[..., None], out["out_img"].shape))
    assert_float_parity(out["final_Ts"], z["ref_final_Ts"].reshape(out["final_Ts"].shape), "final_Ts", mask=clean)
    # conics: off-diagonal terms pass through zero -> absolute tolerance on the scale of the diagonal
    assert_float_parity(out["conics"], z["ref_conics"], "conics", atol=1e-4 * float(np.abs(z["ref_conics"]).max()))
    # integer / index outputs are exact
    for k in ("radii", "num_tiles_hit", "cum_tiles_hit", "gaussian_ids", "gaussian_ids_sorted", "tile_bins"):
        assert_int_equal(out[k], z["ref_" + k].reshape(out[k].shape), k)
    # keys: tile part exact; depth part = float bits, compared as floats (the torch restatement rounds
    # p_view.z differently in the last ulp)
    for k in ("isect_ids", "isect_ids_sorted"):
        a, b = out[k], z["ref_" + k]
        assert_int_equal(a >> 32, b >> 32, k + ".tile")
        da = (a & 0xFFFFFFFF).astype(np.uint32).view(np.float32)
        db = (b & 0xFFFFFFFF).astype(np.uint32).view(np.float32)
        assert_float_parity(da, db, k + ".depth")


@pytest.mark.parametrize("path", REFCUDA, ids=[os.path.basename(p) for p in REFCUDA])
def test_oracle_vs_reference_cuda_ext(oracle, path):
    """Forward + backward of the oracle against what the reference CUDA extension produced on a B200."""
    z = np.load(path)
    s = _scene_from_npz(z)
    out = oracle.render_view(s, s["v_out_img"], s["v_out_alpha"], backward=True)
    vis = z["ref_radii"] > 0
    assert_int_equal(out["radii"], z["ref_radii"], "radii", max_frac_bad=1e-4)
    assert_int_equal(out["num_tiles_hit"], z["ref_num_tiles_hit"], "num_tiles_hit", max_frac_bad=1e-4)
    both = vis & (out["radii"] > 0)
    for k in ("xys", "depths", "compensation", "cov3d"):
        assert_float_parity(out[k], z["ref_" + k], k, mask=both)
    assert_float_parity(out["conics"], z["ref_conics"], "conics", mask=both,
                        atol=1e-4 * float(np.abs(z["ref_conics"][both]).max()))
    assert_float_parity(out["colors"], z["ref_colors"], "colors")
    clean = out["ambiguous"] == 0
    assert clean.mean() > 0.99
    assert_float_parity(out["out_img"], z["ref_out_img"], "out_img", mask=np.broadcast_to(clean[..., None], out["out_img"].shape))
    assert_float_parity(out["final_Ts"], z["ref_final_Ts"], "final_Ts", mask=clean)
    assert_int_equal(out["final_idx"], z["ref_final_idx"], "final_idx", mask=clean, max_frac_bad=1e-4)
    # gradients: the reference sums with order-nondeterministic FP32 atomics
    for k in ("v_xy", "v_conic", "v_colors", "v_opacity", "v_coeffs", "v_mean3d", "v_scale", "v_quat"):
        assert_float_parity(out[k], z["ref_" + k].reshape(out[k].shape), k, max_norm_rel=2e-4, max_frac_bad=2e-3)


@pytest.mark.parametrize("path", MODELSTEP, ids=[os.path.basename(p) for p in MODELSTEP])
def test_oracle_model_step_vs_reference_cuda_ext(oracle, path):
    """The oracle's restatement of the model-level view (raw parameters -> rgb / depth / alpha) and of
    the six raw-parameter gradients against the reference extension + torch autograd glue on a B200."""
    z = np.load(path)
    s = _scene_from_npz(z)
    raw = {k[4:]: z[k] for k in z.files if k.startswith("raw_")}
    out = oracle.render_fused_reference(s, raw, z["up_v_rgb"], z["up_v_depth"], z["up_v_alpha"])
    clean = out["ambiguous"] == 0
    assert_float_parity(out["rgb"], z["ref_rgb"], "rgb", mask=np.broadcast_to(clean[..., None], out["rgb"].shape))
    assert_float_parity(out["depth"], z["ref_depth"], "depth", mask=clean)
    assert_float_parity(out["alpha"], z["ref_alpha"], "alpha", mask=clean, atol=1e-6)
    assert_int_equal(out["radii"], z["ref_radii"], "radii", max_frac_bad=1e-3)
    for k in ("v_means3d", "v_scales_raw", "v_quats_raw", "v_opacities_raw", "v_features_dc", "v_features_rest"):
        assert_float_parity(out[k], z["ref_" + k].reshape(out[k].shape), k, max_norm_rel=2e-4, max_frac_bad=2e-3)

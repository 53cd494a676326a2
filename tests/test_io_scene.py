"""CPU tests of the on-disk formats of SURVEY 8(f4) against fixtures produced by the reference's own code
(tests/golden/gen_golden_io.py): transforms.json parsing vs the reference dataparser, camera -> (viewmat, projmat) vs
the model's prologue, checkpoint reading vs a reference-format checkpoint, and — when the reference tree is present
(build container) — a checkpoint WRITTEN here loaded back by the reference model and torch.optim.Adam."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GROUPS = ("means", "scales", "quats", "features_dc", "features_rest", "opacities")


def test_transforms_json_vs_reference_dataparser():
    from rasterizer.io_scene import load_transforms

    z = np.load(os.path.join(GOLD, "io_transforms_ref.npz"))
    for split in ("train", "val"):
        d = load_transforms(os.path.join(GOLD, "io_transforms.json"), split=split)
        assert [os.path.basename(f) for f in d["image_filenames"]] == list(z[f"{split}_files"])
        np.testing.assert_allclose(d["camera_to_worlds"], z[f"{split}_c2w"], rtol=0, atol=2e-6)
        for k in ("fx", "fy", "cx", "cy"):
            np.testing.assert_array_equal(d[k], z[f"{split}_{k}"].astype(np.float32))
        np.testing.assert_array_equal(d["height"], z[f"{split}_height"])
        np.testing.assert_array_equal(d["width"], z[f"{split}_width"])
        np.testing.assert_allclose(d["transform"], z[f"{split}_transform"], rtol=0, atol=2e-6)
        assert d["scale_factor"] == float(z[f"{split}_scale"])


def test_camera_to_view_proj_vs_reference_model_prologue():
    from rasterizer.io_scene import camera_to_view_proj, load_transforms

    d = load_transforms(os.path.join(GOLD, "io_transforms.json"), split="train")
    z = np.load(os.path.join(GOLD, "io_cameras_ref.npz"))
    ref = np.load(os.path.join(GOLD, "io_transforms_ref.npz"))
    for i in range(len(d["image_filenames"])):
        V, PM, pos = camera_to_view_proj(ref["train_c2w"][i], float(d["fx"][i]), float(d["fy"][i]), int(d["width"][i]),
                                         int(d["height"][i]))
        np.testing.assert_allclose(V, z["viewmat"][i], rtol=0, atol=1e-6)
        np.testing.assert_allclose(PM, z["projmat"][i], rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(pos, ref["train_c2w"][i][:, 3], rtol=0, atol=0)


def test_split_and_errors(tmp_path):
    import json

    from rasterizer.io_scene import load_transforms, train_eval_split_fraction

    tr, ev = train_eval_split_fraction(20, 0.9)
    assert len(tr) == 18 and len(ev) == 2 and tr[0] == 0 and tr[-1] == 19
    meta = json.load(open(os.path.join(GOLD, "io_transforms.json")))
    meta["train_filenames"] = [meta["frames"][0]["file_path"], meta["frames"][3]["file_path"]]
    p = tmp_path / "transforms.json"
    p.write_text(json.dumps(meta))
    d = load_transforms(str(tmp_path), split="train")
    assert len(d["image_filenames"]) == 2
    with pytest.raises(RuntimeError, match="is missing"):
        load_transforms(str(tmp_path), split="val")
    meta["train_filenames"] = ["images/nope.png"]
    p.write_text(json.dumps(meta))
    with pytest.raises(RuntimeError, match="were not found"):
        load_transforms(str(tmp_path), split="train")
    with pytest.raises(AssertionError):
        load_transforms(str(tmp_path / "absent"))


def test_load_reference_format_checkpoint():
    from rasterizer.io_scene import load_checkpoint

    ck = load_checkpoint(os.path.join(GOLD, "io_ref_step-000000123.ckpt"))
    raw = torch.load(os.path.join(GOLD, "io_ref_step-000000123.ckpt"), map_location="cpu", weights_only=False)
    assert ck["step"] == 123 and set(ck["params"]) == set(GROUPS)
    for k in GROUPS:
        assert torch.equal(ck["params"][k], raw["pipeline"]["_model.gauss_params." + k])
        st = ck["optimizers"][k]["state"][0]
        assert st["exp_avg"].shape == ck["params"][k].shape and float(st["step"]) == 2.0
    assert "_model.device_indicator_param" in ck["extra"]
    # DDP-style and pre-gauss_params names
    old = {"step": 7, "pipeline": {"module._model." + k: v for k, v in ck["params"].items()}, "optimizers": {}}
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        torch.save(old, os.path.join(d, "step-000000007.ckpt"))
        ck2 = load_checkpoint(d)
        assert ck2["step"] == 7 and all(torch.equal(ck2["params"][k], ck["params"][k]) for k in GROUPS)


class _FakeOptim:
    """state_dict() of rasterizer.optim.GaussianOptimizers without the GPU (same layout)."""

    def __init__(self, params):
        self.sd = {k: {"state": {0: {"step": torch.tensor(5.0), "exp_avg": torch.randn_like(v), "exp_avg_sq": torch.rand_like(v)}},
                       "param_groups": [{"lr": 1e-3, "betas": (0.9, 0.999), "eps": 1e-15, "weight_decay": 0, "amsgrad": False,
                                         "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                                         "fused": None, "initial_lr": 1e-3, "params": [0]}]} for k, v in params.items()}

    def state_dict(self):
        return self.sd


def test_checkpoint_round_trip_and_reference_loads_it(tmp_path):
    from rasterizer.io_scene import load_checkpoint, save_checkpoint

    g = torch.Generator().manual_seed(0)
    shapes = {"means": (33, 3), "scales": (33, 3), "quats": (33, 4), "features_dc": (33, 3), "features_rest": (33, 15, 3), "opacities": (33, 1)}
    params = {k: torch.randn(s, generator=g) for k, s in shapes.items()}
    opt = _FakeOptim(params)
    (tmp_path / "stale.ckpt").write_text("x")
    path = save_checkpoint(str(tmp_path), 2000, params, opt)
    assert os.path.basename(path) == "step-000002000.ckpt" and os.listdir(tmp_path) == ["step-000002000.ckpt"]
    ck = load_checkpoint(str(tmp_path))
    assert ck["step"] == 2000
    for k in GROUPS:
        assert torch.equal(ck["params"][k], params[k])
        assert torch.equal(ck["optimizers"][k]["state"][0]["exp_avg"], opt.sd[k]["state"][0]["exp_avg"])
    if not os.path.isdir("/root/reference/gs_toolkit"):
        pytest.skip("reference tree not present (GPU box)")
    import sys

    sys.path.insert(0, os.path.join(GOLD))
    import gen_golden_densify as gd

    raw = torch.load(path, map_location="cpu", weights_only=False)
    model = gd.make_model({k: torch.zeros((1,) + s[1:]) for k, s in shapes.items()}, {}, 0, 1)
    state = {k[len("_model."):]: v for k, v in raw["pipeline"].items()}
    state["device_indicator_param"] = torch.empty(0)
    model.load_state_dict(state)          # vanilla_gs.py:236-258 resizes to the stored count
    for k in GROUPS:
        assert torch.equal(model.gauss_params[k].detach(), params[k])
        adam = torch.optim.Adam([model.gauss_params[k]], lr=1.0, eps=1e-15)
        adam.load_state_dict(raw["optimizers"][k])      # engine/optimizers.py:197-204
        assert torch.equal(adam.state[model.gauss_params[k]]["exp_avg_sq"], opt.sd[k]["state"][0]["exp_avg_sq"])
        assert adam.param_groups[0]["lr"] == 1e-3


def test_exponential_decay_lr_vs_reference_scheduler():
    """rasterizer.optim.exponential_decay_lr against the learning rates the reference's ExponentialDecayScheduler +
    LambdaLR produced (tests/golden/io_scheduler_ref.npz, gen_golden_io.py): after the k-th scheduler step lr = f(k)."""
    from rasterizer.optim import default_means_scheduler, exponential_decay_lr

    z = np.load(os.path.join(GOLD, "io_scheduler_ref.npz"))
    fns = {"means": default_means_scheduler(),
           "warm_cos": exponential_decay_lr(1e-3, 1e-4, 500, warmup_steps=100),
           "warm_lin": exponential_decay_lr(1e-3, None, 400, warmup_steps=50, ramp="linear")}
    for tag, fn in fns.items():
        got = np.array([fn(k) for k in range(1, len(z[tag]) + 1)])
        np.testing.assert_allclose(got, z[tag], rtol=1e-12, atol=0, err_msg=tag)


def test_orientation_and_centring_methods_vs_reference_camera_utils():
    """f4 remainder: "pca" / "vertical" orientation and "focus" centring (camera_utils.py:496-660) against outputs of the
    reference's own function (tests/golden/gen_golden_orient.py), all 12 combinations on two pose sets."""
    import os

    import numpy as np

    from rasterizer.io_scene import auto_orient_and_center_poses

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io_orient_ref.npz"))
    for name in ("ring", "array"):
        poses = z[f"{name}_poses"]
        for method in ("pca", "up", "vertical", "none"):
            for center in ("poses", "focus", "none"):
                o, t = auto_orient_and_center_poses(poses.copy(), method=method, center_method=center)
                ro, rt = z[f"{name}_{method}_{center}_oriented"], z[f"{name}_{method}_{center}_transform"]
                assert o.shape == ro.shape and t.shape == rt.shape
                # LAPACK eigenvectors / singular vectors are defined up to sign; the reference fixes the handedness and
                # the sign of the mean up direction afterwards, so the results must agree outright
                assert np.abs(t - rt).max() < 2e-5, (name, method, center, np.abs(t - rt).max())
                assert np.abs(o - ro).max() < 5e-5, (name, method, center, np.abs(o - ro).max())


def test_image_decoding_vs_reference_dataset(tmp_path):
    """load_image_uint8 / load_image against what the reference's InputDataset.get_numpy_image / get_image make of the same
    PNG files (tests/golden/images/*.png, fixtures of gen_golden_images.py): RGB, RGBA and grey, scale factors 1 and 0.5,
    no alpha colour / white / an arbitrary one.  Bit-exact: same Pillow calls, same float32 arithmetic."""
    pytest.importorskip("PIL")
    import torch
    from rasterizer.io_scene import load_depth_image, load_image, load_image_uint8

    z = np.load(os.path.join(GOLD, "io_images_ref.npz"))
    names = [str(n) for n in z["names"]]
    colors = {"none": None, "1.0_1.0_1.0": torch.tensor([1.0, 1.0, 1.0]), "0.2_0.5_0.9": torch.tensor([0.2, 0.5, 0.9])}
    checked = 0
    for scale in (1.0, 0.5):
        for ctag, color in colors.items():
            for n in names:
                path = os.path.join(GOLD, "images", n)
                assert np.array_equal(load_image_uint8(path, scale), z[f"s{scale}_a{ctag}_{n}_u8"]), (scale, ctag, n)
                got = load_image(path, scale, color).numpy()
                ref = z[f"s{scale}_a{ctag}_{n}_f32"]
                assert got.shape == ref.shape and got.dtype == np.float32 and np.array_equal(got, ref), (scale, ctag, n)
                checked += 1
    assert checked == 18
    # depth maps: 16-bit PNG in millimetres, 8-bit monocular, .npy
    from PIL import Image

    d16 = (np.arange(6 * 8, dtype=np.uint16).reshape(6, 8) * 531).astype(np.uint16)
    Image.fromarray(d16).save(str(tmp_path / "d.png"))
    assert np.array_equal(load_depth_image(str(tmp_path / "d.png")).numpy(), d16.astype("float32") / 1000.0)
    np.save(str(tmp_path / "d.npy"), d16)
    assert np.array_equal(load_depth_image(str(tmp_path / "d.npy"), mono_depth=True).numpy(), d16.astype("float32") / 255.0)
    with pytest.raises(ValueError):
        load_depth_image(str(tmp_path / "d.exr"))


def test_load_views_from_a_dataset_directory(tmp_path):
    """transforms.json + image files on disk -> per-view cameras and decoded images (the inputs of a training view)."""
    pytest.importorskip("PIL")
    import json
    import shutil

    from rasterizer.io_scene import camera_to_view_proj, load_transforms, load_views

    meta = json.load(open(os.path.join(GOLD, "io_transforms.json")))
    meta["w"], meta["h"], meta["cx"], meta["cy"] = 14, 10, 7.0, 5.0
    os.makedirs(tmp_path / "images")
    for k, fr in enumerate(meta["frames"]):
        shutil.copy(os.path.join(GOLD, "images", ("rgb.png", "rgba.png", "grey.png")[k % 3]), tmp_path / fr["file_path"])
    json.dump(meta, open(tmp_path / "transforms.json", "w"))
    t = load_transforms(str(tmp_path), split="train")
    views = load_views(str(tmp_path), split="train", alpha_color=torch.tensor([1.0, 1.0, 1.0]))
    assert len(views) == len(t["image_filenames"]) > 0
    for i, v in enumerate(views):
        assert v["image"].shape == (10, 14, 3) and v["image"].dtype == torch.float32
        assert 0.0 <= float(v["image"].min()) and float(v["image"].max()) <= 1.0
        vm, pm, pos = camera_to_view_proj(t["camera_to_worlds"][i], float(t["fx"][i]), float(t["fy"][i]), 14, 10)
        assert np.array_equal(v["viewmat"], vm) and np.array_equal(v["projmat"], pm) and np.array_equal(v["cam_pos"], pos)
    half = load_views(str(tmp_path), split="val", image_scale_factor=0.5, alpha_color=torch.tensor([0.0, 0.0, 0.0]))
    assert half[0]["image"].shape == (5, 7, 3) and half[0]["height"] == 5 and half[0]["width"] == 7
    assert half[0]["fx"] == pytest.approx(0.5 * float(load_transforms(str(tmp_path), split="val")["fx"][0]))
    assert "image" not in load_views(str(tmp_path), load_images=False)[0]

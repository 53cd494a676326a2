"""Generates tests/golden/images/*.png + io_images_ref.npz with the REFERENCE's own dataset code (build container only):
three tiny PNGs (RGB, RGBA, 8-bit grey) and what gs_toolkit/data/datasets/base_dataset.py:48-87 makes of them —
`InputDataset.get_numpy_image` / `get_image` at scale factors 1 and 0.5, with and without an alpha colour.  The methods are
called unbound on a stand-in object that carries exactly the attributes they read.  Missing third-party modules are
stubbed for the import only (see gen_golden_densify.py)."""
import os
import sys
import types

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden_densify as gd  # noqa: E402  (installs the stub finder)

gd._StubFinder.ROOTS = gd._StubFinder.ROOTS + ("OpenEXR", "Imath")
from gs_toolkit.data.datasets.base_dataset import InputDataset  # noqa: E402


def main():
    d = os.path.join(HERE, "images")
    os.makedirs(d, exist_ok=True)
    g = np.random.default_rng(11)
    h, w = 10, 14
    files = {
        "rgb.png": Image.fromarray(g.integers(0, 256, (h, w, 3), dtype=np.uint8), "RGB"),
        "rgba.png": Image.fromarray(g.integers(0, 256, (h, w, 4), dtype=np.uint8), "RGBA"),
        "grey.png": Image.fromarray(g.integers(0, 256, (h, w), dtype=np.uint8), "L"),
    }
    for name, im in files.items():
        im.save(os.path.join(d, name))
    names = sorted(files)
    out = {"names": np.array(names)}
    for scale in (1.0, 0.5):
        for alpha_color in (None, torch.tensor([1.0, 1.0, 1.0]), torch.tensor([0.2, 0.5, 0.9])):
            fake = types.SimpleNamespace()
            fake._dataparser_outputs = types.SimpleNamespace(image_filenames=[os.path.join(d, n) for n in names],
                                                             alpha_color=alpha_color)
            fake.scale_factor = scale
            fake.get_numpy_image = lambda i, fake=fake: InputDataset.get_numpy_image(fake, i)
            tag = f"s{scale}_a{'none' if alpha_color is None else '_'.join(f'{v:.1f}' for v in alpha_color.tolist())}"
            for i, n in enumerate(names):
                out[f"{tag}_{n}_u8"] = InputDataset.get_numpy_image(fake, i)
                out[f"{tag}_{n}_f32"] = InputDataset.get_image(fake, i).numpy()
    np.savez_compressed(os.path.join(HERE, "io_images_ref.npz"), **out)
    print({k: v.shape for k, v in out.items() if k != "names"})


if __name__ == "__main__":
    main()

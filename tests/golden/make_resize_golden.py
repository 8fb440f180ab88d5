"""Generate tests/golden/resize_golden.npz: outputs of the reference's image preparation, frozen.

Run once in the build container:  python tests/golden/make_resize_golden.py

The reference prepares images with PIL (``plip.py:35`` through the CLIP processor; torchvision ``Resize(224, BICUBIC)``
+ ``CenterCrop(224)`` at ``reproducibility/embedders/transform.py:45-52``).  For seeded synthetic images this stores,
per case, the SHA-256 of the 224x224x3 uint8 tile those pipelines produce plus its top-left 24x24 patch:
  * ``round`` cases: ``torchvision.transforms`` on the PIL image (the embedders' ``_transform``);
  * ``floor`` cases: ``transformers.CLIPImageProcessorPil`` (PIL backend; resize + center_crop only).
Inputs are regenerated from the seed at test time (numpy Generator), so only outputs are stored.  Versions at
generation time are recorded in the file."""
import hashlib
import os
import sys

import numpy as np
import PIL
import PIL.Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = [  # (h, w, seed, kind, crop)
    (256, 256, 1, "noise", "floor"), (300, 500, 2, "stripes", "floor"), (1000, 777, 3, "noise", "round"),
    (96, 96, 4, "noise", "round"), (227, 224, 5, "stripes", "round"), (1536, 2048, 6, "noise", "floor"),
    (333, 1999, 7, "stripes", "floor"), (231, 229, 8, "noise", "round"),
]


def make_image(h, w, seed, kind):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    base = ((xx * 7 + yy * 3 + seed) % 256).astype(np.uint8)
    return np.stack([base, 255 - base, ((xx // 8 + yy // 8) % 2 * 255).astype(np.uint8)], axis=-1)


def main():
    import torchvision
    import torchvision.transforms as T
    import transformers
    from transformers import CLIPImageProcessorPil
    tv = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224)])
    hf = CLIPImageProcessorPil()
    out = {"cases": np.array([(h, w, s) for h, w, s, _, _ in CASES], dtype=np.int64),
           "kinds": np.array([k for *_, k, _ in CASES]), "crops": np.array([c for *_, c in CASES]),
           "versions": np.array([f"Pillow {PIL.__version__}", f"torchvision {torchvision.__version__}",
                                 f"transformers {transformers.__version__}"])}
    shas, patches = [], []
    for h, w, seed, kind, crop in CASES:
        img = PIL.Image.fromarray(make_image(h, w, seed, kind))
        if crop == "round":
            tile = np.asarray(tv(img))
        else:
            tile = np.transpose(hf(images=[img], return_tensors="np", do_normalize=False, do_rescale=False)["pixel_values"][0],
                                (1, 2, 0))
        assert tile.shape == (224, 224, 3) and tile.dtype == np.uint8
        shas.append(hashlib.sha256(np.ascontiguousarray(tile).tobytes()).hexdigest())
        patches.append(tile[:24, :24].copy())
    out["sha256"] = np.array(shas)
    out["patches"] = np.stack(patches)
    path = os.path.join(ROOT, "tests", "golden", "resize_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

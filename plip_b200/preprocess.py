"""Host-side image preparation: PIL / path -> 224x224 RGB uint8 tiles.

The arithmetic part of ``CLIPImageProcessor`` (rescale 1/255, normalise by the CLIP mean/std;
TF:models/clip/image_processing_clip.py:50-62, ``reproducibility/embedders/transform.py:45-52``) is fused
into the device im2col kernel (``PLIP_PIX_U8_NHWC``), so the host only has to deliver uint8 tiles:
shortest-edge-224 bicubic resize + centre crop for images that are not already 224x224, which is what
the reference's processor does with PIL before its float conversion.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Union

import numpy as np
import PIL.Image

SIZE = 224
ImageLike = Union[str, PIL.Image.Image, np.ndarray]


def load_rgb(img: ImageLike) -> PIL.Image.Image:
    """Path / PIL image / HxWx3 uint8 array -> RGB PIL image (``Image().decode_example`` + convert_rgb)."""
    if isinstance(img, str):
        img = PIL.Image.open(img)
    elif isinstance(img, np.ndarray):
        img = PIL.Image.fromarray(img)
    if img.mode != "RGB":
        img = img.convert("RGB")
    return img


def resize_center_crop(img: PIL.Image.Image, size: int = SIZE) -> PIL.Image.Image:
    """Shortest edge -> ``size`` (bicubic, aspect preserved, long edge ``int(size * long / short)``),
    then centre crop ``size x size`` (offsets ``(dim - size) // 2``), as CLIPImageProcessor does."""
    w, h = img.size
    if (w, h) == (size, size):
        return img
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    nw, nh = (new_short, new_long) if w <= h else (new_long, new_short)
    if (nw, nh) != (w, h):
        img = img.resize((nw, nh), resample=PIL.Image.BICUBIC)
    left, top = (nw - size) // 2, (nh - size) // 2
    return img.crop((left, top, left + size, top + size))


def _one_tile(im: ImageLike) -> np.ndarray:
    return np.asarray(resize_center_crop(load_rgb(im)))


def to_uint8_tiles(images: Sequence[ImageLike], workers: int = 0) -> np.ndarray:
    """Batch of images -> contiguous uint8 array ``[n,224,224,3]`` (NHWC).

    ``workers > 1`` decodes / resizes in a thread pool (PIL releases the GIL in its C loops) — the host-side
    counterpart of the reference's ``DataLoader(num_workers=…)`` (``embedders/plip.py:39``)."""
    out = np.empty((len(images), SIZE, SIZE, 3), dtype=np.uint8)
    if workers > 1 and len(images) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=workers) as ex:
            for i, tile in enumerate(ex.map(_one_tile, images)):
                out[i] = tile
        return out
    for i, im in enumerate(images):
        out[i] = _one_tile(im)
    return out


def chunks(seq: Sequence, n: int) -> Iterable[Sequence]:
    for i in range(0, len(seq), n):
        yield seq[i:i + n]

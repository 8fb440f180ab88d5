"""TEST INFRASTRUCTURE — CPU restatement of the reference's image resize (bicubic, antialiased, 8 bit).

The reference hands PIL images to ``CLIPProcessor`` (``/root/reference/plip.py:35``) or to torchvision's
``Resize(n_px, interpolation=BICUBIC)`` + ``CenterCrop`` (``reproducibility/embedders/transform.py:45-52``); both
end in ``PIL.Image.resize(..., BICUBIC)``.  That algorithm lives in a third-party dependency (Pillow, 12.2.0 in
this image; un-pinned by the reference): a separable two-pass convolution, horizontal then vertical, with

* per-output-pixel windows ``[xmin, xmin+n)`` around ``center = (xx + 0.5) * scale`` of half-width
  ``support = 2 * max(scale, 1)``, weights ``bicubic((x + xmin - center + 0.5) / max(scale, 1))`` (a = -0.5),
  normalised to sum 1 in double precision;
* weights quantised to 22-bit fixed point (round half away from zero), accumulation in int32 starting from
  ``1 << 21``, arithmetic shift by 22 and a clamp to ``[0,255]``; the intermediate image between the two passes
  is uint8.

Pinned bit-exactly against ``PIL.Image.resize`` itself in ``tests/test_resize.py`` (PIL is in the image here and
on the GPU box).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import ``oracle``.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2
BICUBIC_SUPPORT = 2.0


def bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coefficients(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Fixed-point filter bank for one axis: ``(xmin[out], count[out], k[out, ksize] int32)``."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = BICUBIC_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmins = np.zeros(out_size, np.int32)
    counts = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        xmins[xx], counts[xx] = xmin, xmax
    return xmins, counts, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bicubic_u8(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """``[h,w,c] uint8 -> [new_h,new_w,c] uint8``, equal to ``PIL.Image.resize((new_w,new_h), BICUBIC)``."""
    h, w, c = img.shape
    xm, xc, kh = coefficients(w, new_w)
    ym, yc, kv = coefficients(h, new_h)
    src = img.astype(np.int32)
    tmp = np.empty((h, new_w, c), np.uint8)
    for xx in range(new_w):
        n = xc[xx]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[:, xm[xx]:xm[xx] + n, :], kh[xx, :n], axes=([1], [0]))
        tmp[:, xx, :] = _clip8(acc)
    tmp32 = tmp.astype(np.int32)
    out = np.empty((new_h, new_w, c), np.uint8)
    for yy in range(new_h):
        n = yc[yy]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kv[yy, :n], tmp32[ym[yy]:ym[yy] + n], axes=([0], [0]))
        out[yy] = _clip8(acc)
    return out


def resize_crop_u8(img: np.ndarray, new_w: int, new_h: int, left: int, top: int, size: int = 224) -> np.ndarray:
    """Resize to ``(new_w,new_h)`` then crop ``size x size`` at ``(left, top)``."""
    return resize_bicubic_u8(img, new_w, new_h)[top:top + size, left:left + size]

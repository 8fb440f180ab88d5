"""TEST INFRASTRUCTURE — seeded synthetic inputs of the BASELINE.json configs.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` may import ``oracle``.
"""
from __future__ import annotations

import numpy as np
import torch

from .weights import BOS, EOS


def tiles_u8(n: int, seed: int = 0) -> np.ndarray:
    """cfg1: n synthetic 224x224 RGB tiles, uint8 [n,224,224,3]."""
    return np.random.default_rng(seed).integers(0, 256, (n, 224, 224, 3), dtype=np.uint8)


def pixel_values(n: int, seed: int = 1234) -> torch.Tensor:
    """cfg2: normalised pixels (U[0,1) - mean) / std, fp32 [n,3,224,224]."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, 224, 224, generator=g)
    mean = torch.tensor((0.48145466, 0.4578275, 0.40821073)).view(1, 3, 1, 1)
    std = torch.tensor((0.26862954, 0.26130258, 0.27577711)).view(1, 3, 1, 1)
    return (x - mean) / std


def token_ids(n: int, seed: int = 1235, full_length: bool = False, min_len: int = 8):
    """cfg3: random caption ids [n,77] int64 with bos at 0, first eos at len-1, eos padding after it
    (what the CLIP tokenizer emits), plus the matching attention_mask (1 up to and incl. the eos)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, BOS, (n, 77), generator=g)
    if full_length:
        lens = torch.full((n,), 77)
    else:
        lens = torch.randint(min_len, 78, (n,), generator=g)
    ids[:, 0] = BOS
    ar = torch.arange(77)[None]
    ids = torch.where(ar >= (lens[:, None] - 1), torch.full_like(ids, EOS), ids)
    mask = (ar < lens[:, None]).to(torch.int64)
    return ids, mask

"""TEST / BASELINE INFRASTRUCTURE — the reference's own CPU path, timed by ``bench.py``'s CPU legs only.

The reference's arithmetic is the third-party ``transformers`` package (``plip.py:7,26``: un-pinned by the
reference; 5.5.0 in this image, and therefore on the GPU box).  ``/root/reference`` itself does not travel to the
GPU box, so this module drives the LIVE ``transformers.CLIPModel`` (fp32, host cores) through a restatement of
the reference's 118-line wrapper:

* :func:`plip_encode_images` / :func:`plip_encode_text` follow ``plip.py:31-53`` / ``:55-71`` — batches of
  ``batch_size`` through ``model.get_image_features`` / ``get_text_features`` (with the transformers>=5 shim of
  SURVEY.md §8c: unwrap ``.pooler_output``), ``.detach().cpu().numpy()`` per batch, ``np.stack`` at the end;
* :func:`clip_forward` is the README call ``model(**inputs).logits_per_image`` (``README.md:45-49``).

If ``transformers`` cannot be imported the callers fall back to ``oracle/clip_oracle.py`` (kind "port").
Nothing under ``plip_b200/`` imports this file.
"""
from __future__ import annotations

import time
from typing import Optional

import numpy as np
import torch


def load_model(state_dict):
    """``CLIPModel(CLIPConfig())`` with the given weights (what ``from_pretrained`` yields for a ViT-B/32 checkpoint)."""
    from transformers import CLIPConfig, CLIPModel
    m = CLIPModel(CLIPConfig()).eval()
    m.load_state_dict(state_dict, strict=True)
    return m


def _features(out):
    return out.pooler_output if hasattr(out, "pooler_output") else out        # SURVEY §8c shim 2


@torch.no_grad()
def plip_encode_images(model, pixel_values: torch.Tensor, batch_size: int) -> np.ndarray:
    """plip.py:31-53 with already-preprocessed pixels (a 224x224 tile passes through CLIPProcessor unchanged but
    for the exact rescale/normalise)."""
    rows = []
    for i in range(0, pixel_values.shape[0], batch_size):
        rows.extend(_features(model.get_image_features(pixel_values=pixel_values[i:i + batch_size])).detach().cpu().numpy())
    return np.stack(rows)


@torch.no_grad()
def plip_encode_text(model, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor], batch_size: int) -> np.ndarray:
    """plip.py:55-71 from the tokenizer's output (ids + mask)."""
    rows = []
    for i in range(0, input_ids.shape[0], batch_size):
        am = attention_mask[i:i + batch_size] if attention_mask is not None else None
        rows.extend(_features(model.get_text_features(input_ids=input_ids[i:i + batch_size], attention_mask=am)).detach().cpu().numpy())
    return np.stack(rows)


@torch.no_grad()
def clip_forward(model, input_ids, pixel_values, attention_mask=None) -> torch.Tensor:
    """README.md:45-49 — ``model(**inputs).logits_per_image``."""
    return model(input_ids=input_ids, pixel_values=pixel_values, attention_mask=attention_mask).logits_per_image


def pick_threads(fn, candidates, reps: int = 1):
    """Time ``fn()`` once per candidate thread count and keep the fastest (a 64-thread pool on a batch-32 fp32
    forward was measured SLOWER than 16 threads in round 1).  Returns ``(best_threads, {threads: seconds})``."""
    seen, times = set(), {}
    for t in candidates:
        t = int(t)
        if t < 1 or t in seen:
            continue
        seen.add(t)
        torch.set_num_threads(t)
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        times[t] = (time.perf_counter() - t0) / reps
    best = min(times, key=times.get)
    torch.set_num_threads(best)
    return best, times

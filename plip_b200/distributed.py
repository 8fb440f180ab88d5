"""Multi-GPU data parallelism for the encode path: one process per GPU, batch sharding, and an
all-gather of embeddings only where a cross-batch similarity matrix / global embedding table is needed.

The reference is single-device (``plip.py:15``) — this layer is new.  Every image and
caption is an independent unit and the packed weights (354 MB) are replicated per GPU, so the towers need
no collective at all; the only exchange is ``all_gather`` of ``[n_local,512]`` float32 rows
(NCCL over NVLink on GPUs; gloo in the CPU tests).  Rows are block-partitioned contiguously with the
remainder on the low ranks, so a gather restores the original order.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """``(rank, world_size)``; ``(0, 1)`` when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_counts(n: int, world_size: int) -> List[int]:
    """Rows per rank: contiguous blocks, remainder spread over the low ranks."""
    base, rem = divmod(int(n), int(world_size))
    return [base + (1 if r < rem else 0) for r in range(world_size)]


def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Half-open row range ``[lo, hi)`` owned by ``rank``."""
    counts = shard_counts(n, world_size)
    lo = sum(counts[:rank])
    return lo, lo + counts[rank]


def all_gather_rows(local: torch.Tensor, counts: Optional[Sequence[int]] = None, group=None) -> torch.Tensor:
    """Concatenate per-rank row blocks ``[n_r, d]`` in rank order -> ``[sum n_r, d]`` on every rank.

    Uneven blocks are padded to the largest block for a single ``all_gather_into_tensor`` (one NCCL
    call, NVLS-eligible) and trimmed afterwards.  ``counts`` avoids an extra size exchange."""
    rank, ws = world()
    if ws == 1:
        return local
    local = local.contiguous()
    if counts is None:
        sizes = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
        all_sizes = [torch.zeros_like(sizes) for _ in range(ws)]
        dist.all_gather(all_sizes, sizes, group=group)
        counts = [int(s.item()) for s in all_sizes]
    counts = list(counts)
    if local.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank}: local block has {local.shape[0]} rows, expected {counts[rank]}")
    mx = max(counts)
    d = local.shape[1:]
    if local.shape[0] < mx:
        pad = torch.zeros((mx - local.shape[0], *d), device=local.device, dtype=local.dtype)
        send = torch.cat([local, pad], dim=0)
    else:
        send = local
    recv = torch.empty((ws * mx, *d), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(recv, send, group=group)
    if all(c == mx for c in counts):
        return recv
    return torch.cat([recv[r * mx: r * mx + counts[r]] for r in range(ws)], dim=0)


def all_gather_rows_async(local: torch.Tensor, group=None):
    """Equal-sized row blocks only: start ``all_gather_into_tensor`` on NCCL's stream and return
    ``(gathered, work)``; call ``work.wait()`` before reading ``gathered`` on the current stream.  Lets the
    exchange (and any skew between ranks) overlap with whatever is launched in between — e.g. the vision tower
    while the text embeddings travel."""
    rank, ws = world()
    if ws == 1:
        return local, None
    local = local.contiguous()
    recv = torch.empty((ws * local.shape[0], *local.shape[1:]), device=local.device, dtype=local.dtype)
    work = dist.all_gather_into_tensor(recv, local, group=group, async_op=True)
    return recv, work


class ShardedCLIP:
    """Batch-sharded encode + similarity flows of BASELINE.json's multi-GPU configs.

    ``encode_images`` / ``encode_text`` are callables ``rows -> [n,512]`` for the *local* shard (on GPUs:
    ``Engine.encode_images`` / ``Engine.encode_text``); ``similarity(a, b, scale)`` returns ``scale * a @ b.T``
    on normalised rows (``Engine.similarity``).  The class only does the partitioning / gathering."""

    def __init__(self, encode_images: Callable, encode_text: Callable, similarity: Callable, logit_scale_exp: float):
        self.encode_images = encode_images
        self.encode_text = encode_text
        self.similarity = similarity
        self.logit_scale_exp = float(logit_scale_exp)
        self.rank, self.world_size = world()

    def local_slice(self, n_total: int) -> slice:
        lo, hi = shard_range(n_total, self.rank, self.world_size)
        return slice(lo, hi)

    def zero_shot(self, local_images, class_token_ids, n_total_images: int, gather_embeddings: bool = True):
        """cfg4: images sharded, class prompts replicated (64 x 77 ids: cheaper to recompute than to ship).
        Returns ``(pred_local [n_local], logits_local [n_local, n_classes], image_embeds_all or None)``."""
        txt = self.encode_text(class_token_ids)                      # replicated
        img = self.encode_images(local_images)                       # this rank's block
        logits = self.similarity(img, txt, self.logit_scale_exp)     # [n_local, n_classes]
        pred = logits.argmax(dim=-1)
        all_img = None
        if gather_embeddings:
            all_img = all_gather_rows(img, shard_counts(n_total_images, self.world_size))
        return pred, logits, all_img

    def retrieval(self, local_gallery_images, local_query_ids, n_total_queries: int):
        """cfg5: gallery and queries sharded; query embeddings are all-gathered (small: 10k x 512 fp32 =
        20.5 MB), the gallery stays sharded and each rank returns its row block of the
        ``[n_gallery, n_queries]`` similarity matrix."""
        q_local = self.encode_text(local_query_ids)                  # queries first: their exchange overlaps the gallery
        counts = shard_counts(n_total_queries, self.world_size)
        if self.world_size > 1 and len(set(counts)) == 1:
            q_all, work = all_gather_rows_async(q_local)
            gal = self.encode_images(local_gallery_images)
            work.wait()
        else:
            gal = self.encode_images(local_gallery_images)
            q_all = all_gather_rows(q_local, counts)
        block = self.similarity(gal, q_all, self.logit_scale_exp)    # [n_gallery_local, n_queries]
        return block, gal, q_all

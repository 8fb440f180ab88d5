"""Device versions of the reference's similarity heads (``reproducibility/evaluation``).

``ZeroShotClassifier`` (``evaluation/zero_shot/zero_shot.py:10-13``) and ``ImageRetrieval``
(``evaluation/retrieval/retrieval.py:9-17``) take ``[N,512]`` embedding matrices (numpy or torch, host or
device) like the reference; the ``dot`` + ``argmax`` / ``argsort()[-50:][::-1]`` arithmetic runs in the CUDA
similarity kernels (fp32).  The sklearn metrics of the reference (``metrics.py``) are CPU statistics outside the
hot path: pass ``eval_metrics=`` to reuse them; the p@10 / p@50 retrieval metric is restated here because it is a
pure function of the top-k indices.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from .engine import Engine, similarity_topk


def _t(x) -> torch.Tensor:
    return x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))


class _Head:
    """Constructible with no arguments like the reference's classes (``zero_shot.py:7-8``, ``retrieval.py:6-7``): the
    similarity kernels need no weights, only a device — the current CUDA device, or that of an ``engine`` if given."""

    def __init__(self, engine: Optional[Engine] = None):
        self.engine = engine

    def _topk(self, query, space, k):
        dev = self.engine.device if self.engine is not None else None
        return similarity_topk(query, space, k, scale=1.0, normalize_query=False, normalize_space=False, device=dev)


class ZeroShotClassifier(_Head):

    def predict(self, image_embeddings, text_embeddings, unique_labels: Sequence) -> List:
        """``[unique_labels[np.argmax(i)] for i in image_embeddings.dot(text_embeddings.T)]`` (zero_shot.py:12-13)."""
        idx, _ = self._topk(_t(image_embeddings), _t(text_embeddings), 1)
        return [unique_labels[i] for i in idx[:, 0].cpu().tolist()]

    def zero_shot_classification(self, image_embeddings, text_embeddings, unique_labels, target_labels,
                                 eval_metrics: Optional[Callable] = None):
        predictions = self.predict(image_embeddings, text_embeddings, unique_labels)
        if eval_metrics is None:
            acc = float(np.mean([p == t for p, t in zip(predictions, target_labels)]))
            train_metrics, test_metrics = {"accuracy": acc}, {"accuracy": acc}
        else:
            test_metrics = eval_metrics(target_labels, predictions)
            train_metrics = eval_metrics(target_labels, predictions)
        test_metrics["split"], train_metrics["split"] = "test", "train"
        return train_metrics, test_metrics


def retrieval_metrics(y_target, y_predictions):
    """``reproducibility/metrics.py:5-15``: fraction of queries whose target is in the top 10 / top 50."""
    p10 = sum(1 for t, p in zip(y_target, y_predictions) if t in list(p[:10]))
    p50 = sum(1 for t, p in zip(y_target, y_predictions) if t in list(p[:50]))
    return {"p@10": p10 / len(y_target), "p@50": p50 / len(y_target)}


class ImageRetrieval(_Head):

    def best_scores(self, image_embeddings, text_embeddings, top_k: int = 50) -> np.ndarray:
        """Per text query, indices of the ``top_k`` most similar images, best first
        (``t.dot(image_embeddings.T).argsort()[-50:][::-1]``, retrieval.py:13-16)."""
        imgs, txt = _t(image_embeddings), _t(text_embeddings)
        k = min(top_k, imgs.shape[0])
        idx, _ = self._topk(txt, imgs, k)
        return idx.cpu().numpy()

    def retrieval(self, image_embeddings, text_embeddings):
        best = self.best_scores(image_embeddings, text_embeddings, 50)
        targets = list(range(0, len(image_embeddings)))
        test_metrics, train_metrics = retrieval_metrics(targets, best), retrieval_metrics(targets, best)
        test_metrics["split"], train_metrics["split"] = "test", "train"
        return train_metrics, test_metrics

"""Drop-in for the ``transformers.CLIPModel`` call surface used by the reference.

``PlipCLIPModel`` keeps the three entry points the reference touches —
``get_image_features`` (reference ``plip.py:50``), ``get_text_features`` (``plip.py:68``) and
``model(**inputs).logits_per_image`` (``README.md:45-49``; TF:modeling_clip.py:867-944) — and the OpenAI-clip
``encode_image`` / ``encode_text`` used by ``reproducibility/embedders/plip.py:48,66``.  All compute runs in the
CUDA engine; outputs are torch tensors on the engine's device, like the HF model's.

Return types follow what the reference code expects (transformers v4 semantics): ``get_*_features`` return
the ``[n,512]`` tensor itself — the reference calls ``.detach().cpu().numpy()`` on it directly.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Mapping, Optional, Union

import torch

from .engine import Engine


@dataclass
class CLIPOutput:
    """Fields of ``transformers.models.clip.modeling_clip.CLIPOutput`` (TF:104-135) that the engine produces."""

    logits_per_image: torch.Tensor
    logits_per_text: torch.Tensor
    text_embeds: torch.Tensor
    image_embeds: torch.Tensor
    loss: Optional[torch.Tensor] = None

    def __getitem__(self, k):
        return getattr(self, k)

    def keys(self):
        return ("logits_per_image", "logits_per_text", "text_embeds", "image_embeds")


class PlipCLIPModel:
    """CUDA-engine-backed stand-in for ``CLIPModel`` (ViT-B/32 geometry only, as PLIP ships)."""

    def __init__(self, state_dict: Mapping[str, torch.Tensor], device: Union[int, str, torch.device, None] = None,
                 max_micro_batch: int = 1024, operand_dtype="bf16"):
        self.engine = Engine(state_dict, device=device, max_micro_batch=max_micro_batch, operand_dtype=operand_dtype)
        self.device = self.engine.device
        self.training = False

    # ---- construction -------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, name_or_path: str, device=None, max_micro_batch: int = 1024, operand_dtype="bf16",
                        **hf_kwargs):
        """Read a HuggingFace CLIP checkpoint (e.g. ``vinid/plip`` or a local directory) on the host and pack
        it for the engine.  ``transformers`` is only the checkpoint *reader* here; its forward never runs.
        ``use_auth_token`` (reference ``plip.py:26``) is translated to ``token`` for transformers >= 5."""
        from transformers import CLIPModel  # host-side weight loading only

        tok = hf_kwargs.pop("use_auth_token", None)
        if tok is not None:
            hf_kwargs.setdefault("token", tok)
        hf = CLIPModel.from_pretrained(name_or_path, **hf_kwargs)
        cfg = hf.config
        v, t = cfg.vision_config, cfg.text_config
        # every field the kernels hard-code (common.cuh model constants; TF:configuration_clip.py:47-64,97-109)
        want = {"vision hidden_size": (v.hidden_size, 768), "vision patch_size": (v.patch_size, 32),
                "vision image_size": (v.image_size, 224), "vision num_hidden_layers": (v.num_hidden_layers, 12),
                "vision num_attention_heads": (v.num_attention_heads, 12), "vision intermediate_size": (v.intermediate_size, 3072),
                "vision hidden_act": (v.hidden_act, "quick_gelu"), "vision layer_norm_eps": (float(v.layer_norm_eps), 1e-5),
                "text hidden_size": (t.hidden_size, 512), "text num_hidden_layers": (t.num_hidden_layers, 12),
                "text num_attention_heads": (t.num_attention_heads, 8), "text intermediate_size": (t.intermediate_size, 2048),
                "text hidden_act": (t.hidden_act, "quick_gelu"), "text layer_norm_eps": (float(t.layer_norm_eps), 1e-5),
                "text max_position_embeddings": (t.max_position_embeddings, 77), "text vocab_size": (t.vocab_size, 49408),
                "projection_dim": (cfg.projection_dim, 512)}
        bad = {k: got for k, (got, exp) in want.items() if got != exp}
        if bad:
            raise ValueError("plip_b200 implements the CLIP ViT-B/32 geometry only (PLIP's architecture); this "
                             f"checkpoint differs in {bad}")
        m = cls(hf.state_dict(), device=device, max_micro_batch=max_micro_batch, operand_dtype=operand_dtype)
        # legacy configs (eos_token_id == 2) pool at argmax(input_ids) (TF:564-570); same row whenever an eos exists
        m.engine.set_text_pooling(getattr(t, "eos_token_id", 49407) == 2)
        return m

    @classmethod
    def from_openai_state_dict(cls, state_dict, device=None, max_micro_batch: int = 1024, operand_dtype="bf16"):
        """``clip.load(arch)`` + ``load_state_dict(torch.load(path))`` (``embedders/factory.py:20-27``)."""
        m = cls(state_dict, device=device, max_micro_batch=max_micro_batch, operand_dtype=operand_dtype)
        m.engine.set_text_pooling(True)          # OpenAI clip: x[arange, text.argmax(-1)] (eot has the largest id)
        return m

    # ---- nn.Module-ish no-ops the reference calls -------------------------------------------------
    def to(self, *args, **kwargs):
        return self

    def eval(self):
        return self

    def float(self):
        return self

    @property
    def logit_scale_exp(self) -> float:
        return self.engine.logit_scale_exp

    # ---- HF surface ---------------------------------------------------------------------------
    def get_image_features(self, pixel_values: torch.Tensor = None, **_ignored) -> torch.Tensor:
        """TF:829-863 — vision tower + visual_projection, un-normalised ``[n,512]`` float32."""
        if pixel_values is None:
            raise ValueError("You have to specify pixel_values")
        return self.engine.encode_images(pixel_values)

    def get_text_features(self, input_ids: torch.Tensor = None, attention_mask: Optional[torch.Tensor] = None,
                          **_ignored) -> torch.Tensor:
        """TF:793-825 — text tower + text_projection, un-normalised ``[n,512]`` float32."""
        if input_ids is None:
            raise ValueError("You have to specify input_ids")  # TF:540-541
        return self.engine.encode_text(input_ids, attention_mask)

    def forward(self, input_ids: torch.Tensor = None, pixel_values: torch.Tensor = None,
                attention_mask: Optional[torch.Tensor] = None, return_loss: Optional[bool] = None,
                **_ignored) -> CLIPOutput:
        """TF:867-944 — both towers, L2-normalise, ``exp(logit_scale) * I . T^T``."""
        if input_ids is None:
            raise ValueError("You have to specify input_ids")
        if pixel_values is None:
            raise ValueError("You have to specify pixel_values")
        if not pixel_values.is_cuda:
            # host inputs (an extension: HF would raise on a device mismatch): the pixel upload runs on the engine's
            # copy stream while the text tower computes, so ~3 ms of PCIe time per 1024 uint8 tiles stay hidden
            # ... and batches larger than one micro-batch are uploaded micro-batch by micro-batch, each one computed as
            # soon as it has landed while the next one travels
            mb = self.engine.max_micro_batch
            parts = [self.engine.upload_async(pixel_values[i:i + mb]) for i in range(0, pixel_values.shape[0], mb)]
            txt = self.engine.encode_text(input_ids, attention_mask, normalize=True)
            embs = []
            for chunk, uploaded in parts:
                torch.cuda.current_stream(self.device).wait_event(uploaded)
                embs.append(self.engine.encode_images(chunk, normalize=True))
            img = embs[0] if len(embs) == 1 else torch.cat(embs, dim=0)
        else:
            img = self.engine.encode_images(pixel_values, normalize=True)
            txt = self.engine.encode_text(input_ids, attention_mask, normalize=True)
        lpi = self.engine.similarity(img, txt, normalize_image=False, normalize_text=False)
        loss = None
        if return_loss:  # clip_loss (TF:68-76): symmetric cross entropy; tiny, done with torch on the logits
            lpt = lpi.t()
            tgt = torch.arange(lpt.shape[0], device=lpt.device)
            loss = (torch.nn.functional.cross_entropy(lpt, tgt) + torch.nn.functional.cross_entropy(lpt.t(), tgt)) / 2
        return CLIPOutput(logits_per_image=lpi, logits_per_text=lpi.t(), text_embeds=txt, image_embeds=img, loss=loss)

    __call__ = forward

    # ---- OpenAI-clip surface (reproducibility/embedders/plip.py:48,66) ---------------------------------
    def encode_image(self, images: torch.Tensor) -> torch.Tensor:
        return self.engine.encode_images(images)

    def encode_text(self, idx: torch.Tensor) -> torch.Tensor:
        return self.engine.encode_text(idx)

"""TEST INFRASTRUCTURE — CPU restatement of the PLIP / CLIP ViT-B/32 forward (fp32, torch-CPU ops).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline / ``--impl reference``
legs may import this module; the product path (``plip_b200``) never does and has no CPU fallback.

The arithmetic of the reference's hot path lives in the third-party ``transformers`` package
(un-pinned by the reference: ``requirements.txt`` is empty; installed here: 5.5.0).  Every function
below restates one piece of ``transformers/models/clip/modeling_clip.py`` ("TF:") and cites it.
Pinning: ``tests/golden/make_golden.py`` ran the *real* ``transformers.CLIPModel`` and the reference's
own ``plip.PLIP`` class (with the two compatibility shims of SURVEY.md §8c) in the build container
and committed their outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks this restatement
against those vectors (and against live ``transformers`` when importable).  The reference ships no
tests / golden vectors of its own, so that is the strongest pin available.

``operand_dtype=torch.bfloat16`` emulates the device numerics contract (GEMM / attention operands
rounded to bf16, everything else fp32) and is used to derive tolerances, not as a parity target.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .weights import EOS, TEXT, VISION

SD = Dict[str, torch.Tensor]
LN_EPS = 1e-5  # TF:configuration_clip.py:55,106

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)  # TF:image_processing_clip.py / embedders/transform.py:51
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _r(x: torch.Tensor, dt: Optional[torch.dtype]) -> torch.Tensor:
    """Round to the emulated operand dtype and come back to fp32."""
    return x if dt is None else x.to(dt).to(torch.float32)


def linear(x, w, b=None, dt=None):
    """nn.Linear: x @ w.T + b (operands optionally rounded, fp32 accumulate)."""
    y = _r(x, dt) @ _r(w, dt).t()
    return y if b is None else y + b


def layer_norm(x, w, b):
    """nn.LayerNorm(eps=1e-5) over the last dim (TF:371,380,562,677,686)."""
    return F.layer_norm(x, (x.shape[-1],), w, b, LN_EPS)


def quick_gelu(x):
    """QuickGELUActivation: x * sigmoid(1.702 x) (TF:activations.py:117-123)."""
    return x * torch.sigmoid(1.702 * x)


def preprocess_u8(tiles_u8: torch.Tensor) -> torch.Tensor:
    """CLIPImageProcessor on an already 224x224 RGB uint8 tile [n,224,224,3] -> fp32 [n,3,224,224]:
    rescale 1/255 then normalise by CLIP mean/std (TF:image_processing_clip.py:50-62)."""
    x = tiles_u8.to(torch.float32).permute(0, 3, 1, 2) / 255.0
    mean = torch.tensor(CLIP_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD).view(1, 3, 1, 1)
    return (x - mean) / std


def attention(x, sd: SD, p: str, heads: int, mask: Optional[torch.Tensor], dt=None):
    """CLIPAttention.forward (TF:300-336) with the eager core (TF:261-279):
    softmax(q k^T * dh^-0.5 + mask) v, softmax in fp32."""
    B, S, D = x.shape
    dh = D // heads
    q = linear(x, sd[f"{p}.q_proj.weight"], sd[f"{p}.q_proj.bias"], dt)
    k = linear(x, sd[f"{p}.k_proj.weight"], sd[f"{p}.k_proj.bias"], dt)
    v = linear(x, sd[f"{p}.v_proj.weight"], sd[f"{p}.v_proj.bias"], dt)
    q = q.view(B, S, heads, dh).transpose(1, 2)
    k = k.view(B, S, heads, dh).transpose(1, 2)
    v = v.view(B, S, heads, dh).transpose(1, 2)
    att = (_r(q, dt) @ _r(k, dt).transpose(-1, -2)) * (dh ** -0.5)
    if mask is not None:
        att = att + mask
    att = torch.softmax(att, dim=-1, dtype=torch.float32)
    o = _r(att, dt) @ _r(v, dt)
    o = o.transpose(1, 2).reshape(B, S, D)
    return linear(o, sd[f"{p}.out_proj.weight"], sd[f"{p}.out_proj.bias"], dt)


def mlp(x, sd: SD, p: str, dt=None):
    """CLIPMLP.forward: fc2(quick_gelu(fc1(x))) (TF:347-351)."""
    h = quick_gelu(linear(x, sd[f"{p}.fc1.weight"], sd[f"{p}.fc1.bias"], dt))
    return linear(h, sd[f"{p}.fc2.weight"], sd[f"{p}.fc2.bias"], dt)


def encoder_layer(x, sd: SD, p: str, heads: int, mask, dt=None):
    """CLIPEncoderLayer.forward: pre-LN residual block (TF:363-384)."""
    x = x + attention(layer_norm(x, sd[f"{p}.layer_norm1.weight"], sd[f"{p}.layer_norm1.bias"]),
                      sd, f"{p}.self_attn", heads, mask, dt)
    x = x + mlp(layer_norm(x, sd[f"{p}.layer_norm2.weight"], sd[f"{p}.layer_norm2.bias"]), sd, f"{p}.mlp", dt)
    return x


def encoder(x, sd: SD, prefix: str, heads: int, layers: int, mask, dt=None, hidden: Optional[List] = None):
    """CLIPEncoder.forward (TF:477-507).  ``hidden`` collects the residual stream before each layer
    and after the last (== HF ``output_hidden_states``)."""
    for i in range(layers):
        if hidden is not None:
            hidden.append(x)
        x = encoder_layer(x, sd, f"{prefix}.encoder.layers.{i}", heads, mask, dt)
    if hidden is not None:
        hidden.append(x)
    return x


def vision_embeddings(sd: SD, pixel_values, dt=None):
    """CLIPVisionEmbeddings.forward (TF:202-218): stride-32 conv as a GEMM over 32x32 patches,
    prepend class embedding, add position embedding."""
    B, Cc, H, W = pixel_values.shape
    if H != VISION["image"] or W != VISION["image"]:
        raise ValueError(f"Input image size ({H}*{W}) doesn't match model (224*224).")  # TF:204-207
    w = sd["vision_model.embeddings.patch_embedding.weight"]
    D = w.shape[0]
    patches = pixel_values.reshape(B, Cc, 7, 32, 7, 32).permute(0, 2, 4, 1, 3, 5).reshape(B, 49, Cc * 32 * 32)
    pe = linear(patches, w.reshape(D, -1), None, dt)  # [B,49,768], token order py*7+px (flatten(2).transpose)
    cls = sd["vision_model.embeddings.class_embedding"].expand(B, 1, D)
    x = torch.cat([cls, pe], dim=1)
    return x + sd["vision_model.embeddings.position_embedding.weight"][None]


def vision_transformer(sd: SD, pixel_values, dt=None, hidden: Optional[List] = None):
    """CLIPVisionTransformer.forward (TF:667-691): embeddings -> pre_layrnorm -> encoder ->
    CLS row -> post_layernorm.  Returns pooled [B,768]."""
    x = vision_embeddings(sd, pixel_values, dt)
    x = layer_norm(x, sd["vision_model.pre_layrnorm.weight"], sd["vision_model.pre_layrnorm.bias"])
    x = encoder(x, sd, "vision_model", VISION["heads"], VISION["layers"], None, dt, hidden)
    pooled = x[:, 0, :]
    return layer_norm(pooled, sd["vision_model.post_layernorm.weight"], sd["vision_model.post_layernorm.bias"])


def text_embeddings(sd: SD, input_ids):
    """CLIPTextEmbeddings.forward (TF:234-258)."""
    S = input_ids.shape[-1]
    if S > TEXT["seq"]:
        raise ValueError(f"Sequence length must be less than max_position_embeddings (got {S} and 77)")  # TF:243-247
    tok = sd["text_model.embeddings.token_embedding.weight"][input_ids]
    return tok + sd["text_model.embeddings.position_embedding.weight"][:S][None]


def causal_mask(S: int, attention_mask: Optional[torch.Tensor]):
    """create_causal_mask (TF:546-551): additive [B|1,1,S,S] mask, -inf above the diagonal and on
    padded keys (attention_mask == 0)."""
    neg = torch.finfo(torch.float32).min
    m = torch.full((S, S), neg).triu(1)[None, None]
    if attention_mask is not None:
        pad = (attention_mask == 0)[:, None, None, :]
        m = m.expand(attention_mask.shape[0], 1, S, S).clone()
        m = m.masked_fill(pad, neg)
    return m


def text_transformer(sd: SD, input_ids, attention_mask=None, dt=None, hidden: Optional[List] = None,
                     eos_token_id: int = EOS):
    """CLIPTextTransformer.forward (TF:531-589): embeddings -> causal(+padding) encoder ->
    final_layer_norm -> row of the first eos token (TF:571-584).  Returns pooled [B,512]."""
    x = text_embeddings(sd, input_ids)
    mask = causal_mask(input_ids.shape[-1], attention_mask)
    x = encoder(x, sd, "text_model", TEXT["heads"], TEXT["layers"], mask, dt, hidden)
    x = layer_norm(x, sd["text_model.final_layer_norm.weight"], sd["text_model.final_layer_norm.bias"])
    if eos_token_id == 2:  # legacy configs: argmax of the ids (TF:564-570)
        pos = input_ids.to(torch.int).argmax(dim=-1)
    else:
        pos = (input_ids.to(torch.int) == eos_token_id).int().argmax(dim=-1)
    return x[torch.arange(x.shape[0]), pos]


def get_image_features(sd: SD, pixel_values, dt=None):
    """CLIPModel.get_image_features(...).pooler_output (TF:829-863): [B,512], not normalised."""
    return linear(vision_transformer(sd, pixel_values, dt), sd["visual_projection.weight"], None, dt)


def get_text_features(sd: SD, input_ids, attention_mask=None, dt=None):
    """CLIPModel.get_text_features(...).pooler_output (TF:793-825): [B,512], not normalised."""
    return linear(text_transformer(sd, input_ids, attention_mask, dt), sd["text_projection.weight"], None, dt)


def l2_normalize(x):
    """x / _get_vector_norm(x) (TF:57-65,923-924): no epsilon."""
    return x / x.pow(2).sum(-1, keepdim=True).sqrt()


def similarity(image_embeds, text_embeds, logit_scale_exp: float):
    """logits_per_image = (text @ image.T * exp(logit_scale)).T on normalised embeds (TF:927-930)."""
    return (text_embeds @ image_embeds.t() * logit_scale_exp).t()


def clip_forward(sd: SD, input_ids, pixel_values, attention_mask=None, dt=None):
    """CLIPModel.forward (TF:867-944) -> dict with logits_per_image / logits_per_text and the
    normalised image_embeds / text_embeds."""
    img = l2_normalize(get_image_features(sd, pixel_values, dt))
    txt = l2_normalize(get_text_features(sd, input_ids, attention_mask, dt))
    scale = float(sd["logit_scale"].exp())
    lpi = similarity(img, txt, scale)
    return {"logits_per_image": lpi, "logits_per_text": lpi.t(), "image_embeds": img, "text_embeds": txt}


# ---- host-side heads of the reference (numpy semantics restated with torch) --------------------

def cosine_similarity_keys(key, space, normalize=True):
    """PLIP._cosine_similarity (plip.py:73-76): only the key side is normalised."""
    if normalize:
        key = key / key.norm(dim=-1, keepdim=True)
    return key @ space.t()


def nearest_neighbours(k, key, space, normalize=True):
    """PLIP._nearest_neighbours (plip.py:78-87): argsort()[:, -k:][:, ::-1]."""
    sim = cosine_similarity_keys(key, space, normalize)
    return sim.argsort(dim=-1)[:, -k:].flip(-1)


def cosine(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Per-row cosine between two [n,d] matrices (parity metric; fp64)."""
    a = a.double()
    b = b.double()
    return (a * b).sum(-1) / (a.norm(dim=-1) * b.norm(dim=-1))

"""The oracle (CPU restatement) against the golden vectors produced by the real reference stack
(transformers.CLIPModel + the reference's PLIP class; tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import clip_oracle as O
from oracle import synth

torch.set_grad_enabled(False)


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_image_features_match_hf(golden, state_dict):
    px = synth.pixel_values(8)
    out = O.get_image_features(state_dict, px)
    ref = _t(golden["image_features"])
    assert (1 - O.cosine(out, ref)).max().item() < 1e-10
    assert (out - ref).abs().max().item() < 2e-5


def test_text_features_match_hf(golden, state_dict):
    ids, mask = synth.token_ids(8)
    out = O.get_text_features(state_dict, ids, mask)
    ref = _t(golden["text_features"])
    assert (1 - O.cosine(out, ref)).max().item() < 1e-10
    assert (out - ref).abs().max().item() < 2e-5
    # causality: with eos padding the pooled row does not depend on the padding mask
    out_nomask = O.get_text_features(state_dict, ids, None)
    assert (out_nomask - _t(golden["text_features_nomask"])).abs().max().item() < 2e-5
    assert (out_nomask - out).abs().max().item() < 1e-6


def test_full_length_captions(golden, state_dict):
    ids, mask = synth.token_ids(4, seed=77, full_length=True)
    assert (ids[:, -1] == 49407).all() and (ids[:, :-1] != 49407).all()
    out = O.get_text_features(state_dict, ids, mask)
    assert (out - _t(golden["text_features_full77"])).abs().max().item() < 2e-5


def test_hidden_states_match_hf(golden, state_dict):
    px = synth.pixel_values(8)[:2]
    ids, mask = synth.token_ids(8)
    hv, ht = [], []
    O.vision_transformer(state_dict, px, hidden=hv)
    O.text_transformer(state_dict, ids[:2], mask[:2], hidden=ht)
    for l in (0, 1, 6, 12):
        assert (hv[l][:2, :5] - _t(golden[f"vision_hidden_{l}"])).abs().max().item() < 5e-5
        assert (ht[l][:2, :9] - _t(golden[f"text_hidden_{l}"])).abs().max().item() < 5e-5


def test_clip_forward_logits(golden, state_dict):
    px = synth.pixel_values(8)
    ids, mask = synth.token_ids(8)
    out = O.clip_forward(state_dict, ids, px, mask)
    assert abs(float(state_dict["logit_scale"].exp()) - float(golden["logit_scale_exp"])) < 1e-6
    assert (out["logits_per_image"] - _t(golden["logits_per_image"])).abs().max().item() < 1e-5
    assert (out["image_embeds"] - _t(golden["image_embeds"])).abs().max().item() < 1e-6
    assert (out["text_embeds"] - _t(golden["text_embeds"])).abs().max().item() < 1e-6
    assert torch.equal(out["logits_per_text"], out["logits_per_image"].t())


def test_reference_plip_class_cfg1(golden, state_dict):
    """cfg1: the reference's PLIP.encode_images on 32 synthetic uint8 tiles (batch 8) == oracle on the same tiles."""
    tiles = torch.from_numpy(synth.tiles_u8(32, seed=0))
    out = O.get_image_features(state_dict, O.preprocess_u8(tiles))
    ref = _t(golden["ref_plip_encode_images_bs8"])
    assert ref.shape == (32, 512) and ref.dtype == torch.float32
    assert (1 - O.cosine(out, ref)).max().item() < 1e-9
    assert (out - ref).abs().max().item() < 5e-5


def test_reference_numpy_heads(golden):
    key, space = _t(golden["heads_key"]), _t(golden["heads_space"])
    sim = O.cosine_similarity_keys(key, space)
    assert (sim - _t(golden["ref_cosine_similarity"])).abs().max().item() < 1e-5
    nn = O.nearest_neighbours(5, key, space)
    assert torch.equal(nn, _t(golden["ref_nearest_neighbours_k5"]))


def test_shape_errors_mirror_reference(state_dict):
    with pytest.raises(ValueError, match="doesn't match model"):
        O.vision_embeddings(state_dict, torch.zeros(1, 3, 256, 256))
    with pytest.raises(ValueError, match="Sequence length"):
        O.text_embeddings(state_dict, torch.zeros(1, 78, dtype=torch.long))


def test_bf16_operand_emulation_within_tolerance(state_dict):
    """The device numerics contract (bf16 GEMM/attention operands, fp32 everything else) stays inside the
    north-star cosine bar (>= 1 - 1e-4) — this is what bounds the GPU parity tests' tolerance."""
    px = synth.pixel_values(4)
    ids, mask = synth.token_ids(4)
    ref = O.clip_forward(state_dict, ids, px, mask)
    emu = O.clip_forward(state_dict, ids, px, mask, dt=torch.bfloat16)
    assert (1 - O.cosine(emu["image_embeds"], ref["image_embeds"])).max().item() < 1e-4
    assert (1 - O.cosine(emu["text_embeds"], ref["text_embeds"])).max().item() < 1e-4


def test_live_transformers_if_present(state_dict):
    """Belt and braces: when transformers is importable, compare against the live CLIPModel too."""
    tf = pytest.importorskip("transformers")
    m = tf.CLIPModel(tf.CLIPConfig()).eval()
    m.load_state_dict(state_dict, strict=True)
    px = synth.pixel_values(2, seed=99)
    ids, mask = synth.token_ids(3, seed=98)
    hf = m(input_ids=ids, pixel_values=px, attention_mask=mask)
    out = O.clip_forward(state_dict, ids, px, mask)
    assert (out["logits_per_image"] - hf.logits_per_image).abs().max().item() < 1e-5

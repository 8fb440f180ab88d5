"""Weight packer: HF / OpenAI-clip state dicts -> engine blob (host logic, no GPU)."""
import torch

from plip_b200 import weights as W


def _view(blob, ti):
    raw = blob[ti.offset: ti.offset + ti.numel * (2 if ti.dtype == 1 else 4)]
    return raw.view(torch.bfloat16 if ti.dtype == 1 else torch.float32).view(ti.rows, ti.cols)


def test_pack_layout_and_values(state_dict):
    blob, scale = W.pack_state_dict(state_dict)
    assert abs(scale - float(state_dict["logit_scale"].exp())) < 1e-6
    table = {ti.name.decode(): ti for ti in W.tensor_table()}
    # plain GEMM weight: one RNE rounding to bf16
    ti = table["vision_model.encoder.layers.3.mlp.fc2.weight"]
    assert (ti.rows, ti.cols, ti.dtype, ti.fused) == (768, 3072, 1, 0)
    assert torch.equal(_view(blob, ti), state_dict["vision_model.encoder.layers.3.mlp.fc2.weight"].to(torch.bfloat16))
    # fc1 with layer_norm2 folded in: W' = bf16(gamma o W), bias' = bias + W beta, colsum = rowsum(W')
    L = "vision_model.encoder.layers.3."
    ti = table[L + "mlp.fc1.weight"]
    assert (ti.rows, ti.cols, ti.dtype, ti.fused) == (3072, 768, 1, 2)
    w, b = state_dict[L + "mlp.fc1.weight"], state_dict[L + "mlp.fc1.bias"]
    g, be = state_dict[L + "layer_norm2.weight"], state_dict[L + "layer_norm2.bias"]
    wf = (w * g[None]).to(torch.bfloat16)
    assert torch.equal(_view(blob, ti), wf)
    assert torch.allclose(_view(blob, table[L + "mlp.fc1.bias"]).flatten(), b + w @ be, atol=1e-6)
    assert torch.allclose(_view(blob, table[L + "mlp.fc1.colsum"]).flatten(), wf.float().sum(1), atol=1e-5)
    assert L + "layer_norm2.weight" not in table and L + "layer_norm1.bias" not in table
    # the fold is exact algebra: LN(x) W^T + b == rstd (x W'^T - mean colsum) + bias'  (fp32 check)
    x = torch.randn(5, 768) * 2 + 0.3
    ref = torch.nn.functional.layer_norm(x, (768,), g, be, 1e-5) @ w.t() + b
    mean = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-5)
    wg = w * g[None]
    assert torch.allclose(rstd * (x @ wg.t() - mean * wg.sum(1)) + (b + w @ be), ref, atol=2e-4)
    # fused q|k|v: 0.125 head scale on the q rows (exact: power of two) + layer_norm1 fold
    ti = table["text_model.encoder.layers.0.self_attn.q_proj.weight"]
    assert (ti.rows, ti.cols, ti.fused) == (1536, 512, 3)
    p = "text_model.encoder.layers.0.self_attn."
    g1 = state_dict["text_model.encoder.layers.0.layer_norm1.weight"]
    be1 = state_dict["text_model.encoder.layers.0.layer_norm1.bias"]
    fused = _view(blob, ti)
    assert torch.equal(fused[:512], (state_dict[p + "q_proj.weight"] * 0.125 * g1[None]).to(torch.bfloat16))
    assert torch.equal(fused[512:1024], (state_dict[p + "k_proj.weight"] * g1[None]).to(torch.bfloat16))
    assert torch.equal(fused[1024:], (state_dict[p + "v_proj.weight"] * g1[None]).to(torch.bfloat16))
    bq = _view(blob, table[p + "q_proj.bias"]).flatten()
    assert torch.allclose(bq[:512], 0.125 * (state_dict[p + "q_proj.bias"] + state_dict[p + "q_proj.weight"] @ be1), atol=1e-6)
    assert torch.allclose(bq[512:1024], state_dict[p + "k_proj.bias"] + state_dict[p + "k_proj.weight"] @ be1, atol=1e-6)
    cs = _view(blob, table[p + "q_proj.colsum"]).flatten()
    assert torch.allclose(cs, fused.float().sum(1), atol=1e-5)
    # conv patch embedding viewed as [768, 3*32*32]
    ti = table["vision_model.embeddings.patch_embedding.weight"]
    assert torch.equal(_view(blob, ti), state_dict["vision_model.embeddings.patch_embedding.weight"].reshape(768, 3072).to(torch.bfloat16))
    ti = table["text_model.embeddings.token_embedding.weight"]
    assert ti.dtype == 0 and torch.equal(_view(blob, ti), state_dict["text_model.embeddings.token_embedding.weight"])


def _to_openai(sd):
    out = {"logit_scale": sd["logit_scale"]}
    out["visual.conv1.weight"] = sd["vision_model.embeddings.patch_embedding.weight"]
    out["visual.class_embedding"] = sd["vision_model.embeddings.class_embedding"]
    out["visual.positional_embedding"] = sd["vision_model.embeddings.position_embedding.weight"]
    out["visual.ln_pre.weight"], out["visual.ln_pre.bias"] = sd["vision_model.pre_layrnorm.weight"], sd["vision_model.pre_layrnorm.bias"]
    out["visual.ln_post.weight"], out["visual.ln_post.bias"] = sd["vision_model.post_layernorm.weight"], sd["vision_model.post_layernorm.bias"]
    out["visual.proj"] = sd["visual_projection.weight"].t().contiguous()
    out["token_embedding.weight"] = sd["text_model.embeddings.token_embedding.weight"]
    out["positional_embedding"] = sd["text_model.embeddings.position_embedding.weight"]
    out["ln_final.weight"], out["ln_final.bias"] = sd["text_model.final_layer_norm.weight"], sd["text_model.final_layer_norm.bias"]
    out["text_projection"] = sd["text_projection.weight"].t().contiguous()
    for src, dst in (("vision_model", "visual.transformer"), ("text_model", "transformer")):
        for i in range(12):
            p, q = f"{src}.encoder.layers.{i}", f"{dst}.resblocks.{i}"
            out[f"{q}.attn.in_proj_weight"] = torch.cat([sd[f"{p}.self_attn.{n}.weight"] for n in ("q_proj", "k_proj", "v_proj")])
            out[f"{q}.attn.in_proj_bias"] = torch.cat([sd[f"{p}.self_attn.{n}.bias"] for n in ("q_proj", "k_proj", "v_proj")])
            for a, b in (("self_attn.out_proj", "attn.out_proj"), ("layer_norm1", "ln_1"), ("layer_norm2", "ln_2"),
                         ("mlp.fc1", "mlp.c_fc"), ("mlp.fc2", "mlp.c_proj")):
                out[f"{q}.{b}.weight"], out[f"{q}.{b}.bias"] = sd[f"{p}.{a}.weight"], sd[f"{p}.{a}.bias"]
    return out


def test_openai_clip_names_pack_identically(state_dict):
    """embedders/factory.py:20-27 loads OpenAI-clip checkpoints: same blob as the HF naming."""
    blob_hf, s1 = W.pack_state_dict(state_dict)
    blob_oa, s2 = W.pack_state_dict(_to_openai(state_dict))
    assert s1 == s2 and torch.equal(blob_hf, blob_oa)


def test_prefixed_and_bad_state_dicts(state_dict):
    pref = {"model." + k: v for k, v in state_dict.items()}
    assert set(W.normalize_state_dict(pref)) == set(state_dict)
    import pytest
    with pytest.raises(KeyError):
        W.normalize_state_dict({"foo": torch.zeros(1)})
    bad = dict(state_dict)
    bad["visual_projection.weight"] = torch.zeros(512, 700)
    with pytest.raises(ValueError, match="visual_projection"):
        W.pack_state_dict(bad)

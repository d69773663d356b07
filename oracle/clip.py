"""Oracle: CLIP-ViT-H/14 (PickScore_v1) towers and the PickScore head
(`adv_grpo/pickscore_scorer.py:39-51`), restated from transformers' CLIPModel
architecture (transformers state-dict names).  Cross-checked in
tests/test_oracle_models.py against `transformers.CLIPModel` instantiated from
the same config with the same weights (transformers 5.5 is installed here;
the reference pins 4.54 -- same architecture).
Test infrastructure only (see oracle/__init__.py)."""
import torch
import torch.nn.functional as F


def _ln(p, name, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), p[name + ".weight"], p[name + ".bias"], eps)


def _encoder(p, pre, x, n_layers, heads, causal):
    B, S, D = x.shape
    hd = D // heads
    for i in range(n_layers):
        l = f"{pre}.encoder.layers.{i}"
        h = _ln(p, l + ".layer_norm1", x)
        q, k, v = (F.linear(h, p[f"{l}.self_attn.{n}_proj.weight"], p[f"{l}.self_attn.{n}_proj.bias"])
                   .view(B, S, heads, hd).transpose(1, 2) for n in "qkv")
        o = F.scaled_dot_product_attention(q, k, v, is_causal=causal)
        o = o.transpose(1, 2).reshape(B, S, D)
        x = x + F.linear(o, p[l + ".self_attn.out_proj.weight"], p[l + ".self_attn.out_proj.bias"])
        h = _ln(p, l + ".layer_norm2", x)
        h = F.gelu(F.linear(h, p[l + ".mlp.fc1.weight"], p[l + ".mlp.fc1.bias"]))
        x = x + F.linear(h, p[l + ".mlp.fc2.weight"], p[l + ".mlp.fc2.bias"])
    return x


def image_features(p, cfg, pixel_values, return_tokens=False):
    """PickScoreScorer: `model.get_image_features(pixel_values=)` (pickscore_scorer.py:40-41; transformers CLIPModel vision tower + visual_projection)."""
    pre = "vision_model"
    x = F.conv2d(pixel_values, p[pre + ".embeddings.patch_embedding.weight"], stride=cfg["patch"])
    x = x.flatten(2).transpose(1, 2)
    cls = p[pre + ".embeddings.class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], 1) + p[pre + ".embeddings.position_embedding.weight"][None]
    x = _ln(p, pre + ".pre_layrnorm", x)
    x = _encoder(p, pre, x, cfg["v_layers"], cfg["v_heads"], False)
    pooled = _ln(p, pre + ".post_layernorm", x[:, 0])
    out = F.linear(pooled, p["visual_projection.weight"])
    return (out, x) if return_tokens else out


def text_features(p, cfg, input_ids):
    """PickScoreScorer: `model.get_text_features(input_ids=)` (pickscore_scorer.py:42-43; causal text tower, EOS pooling, text_projection)."""
    pre = "text_model"
    S = input_ids.shape[1]
    x = p[pre + ".embeddings.token_embedding.weight"][input_ids] + \
        p[pre + ".embeddings.position_embedding.weight"][:S][None]
    x = _encoder(p, pre, x, cfg["t_layers"], cfg["t_heads"], True)
    x = _ln(p, pre + ".final_layer_norm", x)
    pooled = x[torch.arange(x.shape[0]), input_ids.argmax(-1)]       # EOS (highest id) pooling
    return F.linear(pooled, p["text_projection.weight"])


def pickscore(p, cfg, input_ids, pixel_values):
    """pickscore_scorer.py:40-51: diag of exp(logit_scale) * T @ I^T, / 26."""
    ie = image_features(p, cfg, pixel_values)
    ie = ie / ie.norm(p=2, dim=-1, keepdim=True)
    te = text_features(p, cfg, input_ids)
    te = te / te.norm(p=2, dim=-1, keepdim=True)
    scores = p["logit_scale"].exp() * (te @ ie.T)
    return scores.diag() / 26

"""Oracle: the text-encoding step of the sampling loop (`compute_text_embeddings`,
`scripts/train_sd3_fast_pickscore.py:186-193` -> `encode_prompt`,
`adv_grpo/diffusers_patch/train_dreambooth_lora_sd3.py:13-144`): two CLIP text encoders with projection
(`hidden_states[-2]` + pooled `text_embeds`, :57-93) and the T5 encoder (`text_encoder(ids)[0]`, :13-55),
concatenated as `cat([pad(cat([clip_l, clip_g], -1), 4096), t5], -2)` / `cat([pooled_l, pooled_g], -1)`
(:112-143).  The encoder bodies are transformers' `CLIPTextModelWithProjection` / `T5EncoderModel`
(transformers==4.54.0 pinned by the reference, setup.py:12) restated from their architecture with transformers
state-dict names; cross-checked against the installed transformers 5.5 classes in tests/test_oracle_models.py.
torch fp32 on CPU.  Test infrastructure only (see oracle/__init__.py)."""
import math

import torch
import torch.nn.functional as F


def _ln(p, name, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), p[name + ".weight"], p[name + ".bias"], eps)


def quick_gelu(x):
    """transformers QuickGELUActivation of the CLIP-L text encoder (encode_prompt, train_dreambooth_lora_sd3.py:59-95)."""
    return x * torch.sigmoid(1.702 * x)


def clip_text_with_projection(p, cfg, input_ids):
    """-> (text_embeds [B, proj], hidden_states: list of L+1 tensors [B, S, W]).  cfg: layers, heads, act, eos_id."""
    pre = "text_model"
    B, S = input_ids.shape
    x = p[pre + ".embeddings.token_embedding.weight"][input_ids] + \
        p[pre + ".embeddings.position_embedding.weight"][:S][None]
    heads = cfg["heads"]
    hd = x.shape[-1] // heads
    act = quick_gelu if cfg["act"] == "quick_gelu" else F.gelu
    hidden = [x]
    for i in range(cfg["layers"]):
        l = f"{pre}.encoder.layers.{i}"
        h = _ln(p, l + ".layer_norm1", x)
        q, k, v = (F.linear(h, p[f"{l}.self_attn.{n}_proj.weight"], p[f"{l}.self_attn.{n}_proj.bias"])
                   .view(B, S, heads, hd).transpose(1, 2) for n in "qkv")
        o = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, S, -1)
        x = x + F.linear(o, p[l + ".self_attn.out_proj.weight"], p[l + ".self_attn.out_proj.bias"])
        h = act(F.linear(_ln(p, l + ".layer_norm2", x), p[l + ".mlp.fc1.weight"], p[l + ".mlp.fc1.bias"]))
        x = x + F.linear(h, p[l + ".mlp.fc2.weight"], p[l + ".mlp.fc2.bias"])
        hidden.append(x)
    last = _ln(p, pre + ".final_layer_norm", x)
    if cfg.get("eos_id", 2) == 2:                     # legacy configs: the highest token id is the EOS token
        pos = input_ids.argmax(-1)
    else:                                             # first occurrence of eos_token_id
        pos = (input_ids == cfg["eos_id"]).int().argmax(-1)
    pooled = last[torch.arange(B), pos]
    return F.linear(pooled, p["text_projection.weight"]), hidden


def t5_relative_position_bucket(relative_position, num_buckets=32, max_distance=128):
    """transformers T5Attention._relative_position_bucket (bidirectional) behind _encode_prompt_with_t5 (train_dreambooth_lora_sd3.py:19-56).
    transformers T5Attention._relative_position_bucket, bidirectional=True."""
    num_buckets //= 2
    ret = (relative_position > 0).long() * num_buckets
    n = relative_position.abs()
    max_exact = num_buckets // 2
    is_small = n < max_exact
    large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).long()
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return ret + torch.where(is_small, n, large)


def t5_position_bias(p, S, num_buckets=32, max_distance=128):
    """transformers T5Attention.compute_bias of block 0, shared by all blocks (train_dreambooth_lora_sd3.py:19-56).
    [H, S, S] additive score bias shared by every layer (block 0 owns the embedding)."""
    ctx = torch.arange(S)[:, None]
    mem = torch.arange(S)[None, :]
    bucket = t5_relative_position_bucket(mem - ctx, num_buckets, max_distance)
    w = p["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]   # [num_buckets, H]
    return w[bucket].permute(2, 0, 1).contiguous()


def _t5_norm(x, w, eps=1e-6):
    var = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def t5_encoder(p, cfg, input_ids):
    """T5 v1.1 encoder stack (gated-GELU feed-forward, RMS layer norm, un-scaled attention scores with the
    relative-position bias, no attention mask: the reference passes none, train_dreambooth_lora_sd3.py:45).
    -> last_hidden_state [B, S, d_model]."""
    x = p["shared.weight"][input_ids]
    B, S, _ = x.shape
    H, dk = cfg["heads"], cfg["d_kv"]
    bias = t5_position_bias(p, S, cfg.get("num_buckets", 32), cfg.get("max_distance", 128))
    for i in range(cfg["layers"]):
        b = f"encoder.block.{i}"
        a = b + ".layer.0.SelfAttention"
        h = _t5_norm(x, p[b + ".layer.0.layer_norm.weight"])
        q, k, v = (F.linear(h, p[f"{a}.{n}.weight"]).view(B, S, H, dk).transpose(1, 2) for n in "qkv")
        scores = q @ k.transpose(-1, -2) + bias[None]
        o = (scores.float().softmax(-1).to(v.dtype) @ v).transpose(1, 2).reshape(B, S, H * dk)
        x = x + F.linear(o, p[a + ".o.weight"])
        f = b + ".layer.1.DenseReluDense"
        h = _t5_norm(x, p[b + ".layer.1.layer_norm.weight"])
        g = F.gelu(F.linear(h, p[f + ".wi_0.weight"]), approximate="tanh") * F.linear(h, p[f + ".wi_1.weight"])
        x = x + F.linear(g, p[f + ".wo.weight"])
    return _t5_norm(x, p["encoder.final_layer_norm.weight"])


def encode_prompt(clip_l, clip_g, t5, ids_l, ids_g, ids_t5):
    """train_dreambooth_lora_sd3.py:96-144 with the tokenisation factored out (`text_input_ids_list`).
    clip_l / clip_g / t5: (params, cfg).  -> (prompt_embeds [B, 77 + S_t5, d_t5], pooled [B, proj_l + proj_g])."""
    embeds, pooled = [], []
    for (p, cfg), ids in ((clip_l, ids_l), (clip_g, ids_g)):
        te, hidden = clip_text_with_projection(p, cfg, ids)
        embeds.append(hidden[-2])                                       # :81
        pooled.append(te)                                               # :80
    clip_embeds = torch.cat(embeds, dim=-1)                             # :125
    pooled = torch.cat(pooled, dim=-1)                                  # :126
    t5_embeds = t5_encoder(t5[0], t5[1], ids_t5)                        # :128-136
    clip_embeds = F.pad(clip_embeds, (0, t5_embeds.shape[-1] - clip_embeds.shape[-1]))   # :138-140
    return torch.cat([clip_embeds, t5_embeds], dim=-2), pooled          # :141

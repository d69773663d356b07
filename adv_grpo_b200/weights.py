"""Architecture configs and seeded random initialisation (true shapes, diffusers / transformers /
timm state-dict names) for the four bodies on the hot path.  No pretrained weights exist on the
box (no network): benchmarks and parity tests run on these seeded tensors; real checkpoints load
through the same names.

Init scales are HF-like (normal 0.02, RMSNorm weights ~1) with small NON-zero adaLN gates so
that every branch carries signal (zero-init gates would hide the attention/MLP paths).
"""
import math

import torch

SD35_MEDIUM = dict(num_layers=24, heads=24, head_dim=64, dual_layers=tuple(range(13)), qk_norm=True,
                   patch_size=2, in_channels=16, pos_embed_max_size=384, base_size=64,
                   joint_dim=4096, pooled_dim=2048)
SD3_MEDIUM = dict(SD35_MEDIUM, dual_layers=(), qk_norm=False, pos_embed_max_size=192)
MMDIT_TINY = dict(num_layers=3, heads=4, head_dim=64, dual_layers=(0, 1), qk_norm=True, patch_size=2,
                  in_channels=16, pos_embed_max_size=96, base_size=16, joint_dim=128, pooled_dim=64)

CLIP_H = dict(patch=14, image=224, v_width=1280, v_layers=32, v_heads=16, v_mlp=5120, t_width=1024,
              t_layers=24, t_heads=16, t_mlp=4096, vocab=49408, ctx=77, proj=1024)
# (tiny test configs keep every width a multiple of 256 and every VAE channel count a multiple of 128: the shapes the
#  native LayerNorm / GroupNorm kernels accept -- there is no PyTorch fallback for other shapes)
CLIP_TINY = dict(patch=14, image=224, v_width=256, v_layers=2, v_heads=4, v_mlp=512, t_width=256,
                 t_layers=2, t_heads=4, t_mlp=512, vocab=1000, ctx=77, proj=64)
DINOV2_B = dict(patch=14, image=518, width=768, layers=12, heads=12, mlp=3072)
DINOV2_TINY = dict(patch=14, image=518, width=256, layers=2, heads=4, mlp=512)
VAE_SD3 = dict(latent_channels=16, block_out=(128, 256, 512, 512), layers_per_block=2)
VAE_TINY = dict(latent_channels=16, block_out=(128, 128, 256, 256), layers_per_block=2)

# text encoders of stabilityai/stable-diffusion-3.5-medium (text_encoder / text_encoder_2 / text_encoder_3)
CLIP_L_TEXT = dict(width=768, layers=12, heads=12, mlp=3072, act="quick_gelu", vocab=49408, ctx=77, proj=768, eos_id=2)
CLIP_G_TEXT = dict(width=1280, layers=32, heads=20, mlp=5120, act="gelu", vocab=49408, ctx=77, proj=1280, eos_id=2)
T5_XXL = dict(d_model=4096, layers=24, heads=64, d_kv=64, d_ff=10240, vocab=32128, num_buckets=32, max_distance=128)
CLIP_L_TEXT_TINY = dict(width=256, layers=3, heads=4, mlp=512, act="quick_gelu", vocab=1000, ctx=77, proj=256, eos_id=2)
CLIP_G_TEXT_TINY = dict(width=512, layers=3, heads=8, mlp=768, act="gelu", vocab=1000, ctx=77, proj=512, eos_id=2)
T5_TINY = dict(d_model=512, layers=2, heads=4, d_kv=64, d_ff=640, vocab=500, num_buckets=32, max_distance=128)

LORA_TARGETS = ("attn.to_q", "attn.to_k", "attn.to_v", "attn.to_out.0",
                "attn.add_q_proj", "attn.add_k_proj", "attn.add_v_proj", "attn.to_add_out")


class _Init:
    def __init__(self, seed, device, dtype):
        self.meta = str(device) == "meta"                 # shape-only inventory (parameter-count tests), no storage
        self.g = None if self.meta else torch.Generator(device=device).manual_seed(seed)
        self.device, self.dtype = device, dtype
        self.p = {}

    def normal(self, name, shape, std=0.02, mean=0.0):
        if self.meta:
            self.p[name] = torch.empty(shape, device="meta", dtype=self.dtype)
            return
        t = torch.randn(shape, generator=self.g, device=self.device, dtype=torch.float32) * std + mean
        self.p[name] = t.to(self.dtype)

    def linear(self, name, out_f, in_f, bias=True, std=0.02):
        self.normal(name + ".weight", (out_f, in_f), std)
        if bias:
            self.normal(name + ".bias", (out_f,), 0.02)


def init_mmdit(cfg, seed=0, device="cpu", dtype=torch.bfloat16):
    d = cfg["heads"] * cfg["head_dim"]
    it = _Init(seed, device, dtype)
    ps, cin = cfg["patch_size"], cfg["in_channels"]
    it.normal("pos_embed.proj.weight", (d, cin, ps, ps), 1.0 / math.sqrt(cin * ps * ps))
    it.normal("pos_embed.proj.bias", (d,), 0.02)
    it.linear("time_text_embed.timestep_embedder.linear_1", d, 256, std=1.0 / 16)
    it.linear("time_text_embed.timestep_embedder.linear_2", d, d)
    it.linear("time_text_embed.text_embedder.linear_1", d, cfg["pooled_dim"])
    it.linear("time_text_embed.text_embedder.linear_2", d, d)
    it.linear("context_embedder", d, cfg["joint_dim"], std=1.0 / math.sqrt(cfg["joint_dim"]))
    L = cfg["num_layers"]
    for i in range(L):
        b = f"transformer_blocks.{i}"
        last, dual = i == L - 1, i in cfg["dual_layers"]
        it.linear(f"{b}.norm1.linear", (9 if dual else 6) * d, d)
        it.linear(f"{b}.norm1_context.linear", (2 if last else 6) * d, d)
        for n in ("to_q", "to_k", "to_v", "to_out.0", "add_q_proj", "add_k_proj", "add_v_proj"):
            it.linear(f"{b}.attn.{n}", d, d)
        if not last:
            it.linear(f"{b}.attn.to_add_out", d, d)
        if cfg["qk_norm"]:
            for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
                it.normal(f"{b}.attn.{n}.weight", (cfg["head_dim"],), 0.1, 1.0)
        if dual:
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                it.linear(f"{b}.attn2.{n}", d, d)
            if cfg["qk_norm"]:
                for n in ("norm_q", "norm_k"):
                    it.normal(f"{b}.attn2.{n}.weight", (cfg["head_dim"],), 0.1, 1.0)
        it.linear(f"{b}.ff.net.0.proj", 4 * d, d)
        it.linear(f"{b}.ff.net.2", d, 4 * d, std=0.01)
        if not last:
            it.linear(f"{b}.ff_context.net.0.proj", 4 * d, d)
            it.linear(f"{b}.ff_context.net.2", d, 4 * d, std=0.01)
    it.linear("norm_out.linear", 2 * d, d)
    it.linear("proj_out", ps * ps * cin, d, std=1.0 / math.sqrt(d))
    return it.p


def init_lora(cfg, rank=32, seed=1, device="cpu", perturb_b=0.0):
    """peft LoraConfig(r, init_lora_weights="gaussian"): A ~ N(0, 1/r), B = 0
    (train_sd3_fast_pickscore.py:500-505).  `perturb_b` emulates a few optimiser steps (B != 0)."""
    d = cfg["heads"] * cfg["head_dim"]
    g = torch.Generator(device=device).manual_seed(seed)
    lora = {}
    L = cfg["num_layers"]
    for i in range(L):
        for t in LORA_TARGETS:
            if t == "attn.to_add_out" and i == L - 1:
                continue
            a = torch.randn((rank, d), generator=g, device=device) / rank
            b = torch.randn((d, rank), generator=g, device=device) * perturb_b
            lora[f"transformer_blocks.{i}.{t}"] = (a, b)
    return lora


def init_vae_decoder(cfg=VAE_SD3, seed=2, device="cpu", dtype=torch.float32):
    it = _Init(seed, device, dtype)
    ch = cfg["block_out"]
    top = ch[-1]

    def conv(name, cout, cin, k=3):
        it.normal(name + ".weight", (cout, cin, k, k), 1.0 / math.sqrt(cin * k * k))
        it.normal(name + ".bias", (cout,), 0.02)

    def gn(name, c):
        it.normal(name + ".weight", (c,), 0.05, 1.0)
        it.normal(name + ".bias", (c,), 0.05)

    def resnet(name, cin, cout):
        gn(name + ".norm1", cin)
        conv(name + ".conv1", cout, cin)
        gn(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout)
        if cin != cout:
            conv(name + ".conv_shortcut", cout, cin, 1)

    conv("decoder.conv_in", top, cfg["latent_channels"])
    resnet("decoder.mid_block.resnets.0", top, top)
    gn("decoder.mid_block.attentions.0.group_norm", top)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        it.linear(f"decoder.mid_block.attentions.0.{n}", top, top, std=1.0 / math.sqrt(top))
    resnet("decoder.mid_block.resnets.1", top, top)
    prev = top
    for i, c in enumerate(reversed(ch)):
        for j in range(cfg["layers_per_block"] + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else c, c)
        prev = c
        if i < len(ch) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c)
    gn("decoder.conv_norm_out", ch[0])
    conv("decoder.conv_out", 3, ch[0])
    return it.p


def init_clip(cfg=CLIP_H, seed=3, device="cpu", dtype=torch.bfloat16):
    it = _Init(seed, device, dtype)

    def ln(name, d):
        it.normal(name + ".weight", (d,), 0.05, 1.0)
        it.normal(name + ".bias", (d,), 0.02)

    def tower(pre, width, layers, mlp):
        for i in range(layers):
            l = f"{pre}.encoder.layers.{i}"
            ln(l + ".layer_norm1", width)
            for n in "qkvo":
                it.linear(f"{l}.self_attn.{'out' if n == 'o' else n}_proj", width, width, std=width ** -0.5)
            ln(l + ".layer_norm2", width)
            it.linear(l + ".mlp.fc1", mlp, width, std=width ** -0.5)
            it.linear(l + ".mlp.fc2", width, mlp, std=0.5 * mlp ** -0.5)

    vw, tw = cfg["v_width"], cfg["t_width"]
    n_tok = (cfg["image"] // cfg["patch"]) ** 2 + 1
    it.normal("vision_model.embeddings.class_embedding", (vw,), 0.5)
    it.normal("vision_model.embeddings.patch_embedding.weight", (vw, 3, cfg["patch"], cfg["patch"]),
              1.0 / math.sqrt(3 * cfg["patch"] ** 2))
    it.normal("vision_model.embeddings.position_embedding.weight", (n_tok, vw), 0.3)
    ln("vision_model.pre_layrnorm", vw)
    tower("vision_model", vw, cfg["v_layers"], cfg["v_mlp"])
    ln("vision_model.post_layernorm", vw)
    it.normal("visual_projection.weight", (cfg["proj"], vw), vw ** -0.5)
    it.normal("text_model.embeddings.token_embedding.weight", (cfg["vocab"], tw), 0.5)
    it.normal("text_model.embeddings.position_embedding.weight", (cfg["ctx"], tw), 0.3)
    tower("text_model", tw, cfg["t_layers"], cfg["t_mlp"])
    ln("text_model.final_layer_norm", tw)
    it.normal("text_projection.weight", (cfg["proj"], tw), tw ** -0.5)
    it.p["logit_scale"] = torch.tensor(math.log(100.0), device=device, dtype=dtype)
    return it.p


def init_dinov2(cfg=DINOV2_B, seed=4, device="cpu", dtype=torch.bfloat16):
    it = _Init(seed, device, dtype)
    w = cfg["width"]
    n_tok = (cfg["image"] // cfg["patch"]) ** 2 + 1
    it.normal("patch_embed.proj.weight", (w, 3, cfg["patch"], cfg["patch"]), 1.0 / math.sqrt(3 * cfg["patch"] ** 2))
    it.normal("patch_embed.proj.bias", (w,), 0.02)
    it.normal("cls_token", (1, 1, w), 0.5)
    it.normal("pos_embed", (1, n_tok, w), 0.3)
    for i in range(cfg["layers"]):
        b = f"blocks.{i}"
        for n in ("norm1", "norm2"):
            it.normal(f"{b}.{n}.weight", (w,), 0.05, 1.0)
            it.normal(f"{b}.{n}.bias", (w,), 0.02)
        it.linear(f"{b}.attn.qkv", 3 * w, w, std=w ** -0.5)
        it.linear(f"{b}.attn.proj", w, w, std=w ** -0.5)
        it.normal(f"{b}.ls1.gamma", (w,), 0.05, 0.3)
        it.linear(f"{b}.mlp.fc1", cfg["mlp"], w, std=w ** -0.5)
        it.linear(f"{b}.mlp.fc2", w, cfg["mlp"], std=0.5 * cfg["mlp"] ** -0.5)
        it.normal(f"{b}.ls2.gamma", (w,), 0.05, 0.3)
    it.normal("norm.weight", (w,), 0.05, 1.0)
    it.normal("norm.bias", (w,), 0.02)
    return it.p


def init_dino_head(in_dim=768, hidden=512, seed=5, device="cpu", dtype=torch.float32):
    """DINOHead (train_sd3_fast_dino_patch.py:592-603): Linear -> GELU -> Linear(1)."""
    it = _Init(seed, device, dtype)
    it.linear("layers.0", hidden, in_dim, std=in_dim ** -0.5)
    it.linear("layers.2", 1, hidden, std=hidden ** -0.5)
    return it.p


def init_clip_text(cfg, seed=5, device="cpu", dtype=torch.bfloat16):
    """transformers `CLIPTextModelWithProjection` state dict (random init at the given sizes)."""
    it = _Init(seed, device, dtype)
    w = cfg["width"]
    it.normal("text_model.embeddings.token_embedding.weight", (cfg["vocab"], w), 0.02)
    it.normal("text_model.embeddings.position_embedding.weight", (cfg["ctx"], w), 0.01)
    for i in range(cfg["layers"]):
        l = f"text_model.encoder.layers.{i}"
        for n in ("layer_norm1", "layer_norm2"):
            it.normal(f"{l}.{n}.weight", (w,), 0.05, 1.0)
            it.normal(f"{l}.{n}.bias", (w,), 0.02)
        for n in "qkv":
            it.linear(f"{l}.self_attn.{n}_proj", w, w, std=w ** -0.5)
        it.linear(f"{l}.self_attn.out_proj", w, w, std=w ** -0.5)
        it.linear(f"{l}.mlp.fc1", cfg["mlp"], w, std=w ** -0.5)
        it.linear(f"{l}.mlp.fc2", w, cfg["mlp"], std=cfg["mlp"] ** -0.5)
    it.normal("text_model.final_layer_norm.weight", (w,), 0.05, 1.0)
    it.normal("text_model.final_layer_norm.bias", (w,), 0.02)
    it.normal("text_projection.weight", (cfg["proj"], w), w ** -0.5)
    return it.p


def init_t5_encoder(cfg, seed=6, device="cpu", dtype=torch.bfloat16):
    """transformers `T5EncoderModel` (v1.1, gated-gelu) state dict (random init at the given sizes)."""
    it = _Init(seed, device, dtype)
    d, inner = cfg["d_model"], cfg["heads"] * cfg["d_kv"]
    it.normal("shared.weight", (cfg["vocab"], d), 1.0)
    it.normal("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight", (cfg["num_buckets"], cfg["heads"]), 0.5)
    for i in range(cfg["layers"]):
        b = f"encoder.block.{i}"
        for n in "qkv":
            it.normal(f"{b}.layer.0.SelfAttention.{n}.weight", (inner, d), (d ** -0.5) * (cfg["d_kv"] ** -0.25))
        it.normal(f"{b}.layer.0.SelfAttention.o.weight", (d, inner), inner ** -0.5)
        it.normal(f"{b}.layer.0.layer_norm.weight", (d,), 0.05, 1.0)
        it.normal(f"{b}.layer.1.layer_norm.weight", (d,), 0.05, 1.0)
        it.normal(f"{b}.layer.1.DenseReluDense.wi_0.weight", (cfg["d_ff"], d), d ** -0.5)
        it.normal(f"{b}.layer.1.DenseReluDense.wi_1.weight", (cfg["d_ff"], d), d ** -0.5)
        it.normal(f"{b}.layer.1.DenseReluDense.wo.weight", (d, cfg["d_ff"]), cfg["d_ff"] ** -0.5)
    it.normal("encoder.final_layer_norm.weight", (d,), 0.05, 1.0)
    return it.p

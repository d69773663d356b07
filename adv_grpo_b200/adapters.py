"""From the objects the UNMODIFIED reference scripts build to the B200 path.

`scripts/train_sd3_fast_pickscore.py:447-449` creates the models with `diffusers.StableDiffusion3Pipeline.
from_pretrained`, `:488-511` wraps the transformer with peft (`get_peft_model` / `PeftModel.from_pretrained`) and
`:663` hands it to `accelerator.prepare`.  Two ways onto the kernels of this package:

  * `from_diffusers(pipeline)`: convert a REAL diffusers pipeline (diffusers / peft installed, weights on disk) into
    `adv_grpo_b200.pipeline.StableDiffusion3Pipeline` from its state dicts -- peft-wrapped parameter names
    (`base_model.model.<m>.base_layer.weight`, `...lora_A.default.weight`) included;
  * `adv_grpo_b200.shims`: `diffusers` / `peft` / `accelerate` / `ml_collections`-shaped modules whose classes ARE this
    package's objects, for boxes where those libraries are absent (`install_as_adv_grpo(shim_third_party=True)`).

Host-side conversion only (state-dict key handling, shape inference); nothing here is on the hot path.
"""
import math
import re

import torch

from . import weights
from .mmdit import SD3Transformer2DModel
from .vae import AutoencoderKL

_PEFT_PREFIX = "base_model.model."


def split_peft_state_dict(sd):
    """{name: tensor} of a (possibly peft-wrapped) module -> (plain state dict, {module: (A [r, in], B [out, r])}).
    Handles `base_model.model.` prefixes, `.base_layer.` infixes and `lora_{A,B}[.<adapter>].weight` keys."""
    base, la, lb = {}, {}, {}
    for k, v in sd.items():
        if k.startswith(_PEFT_PREFIX):
            k = k[len(_PEFT_PREFIX):]
        m = re.match(r"(.*)\.lora_([AB])(?:\.[^.]+)?\.weight$", k)
        if m:
            (la if m.group(2) == "A" else lb)[m.group(1)] = v
            continue
        base[k.replace(".base_layer.", ".")] = v
    lora = {n: (la[n], lb[n]) for n in la if n in lb}
    return base, (lora or None)


def mmdit_config_from_state_dict(sd):
    """Infer the `weights.SD35_MEDIUM`-style config dict from the diffusers `SD3Transformer2DModel` parameter shapes."""
    w = sd["pos_embed.proj.weight"]
    d, cin, ps = w.shape[0], w.shape[1], w.shape[2]
    layers = 1 + max(int(m.group(1)) for k in sd for m in [re.match(r"transformer_blocks\.(\d+)\.", k)] if m)
    head_dim = sd["transformer_blocks.0.attn.norm_q.weight"].shape[0] if "transformer_blocks.0.attn.norm_q.weight" in sd else 64
    dual = tuple(i for i in range(layers) if f"transformer_blocks.{i}.attn2.to_q.weight" in sd)
    if "pos_embed.pos_embed" in sd:
        max_size = int(round(math.sqrt(sd["pos_embed.pos_embed"].shape[1])))
    else:
        max_size = 384 if dual else 192
    return dict(num_layers=layers, heads=d // head_dim, head_dim=head_dim, dual_layers=dual,
                qk_norm="transformer_blocks.0.attn.norm_q.weight" in sd, patch_size=ps, in_channels=cin,
                pos_embed_max_size=max_size, base_size={384: 64, 192: 64}.get(max_size, max(1, max_size // 6)),
                joint_dim=sd["context_embedder.weight"].shape[1],
                pooled_dim=sd["time_text_embed.text_embedder.linear_1.weight"].shape[1])


def transformer_from_state_dict(sd, device="cuda", lora_rank=0, lora_alpha=64, cfg=None):
    """diffusers / peft state dict -> `SD3Transformer2DModel` on the kernels.  LoRA factors found in the dict win over
    `lora_rank` (their rank is read from the shapes)."""
    base, lora = split_peft_state_dict(sd)
    base.pop("pos_embed.pos_embed", None)                      # sincos table: recomputed, not a weight
    cfg = cfg or mmdit_config_from_state_dict(dict(base, **({"pos_embed.pos_embed": sd["pos_embed.pos_embed"]}
                                                            if "pos_embed.pos_embed" in sd else {})))
    if lora:
        lora_rank = next(iter(lora.values()))[0].shape[0]
        lora = {k: (a.float(), b.float()) for k, (a, b) in lora.items()}
    return SD3Transformer2DModel(cfg, base, lora_rank=lora_rank, lora_alpha=lora_alpha, lora=lora, device=device)


def vae_config_from_state_dict(sd):
    n_up = 1 + max(int(m.group(1)) for k in sd for m in [re.match(r"decoder\.up_blocks\.(\d+)\.", k)] if m)
    outs = [sd[f"decoder.up_blocks.{i}.resnets.0.conv1.weight"].shape[0] for i in range(n_up)]
    per = 1 + max(int(m.group(1)) for k in sd for m in [re.match(r"decoder\.up_blocks\.0\.resnets\.(\d+)\.", k)] if m)
    return dict(latent_channels=sd["decoder.conv_in.weight"].shape[1], block_out=tuple(reversed(outs)),
                layers_per_block=per - 1)


def vae_from_state_dict(sd, device="cuda", cfg=None):
    """diffusers `AutoencoderKL` state dict -> decoder on the kernels (the encoder is not on the path)."""
    dec = {k: v for k, v in sd.items() if k.startswith("decoder.")}
    return AutoencoderKL(dec, cfg or vae_config_from_state_dict(dec), device=device)


def from_diffusers(pipeline, device="cuda", use_cuda_graph=True):
    """`StableDiffusion3Pipeline` (diffusers; transformer optionally peft-wrapped) -> this package's pipeline with the
    same weights: MMDiT and VAE decoder on the sm_100a kernels, CLIP-L / CLIP-G / T5 text encoders when present.
    The call a maintainer adds after `train_pick:511`:  `pipeline = adv_grpo_b200.from_diffusers(pipeline)`."""
    from .pipeline import StableDiffusion3Pipeline
    tr = pipeline.transformer
    r, alpha = 0, 64
    pc = getattr(tr, "peft_config", None)
    if pc:
        c = pc.get("default", next(iter(pc.values())))
        r, alpha = int(c.r), float(c.lora_alpha)
    transformer = transformer_from_state_dict(tr.state_dict(), device=device, lora_rank=r, lora_alpha=alpha)
    vae = vae_from_state_dict(pipeline.vae.state_dict(), device=device)
    out = StableDiffusion3Pipeline(transformer, vae, tokenizer=getattr(pipeline, "tokenizer", None), device=device,
                                   use_cuda_graph=use_cuda_graph)
    from . import text_encoders as te
    for name, cls, cfg in (("text_encoder", te.CLIPTextModelWithProjection, weights.CLIP_L_TEXT),
                           ("text_encoder_2", te.CLIPTextModelWithProjection, weights.CLIP_G_TEXT),
                           ("text_encoder_3", te.T5EncoderModel, weights.T5_XXL)):
        enc = getattr(pipeline, name, None)
        setattr(out, name, cls(enc.state_dict(), cfg, device=device) if enc is not None else None)
    for name in ("tokenizer_2", "tokenizer_3"):
        setattr(out, name, getattr(pipeline, name, None))
    return out

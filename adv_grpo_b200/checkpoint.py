"""On-disk LoRA checkpoints in the peft adapter-directory layout the reference writes and reads:
`save_ckpt` (`scripts/train_sd3_fast_pickscore.py:389-398`: rank 0, EMA weights swapped in around
`peft.save_pretrained(<save_dir>/checkpoints/checkpoint-<step>/lora)`) and the resume path
`PeftModel.from_pretrained(transformer, config.train.lora_path)` (`:506-509`), so adapters trained here load
in `inference_t2i.py` / `app.py` and released adapters load here.

Layout: `adapter_config.json` (peft LoraConfig fields) + `adapter_model.safetensors` with keys
`base_model.model.<module>.lora_A.weight` [r, in] and `...lora_B.weight` [out, r] (peft drops the adapter
name "default" when saving).  Host-side I/O only; nothing here is on the hot path.
"""
import json
import os

import torch

ADAPTER_CONFIG = "adapter_config.json"
ADAPTER_WEIGHTS = "adapter_model.safetensors"


def lora_config_dict(rank, alpha, target_modules, base_model="stabilityai/stable-diffusion-3.5-medium"):
    """The fields peft 0.17 writes for `LoraConfig(r=32, lora_alpha=64, init_lora_weights="gaussian",
    target_modules=[...])` (`train_pick:500-505`)."""
    return {
        "peft_type": "LORA", "task_type": None, "auto_mapping": None, "base_model_name_or_path": base_model,
        "revision": None, "inference_mode": True, "r": int(rank), "lora_alpha": int(alpha), "lora_dropout": 0.0,
        "target_modules": sorted(target_modules), "init_lora_weights": "gaussian", "bias": "none",
        "fan_in_fan_out": False, "modules_to_save": None, "layers_to_transform": None, "layers_pattern": None,
        "rank_pattern": {}, "alpha_pattern": {}, "use_rslora": False, "use_dora": False,
    }


def save_adapter_dir(path, state_dict, config):
    """state_dict: {`base_model.model.<module>.lora_{A,B}.weight`: tensor}; config: dict for adapter_config.json."""
    from safetensors.torch import save_file
    os.makedirs(path, exist_ok=True)
    tensors = {k: v.detach().to("cpu").contiguous() for k, v in state_dict.items()}
    save_file(tensors, os.path.join(path, ADAPTER_WEIGHTS), metadata={"format": "pt"})
    with open(os.path.join(path, ADAPTER_CONFIG), "w") as f:
        json.dump(config, f, indent=2, sort_keys=True)


def load_adapter_dir(path):
    """Returns (state_dict, config).  Accepts both the saved key form and keys that still carry the adapter
    name (`...lora_A.default.weight`), and `adapter_model.bin` written by older peft versions."""
    with open(os.path.join(path, ADAPTER_CONFIG)) as f:
        config = json.load(f)
    st_path = os.path.join(path, ADAPTER_WEIGHTS)
    if os.path.exists(st_path):
        from safetensors.torch import load_file
        sd = load_file(st_path)
    else:
        sd = torch.load(os.path.join(path, "adapter_model.bin"), map_location="cpu", weights_only=True)
    return {k.replace(".lora_A.default.", ".lora_A.").replace(".lora_B.default.", ".lora_B."): v for k, v in sd.items()}, config


def save_full_transformer(transformer, path):
    """`transformer.save_pretrained(path)` of a fully fine-tuned model (`config.use_lora = False`): diffusers layout,
    `diffusion_pytorch_model.safetensors` with the diffusers parameter names (fp32 master weights)."""
    from safetensors.torch import save_file
    os.makedirs(path, exist_ok=True)
    tensors = {k: v.detach().to("cpu").contiguous() for k, v in transformer.full_state_dict().items()}
    save_file(tensors, os.path.join(path, "diffusion_pytorch_model.safetensors"), metadata={"format": "pt"})


def save_lora(transformer, path):
    """`peft_model.save_pretrained(path)` for `SD3Transformer2DModel` (fp32 master LoRA factors)."""
    targets = sorted({n.split(".", 2)[2] for n in transformer._lora_names})
    cfg = lora_config_dict(transformer.lora_rank, transformer.lora_rank * transformer.lora_scale, targets)
    save_adapter_dir(path, transformer.lora_state_dict(), cfg)


def load_lora(transformer, path, strict=True):
    """`PeftModel.from_pretrained(transformer, path)` + `set_adapter("default")`: copies the stored factors into the
    model's LoRA parameters (in place, so optimizers / captured CUDA graphs keep their tensors) and checks rank /
    alpha / coverage against the adapter config."""
    sd, cfg = load_adapter_dir(path)
    if int(cfg.get("r", transformer.lora_rank)) != transformer.lora_rank:
        raise ValueError(f"adapter rank {cfg.get('r')} != model LoRA rank {transformer.lora_rank}")
    scale = float(cfg.get("lora_alpha", 0)) / float(cfg.get("r", 1))
    if abs(scale - transformer.lora_scale) > 1e-9:
        raise ValueError(f"adapter scale alpha/r = {scale} != model LoRA scale {transformer.lora_scale}")
    own = transformer.lora_state_dict()
    missing = [k for k in own if k not in sd]
    unexpected = [k for k in sd if k not in own]
    if strict and (missing or unexpected):
        raise KeyError(f"adapter mismatch: missing {missing[:3]} (+{max(0, len(missing) - 3)}), "
                       f"unexpected {unexpected[:3]} (+{max(0, len(unexpected) - 3)})")
    with torch.no_grad():
        for k, dst in own.items():
            if k in sd:
                if tuple(sd[k].shape) != tuple(dst.shape):
                    raise ValueError(f"{k}: shape {tuple(sd[k].shape)} != {tuple(dst.shape)}")
                dst.copy_(sd[k].to(dst.device, dst.dtype))
    transformer.invalidate_lora_cache()
    return missing, unexpected


def save_ckpt(save_dir, transformer, global_step, ema=None, trainable_parameters=None, use_ema=False, is_main_process=True):
    """`save_ckpt` of `train_pick:389-398`: every rank creates the directory, rank 0 writes the adapter with the EMA
    weights swapped in (and back out).  Not saved, as in the reference: optimizer, EMA state, step, RNG, discriminator."""
    root = os.path.join(save_dir, "checkpoints", f"checkpoint-{global_step}", "lora")
    os.makedirs(root, exist_ok=True)
    if not is_main_process:
        return root
    if use_ema and ema is not None:
        ema.copy_ema_to(trainable_parameters, store_temp=True)
        transformer.invalidate_lora_cache()
    try:
        transformer.save_pretrained(root)            # the LoRA adapter directory, or the full weights under use_lora = False
    finally:
        if use_ema and ema is not None:
            ema.copy_temp_to(trainable_parameters)
            transformer.invalidate_lora_cache()
    return root

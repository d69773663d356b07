"""`diffusers` / `peft` / `accelerate` / `ml_collections`-shaped modules whose classes are THIS package's objects, so that
the model-construction lines of the unmodified reference scripts resolve to the B200 path on a box where those
libraries are not installed:

    train_sd3_fast_pickscore.py:447-449   StableDiffusion3Pipeline.from_pretrained(config.pretrained.model)
    train_sd3_fast_pickscore.py:488-511   LoraConfig / get_peft_model / PeftModel.from_pretrained(...).set_adapter
    train_sd3_fast_pickscore.py:415-425   Accelerator(mixed_precision=, project_config=, gradient_accumulation_steps=)
    train_sd3_fast_pickscore.py:663       accelerator.prepare(transformer, optimizer, train_dataloader, test_dataloader)
    train_sd3_fast_pickscore.py:43        config_flags.DEFINE_config_file("config", ...)

`install(force=False)` registers them in `sys.modules` ONLY for libraries that cannot be imported (a real install always
wins unless `force`).  The Accelerator is the thin `torch.distributed` equivalent of what the scripts use (one process
per GPU under torchrun: gather / reduce / accumulate / backward / clip_grad_norm_ / prepare / unwrap_model); DeepSpeed
ZeRO-2 sharding of the optimizer state is not reproduced (18.8 M LoRA parameters: 225 MB of AdamW state).
Host-side plumbing only; nothing here is on the hot path.
"""
import contextlib
import dataclasses
import importlib
import os
import sys
import types

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------ diffusers
_PIPELINE_FACTORY = None


def set_pipeline_factory(fn):
    """`fn(name_or_path, **kw) -> pipeline`: overrides how `from_pretrained` materialises weights (tests: tiny seeded
    models; deployments: a site-specific weight store)."""
    global _PIPELINE_FACTORY
    _PIPELINE_FACTORY = fn


def _load_component_state_dict(root, sub):
    from safetensors.torch import load_file
    d = os.path.join(root, sub)
    if not os.path.isdir(d):
        return None
    sd = {}
    for f in sorted(os.listdir(d)):
        if f.endswith(".safetensors"):
            sd.update(load_file(os.path.join(d, f)))
    return sd or None


def _pipeline_from_pretrained(cls, name_or_path, device=None, use_cuda_graph=True, **kw):
    from . import adapters, weights
    from . import text_encoders as te
    if _PIPELINE_FACTORY is not None:
        return _PIPELINE_FACTORY(name_or_path, **kw)
    device = device or (f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else "cpu")
    if os.path.isdir(str(name_or_path)):                       # a diffusers model directory (safetensors shards)
        tsd = _load_component_state_dict(name_or_path, "transformer")
        vsd = _load_component_state_dict(name_or_path, "vae")
        if tsd is None or vsd is None:
            raise FileNotFoundError(f"{name_or_path}: expected transformer/*.safetensors and vae/*.safetensors")
        pipe = cls(adapters.transformer_from_state_dict(tsd, device=device), adapters.vae_from_state_dict(vsd, device=device),
                   device=device, use_cuda_graph=use_cuda_graph)
        for sub, klass, cfg in (("text_encoder", te.CLIPTextModelWithProjection, weights.CLIP_L_TEXT),
                                ("text_encoder_2", te.CLIPTextModelWithProjection, weights.CLIP_G_TEXT),
                                ("text_encoder_3", te.T5EncoderModel, weights.T5_XXL)):
            sd = _load_component_state_dict(name_or_path, sub)
            setattr(pipe, sub, klass(sd, cfg, device=device) if sd is not None else None)
        return pipe
    if os.environ.get("ADVGRPO_SYNTHETIC_WEIGHTS", "0") == "1":    # no checkpoints on the box: seeded weights, true shapes
        cfg = weights.SD3_MEDIUM if "stable-diffusion-3-medium" in str(name_or_path) else weights.SD35_MEDIUM
        return cls.from_seed(cfg, weights.VAE_SD3, device=device, lora_rank=0,
                             use_cuda_graph=use_cuda_graph).add_seeded_text_encoders()
    raise FileNotFoundError(f"{name_or_path!r} is not a local diffusers model directory (there is no hub access); pass a "
                            "directory, call shims.set_pipeline_factory(...), or set ADVGRPO_SYNTHETIC_WEIGHTS=1")


def _make_diffusers():
    from .pipeline import StableDiffusion3Pipeline
    from .scheduler import FlowMatchEulerDiscreteScheduler
    from .vae import AutoencoderKL
    from .mmdit import SD3Transformer2DModel
    if not hasattr(StableDiffusion3Pipeline, "from_pretrained"):
        StableDiffusion3Pipeline.from_pretrained = classmethod(_pipeline_from_pretrained)
    m = types.ModuleType("diffusers")
    m.__path__ = []
    m.StableDiffusion3Pipeline = StableDiffusion3Pipeline
    m.FlowMatchEulerDiscreteScheduler = FlowMatchEulerDiscreteScheduler
    m.AutoencoderKL = AutoencoderKL
    m.SD3Transformer2DModel = SD3Transformer2DModel
    utils = types.ModuleType("diffusers.utils")
    utils.__path__ = []
    tu = types.ModuleType("diffusers.utils.torch_utils")
    tu.is_compiled_module = lambda module: False
    tu.randn_tensor = lambda shape, generator=None, device=None, dtype=None, layout=None: torch.randn(
        shape, generator=generator, device=device, dtype=dtype)
    utils.torch_utils = tu
    m.utils = utils
    return {"diffusers": m, "diffusers.utils": utils, "diffusers.utils.torch_utils": tu}


# ------------------------------------------------------------------------------------------------ peft
@dataclasses.dataclass
class LoraConfig:
    r: int = 8
    lora_alpha: float = 8
    init_lora_weights: object = True
    target_modules: object = None
    lora_dropout: float = 0.0
    bias: str = "none"


def get_peft_model(model, peft_config, adapter_name="default"):
    """`get_peft_model(pipeline.transformer, LoraConfig(r=32, lora_alpha=64, init_lora_weights="gaussian", ...))`
    (train_pick:500-511): the same frozen weights with a fresh flat LoRA parameter (A ~ N(0, 1/r), B = 0)."""
    from .mmdit import SD3Transformer2DModel
    if not isinstance(model, SD3Transformer2DModel):
        raise TypeError(f"get_peft_model shim: expected adv_grpo_b200 SD3Transformer2DModel, got {type(model)}")
    from .weights import LORA_TARGETS
    tm = peft_config.target_modules
    if tm is not None and set(tm) != set(LORA_TARGETS):
        raise NotImplementedError(f"LoRA target_modules {sorted(tm)} differ from the 8 attention projections the "
                                  f"scripts train ({sorted(LORA_TARGETS)})")
    new = SD3Transformer2DModel(model.cfg, model.p, lora_rank=int(peft_config.r), lora_alpha=float(peft_config.lora_alpha),
                                device=model.device_)
    new.peft_config = {adapter_name: peft_config}
    return new


class PeftModel:
    @staticmethod
    def from_pretrained(model, path, adapter_name="default", **kw):
        """`PeftModel.from_pretrained(pipeline.transformer, config.train.lora_path)` (train_pick:506-509)."""
        from .checkpoint import load_adapter_dir
        _, cfg = load_adapter_dir(path)
        new = get_peft_model(model, LoraConfig(r=int(cfg["r"]), lora_alpha=float(cfg["lora_alpha"]),
                                               target_modules=cfg.get("target_modules")), adapter_name)
        new.load_adapter(path)
        return new


def set_peft_model_state_dict(model, state_dict, adapter_name="default"):
    own = model.lora_state_dict()
    norm = {k.replace(f".lora_A.{adapter_name}.", ".lora_A.").replace(f".lora_B.{adapter_name}.", ".lora_B."): v
            for k, v in state_dict.items()}
    with torch.no_grad():
        for k, dst in own.items():
            src = norm.get(k, norm.get(k.replace("base_model.model.", "")))
            if src is not None:
                dst.copy_(src.to(dst.device, dst.dtype))
    model.invalidate_lora_cache()


def _make_peft():
    m = types.ModuleType("peft")
    m.LoraConfig, m.get_peft_model, m.PeftModel, m.set_peft_model_state_dict = (LoraConfig, get_peft_model, PeftModel,
                                                                               set_peft_model_state_dict)
    return {"peft": m}


# ------------------------------------------------------------------------------------------------ accelerate
@dataclasses.dataclass
class ProjectConfiguration:
    project_dir: str = None
    automatic_checkpoint_naming: bool = False
    total_limit: int = None


class _State:
    deepspeed_plugin = None


class Accelerator:
    """The subset of `accelerate.Accelerator` the two scripts touch, over `torch.distributed` (launch with torchrun)."""
    _last = None

    def __init__(self, mixed_precision="no", project_config=None, gradient_accumulation_steps=1, log_with=None, **kw):
        self.mixed_precision = mixed_precision
        self.project_config = project_config
        self.gradient_accumulation_steps = int(gradient_accumulation_steps)
        self.state = _State()
        if "RANK" in os.environ and not dist.is_initialized():
            dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
        self.process_index = dist.get_rank() if dist.is_initialized() else 0
        self.num_processes = dist.get_world_size() if dist.is_initialized() else 1
        self.local_process_index = int(os.environ.get("LOCAL_RANK", 0))
        if torch.cuda.is_available():
            torch.cuda.set_device(self.local_process_index)
            self.device = torch.device("cuda", self.local_process_index)
        else:
            self.device = torch.device("cpu")
        self._micro = 0
        self.sync_gradients = True
        self.prepared = []
        Accelerator._last = self              # (tests: the instance the script created)

    is_main_process = property(lambda self: self.process_index == 0)
    is_local_main_process = property(lambda self: self.local_process_index == 0)

    def prepare(self, *objs):
        self.prepared.extend(objs)
        return objs if len(objs) != 1 else objs[0]

    def unwrap_model(self, model, **kw):
        return getattr(model, "module", model)

    @contextlib.contextmanager
    def accumulate(self, *models):
        self._micro += 1
        self.sync_gradients = self._micro % self.gradient_accumulation_steps == 0
        yield

    def autocast(self):
        dt = {"bf16": torch.bfloat16, "fp16": torch.float16}.get(self.mixed_precision)
        return torch.autocast(self.device.type, dtype=dt) if dt is not None else contextlib.nullcontext()

    def backward(self, loss, **kw):
        (loss / self.gradient_accumulation_steps).backward(**kw)

    def clip_grad_norm_(self, parameters, max_norm, norm_type=2):
        params = [p for p in parameters if p.grad is not None]
        if self.num_processes > 1:
            for p in params:
                dist.all_reduce(p.grad)
                p.grad /= self.num_processes
        return torch.nn.utils.clip_grad_norm_(params, max_norm, norm_type)

    def gather(self, t):
        if self.num_processes == 1:
            return t
        out = torch.empty((self.num_processes * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous())
        return out

    def reduce(self, t, reduction="sum"):
        if self.num_processes == 1:
            return t
        t = t.clone()
        dist.all_reduce(t)
        return t / self.num_processes if reduction == "mean" else t

    def wait_for_everyone(self):
        if self.num_processes > 1:
            dist.barrier()

    def log(self, values, step=None):
        pass

    def save_state(self, *a, **kw):
        raise NotImplementedError("accelerator.save_state: the scripts checkpoint through save_ckpt (peft adapter dir)")

    def print(self, *a, **kw):
        if self.is_main_process:
            print(*a, **kw)


def _set_seed(seed, device_specific=False):
    import random

    import numpy as np
    if device_specific and dist.is_initialized():
        seed += dist.get_rank()
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def _make_accelerate():
    import logging
    m = types.ModuleType("accelerate")
    m.__path__ = []
    m.Accelerator = Accelerator
    u = types.ModuleType("accelerate.utils")
    u.set_seed, u.ProjectConfiguration = _set_seed, ProjectConfiguration
    lg = types.ModuleType("accelerate.logging")
    lg.get_logger = lambda name, log_level=None: logging.getLogger(name)
    m.utils, m.logging = u, lg
    return {"accelerate": m, "accelerate.utils": u, "accelerate.logging": lg}


# ------------------------------------------------------------------------------------------------ ml_collections
def _make_ml_collections():
    from .config import ConfigDict, load_config
    m = types.ModuleType("ml_collections")
    m.__path__ = []
    m.ConfigDict = ConfigDict
    cf = types.ModuleType("ml_collections.config_flags")

    def DEFINE_config_file(name, default=None, help_string="path to config file", **kw):
        """absl flag whose value is the ConfigDict returned by `<file>.get_config(<preset>)` (`file.py:preset`)."""
        from absl import flags

        class _Parser(flags.ArgumentParser):
            def parse(self, argument):
                if isinstance(argument, ConfigDict):
                    return argument
                path = str(argument).partition(":")[0]
                if path.endswith(".py") and not os.path.exists(path):
                    return argument        # the DEFAULT ("config/base.py", relative to the launch dir) is parsed at
                                           # definition time: left unloaded, as ml_collections defers it to first use
                return load_config(argument)

            def flag_type(self):
                return "config file"

        class _Serializer(flags.ArgumentSerializer):
            def serialize(self, value):
                return str(value)

        return flags.DEFINE(_Parser(), name, None, help_string, serializer=_Serializer(), **kw) if default is None else \
            flags.DEFINE(_Parser(), name, default, help_string, serializer=_Serializer(), **kw)

    cf.DEFINE_config_file = DEFINE_config_file
    m.config_flags = cf
    return {"ml_collections": m, "ml_collections.config_flags": cf}


def _importable(name):
    if name in sys.modules:
        return True
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def install(force=False):
    """Register the shim modules for every library of {diffusers, peft, accelerate, ml_collections} that is not
    installed (all of them with `force`).  Returns the list of shimmed top-level names."""
    done = []
    for top, make in (("diffusers", _make_diffusers), ("peft", _make_peft), ("accelerate", _make_accelerate),
                      ("ml_collections", _make_ml_collections)):
        if force or not _importable(top) or getattr(sys.modules.get(top), "__advgrpo_shim__", False):
            mods = make()
            for name, mod in mods.items():
                mod.__advgrpo_shim__ = True
                sys.modules[name] = mod
            done.append(top)
    return done

"""adv_grpo_b200 -- B200-native (sm_100a) implementation of the Adv-GRPO rollout -> score ->
advantage -> update hot path behind the reference's own Python surface.

`install_as_adv_grpo()` registers this package's modules under the reference's import names
(`adv_grpo.rewards`, `adv_grpo.stat_tracking`, `adv_grpo.ema`, `adv_grpo.ocr`, `adv_grpo.pickscore_scorer`,
`adv_grpo.pick_score_training`, `adv_grpo.diffusers_patch.sd3_sde_with_logprob`,
`adv_grpo.diffusers_patch.sd3_pipeline_with_logprob_fast`, `adv_grpo.diffusers_patch.train_dreambooth_lora_sd3`) so that
`scripts/train_sd3_fast_{pickscore,dino_patch}.py` import the B200 path unchanged.
"""
import importlib
import sys
import types

__version__ = "0.1.0"

_ALIASES = {
    "adv_grpo.rewards": "adv_grpo_b200.rewards",
    "adv_grpo.stat_tracking": "adv_grpo_b200.stat_tracking",
    "adv_grpo.ema": "adv_grpo_b200.ema",
    "adv_grpo.ocr": "adv_grpo_b200.ocr",
    "adv_grpo.pickscore_scorer": "adv_grpo_b200.pickscore_scorer",
    "adv_grpo.pick_score_training": "adv_grpo_b200.pick_score_training",
    "adv_grpo.diffusers_patch": "adv_grpo_b200.diffusers_patch",
    "adv_grpo.diffusers_patch.sd3_sde_with_logprob": "adv_grpo_b200.diffusers_patch.sd3_sde_with_logprob",
    "adv_grpo.diffusers_patch.sd3_pipeline_with_logprob_fast":
        "adv_grpo_b200.diffusers_patch.sd3_pipeline_with_logprob_fast",
    "adv_grpo.diffusers_patch.train_dreambooth_lora_sd3": "adv_grpo_b200.diffusers_patch.train_dreambooth_lora_sd3",
}


def from_diffusers(pipeline, **kw):
    """diffusers `StableDiffusion3Pipeline` (transformer optionally peft-wrapped) -> the B200 pipeline with the same
    weights (`adapters.from_diffusers`)."""
    from .adapters import from_diffusers as _f
    return _f(pipeline, **kw)


def install_as_adv_grpo(shim_third_party=False, force_shims=False):
    """Alias this package's modules as `adv_grpo.*`.  With `shim_third_party=True`, `diffusers` / `peft` / `accelerate` /
    `ml_collections` modules are additionally provided (`adv_grpo_b200.shims`) for every one of those libraries that is
    NOT installed, so `StableDiffusion3Pipeline.from_pretrained`, `get_peft_model` and `accelerator.prepare` of the
    unmodified scripts (train_sd3_fast_pickscore.py:447-449,500-511,663) build and return this package's objects."""
    if shim_third_party:
        from . import shims
        shims.install(force=force_shims)
    root = sys.modules.get("adv_grpo")
    if root is None:
        root = types.ModuleType("adv_grpo")
        root.__path__ = []
        sys.modules["adv_grpo"] = root
    for alias, target in _ALIASES.items():
        mod = importlib.import_module(target)
        sys.modules[alias] = mod
        parent, _, leaf = alias.rpartition(".")
        setattr(sys.modules[parent], leaf, mod)
    return root

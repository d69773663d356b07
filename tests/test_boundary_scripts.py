"""Row (b) of SURVEY.md section 8: the UNMODIFIED reference training script reaches this package's objects.

`install_as_adv_grpo(shim_third_party=True)` provides `adv_grpo.*` plus `diffusers` / `peft` / `accelerate` /
`ml_collections`-shaped modules (none of them is installed here); the test then executes the setup section of
`scripts/train_sd3_fast_pickscore.py` itself (main() from :400 up to the thread-pool creation at :668, i.e. through
`StableDiffusion3Pipeline.from_pretrained` :447, `get_peft_model` :511, `PickScoreScorer` :514, `EMAModuleWrapper` :528,
`multi_score` :556, `DistributedKRepeatSampler` :565, `compute_text_embeddings` :629, `DDP(scorer.model)` :657 and
`accelerator.prepare` :663) against tiny seeded weights on the CPU and checks what the script holds afterwards.
No kernel runs (CPU test); the objects' forward passes are covered by the GPU tests."""
import functools
import importlib.util
import os
import sys
import types

import pytest
import torch

REF_SCRIPT = "/root/reference/scripts/train_sd3_fast_pickscore.py"


class _Stop(Exception):
    pass


class _StubClipText:
    """transformers-convention CLIP text encoder stand-in (pure torch; the CPU test cannot run the kernels)."""

    def __init__(self, width, vocab=49408):
        g = torch.Generator().manual_seed(width)
        self.emb = torch.randn(vocab, width, generator=g) * 0.02
        self.device, self.dtype = torch.device("cpu"), torch.float32

    def requires_grad_(self, flag=True):
        return self

    def to(self, *a, **k):
        return self

    def __call__(self, ids, output_hidden_states=False):
        h = self.emb[ids]
        out = types.SimpleNamespace(hidden_states=[h, h * 0.5, h * 0.25])
        return _Out(h.mean(1), out.hidden_states)


class _Out(tuple):
    def __new__(cls, pooled, hidden_states):
        o = super().__new__(cls, (pooled,))
        o.hidden_states = hidden_states
        return o


class _StubT5(_StubClipText):
    def __call__(self, ids, **kw):
        return (self.emb[ids],)


def _tiny_factory(name_or_path, **kw):
    from adv_grpo_b200 import weights
    from adv_grpo_b200.mmdit import SD3Transformer2DModel
    from adv_grpo_b200.pickscore_scorer import SyntheticCLIPTokenizer
    from adv_grpo_b200.pipeline import StableDiffusion3Pipeline
    from adv_grpo_b200.vae import AutoencoderKL
    assert name_or_path == "stabilityai/stable-diffusion-3.5-medium"
    cfg = dict(weights.MMDIT_TINY, joint_dim=4096)                 # encode_prompt pads the CLIP features to the T5 width
    tr = SD3Transformer2DModel(cfg, weights.init_mmdit(cfg, seed=0), lora_rank=0, device="cpu")
    vae = AutoencoderKL(weights.init_vae_decoder(weights.VAE_TINY, seed=2), weights.VAE_TINY, device="cpu")
    pipe = StableDiffusion3Pipeline(tr, vae, device="cpu", use_cuda_graph=False)
    pipe.text_encoder, pipe.text_encoder_2, pipe.text_encoder_3 = _StubClipText(768), _StubClipText(1280), _StubT5(4096, 32128)
    pipe.tokenizer, pipe.tokenizer_2, pipe.tokenizer_3 = (SyntheticCLIPTokenizer(), SyntheticCLIPTokenizer(),
                                                          SyntheticCLIPTokenizer(32128))
    return pipe


@pytest.mark.skipif(not os.path.exists(REF_SCRIPT), reason="reference checkout not present on this box")
def test_unmodified_reference_script_setup_builds_b200_objects(tmp_path, monkeypatch):
    import adv_grpo_b200
    from adv_grpo_b200 import shims, weights
    from adv_grpo_b200.config import load_config
    monkeypatch.chdir(tmp_path)
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.split(".")[0] in
             ("adv_grpo", "diffusers", "peft", "accelerate", "ml_collections")}
    try:
        shimmed = adv_grpo_b200.install_as_adv_grpo(shim_third_party=True) and shims.install()
        assert {"diffusers", "peft", "accelerate", "ml_collections"} <= set(shimmed)
        shims.set_pipeline_factory(_tiny_factory)
        # the script imports two reference modules this package does not replace (prompt helpers: host code)
        prm = types.ModuleType("adv_grpo.prompts")
        sys.modules["adv_grpo.prompts"] = prm
        sys.modules["adv_grpo"].prompts = prm
        import adv_grpo.pickscore_scorer as ps
        real_scorer = ps.PickScoreScorer
        monkeypatch.setattr(ps, "PickScoreScorer", functools.partial(real_scorer, cfg=weights.CLIP_TINY))
        from absl import flags
        for name in list(flags.FLAGS):
            if name == "config":
                delattr(flags.FLAGS, name)
        spec = importlib.util.spec_from_file_location("ref_train_sd3_fast_pickscore", REF_SCRIPT)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)                               # the file as it is in the reference checkout

        class _DDP(torch.nn.Module):                               # single-process CPU stand-in for DistributedDataParallel
            def __init__(self, module, **kw):
                super().__init__()
                self.module = module

        class _Pool:
            def __init__(self, *a, **k):
                raise _Stop()

        monkeypatch.setattr(mod, "DDP", _DDP)
        monkeypatch.setattr(mod.futures, "ThreadPoolExecutor", _Pool)
        cfg = load_config("pickscore_cotrain_sd3_fast")
        ds = tmp_path / "dataset"
        ds.mkdir()
        (ds / "train.txt").write_text("\n".join(f"prompt {i}" for i in range(16)) + "\n")
        (ds / "test.txt").write_text("a test prompt\n")
        cfg.dataset, cfg.prompt_fn, cfg.wandb_init, cfg.mixed_precision = str(ds), "general_ocr", False, "bf16"
        cfg.sample.num_image_per_prompt, cfg.sample.mini_num_image_per_prompt = 2, 2       # k = 1: one process here
        flags.FLAGS(["prog"])
        flags.FLAGS.config = cfg
        with pytest.raises(_Stop):
            mod.main(None)
        # ---- what the script built ----
        from adv_grpo_b200.mmdit import SD3Transformer2DModel
        assert shims.Accelerator._last is not None
        transformer, optimizer, train_loader, test_loader = shims.Accelerator._last.prepared
        assert type(transformer) is SD3Transformer2DModel                     # pipeline.transformer after get_peft_model
        assert transformer.lora_rank == 32 and transformer.lora_scale == 2.0   # LoraConfig(r=32, lora_alpha=64)
        params = [p for p in transformer.parameters() if p.requires_grad]
        assert len(params) == 1 and params[0] is transformer.lora_flat       # the script's optimizer holds the flat LoRA
        assert optimizer.param_groups[0]["params"][0] is transformer.lora_flat
        prompts, metas = next(iter(train_loader))                             # reference sampler + dataset, untouched
        assert len(prompts) == cfg.sample.train_batch_size and prompts[0].startswith("prompt ")
    finally:
        shims.set_pipeline_factory(None)
        for k in [k for k in sys.modules if k.split(".")[0] in ("adv_grpo", "diffusers", "peft", "accelerate", "ml_collections")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})


def test_adapters_roundtrip_peft_wrapped_state_dict_names():
    """`from_diffusers` path: a peft-wrapped diffusers state dict (base_model.model. prefix, .base_layer. infix,
    lora_A.default.weight keys, the pos_embed buffer) converts to the same model as the plain parameters + LoRA."""
    from adv_grpo_b200 import adapters, weights
    cfg = weights.MMDIT_TINY
    params = weights.init_mmdit(cfg, seed=0)
    lora = weights.init_lora(cfg, rank=32, seed=1, perturb_b=0.02)
    sd = {}
    for k, v in params.items():
        mod = k.rsplit(".", 1)[0]
        sd["base_model.model." + (k.replace(mod, mod + ".base_layer") if mod in lora else k)] = v
    for mod, (a, b) in lora.items():
        sd[f"base_model.model.{mod}.lora_A.default.weight"] = a
        sd[f"base_model.model.{mod}.lora_B.default.weight"] = b
    sd["base_model.model.pos_embed.pos_embed"] = torch.zeros(1, cfg["pos_embed_max_size"] ** 2, 256)
    base, got_lora = adapters.split_peft_state_dict(sd)
    base.pop("pos_embed.pos_embed")
    assert set(base) == set(params) and all(torch.equal(base[k], params[k]) for k in params)
    assert set(got_lora) == set(lora) and all(torch.equal(got_lora[k][0], lora[k][0]) for k in lora)
    inferred = adapters.mmdit_config_from_state_dict(dict(base, **{"pos_embed.pos_embed": sd["base_model.model.pos_embed.pos_embed"]}))
    for key in ("num_layers", "heads", "head_dim", "dual_layers", "qk_norm", "patch_size", "in_channels",
                "pos_embed_max_size", "base_size", "joint_dim", "pooled_dim"):
        assert inferred[key] == cfg[key], key
    model = adapters.transformer_from_state_dict(sd, device="cpu")
    assert model.lora_rank == 32 and len(model._lora_names) == len(lora)
    key = next(iter(lora)).replace(".", "_")
    assert torch.equal(model.lora_A[key].cpu(), lora[next(iter(lora))][0].float())
    vcfg = adapters.vae_config_from_state_dict(weights.init_vae_decoder(weights.VAE_SD3, device="meta"))
    assert vcfg == dict(weights.VAE_SD3)

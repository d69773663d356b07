"""CPU tests of the host-side logic: sampler sharding, config surface, reward registry protocol, EMA,
criterion, scheduler, and the multi-process plumbing (gloo, world_size 2)."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_k_repeat_sampler_partitions_and_is_rank_consistent():
    from adv_grpo_b200.sampler import DistributedKRepeatSampler
    data = list(range(50))
    world, bs, k = 8, 1, 2
    per_rank = [DistributedKRepeatSampler(data, bs, k, world, r, seed=42) for r in range(world)]
    for epoch in (0, 1, 7):
        all_idx = []
        for r, s in enumerate(per_rank):
            s.set_epoch(epoch)
            mine = next(iter(s))
            assert mine == s.indices_for_epoch(epoch)[r] and len(mine) == bs
            all_idx += mine
        vals, counts = np.unique(all_idx, return_counts=True)
        assert len(vals) == world * bs // k and (counts == k).all()
    with pytest.raises(AssertionError):
        DistributedKRepeatSampler(data, 1, 3, 8, 0)


def test_sampler_matches_reference_algorithm():
    """Same draw as train_sd3_fast_pickscore.py:102-126 (randperm(seed+epoch)[:m], repeat k, shuffle, slice)."""
    from adv_grpo_b200.sampler import DistributedKRepeatSampler
    n, world, bs, k, seed, epoch = 99, 4, 2, 2, 42, 5
    g = torch.Generator().manual_seed(seed + epoch)
    idx = torch.randperm(n, generator=g)[: world * bs // k].tolist()
    rep = [i for i in idx for _ in range(k)]
    sh = [rep[i] for i in torch.randperm(len(rep), generator=g).tolist()]
    s = DistributedKRepeatSampler(list(range(n)), bs, k, world, 3, seed)
    assert s.indices_for_epoch(epoch)[3] == sh[6:8]


def test_config_surface_and_reference_config_file_loader():
    from adv_grpo_b200.config import ConfigDict, load_config
    c = load_config("pickscore_cotrain_sd3_fast")
    assert c.sample.num_steps == 10 and c.sample.train_num_steps == 2 and c.train.clip_range == 1e-5
    assert c.sample.noise_level == 0.8 and c.sample.guidance_scale == 4.5 and c.sample.global_std is True
    assert c.sample.num_batches_per_epoch == 12 and c.train.gradient_accumulation_steps == 6
    assert dict(c.reward_fn) == {"pickscore_cotrain": 1}
    d = load_config("dino_patch_cotrain_sd3_fast")
    assert dict(d.reward_fn) == {"dino_patch_cotrain": 1} and d.d_times == 10
    cd = ConfigDict()
    cd.a = {"b": 1}
    assert cd.a.b == 1 and cd["a"]["b"] == 1 and cd.get("zz", 5) == 5 and cd.to_dict() == {"a": {"b": 1}}
    ref = "/root/reference/config/grpo.py"
    if os.path.exists(ref):                      # build container only: execute the reference's own config file
        r = load_config(ref + ":pickscore_cotrain_sd3_fast")
        for path in ("sample.num_steps", "sample.train_num_steps", "sample.guidance_scale", "sample.noise_level",
                     "sample.num_batches_per_epoch", "train.clip_range", "train.gradient_accumulation_steps",
                     "train.adv_clip_max", "train.learning_rate", "resolution", "tune_layer", "d_lr"):
            a, b = c, r
            for k in path.split("."):
                a, b = a[k], b[k]
            assert a == b, path


def test_reward_registry_protocol():
    from adv_grpo_b200 import rewards
    ref_keys = {"deqa", "ocr", "video_ocr", "imagereward", "pickscore", "qwenvl", "aesthetic", "jpeg_compressibility",
                "unifiedreward", "geneval", "clipscore", "image_similarity", "image_similarity_eval",
                "constractive_external", "discriminator", "pickscore_cotrain", "pickscore_patch", "dino_cotrain",
                "dino_multi_cotrain", "dino_patch_cotrain", "siglip_cotrain", "siglip_image_similarity"}
    assert set(rewards.score_functions) == ref_keys           # rewards.py:1013-1036
    with pytest.raises(NotImplementedError):
        rewards.multi_score("cpu", {"aesthetic": 1.0})
    calls = []

    def fake_factory(device):
        def _fn(scorer, images, prompts, metadata):
            calls.append(scorer)
            return torch.arange(len(prompts), dtype=torch.float32), {}
        return _fn

    orig = rewards.score_functions["pickscore_cotrain"]
    rewards.score_functions["pickscore_cotrain"] = fake_factory
    try:
        fn = rewards.multi_score("cpu", {"pickscore_cotrain": 0.5})
        details, extra = fn(torch.zeros(3, 3, 8, 8), ["a", "b", "c"], [{}] * 3, scorer="S")
    finally:
        rewards.score_functions["pickscore_cotrain"] = orig
    assert calls == ["S"] and extra == {}
    assert torch.equal(details["avg"], torch.tensor([0.0, 0.5, 1.0])) and "pickscore_cotrain" in details


def test_ema_wrapper_matches_golden(golden):
    from adv_grpo_b200.ema import EMAModuleWrapper
    params = [torch.nn.Parameter(torch.tensor(p)) for p in golden["G10_init"]]
    ema = EMAModuleWrapper(params, decay=0.9, update_step_interval=8, device="cpu")
    for step in range(40):
        with torch.no_grad():
            for p in params:
                p.add_(0.01 * (step + 1))
        ema.step(params, step)
        np.testing.assert_allclose([e.sum().item() for e in ema.ema_parameters], golden["G10_ema_sums"][step],
                                   rtol=1e-5, atol=1e-5)
    saved = [p.detach().clone() for p in params]
    ema.copy_ema_to(params, store_temp=True)
    assert all(torch.equal(p, e) for p, e in zip(params, ema.ema_parameters))
    ema.copy_temp_to(params)
    assert all(torch.equal(p, s) for p, s in zip(params, saved))


def test_clip_criterion_matches_golden(golden, golden_dir):
    from adv_grpo_b200.pick_score_training import CLIPCriterion, CLIPCriterionConfig
    t = torch.load(os.path.join(golden_dir, "g6_tensors.pt"))
    crit = CLIPCriterion(CLIPCriterionConfig())
    loss = crit.calc_loss(t["t"], t["i0"], t["i1"], torch.tensor(100.0), torch.tensor(1.0), torch.tensor(0.0), torch.tensor(1.0))
    assert abs(loss.item() - golden["G6_loss"]) < 1e-5


def test_scheduler_matches_golden(golden):
    from adv_grpo_b200.scheduler import FlowMatchEulerDiscreteScheduler, retrieve_timesteps
    s = FlowMatchEulerDiscreteScheduler()
    ts, n = retrieve_timesteps(s, 10, "cpu")
    assert n == 10
    np.testing.assert_array_equal(s.sigmas.numpy(), np.array(golden["G4_sigmas"], dtype=np.float32))
    np.testing.assert_array_equal(ts.numpy(), np.array(golden["G4_timesteps"], dtype=np.float32))
    assert s.index_for_timestep(ts[4]) == 4


def test_install_as_adv_grpo_aliases():
    import adv_grpo_b200
    adv_grpo_b200.install_as_adv_grpo()
    from adv_grpo.diffusers_patch.sd3_pipeline_with_logprob_fast import pipeline_with_logprob_random  # noqa: F401
    from adv_grpo.diffusers_patch.sd3_sde_with_logprob import sde_step_with_logprob  # noqa: F401
    from adv_grpo.rewards import multi_score  # noqa: F401
    from adv_grpo.stat_tracking import PerPromptStatTracker  # noqa: F401
    from adv_grpo.ema import EMAModuleWrapper  # noqa: F401


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adv_grpo_b200 import trainer
    # packed all-gather layout: rank-major concatenation, so reshape(world, -1, T)[rank] un-gathers
    t = torch.full((3, 2), float(rank)) + torch.arange(3)[:, None]
    g = trainer.all_gather_cat(t)
    ok = g.shape == (3 * world, 2) and torch.equal(g.reshape(world, 3, 2)[rank], t)
    # flat gradient all-reduce (mean) used at the accumulation boundary
    class T:
        pass
    tr = T()
    tr.world = world
    tr.params = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(2, 3))]
    for p in tr.params:
        p.grad = torch.full_like(p, float(rank + 1))
    trainer.GRPOTrainer._sync_grads(tr)
    ok &= all(torch.allclose(p.grad, torch.full_like(p, (1 + world) / 2)) for p in tr.params)
    # prompt shards: every rank draws from the same permutation, disjoint slices
    from adv_grpo_b200.sampler import DistributedKRepeatSampler
    s = DistributedKRepeatSampler(list(range(30)), 1, 1, world, rank, seed=42)
    mine = torch.tensor(s.indices_for_epoch(3)[rank])
    allv = trainer.all_gather_cat(mine)
    ok &= len(set(allv.tolist())) == world
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def test_lora_adapter_dir_roundtrip_peft_layout(tmp_path):
    """§8f on-disk format: the peft adapter directory of `save_ckpt` (`train_pick:389-398`) / `lora_path` resume
    (`:506-509`): key names, config fields, EMA swap, strictness.  CPU only (no kernels involved)."""
    import json
    from adv_grpo_b200 import checkpoint, weights
    from adv_grpo_b200.ema import EMAModuleWrapper
    from adv_grpo_b200.mmdit import SD3Transformer2DModel
    cfg = weights.MMDIT_TINY
    params = weights.init_mmdit(cfg, seed=0, device="cpu", dtype=torch.bfloat16)
    lora = weights.init_lora(cfg, rank=32, seed=1, perturb_b=0.02)
    m = SD3Transformer2DModel(cfg, params, lora=lora, device="cpu")
    d = tmp_path / "ckpt"
    m.save_pretrained(str(d))
    conf = json.load(open(d / "adapter_config.json"))
    assert conf["peft_type"] == "LORA" and conf["r"] == 32 and conf["lora_alpha"] == 64
    assert conf["target_modules"] == sorted(weights.LORA_TARGETS) and conf["init_lora_weights"] == "gaussian"
    from safetensors.torch import load_file
    sd = load_file(str(d / "adapter_model.safetensors"))
    assert "base_model.model.transformer_blocks.0.attn.to_q.lora_A.weight" in sd
    assert sd["base_model.model.transformer_blocks.0.attn.to_q.lora_A.weight"].shape == (32, cfg["heads"] * cfg["head_dim"])
    assert sd["base_model.model.transformer_blocks.0.attn.to_out.0.lora_B.weight"].shape == (cfg["heads"] * cfg["head_dim"], 32)
    assert all(k.startswith("base_model.model.transformer_blocks.") and ".default." not in k for k in sd)
    # the context-pre-only last block has no to_add_out adapter
    last = cfg["num_layers"] - 1
    assert f"base_model.model.transformer_blocks.{last}.attn.to_add_out.lora_A.weight" not in sd
    # load into a fresh model (different LoRA init) -> identical factors; in place (parameter objects survive)
    m2 = SD3Transformer2DModel(cfg, params, lora=weights.init_lora(cfg, rank=32, seed=9, perturb_b=0.5), device="cpu")
    ids = [id(p) for p in m2.trainable_parameters()]
    missing, unexpected = m2.load_adapter(str(d))
    assert not missing and not unexpected and ids == [id(p) for p in m2.trainable_parameters()]
    for k, v in m.lora_state_dict().items():
        assert torch.equal(v, m2.lora_state_dict()[k])
    # keys carrying the adapter name (in-memory peft state dicts) are accepted too
    from safetensors.torch import save_file
    save_file({k.replace(".weight", ".default.weight"): v for k, v in sd.items()}, str(d / "adapter_model.safetensors"))
    assert checkpoint.load_adapter_dir(str(d))[0].keys() == sd.keys()
    # strictness and rank check
    bad = dict(sd)
    bad.pop("base_model.model.transformer_blocks.0.attn.to_q.lora_A.weight")
    checkpoint.save_adapter_dir(str(tmp_path / "bad"), bad, conf)
    with pytest.raises(KeyError):
        m2.load_adapter(str(tmp_path / "bad"))
    checkpoint.save_adapter_dir(str(tmp_path / "r16"), sd, dict(conf, r=16))
    with pytest.raises(ValueError):
        m2.load_adapter(str(tmp_path / "r16"))
    # save_ckpt: EMA weights are what lands on disk, the live weights come back afterwards
    ps = m.trainable_parameters()
    assert len(ps) == 1 and ps[0] is m.lora_flat                  # one flat master parameter; per-layer factors are views
    ema = EMAModuleWrapper(ps, decay=0.9, update_step_interval=1, device="cpu")
    k0 = "base_model.model.transformer_blocks.0.attn.to_q.lora_A.weight"
    live = [p.detach().clone() for p in ps]
    live_k0 = m.lora_state_dict()[k0].clone()
    with torch.no_grad():
        for p in ps:
            p.add_(1.0)
    assert torch.equal(m.lora_state_dict()[k0], live_k0 + 1.0)    # the views see the update of the flat buffer
    moved = [p.detach().clone() for p in ps]
    root = checkpoint.save_ckpt(str(tmp_path / "run"), m, 7, ema=ema, trainable_parameters=ps, use_ema=True)
    assert root.endswith("checkpoints/checkpoint-7/lora")
    on_disk, _ = checkpoint.load_adapter_dir(root)
    assert torch.equal(on_disk[k0], live_k0)                      # EMA shadow == the weights at wrapper creation
    assert all(torch.equal(a, b) for a, b in zip(moved, ps))      # live weights restored
    assert checkpoint.save_ckpt(str(tmp_path / "run2"), m, 1, is_main_process=False).endswith("lora")
    assert not (tmp_path / "run2" / "checkpoints" / "checkpoint-1" / "lora" / "adapter_config.json").exists()


def test_reference_image_index_matches_reference_preprocessing(tmp_path):
    """§8f on-disk format: the {prompt: [files]} index + Image.open -> Resize((S,S)) -> ToTensor of
    `train_pick:705-707,773-799` (PIL bilinear with antialiasing, [0,1] float32 CHW), fallback image, caching, hook."""
    import json
    from PIL import Image
    from adv_grpo_b200.reference_images import ReferenceImageIndex
    rng = np.random.RandomState(0)
    root = tmp_path / "imgs"
    root.mkdir()
    names = []
    for i, (h, w) in enumerate([(40, 64), (100, 30), (16, 16)]):
        arr = rng.randint(0, 256, size=(h, w, 3), dtype=np.uint8)
        Image.fromarray(arr).save(root / f"r{i}.png")
        names.append(f"r{i}.png")
    Image.fromarray(np.full((8, 8, 3), 7, dtype=np.uint8)).save(tmp_path / "default.png")
    idx_path = tmp_path / "index.json"
    json.dump({"a cat": names[:2], "a dog": [names[2], "missing.png"]}, open(idx_path, "w"))
    idx = ReferenceImageIndex(str(idx_path), str(root), size=32, device="cpu", default_image=str(tmp_path / "default.png"))
    got = idx("a cat")
    assert got.shape == (2, 3, 32, 32) and got.dtype == torch.float32 and 0.0 <= got.min() and got.max() <= 1.0
    for k, name in enumerate(names[:2]):                               # the reference's own expression
        ref = np.asarray(Image.open(root / name).convert("RGB").resize((32, 32), Image.BILINEAR), dtype=np.float32) / 255.0
        assert np.array_equal(got[k].permute(1, 2, 0).numpy(), ref)
    dog = idx("a dog")
    assert dog.shape == (2, 3, 32, 32) and torch.allclose(dog[1], torch.full((3, 32, 32), 7 / 255.0))   # fallback image
    assert idx("a cat") is got                                         # cached
    assert idx("a cat", n=5).shape == (5, 3, 32, 32) and torch.equal(idx("a cat", n=5)[2], got[0])
    with pytest.raises(KeyError):
        idx("a bird")
    fn = idx.as_trainer_fn(["a dog", "a cat"])
    assert torch.equal(fn(1, 2, 32), got)
    strict = ReferenceImageIndex(str(idx_path), str(root), size=32, device="cpu")
    with pytest.raises(FileNotFoundError):
        strict("a dog")


def _mirror(key):
    """Resolves 'adv_grpo/<file>.py::A.b' to the callable of the same dotted name in adv_grpo_b200.
    Inner `_fn` closures are obtained by calling the factory (device='cpu' never touches CUDA at build time)."""
    import importlib
    rel, dotted = key.split("::")
    mod = importlib.import_module("adv_grpo_b200." + rel[len("adv_grpo/"):-3].replace("/", "."))
    parts = dotted.split(".")
    obj = getattr(mod, parts[0])
    for p in parts[1:]:
        if p == "_fn":
            obj = obj("cpu", {"pickscore_cotrain": 1.0}) if parts[0] == "multi_score" else obj("cpu")
        else:
            obj = getattr(obj, p)
    return obj


def test_boundary_signatures_match_reference(golden_dir):
    """every boundary symbol keeps the reference's parameter names, order and defaults
    (tests/golden/signatures.json is extracted from the reference sources by tests/golden/make_signatures.py).
    A mirror may append optional parameters, and may accept **kwargs for reference parameters it ignores."""
    import inspect
    import json
    with open(os.path.join(golden_dir, "signatures.json")) as f:
        ref = json.load(f)
    checked = 0
    for key, spec in sorted(ref.items()):
        if "keys" in spec:
            continue
        factory_name = key.split("::")[1].split(".")[0]
        if key.endswith("._fn") and factory_name in ("pickscore_score", "ocr_score"):
            # these factories build a model / need PaddleOCR: read the closure's parameters off its code object
            from adv_grpo_b200 import rewards
            code = next(c for c in getattr(rewards, factory_name).__code__.co_consts
                        if hasattr(c, "co_name") and c.co_name == "_fn")
            assert list(code.co_varnames[:code.co_argcount]) == [p["name"] for p in spec["params"]], key
            checked += 1
            continue
        fn = _mirror(key)
        sig = inspect.signature(fn)
        ours = list(sig.parameters.values())
        has_varkw = any(p.kind is p.VAR_KEYWORD for p in ours)
        names = [p.name for p in ours]
        ref_params = [p for p in spec["params"]]
        if ref_params and ref_params[0]["name"] == "self" and (not names or names[0] != "self"):
            # bound-method view / the reference's `self` is the pipeline or scheduler passed positionally
            if inspect.ismethod(fn) or key.endswith(("__init__", "__call__")) or "." in key.split("::")[1]:
                ref_params = ref_params[1:]
        pos = 0
        for rp in ref_params:
            if rp["name"] not in sig.parameters:
                assert has_varkw and rp["default"] is not None, f"{key}: parameter {rp['name']!r} is missing"
                continue
            op = sig.parameters[rp["name"]]
            if rp["default"] is None:
                assert op.default is inspect.Parameter.empty, f"{key}: {rp['name']} must stay required"
                if rp["kind"] == "positional":
                    assert names.index(rp["name"]) == pos, f"{key}: positional order of {rp['name']}"
            else:
                assert op.default is not inspect.Parameter.empty, f"{key}: {rp['name']} lost its default"
                want = eval(rp["default"], {"torch": torch})
                got = list(op.default) if isinstance(want, list) else op.default   # immutable tuple for a list default
                assert got == want, f"{key}: default of {rp['name']} is {op.default!r}, reference {want!r}"
            pos += 1
        checked += 1
    assert checked >= 31
    # the registry exposes every reward key of the reference (rewards.py:1014-1038)
    from adv_grpo_b200 import rewards
    assert set(ref["adv_grpo/rewards.py::multi_score.score_functions"]["keys"]) <= set(rewards.score_functions)


def test_ocr_host_plugin_reward_arithmetic():
    """adv_grpo/ocr.py:31-65 with an injected recogniser (PaddleOCR is not installable here): quoted target text,
    blank stripping + lower-casing, containment shortcut, Levenshtein distance capped at len(target), failure =
    maximum penalty, list[float] return; and the registry path `multi_score(device, {"ocr": w})`."""
    import numpy as np
    from adv_grpo_b200 import rewards
    from adv_grpo_b200.ocr import OcrScorer, levenshtein
    assert [levenshtein(*p) for p in (("kitten", "sitting"), ("flaw", "lawn"), ("", "abc"), ("abc", "abc"), ("abc", "abd"))] \
        == [3, 2, 3, 0, 1]
    texts = {0: [("Hello W", 0.9), ("orld", 0.8)], 1: [("HELLO", 0.9), ("junk", 0.0)], 2: [], 3: None,
             4: [("completely unrelated long text", 0.9)]}

    def recognizer(img):
        r = texts[int(img[0, 0, 0])]
        if r is None:
            raise RuntimeError("recogniser failure")
        return r

    imgs = [np.full((8, 8, 3), i, dtype=np.uint8) for i in range(5)]
    prompts = ['a sign that says "Hello World" in neon'] * 5
    got = OcrScorer(recognizer=recognizer)(imgs, prompts)
    # 0: exact after normalisation -> 1; 1: "hello" vs "helloworld": distance 5 of 10 -> 0.5 (zero-confidence line dropped);
    # 2: nothing recognised -> distance 10 -> 0; 3: failure -> 0; 4: distance > len(target) capped -> 0
    assert got == [1.0, 0.5, 0.0, 0.0, 0.0] and isinstance(got, list)
    rewards.OCR_KWARGS["recognizer"] = recognizer
    try:
        fn = rewards.multi_score("cpu", {"ocr": 2.0})
        images = torch.stack([torch.full((3, 8, 8), i / 255.0) for i in range(5)])      # NCHW in [0, 1], rewards.py:680-683
        details, extra = fn(images, prompts, [{}] * 5)
    finally:
        rewards.OCR_KWARGS.clear()
    assert details["ocr"] == got and extra == {}
    assert torch.allclose(details["avg"], 2.0 * torch.tensor(got))
    with pytest.raises(ImportError, match="paddleocr"):
        OcrScorer()


def _criterion_rank(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adv_grpo_b200.pick_score_training import CLIPCriterion, CLIPCriterionConfig
    t = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "g13_tensors.pt"))
    n = t["t"].shape[0] // world
    sl = slice(rank * n, (rank + 1) * n)
    res = {}
    for ibn in (False, True):
        cfg = CLIPCriterionConfig(is_distributed=True, in_batch_negatives=ibn)
        tt = t["t"][sl].clone().requires_grad_(True)
        loss = CLIPCriterion(cfg).calc_loss(tt, t["i0"][sl], t["i1"][sl], torch.tensor(100.0), t["label_0"][sl],
                                            t["label_1"][sl], torch.ones(n))
        loss.backward()
        res[f"loss_ibn{int(ibn)}"] = loss.item()
        res[f"grad_t_sum_ibn{int(ibn)}"] = tt.grad.double().abs().sum().item()
    out[rank] = res
    dist.destroy_process_group()


def test_clip_criterion_in_batch_negatives_ties_and_distributed_g13(golden_dir):
    """The branches of CLIPCriterion.calc_loss the presets leave off (pick_score_training.py:135-170: in-batch
    negatives, per-example labels with ties, features gathered over the ranks) against golden G13 produced by the
    verbatim reference class (tests/golden/make_golden_criterion.py); the distributed case on 2 gloo ranks."""
    import json
    import torch.multiprocessing as mp
    from adv_grpo_b200.pick_score_training import CLIPCriterion, CLIPCriterionConfig
    with open(os.path.join(golden_dir, "golden_criterion.json")) as f:
        g = json.load(f)
    t = torch.load(os.path.join(golden_dir, "g13_tensors.pt"))
    for ibn in (False, True):
        tt = t["t"].clone().requires_grad_(True)
        loss = CLIPCriterion(CLIPCriterionConfig(in_batch_negatives=ibn)).calc_loss(
            tt, t["i0"], t["i1"], torch.tensor(100.0), t["label_0"], t["label_1"], torch.ones(6))
        loss.backward()
        assert abs(loss.item() - g[f"local_loss_ibn{int(ibn)}"]) < 1e-5
        assert abs(tt.grad.double().abs().sum().item() - g[f"local_grad_t_abs_sum_ibn{int(ibn)}"]) < 1e-3
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_criterion_rank, args=(2, 29541, out), nprocs=2, join=True)
        got = {str(k): dict(v) for k, v in out.items()}
    for r in ("0", "1"):
        for k, v in g["distributed_world2"][r].items():
            assert abs(got[r][k] - v) < 1e-3 * max(1.0, abs(v)), (r, k, got[r][k], v)


def test_gemm_tail_split_host_logic():
    """`ops._tail_split_rows` (experimental, off by default): the config-2 dual launch at N = 1536 is 462 pair tiles on
    74 CTA pairs = six full waves + 18 tiles; the split keeps exactly the six full waves in the dual launch, cuts the text
    rows at a whole number of 205-token samples, and leaves a tail that fits one wave of 128x128 tiles.  Shapes whose last
    wave is more than half full, or that have fewer than two full waves, are left alone."""
    from adv_grpo_b200 import ops
    assert ops.GEMM_TAIL_SPLIT is False or os.environ.get("ADVGRPO_GEMM_TAIL_SPLIT") == "1"
    f = ops._tail_split_rows
    s = f(16384, 3280, 1536, 205, 148)
    assert s == 2460 and s % 205 == 0
    main_tiles = ((16384 + 255) // 256 + (s + 255) // 256) * 6
    assert main_tiles == 6 * 74                                            # exactly six full waves
    assert ((3280 - s + 127) // 128) * 12 <= 148                           # the tail is one wave of 128x128 tiles
    assert f(16384, 3280, 1536, 1, 148) == 2560
    assert f(16384, 3280, 4608, 1, 148) is None                            # 18.7 waves: last wave 73 % full
    assert f(16384, 3280, 6144, 1, 148) is None                            # 24.97 waves
    assert f(2048, 410, 1536, 205, 148) is None                            # fewer than two full waves


def test_parameter_inventory_matches_published_checkpoint_sizes():
    """Cheap pin of every weight shape the models are built from: the state-dict inventories of
    `weights.py` reproduce the parameter counts of the published checkpoints the reference loads --
    stabilityai/stable-diffusion-3.5-medium transformer 2.24 B (+ the 384 x 384 x 1536 position table = 2.47 B, the
    size of the released file), stable-diffusion-3-medium 2.03 B, laion CLIP-ViT-H-14 (PickScore_v1's backbone) 986 M,
    timm vit_base_patch14_dinov2.lvd142m 86.6 M."""
    import torch
    from adv_grpo_b200 import weights
    count = lambda p: sum(v.numel() for v in p.values())
    mm = weights.init_mmdit(weights.SD35_MEDIUM, device="meta")
    assert count(mm) == 2_243_171_520
    assert count(mm) + 384 * 384 * 1536 == 2_469_663_936
    assert count(weights.init_mmdit(weights.SD3_MEDIUM, device="meta")) == 2_028_328_000
    assert count(weights.init_clip(weights.CLIP_H, device="meta")) == 986_109_441
    assert count(weights.init_dinov2(weights.DINOV2_B, device="meta")) == 86_579_712
    assert count(weights.init_vae_decoder(weights.VAE_SD3, device="meta")) == 49_545_475
    # LoRA r = 32 on the 8 attention projections of every block (to_add_out is absent in the last block)
    lora = weights.init_lora(weights.SD35_MEDIUM, rank=32, seed=1)
    assert sum(a.numel() + b.numel() for a, b in lora.values()) == (24 * 8 - 1) * 2 * 32 * 1536 == 18_776_064
    assert all(v.dtype == torch.bfloat16 for v in mm.values())

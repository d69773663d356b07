"""Reward path (A7/A8a/A8b) and the stat-tracker mirror on the GPU vs the CPU oracle, tiny tower configs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_pickscore_scorer_matches_oracle():
    from adv_grpo_b200 import weights
    from adv_grpo_b200.pickscore_scorer import PickScoreScorer
    from oracle import clip as clip_o
    from oracle import preprocess as pre_o
    cfg = weights.CLIP_TINY
    params = weights.init_clip(cfg, seed=3, device="cpu", dtype=torch.bfloat16)
    scorer = PickScoreScorer(device=DEV, cfg=cfg, state_dict=params)
    g = torch.Generator().manual_seed(0)
    images = torch.rand(4, 3, 256, 256, generator=g).bfloat16()
    prompts = ["a red cube", "a red cube", "two green spheres on a table", "a red cube"]
    scores = scorer(prompts, images.to(DEV)).float().cpu()
    # oracle: bf16 quantisation -> PIL-exact resize -> CLIP towers in fp32 -> cosine * exp(logit_scale) / 26
    u8 = pre_o.pil_bicubic_resize_u8(pre_o.quantise_bf16(images).numpy(), 224)
    pix = torch.from_numpy(pre_o.clip_pixel_values(u8))
    ids = scorer.processor.tokenizer(prompts, padding=True, max_length=77)["input_ids"]
    p32 = {k: v.float() for k, v in params.items()}
    ocfg = dict(patch=cfg["patch"], v_layers=cfg["v_layers"], v_heads=cfg["v_heads"], t_layers=cfg["t_layers"], t_heads=cfg["t_heads"])
    ref = clip_o.pickscore(p32, ocfg, ids, pix)
    # bf16 towers vs fp32 oracle: scores are O(1) cosines * 100 / 26
    assert torch.allclose(scores, ref, atol=4e-2), (scores, ref)
    # PIL-image entry (the reference's input type) gives the same pixels as the uint8 tensor entry
    from PIL import Image
    q = pre_o.quantise_bf16(images)
    pil = [Image.fromarray(q[i].permute(1, 2, 0).numpy()) for i in range(4)]
    s_pil = scorer(prompts, pil).float().cpu()
    assert torch.allclose(s_pil, scores, atol=1e-6)
    # quirk Q10: the reference does the normalise / dot / scale tail in bf16 and returns bf16 scores
    # (adv_grpo/pickscore_scorer.py:40-51); the flag reproduces that rounding sequence from the same tower outputs
    sc_ref = PickScoreScorer(device=DEV, cfg=cfg, state_dict=params, reference_score_arithmetic=True)
    s16 = sc_ref(prompts, images.to(DEV))
    assert s16.dtype == torch.bfloat16
    bf = torch.bfloat16
    model = sc_ref.model
    ie = sc_ref._last_image_feats_bf16
    ids1 = [scorer.processor.tokenizer([p], padding=True, truncation=True, max_length=77)["input_ids"].to(DEV) for p in prompts]
    with torch.no_grad():                     # the frozen fast path the scorer itself uses (same tower outputs)
        te = torch.cat([model.get_text_features(input_ids=i).to(bf) for i in ids1], 0)
    want = (model.logit_scale.exp().to(bf) * ((te / te.norm(p=2, dim=-1, keepdim=True)) @ (ie / ie.norm(p=2, dim=-1, keepdim=True)).T)).diag() / 26
    assert (s16.float() - want.float()).abs().max().item() <= 2 * 2.0 ** -8 * want.float().abs().max().item()
    assert (s16.float().cpu() - scores).abs().max().item() < 3e-2          # bf16 tail vs fp32 tail on the same features


def test_dino_patch_reward_matches_oracle():
    from adv_grpo_b200 import rewards, weights
    from adv_grpo_b200.dinov2 import DINOHead, DinoV2
    from oracle import dinov2 as dino_o
    cfg = weights.DINOV2_TINY
    params = weights.init_dinov2(cfg, seed=4, device="cpu", dtype=torch.bfloat16)
    scorer = DinoV2(params, cfg, device=DEV)
    head = DINOHead(in_dim=cfg["width"], hidden_dim=64).to(DEV)
    fn = rewards.multi_score(DEV, {"dino_patch_cotrain": 1.0})
    images = torch.rand(3, 3, 128, 128, generator=torch.Generator().manual_seed(1))
    torch.manual_seed(123)
    details, _ = fn(images.to(DEV), ["p"] * 3, [{}] * 3, scorer=scorer, head=head)
    torch.manual_seed(123)
    idx = torch.randint(0, 1369, (3, 64), device=DEV).cpu()             # same CUDA draw as rewards.py:406
    feats = dino_o.forward_features({k: v.float() for k, v in params.items()},
                                    dict(patch=14, heads=cfg["heads"], layers=cfg["layers"]), dino_o.preprocess(images))
    hp = {k: v.detach().float().cpu() for k, v in head.state_dict().items()}
    ref, cls_s, _ = dino_o.patch_reward(hp, feats, idx)
    got = details["dino_patch_cotrain"].float().cpu()
    assert torch.allclose(got, ref, atol=3e-2), (got, ref)
    assert torch.equal(details["avg"].cpu(), got)


def test_stat_tracker_mirror_golden(golden):
    from adv_grpo_b200.stat_tracking import PerPromptStatTracker
    for gs, key in ((False, "G1"), (True, "G2")):
        t = PerPromptStatTracker(global_std=gs)
        a = t.update(['a', 'b', 'a', 'c', 'b', 'a'], [1, 2, 3, 4, 5, 6])
        assert isinstance(a, np.ndarray) and a.dtype == np.float64
        np.testing.assert_allclose(a, golden[key], rtol=1e-12, atol=1e-12)
        size, n_hist = t.get_stats()
        assert abs(size - golden["G1_stats"][0]) < 1e-12 and n_hist == golden["G1_stats"][1]
        t.clear()
        assert t.stats == {}
    t = PerPromptStatTracker(global_std=True)
    a = t.update(['p', 'p', 'q', 'q'], [[1, 1], [2, 2], [3, 3], [4, 4]])
    np.testing.assert_allclose(a, golden["G3"], rtol=1e-12, atol=1e-12)


def test_sde_mirror_signature_and_grad(golden, golden_dir):
    """adv_grpo.diffusers_patch.sd3_sde_with_logprob.sde_step_with_logprob_new drop-in on golden G8."""
    import os
    from adv_grpo_b200.diffusers_patch.sd3_sde_with_logprob import sde_step_with_logprob_new
    from adv_grpo_b200.scheduler import FlowMatchEulerDiscreteScheduler
    sch = FlowMatchEulerDiscreteScheduler()
    sch.set_timesteps(10, device=DEV)
    t = torch.load(os.path.join(golden_dir, "g8_tensors.pt"))
    mo = t["v"].bfloat16().to(DEV).requires_grad_(True)
    ts = sch.timesteps[golden["G8_step_index"]]
    prev, lp, mean, std = sde_step_with_logprob_new(sch, mo, ts, t["x"].bfloat16().to(DEV), noise_level=0.8,
                                                    prev_sample=t["prev"].bfloat16().to(DEV))
    np.testing.assert_allclose(lp.detach().cpu().numpy(), np.array(golden["G8_log_prob"], dtype=np.float32), rtol=2e-6)
    assert torch.equal(mean.cpu(), t["mean"]) and std.shape == (4, 1, 1, 1)
    lp.sum().backward()
    assert mo.grad is not None and torch.isfinite(mo.grad.float()).all()
    # rollout form: fresh noise, returns fp32 prev like the reference
    prev2, lp2, _, _ = sde_step_with_logprob_new(sch, mo.detach(), ts[:1], t["x"].bfloat16().to(DEV), noise_level=0.8,
                                                 generator=torch.Generator(device=DEV).manual_seed(1))
    assert prev2.dtype == torch.float32 and lp2.shape == (4,)


def test_multi_reward_pickscore_plus_host_plugin_config4():
    """BASELINE config 4's reward_fn {"pickscore": 0.5, "ocr": 0.5} (`rewards.py:1043-1093`): the frozen PickScore
    reward on the B200 kernels combined with a HOST plugin that returns a Python list (the reference's OcrScorer is a
    CPU PaddleOCR + Levenshtein plugin, `ocr.py:22-65`; PaddleOCR is not installable here, so a deterministic host
    function with the same call convention stands in).  score_details carries both rewards and their weighted sum."""
    from adv_grpo_b200 import rewards, weights
    from adv_grpo_b200.pickscore_scorer import PickScoreScorer
    g = torch.Generator().manual_seed(0)
    images = torch.rand(4, 3, 96, 96, generator=g).to(DEV)
    prompts = ["a red cube", "a red cube", "two green spheres", "a sign that says hello"]

    def ocr_factory(device):
        def _fn(imgs, prm, metadata):
            assert torch.is_tensor(imgs) and len(prm) == imgs.shape[0]
            return [float(len(p) % 7) / 7.0 for p in prm], {}            # list[float], like OcrScorer.__call__
        return _fn

    orig, orig_kw = rewards.score_functions["ocr"], dict(rewards.PICKSCORE_KWARGS)
    rewards.score_functions["ocr"] = ocr_factory
    rewards.PICKSCORE_KWARGS.update(cfg=weights.CLIP_TINY, seed=3)
    try:
        fn = rewards.multi_score(DEV, {"pickscore": 0.5, "ocr": 0.5})
        details, extra = fn(images.to(torch.bfloat16), prompts, [{}] * 4, only_strict=True)
    finally:
        rewards.score_functions["ocr"] = orig
        rewards.PICKSCORE_KWARGS.clear()
        rewards.PICKSCORE_KWARGS.update(orig_kw)
    assert extra == {} and set(details) == {"pickscore", "ocr", "avg"}
    ocr = torch.tensor([float(len(p) % 7) / 7.0 for p in prompts], device=DEV)
    pick = torch.as_tensor(details["pickscore"]).float()
    assert torch.allclose(details["avg"], 0.5 * pick + 0.5 * ocr, atol=1e-6)
    # the frozen reward equals an identically initialised co-trained scorer (same kernels, same weights)
    ref = PickScoreScorer(device=DEV, cfg=weights.CLIP_TINY, seed=3)(prompts, images.to(torch.bfloat16)).float()
    assert torch.allclose(pick, ref, atol=1e-6)
    # consumed as in train_sd3_fast_pickscore.py:849-856
    assert all(torch.as_tensor(v).float().shape == (4,) for v in details.values())


def test_reward_fn_from_thread_pool_while_main_thread_works():
    """SURVEY.md section 8b threading contract: the scripts call `reward_fn` from an 8-worker ThreadPoolExecutor while
    the main thread keeps sampling (train_sd3_fast_pickscore.py:668,816-817).  Scores computed concurrently (graph
    capture, replay and the text cache included) must equal the serial ones, and the main thread's own kernels and
    allocations must not be disturbed by a capture started in a worker."""
    from concurrent.futures import ThreadPoolExecutor
    from adv_grpo_b200 import ops, rewards, weights
    from adv_grpo_b200.pickscore_scorer import PickScoreScorer
    scorer = PickScoreScorer(device=DEV, cfg=weights.CLIP_TINY, seed=3)
    fn = rewards.multi_score(DEV, {"pickscore_cotrain": 1.0})
    g = torch.Generator().manual_seed(2)
    batches = [torch.rand(4, 3, 96, 96, generator=g).to(DEV).bfloat16() for _ in range(12)]
    prompts = [[f"prompt {i % 3}"] * 4 for i in range(12)]
    serial = [fn(b, p, [{}] * 4, scorer=scorer)[0]["avg"].clone() for b, p in zip(batches, prompts)]
    fresh = PickScoreScorer(device=DEV, cfg=weights.CLIP_TINY, seed=3)       # nothing captured / cached yet
    a = torch.randn(512, 256, device=DEV).bfloat16()
    w = torch.randn(384, 256, device=DEV).bfloat16()
    want = ops.gemm(a, w).clone()
    with ThreadPoolExecutor(max_workers=8) as ex:
        futs = [ex.submit(lambda b, p: fn(b, p, [{}] * 4, scorer=fresh)[0]["avg"].clone(), b, p)
                for b, p in zip(batches, prompts)]
        outs, n_iter = [], 0
        while not all(f.done() for f in futs):                                # the "sampling" thread: kernels + allocations
            o = ops.gemm(a, w) + torch.zeros(512, 384, device=DEV, dtype=torch.bfloat16)
            if n_iter % 64 == 0 and len(outs) < 64:                           # keep a bounded sample of the results
                outs.append(o)
            n_iter += 1
        got = [f.result() for f in futs]
    torch.cuda.synchronize()
    for s, c in zip(serial, got):
        assert torch.equal(s, c)
    assert all(torch.equal(o, want) for o in outs)


def test_stat_tracker_other_types_match_reference_golden(golden_dir):
    """PerPromptStatTracker.update(type='rwr' | 'sft' | 'dpo') (stat_tracking.py:48-70) as modes of the group-advantage
    kernel, bit-exact against the outputs of the verbatim reference file (golden G12), ties and an all-equal group
    included; 2-D rewards for 'rwr' / 'sft'; 'dpo' refuses 2-D rewards (the reference's flat index is only meaningful
    for 1-D)."""
    import json
    import os
    from adv_grpo_b200 import _lib
    from adv_grpo_b200.stat_tracking import PerPromptStatTracker
    with open(os.path.join(golden_dir, "golden_adv_modes.json")) as f:
        g = json.load(f)
    for mode in ("rwr", "sft", "dpo"):
        for gs in (False, True):
            t = PerPromptStatTracker(global_std=gs)
            a = t.update(g["prompts"], g["rewards"], type=mode)
            assert a.dtype == np.float64
            np.testing.assert_array_equal(a, np.array(g[f"{mode}_global{int(gs)}"]))
    for mode in ("rwr", "sft"):
        a = PerPromptStatTracker(global_std=True).update(g["prompts"], g["rewards_2d"], type=mode)
        np.testing.assert_array_equal(a, np.array(g[f"{mode}_2d"]))
    with pytest.raises(_lib.AdvGrpoError, match="dpo"):
        PerPromptStatTracker().update(g["prompts"], g["rewards_2d"], type="dpo")
    with pytest.raises(ValueError):
        PerPromptStatTracker().update(g["prompts"], g["rewards"], type="ppo")

"""MMDiT(-X) on the B200 kernels vs the fp32 CPU oracle (same seeded bf16-valued weights)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(cfg_name="MMDIT_TINY", perturb_b=0.02, B=2, hw=16, n_txt=13):
    from adv_grpo_b200 import weights
    from adv_grpo_b200.mmdit import SD3Transformer2DModel
    from oracle.mmdit import MMDiTOracle
    cfg = getattr(weights, cfg_name)
    params = weights.init_mmdit(cfg, seed=0, device="cpu", dtype=torch.bfloat16)
    lora = weights.init_lora(cfg, rank=32, seed=1, perturb_b=perturb_b)
    lora = {k: (a.bfloat16().float(), b.bfloat16().float()) for k, (a, b) in lora.items()}
    model = SD3Transformer2DModel(cfg, params, lora_rank=32, lora_alpha=64, lora=lora, device=DEV)
    oracle = MMDiTOracle(params, dict(cfg, dual_layers=set(cfg["dual_layers"])), lora=lora, lora_scale=2.0)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, 16, hw, hw, generator=g).bfloat16()
    t = torch.tensor([960.1293, 500.0][:B] if B <= 2 else [700.0] * B)
    ctx = torch.randn(B, n_txt, cfg["joint_dim"], generator=g).bfloat16()
    pooled = torch.randn(B, cfg["pooled_dim"], generator=g).bfloat16()
    return model, oracle, x, t, ctx, pooled


def test_mmdit_forward_matches_oracle():
    model, oracle, x, t, ctx, pooled = _setup()
    with torch.no_grad():
        got = model(x.to(DEV), t.to(DEV), ctx.to(DEV), pooled.to(DEV))[0].float().cpu()
        ref = oracle.forward(x.float(), t, ctx.float(), pooled.float())
    assert got.shape == ref.shape == x.shape
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    # bf16 activations through 3 blocks vs fp32: a few bf16 ulps of the output range
    assert err < 3e-2, err
    cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
    assert cos > 0.9995, cos


def test_mmdit_block_by_block():
    model, oracle, x, t, ctx, pooled = _setup()
    with torch.no_grad():
        for upto in (1, 2):
            gx, gc = model(x.to(DEV), t.to(DEV), ctx.to(DEV), pooled.to(DEV), upto=upto)
            rx, rc = oracle.forward(x.float(), t, ctx.float(), pooled.float(), upto=upto)
            for g_, r_ in ((gx, rx), (gc, rc)):
                err = (g_.float().cpu() - r_).abs().max().item() / r_.abs().max().item()
                assert err < 2e-2, (upto, err)


def test_mmdit_lora_disable_and_grad_mode_consistency():
    model, oracle, x, t, ctx, pooled = _setup()
    args = (x.to(DEV), t.to(DEV), ctx.to(DEV), pooled.to(DEV))
    with torch.no_grad():
        a = model(*args)[0]
        with model.disable_adapter():
            b = model(*args)[0]
    assert not torch.equal(a, b)
    c = model(*args)[0]                      # grad mode: same fused arithmetic -> identical bits
    assert c.requires_grad
    assert torch.equal(a, c.detach())


def test_mmdit_backward_matches_oracle_autograd():
    model, oracle, x, t, ctx, pooled = _setup()
    g = torch.Generator().manual_seed(9)
    w = torch.randn(x.shape, generator=g)
    out = model(x.to(DEV), t.to(DEV), ctx.to(DEV), pooled.to(DEV))[0]
    (out.float() * w.to(DEV)).sum().backward()
    for k in oracle.lora:
        a, b = oracle.lora[k]
        oracle.lora[k] = (a.clone().requires_grad_(True), b.clone().requires_grad_(True))
    ref = oracle.forward(x.float(), t, ctx.float(), pooled.float())
    (ref * w).sum().backward()
    worst = 0.0
    for name in model._lora_names:
        key = name.replace(".", "_")
        ra, rb = oracle.lora[name]
        for got, refg in ((model.lora_A[key].grad, ra.grad), (model.lora_B[key].grad, rb.grad)):
            assert got is not None, name
            if refg.norm().item() == 0:          # e.g. last block's add_q_proj: its context output is discarded
                assert got.float().abs().max().item() == 0, name
                continue
            cos = torch.nn.functional.cosine_similarity(got.float().cpu().flatten(), refg.flatten(), dim=0).item()
            worst = min(worst, cos - 1)
            rel = (got.float().cpu() - refg).norm().item() / refg.norm().item()
            # bf16 forward + bf16 gradients through 3 blocks against the fp32 oracle's autograd
            assert cos > 0.999, (name, cos, rel)
            assert rel < 3e-2, (name, cos, rel)


def test_mmdit_dual_gemm_matches_separate_launches():
    model, oracle, x, t, ctx, pooled = _setup()
    args = (x.to(DEV), t.to(DEV), ctx.to(DEV), pooled.to(DEV))
    with torch.no_grad():
        model.dual_gemm = True
        a = model(*args)[0]
        model.dual_gemm = False
        b = model(*args)[0]
        model.dual_gemm = True
    assert torch.allclose(a.float(), b.float(), atol=2e-2, rtol=2e-2)


def test_mmdit_fused_feed_forward_node_matches_unfused_autograd():
    """The feed-forward pair as one autograd node (GELU backward inside the dX GEMM epilogue, no stored hidden
    activation) gives the same forward bits and the same LoRA gradients (up to bf16 rounding of intermediates) as the
    chain of per-linear nodes."""
    outs, grads = [], []
    for fused in (True, False):
        model, oracle, x, t, ctx, pooled = _setup()
        model.fused_ff = fused
        y = model(x.to(DEV), t.to(DEV), ctx.to(DEV), pooled.to(DEV))[0]
        g = torch.Generator().manual_seed(3)
        (y.float() * torch.randn(y.shape, generator=g).to(DEV)).sum().backward()
        outs.append(y.detach())
        grads.append(torch.cat([p.grad.flatten() for p in model.trainable_parameters()]))
    assert torch.equal(outs[0], outs[1])
    cos = torch.nn.functional.cosine_similarity(grads[0], grads[1], dim=0)
    assert cos > 0.999, cos
    assert (grads[0] - grads[1]).abs().max() <= 0.03 * grads[1].abs().max()


def test_full_finetune_forward_and_all_parameter_grads_match_oracle():
    """`config.use_lora = False` (train_sd3_fast_pickscore.py:488): every transformer parameter
    trains.  Forward of the full-gradient path vs the fp32 oracle, gradients of EVERY diffusers-named parameter (read from
    the flat fp32 master's .grad views) vs the oracle's autograd, and bit-equality of the no-grad (rollout) and grad-mode
    (replay) forwards, which `ratio = 1` at `clip_range = 1e-5` relies on."""
    from adv_grpo_b200 import weights
    from adv_grpo_b200.mmdit import SD3Transformer2DModel
    from oracle.mmdit import MMDiTOracle
    cfg = weights.MMDIT_TINY
    params = weights.init_mmdit(cfg, seed=0, device="cpu", dtype=torch.bfloat16)
    model = SD3Transformer2DModel(cfg, params, lora_rank=32, lora_alpha=64, device=DEV).enable_full_finetune()
    assert [p.numel() for p in model.trainable_parameters()] == [sum(v.numel() for v in params.values())]
    oracle = MMDiTOracle(params, dict(cfg, dual_layers=set(cfg["dual_layers"])), lora=None)
    for k in oracle.p:
        oracle.p[k].requires_grad_(True)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 16, 16, 16, generator=g).bfloat16()
    t = torch.tensor([960.1293, 500.0])
    ctx = torch.randn(2, 13, cfg["joint_dim"], generator=g).bfloat16()
    pooled = torch.randn(2, cfg["pooled_dim"], generator=g).bfloat16()
    wgt = torch.randn(x.shape, generator=g)
    args = (x.to(DEV), t.to(DEV), ctx.to(DEV), pooled.to(DEV))
    with torch.no_grad():
        rollout = model(*args)[0]
    out = model(*args)[0]
    assert out.requires_grad and torch.equal(out.detach(), rollout)
    (out.float() * wgt.to(DEV)).sum().backward()
    ref = oracle.forward(x.float(), t, ctx.float(), pooled.float())
    (ref * wgt).sum().backward()
    err = (out.detach().float().cpu() - ref.detach()).abs().max().item() / ref.abs().max().item()
    assert err < 3e-2, err
    grads = dict(zip(model._full_names, model._full_grad_views))
    checked, worst = 0, (1.0, "")
    for name, p_o in oracle.p.items():
        got = grads[name].float().cpu()
        if p_o.grad is None or p_o.grad.norm().item() == 0:      # e.g. the last block's unused context output projection
            assert got.abs().max().item() == 0, name
            continue
        cos = torch.nn.functional.cosine_similarity(got.flatten(), p_o.grad.flatten(), dim=0).item()
        rel = (got - p_o.grad).norm().item() / p_o.grad.norm().item()
        worst = min(worst, (cos, name))
        # bf16 forward / backward through 3 blocks vs the fp32 oracle's autograd (same bars as the LoRA-gradient test,
        # one notch wider in rel for the small-norm bias / RMS-weight vectors)
        assert cos > 0.995 and rel < 0.1, (name, cos, rel)
        checked += 1
    assert checked >= len(oracle.p) - 8, (checked, len(oracle.p), worst)
    # optimizer step on the master -> in-place refresh of the working weights on the next forward
    from adv_grpo_b200.optim import FlatClipAdamW
    opt = FlatClipAdamW(model.trainable_parameters(), lr=1e-3, weight_decay=0.0, max_grad_norm=1.0)
    opt.step()
    model.invalidate_lora_cache()
    assert model.full_master.grad.abs().max().item() == 0
    with torch.no_grad():
        after = model(*args)[0]
    assert not torch.equal(after, rollout)

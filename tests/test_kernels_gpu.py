"""GPU parity tests (through the C ABI) of the HBM-bound kernels against the CPU oracle and the
golden vectors produced by the reference's own files.  Tolerances are stated per test."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import grpo_loss as loss_o
from oracle import preprocess as pre_o
from oracle import sde as sde_o
from oracle import stat_tracking as st_o
from oracle.scheduler import FlowMatchEulerOracle

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from adv_grpo_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def sched():
    s = FlowMatchEulerOracle()
    s.set_timesteps(10)
    return s


# ------------------------------------------------------------------ A4/A5 SDE step
@pytest.mark.parametrize("B,shape", [(2, (16, 32, 32)), (8, (16, 64, 64)), (3, (16, 8, 8))])
def test_sde_rollout_matches_oracle_bit_exact(ops, sched, B, shape):
    g = torch.Generator().manual_seed(B)
    vu = torch.randn(B, *shape, generator=g).bfloat16()
    vt = torch.randn(B, *shape, generator=g).bfloat16()
    x = torch.randn(B, *shape, generator=g).bfloat16()
    noise = torch.randn(B, *shape, generator=g)
    step = 1
    v = sde_o.cfg_combine(vu, vt, 4.5)                       # bf16 ops on CPU, like the reference on GPU
    prev_o, lp_o, mean_o, std_o = sde_o.sde_step_with_logprob_new(sched.sigmas, [step], v, x, 0.8, noise=noise)
    prev, lp, mean, std = ops.cfg_sde_step_logprob(vu.to(DEV), vt.to(DEV), x.to(DEV), sched.timesteps[step:step + 1],
                                                   sched.timesteps, sched.sigmas, 4.5, 0.8, noise=noise.to(DEV),
                                                   want_mean=True)
    assert torch.equal(mean.cpu(), mean_o), "prev_sample_mean must be bit-identical to the fp32 oracle"
    assert torch.equal(prev.cpu(), prev_o.bfloat16()), "stored next latents = bf16(prev_sample) (fast.py:654-655)"
    np.testing.assert_allclose(lp.cpu().numpy(), lp_o.numpy(), rtol=2e-6, atol=0)   # fp32 sum order only
    np.testing.assert_array_equal(std.cpu().numpy(), np.full(B, std_o.item(), dtype=np.float32))


def test_sde_replay_golden_g8(ops, sched, golden, golden_dir):
    """Replay form with per-sample timesteps against the verbatim reference output (golden G8)."""
    t = torch.load(os.path.join(golden_dir, "g8_tensors.pt"))
    ts = sched.timesteps[golden["G8_step_index"]]
    _, lp, mean, std = ops.cfg_sde_step_logprob(None, t["v"].bfloat16().to(DEV), t["x"].bfloat16().to(DEV), ts,
                                                sched.timesteps, sched.sigmas, 1.0, 0.8,
                                                prev_sample=t["prev"].bfloat16().to(DEV), want_mean=True)
    assert torch.equal(mean.cpu(), t["mean"])
    np.testing.assert_allclose(lp.cpu().numpy(), np.array(golden["G8_log_prob"], dtype=np.float32), rtol=2e-6)
    np.testing.assert_array_equal(std.cpu().numpy(), np.array(golden["G8_std"], dtype=np.float32))


def test_sde_last_step_has_zero_std_and_logprob(ops, sched):
    x = torch.randn(2, 16, 16, 16).bfloat16().to(DEV)
    v = torch.randn(2, 16, 16, 16).bfloat16().to(DEV)
    prev, lp, _, std = ops.cfg_sde_step_logprob(None, v, x, sched.timesteps[9:10], sched.timesteps, sched.sigmas,
                                                1.0, 0.8, seed=1)
    assert lp.tolist() == [0.0, 0.0] and std.tolist() == [0.0, 0.0]      # quirk Q1
    x0 = (x.float() - sched.sigmas[9].item() * v.float())
    assert torch.allclose(prev.float(), x0.bfloat16().float(), atol=1e-2)


def test_sde_unknown_timestep_yields_nan(ops, sched):
    x = torch.zeros(1, 16, 8, 8, dtype=torch.bfloat16, device=DEV)
    _, lp, _, _ = ops.cfg_sde_step_logprob(None, x, x, torch.tensor([123.456]), sched.timesteps, sched.sigmas, 1.0,
                                           0.8, seed=1)
    assert math.isnan(lp.item())


def test_sde_philox_noise_statistics_and_determinism(ops, sched):
    B, n = 4, 16 * 64 * 64
    z = torch.zeros(B, n, dtype=torch.bfloat16, device=DEV)
    # x = v = 0 -> mu = 0 -> prev = std * eps
    kw = dict(timesteps=sched.timesteps[0:1], sched_timesteps=sched.timesteps, sigmas=sched.sigmas,
              guidance_scale=1.0, noise_level=0.8)
    p1, lp1, _, std = ops.cfg_sde_step_logprob(None, z, z, seed=7, offset=0, **kw)
    p2, lp2, _, _ = ops.cfg_sde_step_logprob(None, z, z, seed=7, offset=0, **kw)
    p3, _, _, _ = ops.cfg_sde_step_logprob(None, z, z, seed=7, offset=B * n // 4, **kw)
    assert torch.equal(p1, p2) and torch.equal(lp1, lp2)
    assert not torch.equal(p1, p3)
    eps = p1.float() / std[0]
    assert abs(eps.mean().item()) < 0.01 and abs(eps.std().item() - 1.0) < 0.01
    assert abs((eps ** 4).mean().item() - 3.0) < 0.1                     # Gaussian kurtosis
    assert abs(torch.corrcoef(torch.stack([eps[0], eps[1]]))[0, 1].item()) < 0.02
    np.testing.assert_allclose((-lp1 / std ** 2).cpu().numpy(), np.ones(B), rtol=0.02)


def test_sde_replay_backward_matches_autograd_of_oracle(ops, sched):
    B, shape = 4, (16, 16, 16)
    g = torch.Generator().manual_seed(11)
    npred = torch.randn(2 * B, *shape, generator=g).bfloat16()
    x = torch.randn(B, *shape, generator=g).bfloat16()
    prev = (x.float() + 0.3 * torch.randn(B, *shape, generator=g)).bfloat16()
    idx = [0, 1, 1, 5]
    w = torch.randn(B, generator=g)
    # oracle in fp32 with straight-through bf16 CFG (autograd of the reference expression)
    npo = npred.float().requires_grad_(True)
    vu, vt = npo.chunk(2)
    v = vu + 4.5 * (vt - vu)
    _, lp_o, _, _ = sde_o.sde_step_with_logprob_new(sched.sigmas, idx, v, x, 0.8, prev_sample=prev)
    (lp_o * w).sum().backward()
    npd = npred.to(DEV).requires_grad_(True)
    lp, _, _ = ops.sde_logprob_replay(npd, x.to(DEV), prev.to(DEV), sched.timesteps[idx], sched.timesteps,
                                      sched.sigmas, 4.5, 0.8, cfg=True)
    (lp * w.to(DEV)).sum().backward()
    got, ref = npd.grad.float().cpu(), npo.grad
    # bf16 gradient storage + bf16 CFG rounding in the forward: 2% of the gradient scale
    assert (got - ref).abs().max() <= 0.02 * ref.abs().max()
    cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0)
    assert cos > 0.9995


def test_sde_replay_kl_term_matches_autograd_of_oracle(ops, sched):
    """beta > 0 branch (train_pick:1105-1108,1124-1128): loss = sum_b w_b logp_b + beta * kl, kl from the oracle
    restatement; value and gradient w.r.t. the CFG-batched transformer output."""
    from oracle import grpo_loss as loss_o
    B, shape = 4, (16, 16, 16)
    g = torch.Generator().manual_seed(12)
    npred = torch.randn(2 * B, *shape, generator=g).bfloat16()
    x = torch.randn(B, *shape, generator=g).bfloat16()
    prev = (x.float() + 0.3 * torch.randn(B, *shape, generator=g)).bfloat16()
    mean_ref = x.float() + 0.2 * torch.randn(B, *shape, generator=g)
    idx = [0, 3, 1, 5]
    w = torch.randn(B, generator=g)
    beta = 0.7
    npo = npred.float().requires_grad_(True)
    vu, vt = npo.chunk(2)
    v = vu + 4.5 * (vt - vu)
    _, lp_o, mean_o, _ = sde_o.sde_step_with_logprob_new(sched.sigmas, idx, v, x, 0.8, prev_sample=prev)
    kl_o = loss_o.kl_loss(mean_o, mean_ref)
    ((lp_o * w).sum() + beta * kl_o).backward()
    npd = npred.to(DEV).requires_grad_(True)
    lp, _, _, kl = ops.sde_logprob_replay(npd, x.to(DEV), prev.to(DEV), sched.timesteps[idx], sched.timesteps,
                                          sched.sigmas, 4.5, 0.8, cfg=True, mean_ref=mean_ref.to(DEV))
    assert kl.shape == (B,)
    # forward value: mu carries the bf16 rounding of the CFG combine (the oracle's v is fp32 here)
    assert torch.allclose(kl.mean().cpu(), kl_o.detach(), rtol=2e-2)
    ((lp * w.to(DEV)).sum() + beta * kl.mean()).backward()
    got, ref = npd.grad.float().cpu(), npo.grad
    assert (got - ref).abs().max() <= 0.02 * ref.abs().max()
    assert torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0) > 0.9995
    # the KL gradient alone (no policy term)
    npd2 = npred.to(DEV).requires_grad_(True)
    _, _, _, kl2 = ops.sde_logprob_replay(npd2, x.to(DEV), prev.to(DEV), sched.timesteps[idx], sched.timesteps,
                                          sched.sigmas, 4.5, 0.8, cfg=True, mean_ref=mean_ref.to(DEV))
    kl2.mean().backward()
    npo2 = npred.float().requires_grad_(True)
    vu, vt = npo2.chunk(2)
    _, _, mean_o2, _ = sde_o.sde_step_with_logprob_new(sched.sigmas, idx, vu + 4.5 * (vt - vu), x, 0.8, prev_sample=prev)
    loss_o.kl_loss(mean_o2, mean_ref).backward()
    assert torch.nn.functional.cosine_similarity(npd2.grad.float().cpu().flatten(), npo2.grad.flatten(), dim=0) > 0.9995


def test_flow_sde_variant_golden_g11_and_backward(ops, sched, golden, golden_dir):
    """The Flow-SDE step (`sde_step_with_logprob`, sde.py:13-73) through the same fused kernel: replay and rollout
    against the verbatim reference outputs (golden G11: mean bit-exact, log-prob to fp32 summation order), and the
    backward against autograd of the oracle restatement."""
    t = torch.load(os.path.join(golden_dir, "g11_tensors.pt"))
    idx = golden["G11_step_index"]
    ts = sched.timesteps[idx]
    _, lp, mean, std = ops.cfg_sde_step_logprob(None, t["v"].bfloat16().to(DEV), t["x"].bfloat16().to(DEV), ts,
                                                sched.timesteps, sched.sigmas, 1.0, 0.7,
                                                prev_sample=t["prev"].bfloat16().to(DEV), want_mean=True,
                                                variant=ops.SDE_FLOW_SDE)
    assert torch.equal(mean.cpu(), t["mean"])
    np.testing.assert_allclose(lp.cpu().numpy(), np.array(golden["G11_log_prob"], dtype=np.float32), rtol=3e-6)
    np.testing.assert_array_equal(std.cpu().numpy(), np.array(golden["G11_std"], dtype=np.float32))
    prev, lp_r, mean_r, std_r = ops.cfg_sde_step_logprob(None, t["v"][:2].bfloat16().to(DEV), t["x"][:2].bfloat16().to(DEV),
                                                         sched.timesteps[2:3], sched.timesteps, sched.sigmas, 1.0, 0.7,
                                                         noise=t["noise"].to(DEV), want_mean=True, variant=ops.SDE_FLOW_SDE)
    assert torch.equal(mean_r.cpu(), t["mean_rollout"])
    assert torch.equal(prev.cpu(), t["prev_rollout"].bfloat16())             # stored latents are bf16 (fast.py:654-655)
    np.testing.assert_allclose(lp_r.cpu().numpy(), np.array(golden["G11_rollout_log_prob"], dtype=np.float32), rtol=3e-6)
    np.testing.assert_array_equal(std_r.cpu().numpy(), np.full(2, golden["G11_rollout_std"][0], dtype=np.float32))
    # the reference-surface mirror
    from adv_grpo_b200.diffusers_patch.sd3_sde_with_logprob import sde_step_with_logprob
    _, lp_m, mean_m, std_m = sde_step_with_logprob(sched, t["v"].to(DEV), ts, t["x"].to(DEV), noise_level=0.7,
                                                   prev_sample=t["prev"].to(DEV))
    assert torch.equal(mean_m.cpu(), t["mean"]) and std_m.shape == (4, 1, 1, 1)
    # backward (CFG batch) vs autograd of the oracle
    g = torch.Generator().manual_seed(13)
    B, shape = 4, (16, 16, 16)
    npred = torch.randn(2 * B, *shape, generator=g).bfloat16()
    x, prev_s = t["x"].bfloat16(), t["prev"].bfloat16()
    w = torch.randn(B, generator=g)
    npo = npred.float().requires_grad_(True)
    vu, vt = npo.chunk(2)
    _, lp_o, _, _ = sde_o.sde_step_with_logprob(sched.sigmas, idx, vu + 4.5 * (vt - vu), x, 0.7, prev_sample=prev_s)
    (lp_o * w).sum().backward()
    npd = npred.to(DEV).requires_grad_(True)
    lp_d, _, _ = ops.sde_logprob_replay(npd, x.to(DEV), prev_s.to(DEV), ts, sched.timesteps, sched.sigmas, 4.5, 0.7,
                                        cfg=True, variant=ops.SDE_FLOW_SDE)
    (lp_d * w.to(DEV)).sum().backward()
    got, ref = npd.grad.float().cpu(), npo.grad
    assert (got - ref).abs().max() <= 0.02 * ref.abs().max()
    assert torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0) > 0.9995


# ------------------------------------------------------------------ A9 advantage
def _keys_from_prompts(prompts, L=8):
    table = {}
    rows = []
    for p in prompts:
        if p not in table:
            rng = np.random.RandomState(len(table) + 100)
            table[p] = rng.randint(0, 49407, size=L)
        rows.append(table[p])
    return torch.tensor(np.stack(rows), dtype=torch.int64)


def test_advantage_golden_g1_g2_g3(ops, golden):
    p = ['a', 'b', 'a', 'c', 'b', 'a']
    keys = _keys_from_prompts(p).to(DEV)
    r = torch.tensor([1, 2, 3, 4, 5, 6], dtype=torch.float32, device=DEV)
    a1, st = ops.group_advantage(r, keys, global_std=False)
    a2, _ = ops.group_advantage(r, keys, global_std=True)
    assert a1.dtype == torch.float64
    np.testing.assert_allclose(a1.cpu().numpy(), golden["G1"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(a2.cpu().numpy(), golden["G2"], rtol=1e-12, atol=1e-12)
    assert st[0].item() == 3 and abs(st[1].item() - golden["G1_stats"][0]) < 1e-12
    keys = _keys_from_prompts(['p', 'p', 'q', 'q']).to(DEV)
    r = torch.tensor([[1, 1], [2, 2], [3, 3], [4, 4]], dtype=torch.float32, device=DEV)
    a3, _ = ops.group_advantage(r, keys, global_std=True)
    np.testing.assert_allclose(a3.cpu().numpy(), golden["G3"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("global_std", [True, False])
def test_advantage_seeded_groups_golden_g7(ops, golden, global_std):
    prompts = golden["G7_prompts"]
    r = torch.tensor(golden["G7_rewards"], dtype=torch.float32)
    r2 = r[:, None].repeat(1, 2)
    adv, st = ops.group_advantage(r2.to(DEV), _keys_from_prompts(prompts, L=256).to(DEV), global_std=global_std)
    # tolerance: float64 summation order only
    np.testing.assert_allclose(adv.cpu().numpy(), golden[f"G7_adv_global{int(global_std)}"], rtol=1e-10, atol=1e-10)
    ratio, mean_std = st_o.zero_std_ratio(prompts, r.numpy())
    assert st[0].item() == 6 and st[1].item() == 8
    assert abs(st[2].item() - ratio) < 1e-12
    assert abs(st[3].item() - mean_std) < 1e-6            # the reference computes this one in float32


def test_advantage_large_ragged_groups(ops):
    rng = np.random.RandomState(3)
    N = 1500
    ids = rng.randint(0, 97, size=N)
    prompts = [f"p{i}" for i in ids]
    r = rng.randn(N, 3).astype(np.float32)
    adv, _ = ops.group_advantage(torch.tensor(r).to(DEV), torch.tensor(ids, dtype=torch.int64).to(DEV), True)
    np.testing.assert_allclose(adv.cpu().numpy(), st_o.grpo_advantages(prompts, r, True), rtol=1e-9, atol=1e-9)
    adv, _ = ops.group_advantage(torch.tensor(r).to(DEV), torch.tensor(ids, dtype=torch.int64).to(DEV), False)
    np.testing.assert_allclose(adv.cpu().numpy(), st_o.grpo_advantages(prompts, r, False), rtol=1e-9, atol=1e-9)


# ------------------------------------------------------------------ A11 loss
@pytest.mark.parametrize("B", [1, 8, 33])
def test_grpo_clip_loss_and_grad(ops, B):
    g = torch.Generator().manual_seed(B)
    lp_old = -torch.rand(B, generator=g)
    lp = lp_old + 4e-5 * torch.randn(B, generator=g)          # ratios straddle 1 +- 1e-5
    adv = 3 * torch.randn(B, generator=g, dtype=torch.float64)
    adv[0] = 7.5                                              # exercises adv_clip_max
    if B > 2:
        adv[2] = 0.0
    lpo = lp.clone().requires_grad_(True)
    loss_ref, info = loss_o.grpo_clip_loss(lpo, lp_old, adv, 1e-5, 5.0)
    loss_ref.backward()
    lpd = lp.to(DEV).requires_grad_(True)
    loss, stats = ops.grpo_clip_loss(lpd, lp_old.to(DEV), adv.to(DEV), 1e-5, 5.0)
    loss.backward()
    assert loss.dtype == torch.float64
    # advantages/step losses within 1e-3 rel (north_star); we are far inside it
    np.testing.assert_allclose(loss.item(), loss_ref.item(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(stats[1].item(), info["approx_kl"].item(), rtol=1e-5, atol=1e-15)
    for k, name in ((2, "clipfrac"), (3, "clipfrac_gt_one"), (4, "clipfrac_lt_one")):
        assert abs(stats[k].item() - info[name].item()) < 1e-6   # the reference holds these in float32
    np.testing.assert_allclose(lpd.grad.cpu().numpy(), lpo.grad.numpy(), rtol=1e-5, atol=1e-9)


# ------------------------------------------------------------------ adaLN LN-modulate, QK norm
@pytest.mark.parametrize("D,S,dual", [(1536, 77, True), (1536, 33, False), (256, 5, True), (768, 9, False)])
def test_ln_modulate_fwd_bwd(ops, D, S, dual):
    B = 3
    g = torch.Generator().manual_seed(D + S)
    x = (torch.randn(B, S, D, generator=g) * 2 + 0.5).bfloat16()
    emb = (0.3 * torch.randn(B, 9 * D, generator=g)).bfloat16()
    sh, sc, sh2, sc2 = emb[:, :D], emb[:, D:2 * D], emb[:, 6 * D:7 * D], emb[:, 7 * D:8 * D]
    xr = x.float().requires_grad_(True)
    n = torch.nn.functional.layer_norm(xr, (D,), eps=1e-6)
    y_ref = n * (1 + sc.float()[:, None]) + sh.float()[:, None]
    y2_ref = n * (1 + sc2.float()[:, None]) + sh2.float()[:, None]
    gy = torch.randn(B, S, D, generator=g).bfloat16()
    gy2 = torch.randn(B, S, D, generator=g).bfloat16()
    ((y_ref * gy.float()).sum() + ((y2_ref * gy2.float()).sum() if dual else 0)).backward()
    xd = x.to(DEV).requires_grad_(True)
    embd = emb.to(DEV)
    chunks = (embd[:, :D], embd[:, D:2 * D], embd[:, 6 * D:7 * D], embd[:, 7 * D:8 * D])
    if dual:
        y, y2 = ops.ln_modulate(xd, chunks[0], chunks[1], chunks[2], chunks[3])
        ((y.float() * gy.to(DEV).float()).sum() + (y2.float() * gy2.to(DEV).float()).sum()).backward()
        assert (y2.float().cpu() - y2_ref).abs().max() <= 2 ** -7 * y2_ref.abs().max()
    else:
        y = ops.ln_modulate(xd, chunks[0], chunks[1])
        (y.float() * gy.to(DEV).float()).sum().backward()
    # one bf16 rounding of the output: half an ulp relative to the largest magnitude
    assert (y.float().cpu() - y_ref.detach()).abs().max() <= 2 ** -7 * y_ref.abs().max()
    gref = xr.grad
    assert (xd.grad.float().cpu() - gref).abs().max() <= 2 ** -6 * gref.abs().max()


@pytest.mark.parametrize("S_txt", [0, 13])
def test_qk_norm_concat_fwd_bwd(ops, S_txt):
    B, S_img, H, D = 2, 20, 8, 64
    g = torch.Generator().manual_seed(5 + S_txt)
    qi = torch.randn(B, S_img, 3 * H * D, generator=g).bfloat16()
    qt = torch.randn(B, S_txt, 3 * H * D, generator=g).bfloat16() if S_txt else None
    ws = [(1 + 0.1 * torch.randn(D, generator=g)).bfloat16() for _ in range(4)]

    def ref(qi, qt):
        def norm(t, wq, wk):
            t = t.view(t.shape[0], t.shape[1], 3, H, D)
            q, k, v = t[:, :, 0], t[:, :, 1], t[:, :, 2]
            rn = lambda z, w: z * torch.rsqrt(z.pow(2).mean(-1, keepdim=True) + 1e-6) * w.float()
            return torch.stack([rn(q, wq), rn(k, wk), v], dim=2)
        parts = [norm(qi, ws[0], ws[1])]
        if qt is not None:
            parts.append(norm(qt, ws[2], ws[3]))
        return torch.cat(parts, dim=1)

    qir = qi.float().requires_grad_(True)
    qtr = qt.float().requires_grad_(True) if S_txt else None
    out_ref = ref(qir, qtr)
    go = torch.randn(out_ref.shape, generator=g).bfloat16()
    (out_ref * go.float()).sum().backward()
    qid = qi.to(DEV).requires_grad_(True)
    qtd = qt.to(DEV).requires_grad_(True) if S_txt else None
    wd = [w.to(DEV) for w in ws]
    out = ops.qk_norm_concat(qid, qtd, wd[0], wd[1], wd[2] if S_txt else None, wd[3] if S_txt else None, H, D)
    (out.float() * go.to(DEV).float()).sum().backward()
    assert out.shape == (B, S_img + S_txt, 3, H, D)
    assert (out.float().cpu() - out_ref.detach()).abs().max() <= 2 ** -7 * out_ref.abs().max()
    assert (qid.grad.float().cpu() - qir.grad).abs().max() <= 2 ** -6 * qir.grad.abs().max()
    if S_txt:
        assert (qtd.grad.float().cpu() - qtr.grad).abs().max() <= 2 ** -6 * qtr.grad.abs().max()
    # no-norm variant (SD3-medium): pure concat
    out2 = ops.qk_norm_concat(qid.detach(), qtd.detach() if S_txt else None, None, None, None, None, H, D)
    assert torch.equal(out2[:, :S_img].reshape(B, S_img, -1), qid.detach())


# ------------------------------------------------------------------ reward preprocessing
@pytest.mark.parametrize("H", [512, 256, 128])
def test_clip_preprocess_bit_exact_with_pillow(ops, H):
    from PIL import Image
    g = torch.Generator().manual_seed(H)
    img = torch.rand(2, 3, H, H, generator=g)
    img[0, :, : H // 2] = torch.linspace(0, 1, H)[None, None, :].expand(3, H // 2, H)   # smooth region
    img = img.bfloat16()
    pix, u8 = ops.clip_preprocess(img.to(DEV), 224, dtype=torch.float32, want_u8=True)
    q = pre_o.quantise_bf16(img)                                   # rewards.py:581 on the bf16 tensor
    ref_u8 = pre_o.pil_bicubic_resize_u8(q.numpy(), 224)
    assert np.array_equal(u8.cpu().numpy(), ref_u8), "resized bytes must be bit-exact (integer work)"
    for b in range(2):                                             # and against Pillow itself
        pil = np.array(Image.fromarray(q[b].permute(1, 2, 0).numpy()).resize((224, 224), resample=Image.BICUBIC))
        assert np.array_equal(u8[b].permute(1, 2, 0).cpu().numpy(), pil)
    np.testing.assert_allclose(pix.cpu().numpy(), pre_o.clip_pixel_values(ref_u8), rtol=0, atol=2e-7)
    pix_bf16 = ops.clip_preprocess(img.to(DEV), 224)
    assert torch.equal(pix_bf16.cpu(), torch.from_numpy(pre_o.clip_pixel_values(ref_u8)).bfloat16()) or \
        (pix_bf16.float().cpu() - torch.from_numpy(pre_o.clip_pixel_values(ref_u8))).abs().max() < 2 ** -6


def test_dino_preprocess_matches_torch_bicubic(ops):
    from oracle import dinov2 as dino_o
    img = torch.rand(2, 3, 512, 512, generator=torch.Generator().manual_seed(1))
    ref = dino_o.preprocess(img)
    got = ops.dino_preprocess(img.to(DEV), 518).float().cpu()
    assert got.shape == (2, 3, 518, 518)
    # fp32 interpolation + one bf16 rounding of values up to ~2.7
    assert (got - ref).abs().max() < 2 ** -6
    got_bf = ops.dino_preprocess(img.bfloat16().to(DEV), 518).float().cpu()
    ref_bf = dino_o.preprocess(img.bfloat16().float())
    assert (got_bf - ref_bf).abs().max() < 0.05


# ------------------------------------------------------------------ VAE GroupNorm + SiLU (A6 glue)
@pytest.mark.parametrize("C,H,W,silu", [(128, 24, 20, True), (256, 16, 16, True), (512, 9, 7, False), (512, 32, 32, True)])
def test_group_norm_silu_nhwc(ops, C, H, W, silu):
    g = torch.Generator().manual_seed(C + H)
    x = (torch.randn(3, C, H, W, generator=g) * 2 + 0.7)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ref = torch.nn.functional.group_norm(x, 32, gamma, beta, eps=1e-6)
    if silu:
        ref = torch.nn.functional.silu(ref)
    xd = x.to(DEV).contiguous(memory_format=torch.channels_last)
    got = ops.group_norm_silu_nhwc(xd, gamma.to(DEV), beta.to(DEV), 32, 1e-6, silu)
    assert got.shape == x.shape and got.is_contiguous(memory_format=torch.channels_last)
    # fp32 statistics accumulated in double; __expf in SiLU
    assert torch.allclose(got.cpu(), ref, rtol=2e-5, atol=2e-5)


def test_vae_decoder_matches_oracle_at_true_widths():
    from adv_grpo_b200 import weights
    from adv_grpo_b200.vae import AutoencoderKL, VaeImageProcessor
    from oracle import vae as vae_o
    vp = weights.init_vae_decoder(weights.VAE_SD3, seed=2, device="cpu")
    vae = AutoencoderKL(vp, weights.VAE_SD3, device=DEV)
    z = torch.randn(2, 16, 8, 8, generator=torch.Generator().manual_seed(0))
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        got = VaeImageProcessor().postprocess(vae.decode(z.to(DEV) / 1.5305 + 0.0609)[0], output_type="pt").cpu()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    ref = vae_o.decode_latents_to_image(vp, z)
    assert got.shape == (2, 3, 64, 64)
    # the convolutions run in TF32 (tcgen05 kind::tf32 here, cuDNN TF32 in the reference's fp32 VAE under PyTorch's
    # default cudnn.allow_tf32): bound our deviation from the fp32 oracle by that of the library TF32 path on the
    # same decoder, and report both
    # library TF32 path: the oracle's own decoder (torch convolutions / matmuls) on the GPU with TF32 allowed
    prev_mm = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        lib = vae_o.decode_latents_to_image({k: v.to(DEV) for k, v in vp.items()}, z.to(DEV)).cpu()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
        torch.backends.cuda.matmul.allow_tf32 = prev_mm
    err, err_lib = (got - ref).abs().max().item(), (lib - ref).abs().max().item()
    print(f"VAE decode max |err| vs fp32 oracle: tcgen05 TF32 {err:.2e}, cuDNN TF32 {err_lib:.2e}")
    assert err < max(2e-3, 2.0 * err_lib), (err, err_lib)
    assert (got - ref).abs().mean().item() < 5e-4


def test_vae_fused_helpers(ops):
    g = torch.Generator().manual_seed(0)
    a = torch.randn(2, 128, 6, 10, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    b = torch.randn(2, 128, 6, 10, generator=g).to(DEV).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(128, generator=g).to(DEV)
    assert torch.allclose(ops.add_bias_nhwc(a, b, bias), a + b + bias[None, :, None, None], atol=1e-6)
    assert torch.equal(ops.add_bias_nhwc(a, b), a + b)
    up = ops.upsample_nearest2x_nhwc(a)
    # nearest-2x, handed to the upsampler's TF32 convolution as TF32 values (round to nearest, ties away: the low 13
    # mantissa bits are zero)
    a_tf32 = ((a.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    assert torch.equal(up, torch.nn.functional.interpolate(a_tf32, scale_factor=2.0, mode="nearest"))
    assert up.is_contiguous(memory_format=torch.channels_last)
    gamma, beta = torch.randn(128, generator=g).to(DEV), torch.randn(128, generator=g).to(DEV)
    got = ops.group_norm_silu_nhwc(a, gamma, beta, 32, 1e-6, True, in_bias=bias)
    ref = torch.nn.functional.silu(torch.nn.functional.group_norm(a + bias[None, :, None, None], 32, gamma, beta, eps=1e-6))
    assert torch.allclose(got, ref, rtol=2e-5, atol=2e-5)
    got_r = ops.group_norm_silu_nhwc(a, gamma, beta, 32, 1e-6, True, in_bias=bias, round_tf32=True)
    assert torch.equal(got_r, ((got.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32))


# ------------------------------------------------------------------ A6 convolution (tcgen05 TF32 implicit GEMM)
@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [(2, 16, 16, 64, 128, 3), (1, 64, 64, 512, 512, 3), (2, 5, 20, 32, 96, 3),
                                              (1, 128, 128, 256, 128, 3), (2, 32, 32, 512, 256, 1), (1, 7, 9, 64, 32, 1),
                                              (3, 8, 256, 128, 128, 3), (2, 17, 40, 64, 320, 3)])
@pytest.mark.parametrize("variant", [0, 1])
def test_conv2d_nhwc_tf32_matches_torch_fp32(ops, B, H, W, Cin, Cout, k, variant):
    """3x3 (zero padding from the TMA out-of-bounds fill, incl. negative start coordinates) and 1x1 convolutions,
    ragged patches (W, H not multiples of the patch), vs an fp32 torch convolution.  TF32 inputs (10-bit mantissa),
    fp32 accumulation: error bound ~2^-10 per operand relative to the accumulated magnitude."""
    g = torch.Generator(device=DEV).manual_seed(B + H + W + Cin + Cout)
    x = torch.randn(B, Cin, H, W, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    w = torch.randn(Cout, Cin, k, k, device=DEV, generator=g) / (Cin * k * k) ** 0.5
    bias = torch.randn(Cout, device=DEV, generator=g)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = torch.nn.functional.conv2d(x.double(), w.double(), bias.double(), padding=k // 2).float()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    wp = ops.pack_conv_weight_tf32(w)
    from adv_grpo_b200 import _lib
    _lib.load().advgrpo_debug_set_conv_variant(variant)      # 0 = auto (CTA pairs for >= 256 output channels), 1 = single CTA
    try:
        got = ops.conv2d_nhwc_tf32(x, wp, bias, k)
        got_nb = ops.conv2d_nhwc_tf32(x, wp, None, k)
    finally:
        _lib.load().advgrpo_debug_set_conv_variant(0)
    assert got.shape == ref.shape and got.is_contiguous(memory_format=torch.channels_last)
    err = (got - ref).abs().max().item()
    assert err <= 4e-3 * ref.abs().max().item(), (err, ref.abs().max().item())
    assert torch.allclose(got_nb + bias[None, :, None, None], got, atol=1e-5)
    # borders are where the out-of-bounds fill matters: check them separately
    for sl in ((slice(None), slice(None), 0), (slice(None), slice(None), H - 1), (slice(None), slice(None), slice(None), 0),
               (slice(None), slice(None), slice(None), W - 1)):
        assert (got[sl] - ref[sl]).abs().max().item() <= 4e-3 * ref.abs().max().item()


# ------------------------------------------------------------------ A12 clip + AdamW
@pytest.mark.parametrize("n", [5, 1_000_003, 18_776_064])
def test_clip_adamw_matches_torch_adamw(ops, n):
    """train_sd3_fast_pickscore.py:1165-1171 with the optimizer of :515-521: clip_grad_norm_(max_norm) ->
    torch.optim.AdamW.step() -> zero_grad().  The reference's own optimizer (torch.optim.AdamW, CPU fp32,
    single-tensor path) is the checker.  Three steps: clipped (norm >> 1), clipped, not clipped (tiny gradient).
    Tolerance: 1e-5 relative on p / m / v (fp32; FMA contraction differs by an ulp), 2e-6 on the norm (vs float64)."""
    from adv_grpo_b200.optim import FlatClipAdamW
    g = torch.Generator().manual_seed(n % 1000)
    p0 = torch.randn(n, generator=g) * 0.18
    hp = dict(lr=3e-4, betas=(0.9, 0.999), weight_decay=1e-4, eps=1e-8)
    ref_p = torch.nn.Parameter(p0.clone())
    ref_opt = torch.optim.AdamW([ref_p], foreach=False, fused=False, **hp)
    our_p = torch.nn.Parameter(p0.clone().to(DEV))
    our_opt = FlatClipAdamW([our_p], max_grad_norm=1.0, **hp)
    for step, scale in enumerate((1.0, 0.05, 1e-6)):
        grad = torch.randn(n, generator=g) * scale
        # clip_grad_norm_ with the norm taken in float64: torch's own fp32 CPU reduction is 1e-5 off at 1e6 elements
        # and 9e-4 off at 1.9e7 (measured), the kernel accumulates its block partials in f64
        ref_norm = grad.double().norm().float()
        ref_p.grad = grad * torch.clamp(1.0 / (ref_norm + 1e-6), max=1.0)
        torch.testing.assert_close(ref_norm, grad.norm(), rtol=2e-3, atol=0)
        ref_opt.step()
        our_p.grad = grad.clone().to(DEV)
        norm = our_opt.step()
        assert torch.count_nonzero(our_p.grad).item() == 0            # cleared in the same pass
        torch.testing.assert_close(norm.cpu()[0], ref_norm, rtol=2e-6, atol=0)
        st = ref_opt.state[ref_p]
        torch.testing.assert_close(our_p.detach().cpu(), ref_p.detach(), rtol=1e-5, atol=1e-7)
        for name in ("exp_avg", "exp_avg_sq"):           # atol scaled to the tensor: sums that cancel keep ulp-level errors
            torch.testing.assert_close(our_opt.state[our_p][name].cpu(), st[name], rtol=1e-5,
                                       atol=1e-6 * st[name].abs().max().item())
    # zero_grad=False leaves the clipped gradient in place, like clip_grad_norm_
    grad = torch.randn(n, generator=g)
    our_p.grad = grad.clone().to(DEV)
    norm = our_opt.step(zero_grad=False)
    want = grad * min(1.0 / (grad.double().norm().item() + 1e-6), 1.0)
    torch.testing.assert_close(our_p.grad.cpu(), want, rtol=1e-5, atol=1e-9)
    # state dict round trip keeps the torch AdamW names
    sd = our_opt.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and int(sd["state"][0]["step"]) == 4


# ------------------------------------------------------------------ A8 glue: affine LayerNorm of the reward towers
@pytest.mark.parametrize("rows,D,eps", [(2 * 257, 1280, 1e-5), (3 * 1370, 768, 1e-6), (77, 1024, 1e-5), (1, 256, 1e-5)])
def test_layer_norm_affine_matches_torch_fp32(ops, rows, D, eps):
    """CLIPEncoderLayer.layer_norm1/2 (transformers, pickscore_scorer.py:40-43) / timm Block.norm1/2
    (rewards.py:397-399): fp32 statistics over the bf16 row, bf16 output.  Checked against torch's fp32
    LayerNorm of the same bf16 values: at most one bf16 rounding step apart (2^-8 relative) per element."""
    g = torch.Generator().manual_seed(rows + D)
    x = (torch.randn(rows, D, generator=g) * 3 + 0.5).bfloat16()
    w = (1 + 0.2 * torch.randn(D, generator=g)).bfloat16()
    b = (0.1 * torch.randn(D, generator=g)).bfloat16()
    y = ops.layer_norm(x.to(DEV), w.to(DEV), b.to(DEV), eps)
    assert y.dtype == torch.bfloat16 and y.shape == x.shape
    ref = torch.nn.functional.layer_norm(x.float(), (D,), w.float(), b.float(), eps)
    err = (y.float().cpu() - ref).abs()
    assert (err <= ref.abs() * 2.0 ** -8 + 1e-6).all(), err.max()
    # 3-D input keeps its shape
    x3 = x.to(DEV).view(1, rows, D)
    assert ops.layer_norm(x3, w.to(DEV), b.to(DEV), eps).shape == x3.shape
    # native backward (A14: the LayerNorms of the trainable CLIP blocks): dx, d weight, d bias vs torch fp32 autograd
    xg, wg, bg = (t.to(DEV).clone().requires_grad_() for t in (x, w, b))
    dy = torch.randn(rows, D, generator=g).bfloat16()
    ops.layer_norm(xg, wg, bg, eps).backward(dy.to(DEV))
    xr, wr, br = (t.float().clone().requires_grad_() for t in (x, w, b))
    torch.nn.functional.layer_norm(xr, (D,), wr, br, eps).backward(dy.float())
    for got, ref_g in ((xg.grad, xr.grad), (wg.grad, wr.grad), (bg.grad, br.grad)):
        assert got.dtype == torch.bfloat16
        rel = (got.float().cpu() - ref_g).norm().item() / max(ref_g.norm().item(), 1e-12)
        assert rel < 6e-3, rel                                   # one bf16 rounding of the result
    # frozen parameters (post_layernorm during the D step): only dx
    xg2 = x.to(DEV).clone().requires_grad_()
    ops.layer_norm(xg2, w.to(DEV), b.to(DEV), eps).backward(dy.to(DEV))
    assert torch.equal(xg2.grad, xg.grad)


def test_stat_tracker_device_path_counts_distinct_prompts_and_handles_long_T(ops):
    """`update_device` (gathered prompt-id rows on the device): `get_stats()` reports the reference's
    (avg group size, trained_prompt_num = distinct prompts ever seen, stat_tracking.py:36-37,72-75) across epochs, and the
    fast grpo kernel accepts T = 20 columns (config 4 trains on up to num_steps timesteps)."""
    from adv_grpo_b200.stat_tracking import PerPromptStatTracker
    tr = PerPromptStatTracker(global_std=True, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(5)
    T = 20
    for epoch, base in enumerate((0, 2)):                                # prompts {0,1,2,3} then {2,3,4,5}
        ids = (torch.arange(32, device=DEV) // 8 + base)[:, None].expand(32, 256).contiguous()
        r = torch.randn(32, device=DEV, generator=g)[:, None].repeat(1, T)
        adv = tr.update_device(ids, r)
        rr = r.double().reshape(4, 8, T)
        ref = ((rr - rr.mean(1, keepdim=True)) / (r.double().std(0, unbiased=False) + 1e-4)).reshape(32, T)
        assert (adv - ref).abs().max().item() < 1e-9
        size, seen = tr.get_stats()
        assert size == 8.0 and seen == (4 if epoch == 0 else 6)
        tr.clear()



# ------------------------------------------------------------------ 8f-3 reference-image resize (Pillow BILINEAR + ToTensor)
@pytest.mark.parametrize("H,W,out", [(480, 640, 512), (200, 300, 512), (512, 512, 512), (1500, 1000, 512), (37, 53, 64),
                                     (2048, 3072, 1024)])
def test_pil_resize_bilinear_bit_exact_with_pillow(ops, H, W, out):
    """transforms.Resize((S, S)) on a PIL image + ToTensor() (train_sd3_fast_pickscore.py:791-797) on the device: every
    resized byte equals Pillow's antialiased BILINEAR resize (down- and up-scaling, non-square inputs), and the float
    output is byte / 255 in float32."""
    from PIL import Image
    rng = np.random.default_rng(H + W)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((out, out), Image.BILINEAR))            # [out, out, 3]
    got, u8 = ops.pil_resize_bilinear(torch.from_numpy(img).to(DEV), out, out, want_u8=True)
    assert torch.equal(u8.cpu(), torch.from_numpy(ref.copy()).permute(2, 0, 1))
    assert torch.equal(got.cpu(), torch.from_numpy(ref.copy()).permute(2, 0, 1).float().div(255.0))


def test_reference_image_index_gpu_path_equals_pil_path(tmp_path):
    """ReferenceImageIndex on a CUDA device (host decode, device resize) returns exactly what the PIL path returns, for a
    JPEG and a PNG of different sizes, including the missing-file fallback."""
    import json
    from PIL import Image
    from adv_grpo_b200.reference_images import ReferenceImageIndex
    rng = np.random.default_rng(3)
    Image.fromarray(rng.integers(0, 256, (300, 420, 3), dtype=np.uint8)).save(tmp_path / "a.jpg", quality=90)
    Image.fromarray(rng.integers(0, 256, (640, 512, 3), dtype=np.uint8)).save(tmp_path / "b.png")
    Image.fromarray(rng.integers(0, 256, (64, 64, 3), dtype=np.uint8)).save(tmp_path / "default.png")
    Image.fromarray(rng.integers(0, 256, (200, 333, 3), dtype=np.uint8)).save(tmp_path / "c.jpg", quality=80, progressive=True)
    Image.fromarray(rng.integers(0, 256, (90, 70), dtype=np.uint8)).save(tmp_path / "d.png")
    Image.fromarray(np.full((16, 16, 4), 100, dtype=np.uint8), mode="CMYK").save(tmp_path / "e.jpg")      # Pillow's job
    bad = bytearray((tmp_path / "b.png").read_bytes())
    bad[-14] ^= 0xFF                                                   # damaged IDAT checksum: refused here, Pillow does not check it
    (tmp_path / "f.png").write_bytes(bytes(bad))
    files = ["a.jpg", "b.png", "missing.jpg", "c.jpg", "d.png", "e.jpg", "f.png", "a.jpg"]
    (tmp_path / "index.json").write_text(json.dumps({"a cat": files}))
    kw = dict(size=256, default_image=str(tmp_path / "default.png"))
    gpu = ReferenceImageIndex(str(tmp_path / "index.json"), str(tmp_path), device=DEV, **kw)("a cat")      # host stages on a thread pool
    one = ReferenceImageIndex(str(tmp_path / "index.json"), str(tmp_path), device=DEV, host_threads=1, **kw)("a cat")
    cpu = ReferenceImageIndex(str(tmp_path / "index.json"), str(tmp_path), device="cpu", **kw)("a cat")
    assert gpu.is_cuda and gpu.shape == (len(files), 3, 256, 256) and gpu.dtype == torch.float32
    assert torch.equal(gpu.cpu(), cpu) and torch.equal(one.cpu(), cpu)


# ------------------------------------------------------------------ 8f-3 JPEG decode (host Huffman + device IDCT / colour)
@pytest.mark.parametrize("h,w,kw", [(512, 512, dict(quality=90, subsampling=2)), (768, 1024, dict(quality=85, subsampling=2)),
                                    (333, 517, dict(quality=75, subsampling=1)), (600, 401, dict(quality=95, subsampling=0)),
                                    (480, 640, dict(quality=60, subsampling=2, restart_marker_blocks=4)),
                                    (257, 129, dict(quality=80, gray=True)), (1, 1, dict(quality=90, subsampling=2)),
                                    (17, 9, dict(quality=100, subsampling=2)),
                                    (512, 768, dict(quality=85, subsampling=2, progressive=True)),
                                    (301, 203, dict(quality=92, subsampling=1, progressive=True)),
                                    (480, 640, dict(quality=70, subsampling=0, progressive=True, restart_marker_blocks=8)),
                                    # narrow planes: libjpeg replicates instead of the triangle filter when ceil(W / 2) <= 2
                                    (47, 3, dict(quality=100, subsampling=2)), (16, 4, dict(quality=90, subsampling=1)),
                                    (30, 2, dict(quality=80, subsampling=2, progressive=True)), (9, 5, dict(quality=90, subsampling=2))])
def test_jpeg_decode_bit_exact_with_pillow(h, w, kw):
    """`Image.open(path).convert("RGB")` (train_sd3_fast_pickscore.py:779) on the device: every byte equals Pillow's decode."""
    import io
    from PIL import Image
    from adv_grpo_b200 import jpeg as jpeg_b
    from jpeg_util import _jpeg_bytes
    data = _jpeg_bytes(h, w, seed=h + w, **kw)
    ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    got = jpeg_b.decode_jpeg_to_device(data, DEV)
    assert got.dtype == torch.uint8 and got.shape == (h, w, 3)
    assert torch.equal(got.cpu(), torch.from_numpy(ref.copy()))
    from jpeg_util import _cmyk_jpeg
    assert jpeg_b.decode_jpeg_to_device(_cmyk_jpeg(), DEV) is None


# ------------------------------------------------------------------ 8f-3 PNG decode (host inflate + device wavefront unfilter)
@pytest.mark.parametrize("h,w,mode", [(512, 512, "RGB"), (300, 1500, "RGB"), (1500, 300, "RGB"), (2100, 64, "RGBA"),
                                      (257, 129, "L"), (100, 100, "LA"), (333, 517, "P"), (1, 1, "RGB"), (1025, 3, "RGB")])
def test_png_decode_bit_exact_with_pillow(h, w, mode):
    """`Image.open(path).convert("RGB")` (train_sd3_fast_pickscore.py:779; the reference images are PNG files) on the device:
    every byte equals Pillow's decode -- all colour types, images taller than one 1024-row wavefront band."""
    import io
    from PIL import Image
    from adv_grpo_b200 import png as png_b
    from png_util import pillow_png
    data = pillow_png(h, w, mode, seed=h + w)
    ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    got = png_b.decode_png_to_device(data, DEV)
    assert got is not None and got.dtype == torch.uint8 and got.shape == (h, w, 3)
    assert torch.equal(got.cpu(), torch.from_numpy(ref.copy()))


@pytest.mark.parametrize("ct", [0, 2, 3, 4, 6])
def test_png_decode_all_filter_types(ct):
    """Hand-assembled files: random filter type per row (None / Sub / Up / Average / Paeth -- Pillow's encoder never emits
    Average), stored and compressed deflate blocks, IDAT split into 100-byte chunks; 1300 rows = two wavefront bands."""
    import io
    from PIL import Image
    from adv_grpo_b200 import png as png_b
    from png_util import handmade_png
    for h, w, level, split, kind in ((1300, 37, 6, 100, 0), (64, 200, 0, None, 1), (5, 1, 9, 5, 2)):
        data, _ = handmade_png(h, w, ct, seed=ct + h, level=level, split=split, kind=kind)
        ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        got = png_b.decode_png_to_device(data, DEV)
        assert torch.equal(got.cpu(), torch.from_numpy(ref.copy())), (h, w, ct)


@pytest.mark.parametrize("ct,bd", [(0, 1), (0, 2), (0, 4), (0, 16), (2, 16), (6, 16), (3, 1), (3, 2), (3, 4), (3, 8), (4, 8), (4, 16)])
def test_png_decode_every_bit_depth(ct, bd):
    """Sub-byte greyscale (scaled to 0..255) and palette samples, 16-bit greyscale (Pillow clips I;16 to 255) and 16-bit
    truecolour (+ alpha) and greyscale + alpha (the high byte of every sample): what Pillow's convert("RGB") returns, with random filter types
    (filter pixels of 1, 2, 6 and 8 bytes) and 1100 rows = two wavefront bands."""
    import io
    from PIL import Image
    from adv_grpo_b200 import png as png_b
    from png_util import handmade_png
    for h, w, kind in ((1100, 29, 0), (7, 130, 1)):
        data, _ = handmade_png(h, w, ct, seed=ct * 17 + bd + h, bd=bd, kind=kind)
        ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        got = png_b.decode_png_to_device(data, DEV)
        assert got is not None and torch.equal(got.cpu(), torch.from_numpy(ref.copy())), (ct, bd, h, w)
    assert png_b.decode_png_to_device(handmade_png(8, 8, 2, bd=4)[0], DEV) is None      # 4-bit truecolour does not exist in PNG


@pytest.mark.parametrize("ct,bd", [(2, 8), (6, 8), (0, 8), (0, 2), (3, 4), (3, 8), (2, 16), (4, 8), (4, 16), (6, 16)])
def test_png_decode_adam7_interlaced(ct, bd):
    """Adam7-interlaced files: each of the seven reduced images is unfiltered by its own wavefront and scattered into place;
    sizes from 1 x 1 (one pass) to 1300 rows (the last pass alone spans 650 rows)."""
    import io
    from PIL import Image
    from adv_grpo_b200 import png as png_b
    from png_util import handmade_png
    for h, w in ((1300, 21), (37, 150), (1, 1), (2, 3), (8, 8), (5, 4)):
        data, _ = handmade_png(h, w, ct, seed=ct * 31 + bd + h, bd=bd, interlace=1)
        ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        got = png_b.decode_png_to_device(data, DEV)
        assert got is not None and torch.equal(got.cpu(), torch.from_numpy(ref.copy())), (ct, bd, h, w)

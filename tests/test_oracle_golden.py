"""Pin the oracle restatements to the outputs of the reference's own files
(tests/golden/golden.json, produced by tests/golden/make_golden.py from /root/reference)."""
import os

import numpy as np
import torch

from oracle import clip_criterion, ema as ema_o, sde as sde_o, stat_tracking as st_o
from oracle.scheduler import FlowMatchEulerOracle


def test_advantages_g1_g2_g3(golden):
    p = ['a', 'b', 'a', 'c', 'b', 'a']
    r = [1, 2, 3, 4, 5, 6]
    np.testing.assert_allclose(st_o.grpo_advantages(p, r, False), golden["G1"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(st_o.grpo_advantages(p, r, True), golden["G2"], rtol=0, atol=1e-15)
    a = st_o.grpo_advantages(['p', 'p', 'q', 'q'], [[1, 1], [2, 2], [3, 3], [4, 4]], True)
    assert a.dtype == np.float64 and str(a.dtype) == golden["G3_dtype"]
    np.testing.assert_allclose(a, golden["G3"], rtol=0, atol=1e-15)


def test_advantages_seeded_groups(golden):
    r = np.repeat(np.array(golden["G7_rewards"], dtype=np.float32)[:, None], 2, axis=1)
    for gs in (0, 1):
        a = st_o.grpo_advantages(golden["G7_prompts"], r, bool(gs))
        np.testing.assert_allclose(a, golden[f"G7_adv_global{gs}"], rtol=1e-13, atol=1e-13)
    ratio, mean_std = st_o.zero_std_ratio(golden["G7_prompts"], np.array(golden["G7_rewards"], dtype=np.float32))
    assert abs(ratio - 1 / 6) < 1e-12 and mean_std > 0


def test_scheduler_g4(golden):
    s = FlowMatchEulerOracle()
    ts = s.set_timesteps(10)
    np.testing.assert_allclose(s.sigmas.numpy(), golden["G4_sigmas"], rtol=0, atol=0)
    np.testing.assert_allclose(ts.numpy(), golden["G4_timesteps"], rtol=0, atol=0)
    # closed form: sigma_i = 3 s / (1 + 2 s), s = linspace(1, sigma_min, T)
    smin = 3 * 0.001 / (1 + 2 * 0.001)
    lin = np.linspace(1.0, smin, 10)
    np.testing.assert_allclose(s.sigmas[:-1].numpy(), 3 * lin / (1 + 2 * lin), rtol=1e-6)
    assert s.index_for_timestep(ts[3]) == 3


def test_sde_step_g5(golden, golden_dir):
    t = torch.load(os.path.join(golden_dir, "g5_tensors.pt"))
    s = FlowMatchEulerOracle()
    s.set_timesteps(10)
    gen = torch.Generator().manual_seed(1)
    prev, lp, mean, std = sde_o.sde_step_with_logprob_new(s.sigmas, [0], t["v"], t["x"], 0.8, generator=gen)
    assert torch.equal(prev, t["prev"]) and torch.equal(mean, t["mean"])
    np.testing.assert_array_equal(lp.numpy(), np.array(golden["G5_log_prob"], dtype=np.float32))
    np.testing.assert_array_equal(std.flatten().numpy(), np.array(golden["G5_std"], dtype=np.float32))
    _, lp2, _, _ = sde_o.sde_step_with_logprob_new(s.sigmas, [0], t["v"], t["x"], 0.8, prev_sample=prev)
    np.testing.assert_array_equal(lp2.numpy(), np.array(golden["G5_replay_log_prob"], dtype=np.float32))


def test_sde_step_replay_per_sample_g8_g9(golden, golden_dir):
    t = torch.load(os.path.join(golden_dir, "g8_tensors.pt"))
    s = FlowMatchEulerOracle()
    s.set_timesteps(10)
    _, lp, mean, std = sde_o.sde_step_with_logprob_new(s.sigmas, golden["G8_step_index"], t["v"], t["x"], 0.8,
                                                        prev_sample=t["prev"])
    assert torch.equal(mean, t["mean"])
    np.testing.assert_array_equal(lp.numpy(), np.array(golden["G8_log_prob"], dtype=np.float32))
    np.testing.assert_array_equal(std.flatten().numpy(), np.array(golden["G8_std"], dtype=np.float32))
    t5 = torch.load(os.path.join(golden_dir, "g5_tensors.pt"))
    _, lp9, _, std9 = sde_o.sde_step_with_logprob_new(s.sigmas, [9], t5["v"], t5["x"], 0.8,
                                                      generator=torch.Generator().manual_seed(3))
    assert lp9.tolist() == golden["G9_last_step_log_prob"] == [0.0, 0.0]
    assert std9.flatten().tolist() == golden["G9_last_step_std"]


def test_clip_criterion_g6(golden, golden_dir):
    t = torch.load(os.path.join(golden_dir, "g6_tensors.pt"))
    loss = clip_criterion.clip_pair_loss(t["t"], t["i0"], t["i1"], torch.tensor(100.0), torch.tensor(1.0),
                                         torch.tensor(0.0))
    assert abs(loss.item() - golden["G6_loss"]) < 1e-6


def test_ema_g10(golden):
    params = [torch.tensor(p) for p in golden["G10_init"]]
    ema = [p.clone() for p in params]
    for step in range(40):
        for p in params:
            p.add_(0.01 * (step + 1))
        ema_o.ema_step(ema, params, 0.9, 8, step)
        got = [e.sum().item() for e in ema]
        np.testing.assert_allclose(got, golden["G10_ema_sums"][step], rtol=1e-5, atol=1e-5)


def test_flow_sde_step_g11(golden, golden_dir):
    """Flow-SDE variant (sde.py:13-73) against the verbatim reference output: replay with per-sample timesteps
    (incl. sigma == 1 -> sigma_max) and the rollout form with injected noise, bit for bit."""
    t = torch.load(os.path.join(golden_dir, "g11_tensors.pt"))
    s = FlowMatchEulerOracle()
    s.set_timesteps(10)
    _, lp, mean, std = sde_o.sde_step_with_logprob(s.sigmas, golden["G11_step_index"], t["v"], t["x"], 0.7,
                                                   prev_sample=t["prev"])
    assert torch.equal(mean, t["mean"])
    np.testing.assert_array_equal(lp.numpy(), np.array(golden["G11_log_prob"], dtype=np.float32))
    np.testing.assert_array_equal(std.flatten().numpy(), np.array(golden["G11_std"], dtype=np.float32))
    prev, lp_r, mean_r, std_r = sde_o.sde_step_with_logprob(s.sigmas, [2], t["v"][:2], t["x"][:2], 0.7, noise=t["noise"])
    assert torch.equal(prev, t["prev_rollout"]) and torch.equal(mean_r, t["mean_rollout"])
    np.testing.assert_array_equal(lp_r.numpy(), np.array(golden["G11_rollout_log_prob"], dtype=np.float32))
    np.testing.assert_array_equal(std_r.flatten().numpy(), np.array(golden["G11_rollout_std"], dtype=np.float32))


def test_advantage_modes_g12(golden_dir):
    """'rwr' / 'sft' / 'dpo' of PerPromptStatTracker.update: oracle restatement vs the verbatim reference file
    (tests/golden/make_golden_adv_modes.py)."""
    import json
    import os
    with open(os.path.join(golden_dir, "golden_adv_modes.json")) as f:
        g = json.load(f)
    for mode in ("rwr", "sft", "dpo"):
        got = st_o.mode_advantages(g["prompts"], g["rewards"], mode)
        for gs in (0, 1):                                   # global_std does not enter these modes
            np.testing.assert_array_equal(got, np.array(g[f"{mode}_global{gs}"]))
    for mode in ("rwr", "sft"):
        np.testing.assert_array_equal(st_o.mode_advantages(g["prompts"], g["rewards_2d"], mode), np.array(g[f"{mode}_2d"]))

"""Generate tests/golden/*.json / *.pt by EXECUTING THE REFERENCE'S OWN FILES.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The GPU box has no /root/reference; tests read only the committed outputs.

Reference files executed verbatim:
  adv_grpo/stat_tracking.py                      -> G1, G2, G3 (+ a seeded [N,T] case)
  adv_grpo/diffusers_patch/sd3_sde_with_logprob.py (behind a ~40-line diffusers stub)
                                                 -> G5 (+ a replay / multi-timestep case)
  adv_grpo/pick_score_training.py CLIPCriterion.calc_loss -> G6
  adv_grpo/ema.py                                -> EMA trajectory
The FlowMatchEuler scheduler itself is third-party (diffusers, absent): G4 is the restated
schedule and is pinned only against the closed form sigma = 3 s / (1 + 2 s).
"""
import importlib.util
import json
import math
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(OUT, "..", ".."))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def install_diffusers_stub():
    """Minimal stand-ins for the two diffusers symbols sde.py imports (sde.py:8-9)."""
    from oracle.scheduler import FlowMatchEulerOracle

    def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
        return torch.randn(shape, generator=generator, device=device, dtype=dtype)

    d = types.ModuleType("diffusers")
    du = types.ModuleType("diffusers.utils")
    dtu = types.ModuleType("diffusers.utils.torch_utils")
    dtu.randn_tensor = randn_tensor
    ds = types.ModuleType("diffusers.schedulers")
    dsf = types.ModuleType("diffusers.schedulers.scheduling_flow_match_euler_discrete")
    dsf.FlowMatchEulerDiscreteScheduler = FlowMatchEulerOracle
    for n, m in [("diffusers", d), ("diffusers.utils", du), ("diffusers.utils.torch_utils", dtu),
                 ("diffusers.schedulers", ds),
                 ("diffusers.schedulers.scheduling_flow_match_euler_discrete", dsf)]:
        sys.modules[n] = m


def main():
    gold = {}
    # ---------------- stat_tracking (verbatim reference) ----------------
    st = _load("ref_stat_tracking", f"{REF}/adv_grpo/stat_tracking.py")
    tr = st.PerPromptStatTracker(global_std=False)
    gold["G1"] = tr.update(['a', 'b', 'a', 'c', 'b', 'a'], [1, 2, 3, 4, 5, 6]).tolist()
    gold["G1_stats"] = list(tr.get_stats())
    tr = st.PerPromptStatTracker(global_std=True)
    gold["G2"] = tr.update(['a', 'b', 'a', 'c', 'b', 'a'], [1, 2, 3, 4, 5, 6]).tolist()
    tr = st.PerPromptStatTracker(global_std=True)
    g3 = tr.update(['p', 'p', 'q', 'q'], [[1, 1], [2, 2], [3, 3], [4, 4]])
    gold["G3"] = g3.tolist()
    gold["G3_dtype"] = str(g3.dtype)
    rng = np.random.RandomState(0)
    n_prompts, G, T = 6, 8, 2
    prompts = [f"prompt {i}" for i in range(n_prompts) for _ in range(G)]
    perm = rng.permutation(len(prompts))
    prompts = [prompts[i] for i in perm]
    rewards = rng.rand(len(prompts)).astype(np.float32)
    rewards[[i for i, p in enumerate(prompts) if p == "prompt 3"]] = 0.25     # a zero-std group
    rew2 = np.repeat(rewards[:, None], T, axis=1)
    gold["G7_prompts"] = prompts
    gold["G7_rewards"] = rewards.tolist()
    for gs in (True, False):
        tr = st.PerPromptStatTracker(global_std=gs)
        gold[f"G7_adv_global{int(gs)}"] = tr.update(prompts, rew2).tolist()

    # ---------------- sde_step_with_logprob_new (verbatim reference) ----------------
    install_diffusers_stub()
    sde = _load("ref_sde", f"{REF}/adv_grpo/diffusers_patch/sd3_sde_with_logprob.py")
    from oracle.scheduler import FlowMatchEulerOracle
    sch = FlowMatchEulerOracle()
    sch.set_timesteps(10)
    gold["G4_sigmas"] = sch.sigmas.tolist()
    gold["G4_timesteps"] = sch.timesteps.tolist()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 16, 8, 8, generator=g)
    v = torch.randn(2, 16, 8, 8, generator=g)
    gen = torch.Generator().manual_seed(1)
    prev, lp, mean, std = sde.sde_step_with_logprob_new(sch, v, sch.timesteps[0:1], x, noise_level=0.8,
                                                        generator=gen)
    gold["G5_log_prob"] = lp.tolist()
    gold["G5_std"] = std.flatten().tolist()
    prev2, lp2, _, _ = sde.sde_step_with_logprob_new(sch, v, sch.timesteps[0:1], x, noise_level=0.8,
                                                     prev_sample=prev)
    gold["G5_replay_log_prob"] = lp2.tolist()
    torch.save({"x": x, "v": v, "prev": prev, "mean": mean}, os.path.join(OUT, "g5_tensors.pt"))
    # per-sample timesteps (the replay form, train_sd3_fast_pickscore.py:258-265), bf16-valued inputs
    g = torch.Generator().manual_seed(2)
    xb = torch.randn(4, 16, 16, 16, generator=g).bfloat16().float()
    vb = torch.randn(4, 16, 16, 16, generator=g).bfloat16().float()
    pb = torch.randn(4, 16, 16, 16, generator=g).bfloat16().float()
    ts = sch.timesteps[[1, 1, 3, 8]]
    _, lp3, mean3, std3 = sde.sde_step_with_logprob_new(sch, vb, ts, xb, noise_level=0.8, prev_sample=pb)
    gold["G8_log_prob"] = lp3.tolist()
    gold["G8_std"] = std3.flatten().tolist()
    gold["G8_step_index"] = [1, 1, 3, 8]
    torch.save({"x": xb, "v": vb, "prev": pb, "mean": mean3}, os.path.join(OUT, "g8_tensors.pt"))
    # last step: sigma_prev = 0 -> std = 0, log_prob = 0 (quirk Q1)
    _, lp4, _, std4 = sde.sde_step_with_logprob_new(sch, v, sch.timesteps[9:10], x, noise_level=0.8,
                                                    generator=torch.Generator().manual_seed(3))
    gold["G9_last_step_log_prob"] = lp4.tolist()
    gold["G9_last_step_std"] = std4.flatten().tolist()

    # ---------------- CLIPCriterion.calc_loss (verbatim reference) ----------------
    psrc = open(f"{REF}/adv_grpo/pick_score_training.py").read()
    # the module imports PIL/transformers/tqdm at the top; only the criterion classes are needed
    start = psrc.index("@dataclass\nclass CLIPCriterionConfig")
    end = psrc.index("# ====== 数据准备 ======")
    ns = {}
    exec("import torch\nfrom dataclasses import dataclass\nfrom torch.nn.modules.loss import _Loss\n"
         + psrc[start:end], ns)
    crit = ns["CLIPCriterion"](ns["CLIPCriterionConfig"]())
    torch.manual_seed(1)
    t, i0, i1 = (torch.nn.functional.normalize(torch.randn(3, 32), dim=-1) for _ in range(3))
    loss = crit.calc_loss(t, i0, i1, torch.tensor(100.0), torch.tensor(1.0), torch.tensor(0.0),
                          torch.tensor(1.0))
    gold["G6_loss"] = loss.item()
    torch.save({"t": t, "i0": i0, "i1": i1}, os.path.join(OUT, "g6_tensors.pt"))

    # ---------------- EMA (verbatim reference) ----------------
    ema_mod = _load("ref_ema", f"{REF}/adv_grpo/ema.py")
    torch.manual_seed(5)
    params = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    ema = ema_mod.EMAModuleWrapper(params, decay=0.9, update_step_interval=8, device="cpu")
    traj = []
    for step in range(40):
        with torch.no_grad():
            for p in params:
                p.add_(0.01 * (step + 1))
        ema.step(params, step)
        traj.append([e.sum().item() for e in ema.ema_parameters])
    gold["G10_ema_sums"] = traj
    gold["G10_init"] = [p.detach().clone().sub(sum(0.01 * (s + 1) for s in range(40))).tolist() for p in params]

    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote golden.json with keys", sorted(gold))


if __name__ == "__main__":
    main()

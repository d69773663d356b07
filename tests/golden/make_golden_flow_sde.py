"""Golden vectors G11 for the Flow-SDE step `sde_step_with_logprob` (adv_grpo/diffusers_patch/sd3_sde_with_logprob.py:13-73),
produced by executing the reference's own file behind the diffusers stub of make_golden.py.  Appends to golden.json and
writes g11_tensors.pt; the other goldens are left untouched.  Run in the build container (needs /root/reference)."""
import json
import os
import sys

import torch

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
sys.path.insert(0, os.path.join(OUT, "..", ".."))
import make_golden as mg  # noqa: E402


def main():
    mg.install_diffusers_stub()
    sde = mg._load("ref_sde", f"{mg.REF}/adv_grpo/diffusers_patch/sd3_sde_with_logprob.py")
    from oracle.scheduler import FlowMatchEulerOracle
    sch = FlowMatchEulerOracle()
    sch.set_timesteps(10)
    gold = json.load(open(os.path.join(OUT, "golden.json")))
    # replay form, per-sample timesteps incl. step 0 (sigma == 1 -> the sigma_max substitution of sde.py:47), bf16-valued inputs
    g = torch.Generator().manual_seed(21)
    xb = torch.randn(4, 16, 16, 16, generator=g).bfloat16().float()
    vb = torch.randn(4, 16, 16, 16, generator=g).bfloat16().float()
    pb = (xb + 0.3 * torch.randn(4, 16, 16, 16, generator=g)).bfloat16().float()
    idx = [0, 1, 3, 8]
    _, lp, mean, std = sde.sde_step_with_logprob(sch, vb, sch.timesteps[idx], xb, noise_level=0.7, prev_sample=pb)
    gold["G11_step_index"] = idx
    gold["G11_log_prob"] = lp.tolist()
    gold["G11_std"] = std.flatten().tolist()
    # rollout form (one broadcast timestep, noise from a seeded generator)
    gen = torch.Generator().manual_seed(22)
    prev_r, lp_r, mean_r, std_r = sde.sde_step_with_logprob(sch, vb[:2], sch.timesteps[2:3], xb[:2], noise_level=0.7,
                                                             generator=gen)
    noise = torch.randn(vb[:2].shape, generator=torch.Generator().manual_seed(22))
    gold["G11_rollout_log_prob"] = lp_r.tolist()
    gold["G11_rollout_std"] = std_r.flatten().tolist()
    torch.save({"x": xb, "v": vb, "prev": pb, "mean": mean, "noise": noise, "prev_rollout": prev_r, "mean_rollout": mean_r},
               os.path.join(OUT, "g11_tensors.pt"))
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("G11 written:", gold["G11_log_prob"], gold["G11_std"])


if __name__ == "__main__":
    main()

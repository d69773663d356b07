"""Calibration for tests/test_fullsize_gpu.py (CPU only, needs no GPU and no reference checkout).

How far does the reference's OWN dtype regime (bf16 autocast: every matmul / norm output rounded to bf16) sit
from an fp32 evaluation of the same rollout at the true SD3.5-medium size with the seeded random weights the
test uses?  Runs the oracle twice (dtype=float32 and dtype=bfloat16) on the inputs of
test_config1_rollout_sd35_medium_true_size_matches_oracle and prints per-latent max / mean deviations.  The
numbers are recorded in tests/golden/bf16_regime_calibration.json; the GPU test bounds the CUDA path's deviation
from the fp32 oracle by a small multiple of them."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from adv_grpo_b200 import weights                      # noqa: E402
from oracle import pipeline as pipe_o                  # noqa: E402
from oracle.mmdit import MMDiTOracle                   # noqa: E402


def main():
    cfg = weights.SD35_MEDIUM
    params = weights.init_mmdit(cfg, seed=0, device="cpu", dtype=torch.bfloat16)
    lora = weights.init_lora(cfg, rank=32, seed=1, perturb_b=0.01)
    lora = {k: (a.bfloat16().float(), b.bfloat16().float()) for k, (a, b) in lora.items()}
    G, steps, T_train = 2, 4, 2
    g = torch.Generator().manual_seed(7)
    pe = torch.randn(1, 205, 4096, generator=g).bfloat16()
    pp = torch.randn(1, 2048, generator=g).bfloat16()
    ne = torch.randn(1, 205, 4096, generator=g).bfloat16()
    npool = torch.randn(1, 2048, generator=g).bfloat16()
    lat = torch.randn(G, 16, 32, 32, generator=g).bfloat16()
    noises = [torch.randn(G, 16, 32, 32, generator=g) for _ in range(steps)]
    out = {}
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        t0 = time.time()
        o = MMDiTOracle(params, dict(cfg, dual_layers=set(cfg["dual_layers"])), lora=lora, lora_scale=2.0, dtype=dt)
        with torch.no_grad():
            _, lats, lps, _, _ = pipe_o.rollout(o, None, pe.repeat(G, 1, 1), pp.repeat(G, 1), ne.repeat(G, 1, 1),
                                                npool.repeat(G, 1), lat, steps, 4.5, 0.8, T_train, 0, noises, decode=False)
        out[name] = (lats, lps)
        print(name, f"{time.time() - t0:.0f}s", flush=True)
    res = {"latents": [], "log_probs": []}
    for a, b in zip(out["bf16"][0], out["fp32"][0]):
        d = (a.float() - b.float()).abs()
        res["latents"].append(dict(max=d.max().item(), mean=d.mean().item(), range=b.float().abs().max().item()))
    for a, b in zip(out["bf16"][1], out["fp32"][1]):
        res["log_probs"].append(dict(max_abs=(a - b).abs().max().item(), ref=b.abs().max().item()))
    print(json.dumps(res, indent=1))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bf16_regime_calibration.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()

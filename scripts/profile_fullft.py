"""Kernel table of one full-fine-tuning replay micro-step (forward + loss + backward at the config-2 shape, B = 16 CFG batch):
torch.profiler CUDA-activity totals per kernel.  Usage: python scripts/profile_fullft.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adv_grpo_b200 import weights  # noqa: E402
from adv_grpo_b200.mmdit import SD3Transformer2DModel  # noqa: E402

DEV = "cuda"
cfg = weights.SD35_MEDIUM
model = SD3Transformer2DModel(cfg, weights.init_mmdit(cfg, seed=0, device=DEV, dtype=torch.bfloat16), device=DEV).enable_full_finetune()
g = torch.Generator(device=DEV).manual_seed(0)
B = 16
x = torch.randn(B, 16, 64, 64, device=DEV, generator=g).bfloat16()
t = torch.full((B,), 500.0, device=DEV)
ctx = torch.randn(B, 205, 4096, device=DEV, generator=g).bfloat16()
pooled = torch.randn(B, 2048, device=DEV, generator=g).bfloat16()


def step():
    out = model(x, t, ctx, pooled)[0]
    out.float().square().mean().backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
torch.cuda.synchronize()
print(f"full fine-tuning micro-step (fwd + bwd, B = {B}): {e0.elapsed_time(e1) / 3:.1f} ms")
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
tot = sum(r.device_time_total for r in rows)
print(f"{'kernel':100s} {'calls':>6s} {'ms':>9s} {'share':>7s}")
for r in rows[:28]:
    print(f"{r.key[:100]:100s} {r.count:6d} {r.device_time_total / 1e3:9.2f} {100 * r.device_time_total / tot:6.1f}%")
native = sum(r.device_time_total for r in rows if "advgrpo" in r.key)
print(f"total {tot / 1e3:.1f} ms of device time, {100 * native / tot:.1f}% in advgrpo kernels")

"""Kernel table of one PickScore discriminator step at the true CLIP-ViT-H/14 size (BASELINE config 5: 16 real + 16 fake
images, tune_layer = -1): torch.profiler CUDA-activity totals per kernel, so the share of native vs library kernels and
the cost of the short-sequence attention / weight-gradient / Adam kernels are visible.  Usage: python scripts/profile_dstep.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adv_grpo_b200 import weights  # noqa: E402
from adv_grpo_b200.optim import TorchOrderAdam  # noqa: E402
from adv_grpo_b200.pick_score_training import CLIPCriterion, CLIPCriterionConfig  # noqa: E402
from adv_grpo_b200.pickscore_scorer import PickScoreScorer, images_to_pixel_values  # noqa: E402

DEV = "cuda"
N = int(os.environ.get("N", 16))
scorer = PickScoreScorer(device=DEV, cfg=weights.CLIP_H)
model = scorer.model
for p in model.parameters():
    p.requires_grad = False
for p in model.vision_model.encoder.layers[-1:].parameters():
    p.requires_grad = True
opt = TorchOrderAdam(model.parameters(), lr=5e-6, betas=(0.5, 0.999))
g = torch.Generator().manual_seed(0)
real = (torch.rand(N, 3, 512, 512, generator=g) * 255).to(torch.uint8)
fake = (torch.rand(N, 3, 512, 512, generator=g) * 255).to(torch.uint8)
ids = scorer.processor.tokenizer(["a photo of a cat"] * N, padding="max_length", max_length=77)["input_ids"].to(DEV)
crit = CLIPCriterion(CLIPCriterionConfig())


def step():
    batch = {"input_ids": ids, "pixels_0": images_to_pixel_values(real, DEV), "pixels_1": images_to_pixel_values(fake, DEV),
             "label_0": torch.tensor(1.0, device=DEV), "label_1": torch.tensor(0.0, device=DEV),
             "num_examples_per_prompt": torch.tensor(1.0, device=DEV)}
    opt.zero_grad()
    loss = crit(model, batch)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
print(f"D step ({N} real + {N} fake images): {e0.elapsed_time(e1) / 5:.2f} ms")
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
tot = sum(r.device_time_total for r in rows)
print(f"{'kernel':90s} {'calls':>6s} {'us':>10s} {'share':>7s}")
for r in rows[:32]:
    print(f"{r.key[:90]:90s} {r.count:6d} {r.device_time_total:10.1f} {100 * r.device_time_total / tot:6.1f}%")
native = sum(r.device_time_total for r in rows if "advgrpo" in r.key)
print(f"total {tot / 1e3:.2f} ms of device time, {100 * native / tot:.1f}% in advgrpo kernels")

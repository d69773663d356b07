"""Per-kernel CUDA time of one eager MMDiT-X forward at B=16, 512x512 (torch.profiler, kernel names)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from adv_grpo_b200 import weights
from adv_grpo_b200.mmdit import SD3Transformer2DModel
dev = "cuda:0"
cfg = weights.SD35_MEDIUM
m = SD3Transformer2DModel(cfg, weights.init_mmdit(cfg, device=dev), lora=None, device=dev)
x = torch.randn(16, 16, 64, 64, device=dev).bfloat16(); t = torch.full((16,), 500.0, device=dev)
e = torch.randn(16, 205, 4096, device=dev).bfloat16(); p = torch.randn(16, 2048, device=dev).bfloat16()
with torch.no_grad():
    for _ in range(3):
        m(x, t, e, p)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            m(x, t, e, p)
        torch.cuda.synchronize()
rows = []
for ev in prof.key_averages():
    rows.append((ev.key[:90], ev.count / 3, ev.device_time_total / 3 / 1e3))
tot = sum(r[2] for r in rows)
print(f"total CUDA kernel time per forward: {tot:.2f} ms")
for k, n, ms in sorted(rows, key=lambda r: -r[2])[:25]:
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  x{n:6.1f}  {k}")

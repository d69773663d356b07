import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from adv_grpo_b200 import weights
from adv_grpo_b200.vae import AutoencoderKL
dev = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
vae = AutoencoderKL(weights.init_vae_decoder(weights.VAE_SD3, device=dev), weights.VAE_SD3, device=dev)
z = torch.randn(8, 16, 64, 64, device=dev)
for _ in range(3):
    vae.decode(z)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        vae.decode(z)
    torch.cuda.synchronize()
rows = [(ev.key[:100], ev.count / 3, ev.device_time_total / 3 / 1e3) for ev in prof.key_averages()]
tot = sum(r[2] for r in rows)
print(f"VAE decode: total CUDA kernel time {tot:.2f} ms")
for k, n, ms in sorted(rows, key=lambda r: -r[2])[:16]:
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  x{n:6.1f}  {k}")

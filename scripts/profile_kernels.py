"""One launch set of the four tensor-core kernels at the config-2 shapes, for `ncu --set full -k regex:...`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adv_grpo_b200 import ops
B, S, H, D = 16, 1229, 24, 64
qkv = torch.randn(B, S, 3, H, D, device="cuda").bfloat16()
dout = torch.randn(B, S, H, D, device="cuda").bfloat16()
M, N, K = 16384, 4608, 1536
a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
t = torch.randn(M, 128, device="cuda").bfloat16(); w2 = torch.randn(N, 128, device="cuda").bfloat16()
x = torch.randn(8, 256, 256, 256, device="cuda").contiguous(memory_format=torch.channels_last)
wc = ops.pack_conv_weight_tf32(torch.randn(256, 256, 3, 3, device="cuda") / 48.0)
for _ in range(3):
    out, lse = ops.attention_fwd(qkv)
    dq = ops.attention_bwd(qkv, out, dout, lse)
    ops.gemm(a, w, a2=t, w2=w2)
    ops.conv2d_nhwc_tf32(x, wc, None, 3)
torch.cuda.synchronize()

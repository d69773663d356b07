import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adv_grpo_b200 import ops
B, S, H, D = 16, 1229, 24, 64
qkv = torch.randn(B, S, 3, H, D, device="cuda").bfloat16()
dout = torch.randn(B, S, H, D, device="cuda").bfloat16()
v = int(os.environ.get("VARIANT", "1"))
for _ in range(2):
    out, lse = ops.attention_fwd(qkv, variant=v)
    dq = ops.attention_bwd(qkv, out, dout, lse)
M, N, K = 16384, 4608, 1536
a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
for _ in range(2):
    ops.gemm(a, w)
torch.cuda.synchronize()

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adv_grpo_b200 import ops
B, S, H, D = 16, 1229, 24, 64
qkv = torch.randn(B, S, 3, H, D, device="cuda").bfloat16()
for _ in range(4):
    ops.attention_fwd(qkv, want_lse=False, split=1024)
torch.cuda.synchronize()

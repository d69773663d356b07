import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adv_grpo_b200 import ops
B, S, H, D = 16, 1229, 24, 64
qkv = torch.randn(B, S, 3, H, D, device="cuda").bfloat16()
for v in (1, 1, 1, 2, 2, 2):
    ops.attention_fwd(qkv, variant=v, want_lse=False)
torch.cuda.synchronize()

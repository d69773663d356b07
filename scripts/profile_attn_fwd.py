"""ncu target: a few launches of the attention forward at the config-2 joint shape.  VARIANT picks the kernel
(0 = library default), SEQ the sequence length."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adv_grpo_b200 import ops
B, S, H, D = int(os.environ.get("BATCH", 16)), int(os.environ.get("SEQ", 1229)), 24, 64
variant = int(os.environ.get("VARIANT", 0))
qkv = torch.randn(B, S, 3, H, D, device="cuda").bfloat16()
for _ in range(4):
    ops.attention_fwd(qkv, want_lse=False, variant=variant)
torch.cuda.synchronize()

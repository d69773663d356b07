"""ncu target: the fused QKV projection + LoRA second product GEMM of bench.py's roofline entry (M=16384 N=4608 K=1536+128)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adv_grpo_b200 import ops
M, K, N, K2 = 16384, 1536, 4608, 128
A = torch.randn(M, K, device="cuda").bfloat16(); W = torch.randn(N, K, device="cuda").bfloat16()
A2 = torch.randn(M, K2, device="cuda").bfloat16(); W2 = torch.randn(N, K2, device="cuda").bfloat16()
bias = torch.randn(N, device="cuda").bfloat16()
for _ in range(4):
    ops.gemm(A, W, bias=bias, a2=A2, w2=W2)
torch.cuda.synchronize()

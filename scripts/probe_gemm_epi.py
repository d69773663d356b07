"""GEMM epilogue / cache-state probe: the block's projection shapes with every epilogue the MMDiT uses, warm
(same operands re-used, L2 resident) and cold (a > L2 buffer written between launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adv_grpo_b200 import ops

dev = "cuda:0"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, cold, iters=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        if cold:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters


def run(tag, M, N, K, epi=ops.EPI_NONE, res=False, preact=False, lora=0):
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    b = torch.randn(N, device=dev).bfloat16()
    kw = {}
    if res:
        kw.update(residual=torch.randn(M, N, device=dev).bfloat16(), gate=torch.randn(M // 1024, N, device=dev).bfloat16(),
                  rows_per_gate=1024)
    if preact:
        kw.update(preact_out=torch.empty(M, N, device=dev).bfloat16())
    if lora:
        kw.update(a2=torch.randn(M, lora, device=dev).bfloat16(), w2=torch.randn(N, lora, device=dev).bfloat16())
    out = torch.empty(M, N, device=dev).bfloat16()
    fn = lambda: ops.gemm(a, w, b, epilogue=epi, out=out, **kw)
    fl = 2.0 * M * N * (K + lora)
    tw, tc = timeit(fn, False), timeit(fn, True)
    print(f"{tag:34s} {M}x{N}x{K}: warm {tw*1e3:7.1f} us {fl/tw/1e9:7.1f} TF/s | cold {tc*1e3:7.1f} us {fl/tc/1e9:7.1f} TF/s", flush=True)


M = 16384
run("qkv none", M, 4608, 1536)
run("qkv +lora128", M, 4608, 1536, lora=128)
run("out gate+res", M, 1536, 1536, epi=ops.EPI_GATE_RESIDUAL, res=True)
run("out gate+res +lora64", M, 1536, 1536, epi=ops.EPI_GATE_RESIDUAL, res=True, lora=64)
run("ff1 none", M, 6144, 1536)
run("ff1 gelu_tanh", M, 6144, 1536, epi=ops.EPI_GELU_TANH)
run("ff1 gelu_tanh+preact", M, 6144, 1536, epi=ops.EPI_GELU_TANH, preact=True)
run("ff2 none", M, 1536, 6144)
run("ff2 gate+res", M, 1536, 6144, epi=ops.EPI_GATE_RESIDUAL, res=True)
M = 3280
run("txt qkv none", M, 4608, 1536)
run("txt ff1 gelu", M, 6144, 1536, epi=ops.EPI_GELU_TANH)
run("txt ff2 gate+res", M, 1536, 6144, epi=ops.EPI_GATE_RESIDUAL, res=True)

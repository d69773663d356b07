"""CUDA-event timing of the fused QKV-norm GEMM against the two-kernel path at the config-2 shape."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adv_grpo_b200 import ops
DEV = "cuda"
g = torch.Generator(device=DEV).manual_seed(1)
B, S_img, S_txt, H, D, K = 16, 1024, 205, 24, 64, 1536
N = 3 * H * D
x = torch.randn(B, S_img, K, device=DEV, generator=g).bfloat16()
c = torch.randn(B, S_txt, K, device=DEV, generator=g).bfloat16()
w = [(torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16() for _ in range(2)]
bias = [torch.randn(N, device=DEV, generator=g).bfloat16() for _ in range(2)]
nq = [(1 + 0.2 * torch.randn(D, device=DEV, generator=g)).bfloat16() for _ in range(2)]
nk = [(1 + 0.2 * torch.randn(D, device=DEV, generator=g)).bfloat16() for _ in range(2)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)

def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

cases = {
    "dual NONE (img+txt)": lambda: ops.gemm_dual((x, c), w, bias=bias),
    "dual NONE + qk_norm_concat": lambda: ops.qk_norm_concat(*ops.gemm_dual((x, c), w, bias=bias), nq[0], nk[0], nq[1], nk[1], H, D),
    "fused qkv_norm (img+txt)": lambda: ops.gemm_qkv_norm(x, c, w, bias, nq, nk, H, D),
    "fused qkv_norm no norm weights": lambda: ops.gemm_qkv_norm(x, c, w, bias, (None, None), (None, None), H, D),
    "single NONE (img)": lambda: ops.gemm(x, w[0], bias=bias[0]),
    "fused qkv_norm (img only)": lambda: ops.gemm_qkv_norm(x, None, w[:1], bias[:1], nq[:1], nk[:1], H, D),
    "single GELU (img)": lambda: ops.gemm(x, w[0], bias=bias[0], epilogue=ops.EPI_GELU_TANH),
    "single NONE (txt)": lambda: ops.gemm(c, w[1], bias=bias[1]),
}
la = [(torch.randn(128, K, device=DEV, generator=g) / 32).bfloat16() for _ in range(2)]
cases["LoRA down-projection dual (N=128)"] = lambda: ops.gemm_dual((x, c), la)
with torch.no_grad():
    for name, fn in cases.items():
        print(f"{name:40s} {timeit(fn):8.1f} us", flush=True)
    sh, sc = bias[0][:K].repeat(B, 1).contiguous(), bias[0][K:2 * K].repeat(B, 1).contiguous()
    print(f"ln_modulate: img {timeit(lambda: ops.ln_modulate(x, sh, sc)):6.1f} us  "
          f"img dual {timeit(lambda: ops.ln_modulate(x, sh, sc, sc, sh)):6.1f} us  "
          f"txt {timeit(lambda: ops.ln_modulate(c, sh, sc)):6.1f} us", flush=True)
    y = torch.empty_like(x)
    print(f"torch copy img (same bytes): {timeit(lambda: y.copy_(x)):6.1f} us")

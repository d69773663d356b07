"""Diagnostics for the tcgen05 kernels on a real B200: each probe runs in its own process so a trapped
kernel (sticky CUDA error) cannot poison the others.  Usage: python scripts/gpu_probe.py [probe ...]"""
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def probe_gemm_small():
    import torch
    from adv_grpo_b200 import ops
    for (M, N, K) in [(128, 128, 64), (256, 256, 64), (256, 256, 128), (512, 512, 256), (1229, 1536, 1536), (300, 256, 64)]:
        g = torch.Generator(device="cuda").manual_seed(1)
        a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        w = torch.randn(N, K, device="cuda", generator=g).bfloat16() / 8
        c = ops.gemm(a, w)
        torch.cuda.synchronize()
        ref = a.float() @ w.float().T
        err = (c.float() - ref).abs().max().item() / ref.abs().max().item()
        print(f"gemm {M}x{N}x{K}: rel err {err:.3e}", flush=True)
        if err > 1e-2:
            print(" got", c[0, :8].float().tolist())
            print(" ref", ref[0, :8].tolist())
            print(" got row1", c[1, :4].float().tolist(), "ref", ref[1, :4].tolist())
            # which K-slices / rows are right?  identity-like probes
            a2 = torch.zeros_like(a); a2[:, 0] = 1
            c2 = ops.gemm(a2, w).float()
            print(" col0 probe err", (c2 - w.float()[:, 0][None]).abs().max().item())
            a3 = torch.zeros_like(a); a3[:, min(17, K - 1)] = 1
            c3 = ops.gemm(a3, w).float()
            print(" col17 probe err", (c3 - w.float()[:, min(17, K - 1)][None]).abs().max().item())


def _attn(variant, D=64, S=256, causal=False):
    import torch
    from adv_grpo_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    B, H = 1, 2
    qkv = torch.randn(B, S, 3, H, D, device="cuda", generator=g).bfloat16()
    out, lse = ops.attention_fwd(qkv, variant=variant, causal=causal)
    torch.cuda.synchronize()
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3).float() for i in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal).permute(0, 2, 1, 3)
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    s = (q @ k.transpose(-1, -2)) / math.sqrt(D)
    if causal:
        s = s.masked_fill(torch.ones(S, S, device="cuda", dtype=torch.bool).triu(1), float("-inf"))
    lse_ref = torch.logsumexp(s, -1)
    print(f"attn variant={variant} D={D} S={S} causal={causal}: rel err {err:.3e}  lse err {(lse - lse_ref).abs().max().item():.3e}", flush=True)
    if err > 2e-2:
        print(" got", out[0, 0, 0, :6].float().tolist())
        print(" ref", ref[0, 0, 0, :6].tolist())
        print(" got r77", out[0, 77, 1, :4].float().tolist(), "ref", ref[0, 77, 1, :4].tolist())


def probe_attn_v1():
    _attn(1, S=128)
    _attn(1, S=256)
    _attn(1, S=1229)


def probe_attn_v2():
    _attn(2, S=256)
    _attn(2, S=1229)


def probe_attn_variants():
    for v in (7, 8, 9, 10):
        _attn(v, S=1229)
        _attn(v, S=300, causal=True)


def probe_attn_d128():
    _attn(1, D=128, S=257)


def probe_attn_causal():
    _attn(1, S=77, causal=True)
    _attn(1, S=300, causal=True)


def probe_attn_bwd():
    import torch
    from adv_grpo_b200 import ops
    for (B, S, H) in [(1, 128, 1), (1, 256, 1), (1, 1229, 2)]:
        g = torch.Generator(device="cuda").manual_seed(1)
        qkv = torch.randn(B, S, 3, H, 64, device="cuda", generator=g).bfloat16()
        dout = torch.randn(B, S, H, 64, device="cuda", generator=g).bfloat16()
        out, lse = ops.attention_fwd(qkv)
        dqkv = ops.attention_bwd(qkv, out, dout, lse)
        torch.cuda.synchronize()
        ref_in = qkv.float().requires_grad_(True)
        q, k, v = (ref_in[:, :, i].permute(0, 2, 1, 3) for i in range(3))
        o = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3)
        (o * dout.float()).sum().backward()
        for i, n in enumerate("qkv"):
            got, ref = dqkv[:, :, i].float(), ref_in.grad[:, :, i]
            err = (got - ref).abs().max().item() / ref.abs().max().item()
            print(f"attn bwd S={S} d{n}: rel err {err:.3e}", flush=True)
            if err > 3e-2:
                print("  got", got[0, 0, 0, :4].tolist(), "ref", ref[0, 0, 0, :4].tolist())
                print("  got r70", got[0, 70 % S, 0, :4].tolist(), "ref", ref[0, 70 % S, 0, :4].tolist())


def probe_perf_bwd():
    import torch
    from adv_grpo_b200 import ops
    B, S, H, D = 16, 1229, 24, 64
    qkv = torch.randn(B, S, 3, H, D, device="cuda").bfloat16()
    dout = torch.randn(B, S, H, D, device="cuda").bfloat16()
    out, lse = ops.attention_fwd(qkv)
    flops = 10 * B * H * S * S * D
    for _ in range(3):
        ops.attention_bwd(qkv, out, dout, lse)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.attention_bwd(qkv, out, dout, lse)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"attn bwd: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3).detach().requires_grad_(True) for i in range(3))
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v)
    go = dout.permute(0, 2, 1, 3)
    for _ in range(3):
        torch.autograd.grad(o, (q, k, v), go, retain_graph=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        torch.autograd.grad(o, (q, k, v), go, retain_graph=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"torch SDPA bwd (library baseline): {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)


def probe_perf():
    import torch
    from adv_grpo_b200 import ops
    B, S, H, D = 16, 1229, 24, 64
    qkv = torch.randn(B, S, 3, H, D, device="cuda").bfloat16()
    flops = 4 * B * H * S * S * D
    for variant in (3, 0, 14, 15):
        for _ in range(3):
            ops.attention_fwd(qkv, variant=variant, want_lse=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.attention_fwd(qkv, variant=variant, want_lse=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"attn fwd variant {variant}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    for _ in range(3):
        torch.nn.functional.scaled_dot_product_attention(q, k, v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch.nn.functional.scaled_dot_product_attention(q, k, v)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"torch SDPA (library baseline): {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
    for (M, N, K) in [(16384, 4608, 1536), (16384, 1536, 1536), (16384, 6144, 1536), (16384, 1536, 6144), (3280, 1536, 1536), (3280, 4608, 1536), (3280, 1536, 6144)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = torch.randn(N, K, device="cuda").bfloat16()
        for name, fn in (("ours-1cta", lambda: ops.gemm(a, w)), ("ours-2cta-bn256", lambda: ops.gemm(a, w)),
                         ("ours-auto", lambda: ops.gemm(a, w)), ("cublas", lambda: torch.nn.functional.linear(a, w))):
            ops.set_gemm_variant({"ours-1cta": 1, "ours-2cta-bn256": 3}.get(name, 0))
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"gemm {M}x{N}x{K} {name}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)


def probe_attn_split():
    """Default D=64 path with the image / text output split (TMA-store epilogue) against the unsplit output."""
    import torch
    from adv_grpo_b200 import ops
    for (B, S, H, split) in ((2, 1229, 3, 1024), (1, 461, 2, 256), (2, 1229, 2, 1000), (1, 4301, 2, 4096)):
        qkv = torch.randn(B, S, 3, H, 64, device="cuda").bfloat16()
        ref, lse_ref = ops.attention_fwd(qkv, variant=3)
        (o1, o2), lse = ops.attention_fwd(qkv, split=split)
        full, _ = ops.attention_fwd(qkv)
        torch.cuda.synchronize()
        got = torch.cat([o1, o2], 1)
        e1 = (got.float() - ref.float()).abs().max().item()
        e2 = (full.float() - ref.float()).abs().max().item()
        print(f"split B={B} S={S} H={H} split={split}: |split - v3| {e1:.3e}  |default - v3| {e2:.3e}  lse {(lse - lse_ref).abs().max().item():.3e}", flush=True)


def probe_attn_quad():
    """Correctness of the 8-softmax-warp (quad TMEM layout) forward, variants 16-18."""
    for v in (16, 17, 18):
        for S in (128, 77, 461, 1229, 1370):
            _attn(v, S=S)
        _attn(v, S=300, causal=True)
        _attn(v, S=1024, causal=True)


def probe_attn_pair():
    """Correctness of the pair kernel (one CTA per SM, two query tiles in antiphase), variants 21-25."""
    import torch
    from adv_grpo_b200 import ops
    for v in (21, 22, 23, 24, 25):
        for S in (128, 77, 256, 461, 1229, 1370):
            _attn(v, S=S)
    # many items per CTA + split output (image / text row ranges), vs the quad kernel bit for bit?  (no: different
    # summation order of the row sums is identical, the exp token does not change arithmetic) -> compare numerically
    g = torch.Generator(device="cuda").manual_seed(3)
    B, S, H = 16, 1229, 24
    qkv = torch.randn(B, S, 3, H, 64, device="cuda", generator=g).bfloat16()
    ref, lse_ref = ops.attention_fwd(qkv, variant=17)
    for v in (21, 22):
        out, lse = ops.attention_fwd(qkv, variant=v)
        torch.cuda.synchronize()
        print(f"pair variant {v} vs quad at B=16 S=1229 H=24: max |d out| {(out.float() - ref.float()).abs().max().item():.3e} "
              f"max |d lse| {(lse - lse_ref).abs().max().item():.3e}  equal {torch.equal(out, ref)}", flush=True)


def _time_attn(B, S, H, D, variant, iters=10):
    import torch
    from adv_grpo_b200 import ops
    qkv = torch.randn(B, S, 3, H, D, device="cuda").bfloat16()
    for _ in range(3):
        ops.attention_fwd(qkv, variant=variant, want_lse=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.attention_fwd(qkv, variant=variant, want_lse=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, 4 * B * H * S * S * D / ms / 1e9


def probe_perf_attn():
    import torch
    for (B, S, H) in ((16, 1229, 24), (16, 1024, 24), (4, 4301, 24), (32, 1370, 12)):
        for variant in (0, 17, 21, 22, 23, 24, 25):
            ms, tf = _time_attn(B, S, H, 64, variant)
            print(f"attn fwd B={B} S={S} H={H} variant {variant}: {ms:.3f} ms  {tf:.1f} TFLOP/s", flush=True)
        qkv = torch.randn(B, S, 3, H, 64, device="cuda").bfloat16()
        q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
        for _ in range(3):
            torch.nn.functional.scaled_dot_product_attention(q, k, v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            torch.nn.functional.scaled_dot_product_attention(q, k, v)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"torch SDPA B={B} S={S} H={H} (library baseline): {ms:.3f} ms  {4 * B * H * S * S * 64 / ms / 1e9:.1f} TFLOP/s", flush=True)


PROBES = {k[6:]: v for k, v in globals().items() if k.startswith("probe_")}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        PROBES[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(PROBES)
    for n in names:
        print(f"===== probe {n} =====", flush=True)
        r = subprocess.run([sys.executable, __file__, "--one", n], capture_output=True, text=True, timeout=300)
        print(r.stdout[-4000:])
        if r.returncode != 0:
            print(f"  [probe {n} exit code {r.returncode}]\n{r.stderr[-1500:]}")

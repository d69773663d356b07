"""Debug probe: fused QKV-norm GEMM epilogue cases, each in its own process (sticky CUDA errors)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def one(B, S_img, S_txt, H, pre):
    import math, torch
    from adv_grpo_b200 import ops
    DEV = "cuda"
    g = torch.Generator(device=DEV).manual_seed(1)
    K, D = 256, 64
    N = 3 * H * D
    x = torch.randn(B, S_img, K, device=DEV, generator=g).bfloat16()
    c = torch.randn(B, S_txt, K, device=DEV, generator=g).bfloat16() if S_txt else None
    w = [(torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16() for _ in range(2)]
    bias = [torch.randn(N, device=DEV, generator=g).bfloat16() for _ in range(2)]
    nq = [(1 + 0.2 * torch.randn(D, device=DEV, generator=g)).bfloat16() for _ in range(2)]
    nk = [(1 + 0.2 * torch.randn(D, device=DEV, generator=g)).bfloat16() for _ in range(2)]
    n = 2 if S_txt else 1
    pres = tuple(torch.zeros(B * s, N, device=DEV, dtype=torch.bfloat16) for s in (S_img, S_txt)[:n]) if pre else (None, None)
    joint = ops.gemm_qkv_norm(x, c, w[:n], bias[:n], nq[:n], nk[:n], H, D, prenorm_out=pres)
    torch.cuda.synchronize()
    if S_txt:
        qx, qc = ops.gemm_dual((x, c), w, bias=bias)
    else:
        qx, qc = ops.gemm(x, w[0], bias=bias[0]), None
    ref = ops.qk_norm_concat(qx, qc, nq[0], nk[0], nq[1] if S_txt else None, nk[1] if S_txt else None, H, D)
    torch.cuda.synchronize()
    bad = (joint != ref).reshape(B, S_img + S_txt, -1).any(-1)
    print(f"B={B} S=({S_img},{S_txt}) H={H} pre={pre}: equal={torch.equal(joint, ref)} bad rows={int(bad.sum())}"
          + (f" first bad {bad.nonzero()[:4].tolist()}" if bad.any() else ""), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(*[int(a) for a in sys.argv[1:]])
    else:
        for case in [(1, 256, 0, 4, 0), (2, 256, 0, 4, 0), (2, 256, 0, 4, 1), (2, 300, 0, 4, 0), (2, 1024, 205, 24, 0), (3, 64, 13, 4, 0)]:
            r = subprocess.run([sys.executable, __file__] + [str(a) for a in case], capture_output=True, text=True, timeout=300)
            print(r.stdout.strip()[-400:] or ("FAIL " + str(case) + " " + r.stderr.strip()[-300:]), flush=True)

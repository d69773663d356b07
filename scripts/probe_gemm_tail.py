"""Round-2 probe for DESIGN.md section 7 item 2: the N = 1536 dual GEMMs of an MMDiT block (image 16384 rows + text 3280
rows) with and without the tail split of `ops.gemm_dual` (ADVGRPO_GEMM_TAIL_SPLIT), for the three K of the path
(out-projection / dX of QKV / FF2), plain and GATE_RESIDUAL epilogues: time (CUDA events) and bitwise equality."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adv_grpo_b200 import ops  # noqa: E402

dev = "cuda"


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    B, S_img, S_txt, N = 16, 1024, 205, 1536
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.5).bfloat16()
    for K in (1536, 4608, 6144):
        x0, x1 = rnd(B * S_img, K), rnd(B * S_txt, K)
        w0, w1 = rnd(N, K) * 0.05, rnd(N, K) * 0.05
        b0, b1 = rnd(N), rnd(N)
        r0, r1 = rnd(B * S_img, N), rnd(B * S_txt, N)
        gate = rnd(B, 2 * N)
        for name, kw in (("plain", dict(bias=(b0, b1))),
                         ("gate_residual", dict(bias=(b0, b1), epilogue=ops.EPI_GATE_RESIDUAL, residual=(r0, r1),
                                                gate=(gate[:, :N], gate[:, N:]), rows_per_gate=(S_img, S_txt)))):
            run = lambda: ops.gemm_dual((x0, x1), (w0, w1), **kw)
            ops.GEMM_TAIL_SPLIT = False
            ref = [t.clone() for t in run()]
            t_off = timeit(run)
            ops.GEMM_TAIL_SPLIT = True
            got = run()
            same = all(torch.equal(a, b) for a, b in zip(ref, got))
            t_on = timeit(run)
            ops.GEMM_TAIL_SPLIT = False
            fl = 2.0 * (B * (S_img + S_txt)) * N * K
            print(f"K={K:5d} {name:14s} off {t_off * 1e3:7.1f} us ({fl / t_off / 1e9:6.0f} TFLOP/s)  "
                  f"split {t_on * 1e3:7.1f} us ({fl / t_on / 1e9:6.0f} TFLOP/s)  bitwise equal: {same}")


if __name__ == "__main__":
    main()

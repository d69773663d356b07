"""HBM-bound kernels of the path at BASELINE config-2 (512x512, G=8) and config-4 (1024x1024, G=16 in micro-batches of 8)
sizes, plus one streaming size that is not launch-latency-bound: CUDA-event time per launch, ALGORITHMIC bytes per launch
 and the achieved GB/s against the measured HBM peak (MEASURED_PEAKS.json).

  python scripts/profile_hbm_kernels.py                 # table (CUDA events, L2 flushed between launches)
  ncu --set full -k regex:'sde_|group_advantage|grpo_clip|ln_modulate|qk_norm' ... python scripts/profile_hbm_kernels.py --once
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adv_grpo_b200 import ops

ONCE = "--once" in sys.argv
dev = "cuda"
try:
    HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    HBM_SRC = "measured"
except Exception:
    HBM, HBM_SRC = 6551.0, "fallback"
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2


def timed(fn, iters=20):
    if ONCE:
        fn()
        torch.cuda.synchronize()
        return float("nan")
    for _ in range(3):
        fn()
    ms = 0.0
    for _ in range(iters):
        flush_buf.zero_()                                             # cold L2 for every launch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    return ms / iters


def row(name, size, nbytes, ms):
    gbs = nbytes / ms / 1e6 if ms == ms else float("nan")
    print(f"{name:34s} {size:34s} {nbytes / 1e6:9.2f} MB {ms * 1e3:9.1f} us {gbs:8.0f} GB/s  {100 * gbs / HBM:5.1f} % of {HBM:.0f} ({HBM_SRC})",
          flush=True)


def sde(tag, B, res):
    n = 16 * (res // 8) ** 2
    g = torch.Generator(device=dev).manual_seed(0)
    shape = (B, 16, res // 8, res // 8)
    vu, vt, x = (torch.randn(shape, device=dev, generator=g).bfloat16() for _ in range(3))
    sig = torch.linspace(1.0, 0.0, 11, device=dev)
    sched = sig[:-1] * 1000
    t = sched[3].expand(B).contiguous()
    ms = timed(lambda: ops.cfg_sde_step_logprob(vu, vt, x, t, sched, sig, 4.5, 0.7, seed=1))
    row("sde_fwd_kernel (rollout step)", f"{tag}: B={B} n={n}", B * (8 * n + 4), ms)
    prev, _, _, _ = ops.cfg_sde_step_logprob(vu, vt, x, t, sched, sig, 4.5, 0.7, seed=1)
    ms = timed(lambda: ops.cfg_sde_step_logprob(vu, vt, x, t, sched, sig, 4.5, 0.7, prev_sample=prev))
    row("sde_fwd_kernel (replay log-prob)", f"{tag}: B={B} n={n}", B * (8 * n + 4), ms)
    npred = torch.cat([vu, vt]).requires_grad_(True)

    def fb():
        lp, _, _ = ops.sde_logprob_replay(npred, x, prev, t, sched, sig, 4.5, 0.7)
        lp.sum().backward()
    fb()
    from adv_grpo_b200 import _lib
    gl = torch.ones(B, device=dev)
    gvu, gvt = torch.empty_like(vu), torch.empty_like(vt)
    ms = timed(lambda: _lib.call("advgrpo_cfg_sde_logprob_bwd", vu.data_ptr(), vt.data_ptr(), x.data_ptr(), prev.data_ptr(),
                                 t.data_ptr(), B, sched.data_ptr(), sig.data_ptr(), 10, gl.data_ptr(), gvu.data_ptr(),
                                 gvt.data_ptr(), B, n, 4.5, 0.7, torch.cuda.current_stream().cuda_stream))
    row("sde_bwd_kernel", f"{tag}: B={B} n={n}", B * 12 * n, ms)


def advantage(tag, N, T):
    r = torch.randn(N, T, device=dev)
    keys = torch.arange(N, device=dev).div(8, rounding_mode="floor")[:, None].expand(N, 256).contiguous()
    ms = timed(lambda: ops.group_advantage(r, keys))
    row("group_advantage_kernel", f"{tag}: N={N} T={T} L=256", N * T * 12 + N * 256 * 8, ms)
    lp, old = torch.randn(N, device=dev), torch.randn(N, device=dev)
    adv = torch.randn(N, device=dev, dtype=torch.float64)
    ms = timed(lambda: ops.grpo_clip_loss(lp, old, adv, 1e-5, 5.0))
    row("grpo_clip_loss_kernel", f"{tag}: N={N}", N * 16 + 56, ms)


def norms(tag, B, S_img, S_txt=205, W=1536, H=24):
    x = torch.randn(B, S_img, W, device=dev).bfloat16().requires_grad_(True)
    mod = torch.randn(B, 6 * W, device=dev).bfloat16()
    sh, sc = mod[:, :W], mod[:, W:2 * W]
    ms = timed(lambda: ops.ln_modulate(x.detach(), sh, sc))
    row("ln_modulate_fwd_kernel", f"{tag}: [{B},{S_img},{W}]", 2 * B * S_img * W * 2, ms)
    y = ops.ln_modulate(x, sh, sc)
    dy = torch.randn_like(y)
    ms = timed(lambda: torch.autograd.grad(y, x, dy, retain_graph=True))
    row("ln_modulate_bwd_kernel", f"{tag}: [{B},{S_img},{W}]", 3 * B * S_img * W * 2, ms)
    qi = torch.randn(B, S_img, 3 * W, device=dev).bfloat16().requires_grad_(True)
    qt = torch.randn(B, S_txt, 3 * W, device=dev).bfloat16().requires_grad_(True)
    w = [torch.randn(64, device=dev).bfloat16() for _ in range(4)]
    ms = timed(lambda: ops.qk_norm_concat(qi.detach(), qt.detach(), *w, H))
    row("qk_norm_concat_fwd_kernel", f"{tag}: [{B},{S_img}+{S_txt},3,{H},64]", 2 * B * (S_img + S_txt) * 3 * W * 2, ms)
    o = ops.qk_norm_concat(qi, qt, *w, H)
    do = torch.randn_like(o)
    ms = timed(lambda: torch.autograd.grad(o, (qi, qt), do, retain_graph=True))
    row("qk_norm_concat_bwd_kernel", f"{tag}: [{B},{S_img}+{S_txt},3,{H},64]", 3 * B * (S_img + S_txt) * 3 * W * 2, ms)


print(f"# HBM-bound kernels, CUDA events, L2 flushed before every launch; peak = {HBM:.0f} GB/s ({HBM_SRC})")
sde("cfg2", 8, 512)
sde("cfg4", 8, 1024)
if not ONCE:
    sde("stream", 64, 1024)
advantage("cfg2 1 rank", 16, 10)
advantage("cfg4 8 ranks", 8 * 16 * 2, 20)
advantage("large epoch", 2048, 2)
norms("cfg2", 16, 1024)
norms("cfg4", 16, 4096)

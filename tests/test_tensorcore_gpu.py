"""GPU parity tests of the tcgen05 kernels (GEMM with fused epilogues, flash attention) against
plain PyTorch fp32 references of the same op on the same bf16 inputs."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from adv_grpo_b200 import ops as _ops
    return _ops


def _rel_err(got, ref):
    return ((got.float() - ref.float()).abs().max() / ref.float().abs().max()).item()


# ------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 256, 128), (300, 1536, 1536), (1229, 4608, 1536),
                                   (77, 128, 256), (2458, 6144, 1536), (16, 9216, 1536), (500, 64, 1536)])
def test_gemm_plain_and_bias(ops, M, N, K):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g).bfloat16()
    ref = a.float() @ w.float().T
    got = ops.gemm(a, w)
    # bf16 output rounding (2^-8 relative) on fp32-accumulated products
    assert _rel_err(got, ref) < 6e-3
    got_b = ops.gemm(a, w, bias=bias)
    assert _rel_err(got_b, ref + bias.float()) < 6e-3


def test_gemm_epilogues_and_lora(ops):
    B, S, K, N, R = 2, 333, 1536, 1536, 64
    g = torch.Generator(device=DEV).manual_seed(0)
    a = torch.randn(B * S, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g).bfloat16()
    ref = a.float() @ w.float().T + bias.float()
    got = ops.gemm(a, w, bias=bias, epilogue=ops.EPI_GELU_TANH)
    assert _rel_err(got, torch.nn.functional.gelu(ref, approximate="tanh")) < 6e-3
    got = ops.gemm(a, w, bias=bias, epilogue=ops.EPI_GELU_ERF)
    assert _rel_err(got, torch.nn.functional.gelu(ref)) < 6e-3
    res = torch.randn(B * S, N, device=DEV, generator=g).bfloat16()
    gate_mat = torch.randn(B, 3 * N, device=DEV, generator=g).bfloat16()
    gate = gate_mat[:, N:2 * N]
    got = ops.gemm(a, w, bias=bias, epilogue=ops.EPI_GATE_RESIDUAL, residual=res, gate=gate, rows_per_gate=S)
    ref_g = res.float() + gate.float().repeat_interleave(S, 0) * ref
    assert _rel_err(got, ref_g) < 6e-3
    # LoRA second product
    a2 = torch.randn(B * S, R, device=DEV, generator=g).bfloat16()
    w2 = (0.1 * torch.randn(N, R, device=DEV, generator=g)).bfloat16()
    got = ops.gemm(a, w, bias=bias, a2=a2, w2=w2)
    assert _rel_err(got, ref + a2.float() @ w2.float().T) < 6e-3


def test_gemm_strided_operands(ops):
    g = torch.Generator(device=DEV).manual_seed(3)
    big = torch.randn(200, 3 * 256, device=DEV, generator=g).bfloat16()
    a = big[:, 256:512]                                   # row stride 768
    w = (torch.randn(128, 256, device=DEV, generator=g) / 16).bfloat16()
    assert _rel_err(ops.gemm(a, w), a.float() @ w.float().T) < 6e-3


# ------------------------------------------------------------------ attention
def _ref_attention(qkv, scale, causal):
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3).float() for i in range(3))
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal, scale=scale)
    s = (q @ k.transpose(-1, -2)) * scale
    if causal:
        S = s.shape[-1]
        s = s.masked_fill(torch.ones(S, S, device=s.device, dtype=torch.bool).triu(1), float("-inf"))
    return o.permute(0, 2, 1, 3), torch.logsumexp(s, dim=-1)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 16, 17, 18, 21, 22, 23])
@pytest.mark.parametrize("B,S,H", [(1, 128, 1), (2, 461, 4), (1, 1024, 2), (2, 1229, 3), (1, 1370, 2), (1, 77, 2),
                                   (1, 4301, 1)])
def test_attention_fwd_d64(ops, variant, B, S, H):
    g = torch.Generator(device=DEV).manual_seed(S + H)
    qkv = torch.randn(B, S, 3, H, 64, device=DEV, generator=g).bfloat16()
    out, lse = ops.attention_fwd(qkv, variant=variant)
    ref, lse_ref = _ref_attention(qkv, 0.125, False)
    # P is rounded to bf16 before P@V and O to bf16 on store: ~2^-8 relative to max |O|
    assert (out.float() - ref).abs().max().item() < 1.5e-2 * ref.abs().max().item()
    assert (lse - lse_ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("S,S_split", [(1229, 1024), (461, 256), (1229, 1000), (300, 77)])
def test_attention_fwd_pair_kernel_split_output_many_items(ops, S, S_split):
    """Default D=64 path (pair kernel: two query tiles per persistent CTA): more work items than SMs, so every CTA walks
    several items (Q double buffer, deferred epilogue, O staged in the Q buffers); image / text split output by TMA when
    the split is tile-aligned and by direct stores when it is not.  Bit-identical to the quad kernel (same arithmetic)."""
    g = torch.Generator(device=DEV).manual_seed(S)
    B, H = 6, 16                                              # 6 * 16 * ceil(S / 256) items on 148 CTAs
    qkv = torch.randn(B, S, 3, H, 64, device=DEV, generator=g).bfloat16()
    (o1, o2), lse = ops.attention_fwd(qkv, split=S_split)
    ref, lse_ref = ops.attention_fwd(qkv, variant=17)
    got = torch.cat([o1, o2], dim=1)
    assert torch.equal(got, ref)
    assert torch.equal(lse, lse_ref)
    ref32, _ = _ref_attention(qkv, 0.125, False)
    assert (got.float() - ref32).abs().max().item() < 1.5e-2 * ref32.abs().max().item()


def test_attention_fwd_large_logits_and_lazy_rescale(ops):
    """Rows whose running max grows by more than the lazy-rescale threshold across KV tiles."""
    B, S, H = 1, 640, 2
    g = torch.Generator(device=DEV).manual_seed(9)
    qkv = torch.randn(B, S, 3, H, 64, device=DEV, generator=g)
    qkv[:, :, 1] *= torch.linspace(0.2, 6.0, S, device=DEV)[None, :, None, None]   # later keys dominate
    qkv = qkv.bfloat16()
    out, lse = ops.attention_fwd(qkv)
    ref, lse_ref = _ref_attention(qkv, 0.125, False)
    assert (out.float() - ref).abs().max().item() < 2e-2 * ref.abs().max().item()
    assert (lse - lse_ref).abs().max().item() < 5e-3


@pytest.mark.parametrize("S", [77, 300])
def test_attention_fwd_causal(ops, S):
    g = torch.Generator(device=DEV).manual_seed(S)
    qkv = torch.randn(2, S, 3, 4, 64, device=DEV, generator=g).bfloat16()
    out, lse = ops.attention_fwd(qkv, causal=True)
    ref, lse_ref = _ref_attention(qkv, 0.125, True)
    assert (out.float() - ref).abs().max().item() < 1.5e-2 * ref.abs().max().item()
    assert (lse - lse_ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("S", [257, 50])
def test_attention_fwd_d128_padded_clip_heads(ops, S):
    """CLIP ViT-H heads are 80 wide: zero-padded to 128 with the 1/sqrt(80) scale."""
    g = torch.Generator(device=DEV).manual_seed(S)
    qkv = torch.zeros(2, S, 3, 4, 128, device=DEV)
    qkv[..., :80] = torch.randn(2, S, 3, 4, 80, device=DEV, generator=g)
    qkv = qkv.bfloat16()
    scale = 80 ** -0.5
    out, lse = ops.attention_fwd(qkv, scale=scale)
    ref, lse_ref = _ref_attention(qkv, scale, False)
    assert (out.float() - ref).abs().max().item() < 1.5e-2 * ref.abs().max().item()
    assert out[..., 80:].abs().max().item() == 0
    assert (lse - lse_ref).abs().max().item() < 2e-3


# ------------------------------------------------------------------ attention backward
@pytest.mark.parametrize("B,S,H,causal", [(1, 128, 1, False), (1, 256, 2, False), (2, 461, 3, False),
                                          (1, 1229, 2, False), (1, 1024, 1, False), (2, 300, 2, True)])
def test_attention_bwd_d64(ops, B, S, H, causal):
    g = torch.Generator(device=DEV).manual_seed(S + H)
    qkv = torch.randn(B, S, 3, H, 64, device=DEV, generator=g).bfloat16()
    dout = torch.randn(B, S, H, 64, device=DEV, generator=g).bfloat16()
    out, lse = ops.attention_fwd(qkv, causal=causal)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, causal=causal)
    ref_in = qkv.float().requires_grad_(True)
    q, k, v = (ref_in[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal).permute(0, 2, 1, 3)
    (o * dout.float()).sum().backward()
    ref = ref_in.grad
    for i, name in enumerate("qkv"):
        got_i, ref_i = dqkv[:, :, i].float(), ref[:, :, i]
        # P and dS are rounded to bf16 before the dV / dK / dQ products, grads stored in bf16
        err = (got_i - ref_i).abs().max().item() / ref_i.abs().max().item()
        assert err < 2e-2, (name, err)
        cos = torch.nn.functional.cosine_similarity(got_i.flatten(), ref_i.flatten(), dim=0).item()
        assert cos > 0.9995, (name, cos)


def test_attention_autograd_function(ops):
    g = torch.Generator(device=DEV).manual_seed(0)
    qkv = torch.randn(2, 333, 3, 4, 64, device=DEV, generator=g).bfloat16().requires_grad_(True)
    w = torch.randn(2, 333, 4, 64, device=DEV, generator=g)
    out = ops.attention(qkv)
    (out.float() * w).sum().backward()
    assert qkv.grad is not None and qkv.grad.shape == qkv.shape and torch.isfinite(qkv.grad.float()).all()


@pytest.mark.parametrize("variant", [0, 1, 3])
@pytest.mark.parametrize("M,N,K", [(1000, 1536, 256), (520, 192, 64), (3280, 1536, 1536), (256, 4608, 128), (700, 384, 192)])
def test_gemm_tile_variants(ops, variant, M, N, K):
    """CTA pairs / single CTA and 256- / 192- / 128-wide tiles must agree with fp32 matmul (incl. ragged M, N)."""
    g = torch.Generator(device=DEV).manual_seed(M + N)
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g).bfloat16()
    res = torch.randn(M, N, device=DEV, generator=g).bfloat16()
    gate = torch.randn(1, N, device=DEV, generator=g).bfloat16()
    ops.set_gemm_variant(variant)
    try:
        got = ops.gemm(a, w, bias=bias)
        got_r = ops.gemm(a, w, bias=bias, epilogue=ops.EPI_GATE_RESIDUAL, residual=res, gate=gate, rows_per_gate=M)
    finally:
        ops.set_gemm_variant(0)
    ref = a.float() @ w.float().T + bias.float()
    assert _rel_err(got, ref) < 6e-3
    assert _rel_err(got_r, res.float() + gate.float() * ref) < 6e-3


def test_gemm_dual_problem_launch(ops):
    """Image-stream + text-stream projection in one persistent launch == two separate launches."""
    g = torch.Generator(device=DEV).manual_seed(5)
    B, N_img, N_txt, K, N, R = 2, 1024, 205, 1536, 1536, 64
    a = [torch.randn(B * n, K, device=DEV, generator=g).bfloat16() for n in (N_img, N_txt)]
    w = [(torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16() for _ in range(2)]
    bias = [torch.randn(N, device=DEV, generator=g).bfloat16() for _ in range(2)]
    res = [torch.randn(B * n, N, device=DEV, generator=g).bfloat16() for n in (N_img, N_txt)]
    gate = [torch.randn(B, N, device=DEV, generator=g).bfloat16() for _ in range(2)]
    a2 = [torch.randn(B * n, R, device=DEV, generator=g).bfloat16() for n in (N_img, N_txt)]
    w2 = [(0.1 * torch.randn(N, R, device=DEV, generator=g)).bfloat16() for _ in range(2)]
    y = ops.gemm_dual(a, w, bias=bias, a2=a2, w2=w2, epilogue=ops.EPI_GATE_RESIDUAL, residual=res, gate=gate,
                      rows_per_gate=(N_img, N_txt))
    for i, n in enumerate((N_img, N_txt)):
        single = ops.gemm(a[i], w[i], bias=bias[i], a2=a2[i], w2=w2[i], epilogue=ops.EPI_GATE_RESIDUAL, residual=res[i],
                          gate=gate[i], rows_per_gate=n)
        ref = res[i].float() + gate[i].float().repeat_interleave(n, 0) * (a[i].float() @ w[i].float().T + bias[i].float()
                                                                         + a2[i].float() @ w2[i].float().T)
        assert _rel_err(y[i], ref) < 6e-3
        assert torch.allclose(y[i].float(), single.float(), atol=1e-2, rtol=1e-2)
    pre = [torch.empty(B * n, N, device=DEV, dtype=torch.bfloat16) for n in (N_img, N_txt)]
    h = ops.gemm_dual(a, w, bias=bias, epilogue=ops.EPI_GELU_TANH, preact_out=pre)
    for i in range(2):
        z = a[i].float() @ w[i].float().T + bias[i].float()
        assert _rel_err(pre[i], z) < 6e-3
        assert _rel_err(h[i], torch.nn.functional.gelu(z, approximate="tanh")) < 6e-3


@pytest.mark.parametrize("variant", [0, 1, 3])
@pytest.mark.parametrize("M,N,K", [(777, 1544, 128), (300, 200, 64), (2600, 6144, 256)])
def test_gemm_epilogue_staging_ring_ragged(ops, variant, M, N, K):
    """The epilogue staging ring (residual prefetch two column groups ahead, pre-activation + activation tiles)
    at ragged M / N (N not a multiple of the 64-column group, last tile partially outside) for every tile shape."""
    g = torch.Generator(device=DEV).manual_seed(M + N)
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g).bfloat16()
    z = a.float() @ w.float().T + bias.float()
    ops.set_gemm_variant(variant)
    try:
        pre = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
        h = ops.gemm(a, w, bias=bias, epilogue=ops.EPI_GELU_TANH, preact_out=pre)
        assert torch.equal(pre, z.bfloat16()) or _rel_err(pre, z) < 6e-3
        # the activation is evaluated on the bf16-rounded pre-activation (training replay == rollout)
        assert _rel_err(h, torch.nn.functional.gelu(pre.float(), approximate="tanh")) < 6e-3
        h2 = ops.gemm(a, w, bias=bias, epilogue=ops.EPI_GELU_TANH)
        assert torch.equal(h, h2)
        he = ops.gemm(a, w, bias=bias, epilogue=ops.EPI_GELU_ERF)
        assert _rel_err(he, torch.nn.functional.gelu(pre.float())) < 6e-3
        rpg = 100
        res = torch.randn(M, N, device=DEV, generator=g).bfloat16()
        gate = torch.randn((M + rpg - 1) // rpg, N, device=DEV, generator=g).bfloat16()
        y = ops.gemm(a, w, bias=bias, epilogue=ops.EPI_GATE_RESIDUAL, residual=res, gate=gate, rows_per_gate=rpg)
        ref = res.float() + gate.float().repeat_interleave(rpg, 0)[:M] * z
        assert _rel_err(y, ref) < 6e-3
    finally:
        ops.set_gemm_variant(0)


@pytest.mark.parametrize("variant", [0, 1, 3])
@pytest.mark.parametrize("B,S_img,S_txt,H", [(2, 1024, 205, 24), (3, 64, 13, 4), (16, 256, 205, 24), (2, 300, 0, 4),
                                             (5, 100, 77, 4)])
def test_gemm_qkv_norm_fused_epilogue_bit_exact(ops, variant, B, S_img, S_txt, H):
    """Fused QKV projection + per-head q/k RMSNorm + [image, text] concat (one launch, clipped 3-D TMA stores into
    the joint buffer) == dual GEMM followed by qk_norm_concat, bit for bit; also vs an fp32 torch reference."""
    g = torch.Generator(device=DEV).manual_seed(B + S_img + S_txt)
    K, D, R = 256, 64, 64
    N = 3 * H * D
    x = torch.randn(B, S_img, K, device=DEV, generator=g).bfloat16()
    c = torch.randn(B, S_txt, K, device=DEV, generator=g).bfloat16() if S_txt else None
    w = [(torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16() for _ in range(2)]
    bias = [torch.randn(N, device=DEV, generator=g).bfloat16() for _ in range(2)]
    nq = [(1 + 0.2 * torch.randn(D, device=DEV, generator=g)).bfloat16() for _ in range(2)]
    nk = [(1 + 0.2 * torch.randn(D, device=DEV, generator=g)).bfloat16() for _ in range(2)]
    a2 = [torch.randn(B, s, R, device=DEV, generator=g).bfloat16() for s in (S_img, S_txt)]
    w2 = [(0.1 * torch.randn(N, R, device=DEV, generator=g)).bfloat16() for _ in range(2)]
    ops.set_gemm_variant(variant)
    try:
        for lora in (False, True):
            kw = dict(a2=tuple(a2), w2=tuple(w2)) if lora else {}
            pre = [torch.full((B * s, N), float("nan"), device=DEV, dtype=torch.bfloat16) for s in (S_img, S_txt)]
            if S_txt:
                joint = ops.gemm_qkv_norm(x, c, w, bias, nq, nk, H, D, prenorm_out=tuple(pre), **kw)
                joint_np = ops.gemm_qkv_norm(x, c, w, bias, nq, nk, H, D, **kw)
                qx, qc = ops.gemm_dual((x, c), w, bias=bias, **kw)
                ref = ops.qk_norm_concat(qx, qc, nq[0], nk[0], nq[1], nk[1], H, D)
            else:
                kw1 = dict(a2=(a2[0],), w2=(w2[0],)) if lora else {}
                joint = ops.gemm_qkv_norm(x, None, w[:1], bias[:1], nq[:1], nk[:1], H, D, prenorm_out=(pre[0],), **kw1)
                joint_np = ops.gemm_qkv_norm(x, None, w[:1], bias[:1], nq[:1], nk[:1], H, D, **kw1)
                qx, qc = ops.gemm(x, w[0], bias=bias[0], **({"a2": a2[0], "w2": w2[0]} if lora else {})), None
                ref = ops.qk_norm_concat(qx, None, nq[0], nk[0], None, None, H, D)
            assert joint.shape == (B, S_img + S_txt, 3, H, D)
            assert torch.equal(joint, ref) and torch.equal(joint_np, ref)
            assert torch.equal(pre[0].reshape(qx.shape), qx)
            if S_txt:
                assert torch.equal(pre[1].reshape(qc.shape), qc)
            # fp32 reference of the same op on the image stream
            z = x.float().reshape(-1, K) @ w[0].float().T + bias[0].float()
            if lora:
                z = z + a2[0].float().reshape(-1, R) @ w2[0].float().T
            z = z.reshape(B, S_img, 3, H, D)
            rn = torch.rsqrt(z[:, :, :2].pow(2).mean(-1, keepdim=True) + 1e-6)
            wn = torch.stack([nq[0].float(), nk[0].float()])[None, None, :, None, :]
            zr = torch.cat([z[:, :, :2] * rn * wn, z[:, :, 2:]], 2)
            assert _rel_err(joint[:, :S_img], zr) < 1.2e-2
    finally:
        ops.set_gemm_variant(0)


@pytest.mark.parametrize("variant", [0, 1, 3])
def test_gemm_gelu_grad_epilogue_and_row_gate_mul(ops, variant):
    """Backward of a GELU feed-forward without elementwise passes: dz = ((dy * gate) W) * gelu_tanh'(z) with z
    prefetched like the residual tile; and the adaLN-gate multiply as one HBM pass."""
    g = torch.Generator(device=DEV).manual_seed(21 + variant)
    B, S, K, N = 3, 333, 256, 1544
    dy = torch.randn(B * S, K, device=DEV, generator=g).bfloat16()
    gate = torch.randn(B, 4 * K, device=DEV, generator=g).bfloat16()[:, K:2 * K]          # strided rows, like an adaLN chunk
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    z = (1.5 * torch.randn(B * S, N, device=DEV, generator=g)).bfloat16()
    d = ops.row_gate_mul(dy, gate, S)
    ref_d = (dy.float().view(B, S, K) * gate.float()[:, None]).view(B * S, K)
    assert torch.equal(d, ref_d.bfloat16())
    ops.set_gemm_variant(variant)
    try:
        got = ops.gemm(d, w, epilogue=ops.EPI_GELU_TANH_GRAD, residual=z)
        got2 = ops.gemm_dual((d, d[:100]), (w, w), epilogue=ops.EPI_GELU_TANH_GRAD, residual=(z, z[:100]))
    finally:
        ops.set_gemm_variant(0)
    acc = d.float() @ w.float().T
    zz = z.float().requires_grad_(True)
    torch.nn.functional.gelu(zz, approximate="tanh").backward(acc)
    assert _rel_err(got, zz.grad) < 6e-3
    assert _rel_err(got2[0], zz.grad) < 6e-3 and _rel_err(got2[1], zz.grad[:100]) < 6e-3


# ------------------------------------------------------------------ skinny TN GEMM (LoRA weight gradients)
@pytest.mark.parametrize("Kt,Ms,Nb", [(16384, 128, 1536), (3280, 64, 1536), (16384, 64, 6144), (2460, 128, 4608),
                                      (100, 64, 64), (77, 32, 136), (5000, 256, 264)])
@pytest.mark.parametrize("transpose", [False, True])
def test_gemm_tn_skinny_matches_fp32(ops, Kt, Ms, Nb, transpose):
    """dt^T x and (t^T dy)^T of the LoRA backward (peft lora.Linear autograd behind train_sd3_fast_pickscore.py:1165):
    contraction over the token axis with deterministic split-K, vs the fp32 product of the same bf16 operands."""
    g = torch.Generator(device=DEV).manual_seed(Kt + Ms)
    a = torch.randn(Kt, Ms, device=DEV, generator=g).bfloat16()
    b = torch.randn(Kt, Nb, device=DEV, generator=g).bfloat16()
    out = ops.gemm_tn_skinny(a, b, transpose_out=transpose)
    out2 = ops.gemm_tn_skinny(a, b, transpose_out=transpose)
    assert torch.equal(out, out2)                               # fixed summation order: bit-reproducible
    ref = a.float().t() @ b.float()
    if transpose:
        ref = ref.t()
    assert out.shape == ref.shape
    # fp32 accumulation, one bf16 rounding of the result
    assert (out.float() - ref).abs().max().item() < 6e-3 * ref.abs().max().item() + 1e-3


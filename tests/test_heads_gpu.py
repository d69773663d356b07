"""GPU tests of the score-head / discriminator-step kernels (csrc/heads.cu) and of the native autograd ops the trainable
discriminator blocks run on, each against plain PyTorch fp32 arithmetic
on the same inputs.  Tolerances are stated per test."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from adv_grpo_b200 import ops as _ops
    return _ops


def _rel(got, ref):
    return (got.float() - ref.float()).norm().item() / max(ref.float().norm().item(), 1e-12)


@pytest.mark.parametrize("rows,C", [(1, 8), (2080, 512), (8224, 1280), (70, 5120), (333, 264)])
def test_col_sum_matches_fp64(ops, rows, C):
    """Bias gradients / the head's dw2: sum_r s[r] a[r, c] b[r, c] with fp32 accumulation, deterministic."""
    g = torch.Generator(device=DEV).manual_seed(rows + C)
    a = torch.randn(rows, C, device=DEV, generator=g).bfloat16()
    b = torch.randn(rows, C, device=DEV, generator=g).bfloat16()
    s = torch.randn(rows, device=DEV, generator=g)
    for got, ref in ((ops.col_sum(a), a.double().sum(0)), (ops.col_sum(a, b=b), (a.double() * b.double()).sum(0)),
                     (ops.col_sum(a, row_scale=s), (a.double() * s.double()[:, None]).sum(0))):
        assert got.dtype == torch.float32 and got.shape == (C,)
        assert (got.double() - ref).abs().max().item() <= 1e-4 * max(1.0, math.sqrt(rows)) + 1e-5 * ref.abs().max().item()
    assert torch.equal(ops.col_sum(a, b=b), ops.col_sum(a, b=b))


@pytest.mark.parametrize("M,N,K", [(514, 1280, 1280), (1040, 512, 768), (300, 5120, 1280), (64, 64, 64)])
def test_gemm_tn_wide_matches_fp32(ops, M, N, K):
    """dW = dy^T x for full Linear layers (output wider than the 256 rows of the LoRA case)."""
    g = torch.Generator(device=DEV).manual_seed(M + N)
    dy = torch.randn(M, N, device=DEV, generator=g).bfloat16()
    x = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    got = ops.gemm_tn(dy, x)
    ref = dy.float().t() @ x.float()
    assert got.shape == (N, K) and _rel(got, ref) < 4e-3


@pytest.mark.parametrize("M", [300, 1028])
def test_gemm_erf_gelu_grad_epilogue(ops, M):
    """C = (d W^T) * gelu_erf'(z): the fc2 input-gradient GEMM of a ViT MLP with the activation derivative fused."""
    g = torch.Generator(device=DEV).manual_seed(M)
    K, N = 256, 512
    d = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    z = (1.5 * torch.randn(M, N, device=DEV, generator=g)).bfloat16()
    got = ops.gemm(d, w, epilogue=ops.EPI_GELU_ERF_GRAD, residual=z)
    zz = z.float().requires_grad_(True)
    torch.nn.functional.gelu(zz).backward(d.float() @ w.float().T)
    assert _rel(got, zz.grad) < 6e-3


def test_linear_and_mlp_autograd_match_torch_fp32(ops):
    """ops.linear / ops.mlp_gelu (forward + native backward) vs fp32 torch autograd on the same bf16 values: every
    result is one bf16 rounding of an fp32-accumulated product (rel 2^-8 per element, ~4e-3 in norm)."""
    g = torch.Generator(device=DEV).manual_seed(0)
    B, S, W, Mh = 2, 257, 256, 512
    x = torch.randn(B, S, W, device=DEV, generator=g).bfloat16()
    mk = lambda *s: (torch.randn(*s, device=DEV, generator=g) / math.sqrt(s[-1])).bfloat16()
    w, b, w1, b1, w2, b2 = mk(W, W), mk(W), mk(Mh, W), mk(Mh), mk(W, Mh), mk(W)
    dy = torch.randn(B, S, W, device=DEV, generator=g).bfloat16()
    leaves = [t.clone().requires_grad_() for t in (x, w, b, w1, b1, w2, b2)]
    xg, wg, bg, w1g, b1g, w2g, b2g = leaves
    y = ops.mlp_gelu(ops.linear(xg, wg, bg), w1g, b1g, w2g, b2g)
    y.backward(dy)
    ref_leaves = [t.float().clone().requires_grad_() for t in (x, w, b, w1, b1, w2, b2)]
    xr, wr, br, w1r, b1r, w2r, b2r = ref_leaves
    F = torch.nn.functional
    yr = F.linear(F.gelu(F.linear(F.linear(xr, wr, br), w1r, b1r)), w2r, b2r)
    yr.backward(dy.float())
    assert y.dtype == torch.bfloat16 and _rel(y, yr) < 8e-3
    for got, ref in zip(leaves, ref_leaves):
        assert got.grad.dtype == torch.bfloat16 and got.grad.shape == ref.grad.shape
        assert _rel(got.grad, ref.grad) < 1.5e-2, _rel(got.grad, ref.grad)
    # frozen weight: only the input gradient is produced
    xg2 = x.clone().requires_grad_()
    ops.linear(xg2, w).backward(dy)
    assert _rel(xg2.grad, dy.float() @ w.float()) < 6e-3


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("betas", [(0.5, 0.999), (0.9, 0.999)])
def test_adam_torch_order_matches_torch_adam(ops, dtype, betas):
    """TorchOrderAdam vs torch.optim.Adam (multi-tensor path) on parameters of the discriminator's dtypes: bf16
    parameters + bf16 moments must match to the last bit on (almost) every element (torch's own fused-multiply-add
    contraction inside each op is the only freedom), fp32 ones to fp32 rounding."""
    from adv_grpo_b200.optim import TorchOrderAdam
    g = torch.Generator(device=DEV).manual_seed(1)
    shapes = [(512, 768), (512,), (1, 512), (1,), (1000, 37)]
    p_nat = [(0.05 * torch.randn(*s, device=DEV, generator=g)).to(dtype).requires_grad_() for s in shapes]
    p_ref = [p.detach().clone().requires_grad_() for p in p_nat]
    frozen = torch.zeros(3, device=DEV, dtype=dtype, requires_grad=True)          # no gradient: skipped by both
    opt_n = TorchOrderAdam(p_nat + [frozen], lr=5e-4, betas=betas)
    opt_r = torch.optim.Adam(p_ref, lr=5e-4, betas=betas, foreach=True)
    for step in range(4):
        for a, b in zip(p_nat, p_ref):
            gr = (torch.randn(a.shape, device=DEV, generator=g) * (0.1 if step else 3.0)).to(dtype)
            a.grad, b.grad = gr.clone(), gr.clone()
        v0 = p_nat[0]._version
        opt_n.step()
        opt_r.step()
        assert p_nat[0]._version > v0
        opt_n.zero_grad()
        opt_r.zero_grad()
    for a, b in zip(p_nat, p_ref):
        if dtype == torch.bfloat16:
            diff = (a.detach().float() - b.detach().float()).abs()
            ulp = torch.clamp(b.detach().float().abs(), min=5e-4) * 2.0 ** -7       # of the parameter or of one lr-sized step
            assert (diff <= ulp).all()
            exact = (diff == 0).float().mean().item()
            assert exact > 0.999, exact
        else:
            assert torch.allclose(a.detach(), b.detach(), rtol=2e-6, atol=1e-9)
    st_n, st_r = opt_n.state[p_nat[0]], opt_r.state[p_ref[0]]
    assert st_n["step"] == 4 and st_n["exp_avg"].dtype == dtype
    tol = dict(rtol=1e-2, atol=1e-6) if dtype == torch.bfloat16 else dict(rtol=2e-6, atol=1e-12)
    assert torch.allclose(st_n["exp_avg"].float(), st_r["exp_avg"].float(), **tol)
    assert torch.allclose(st_n["exp_avg_sq"].float(), st_r["exp_avg_sq"].float(), **tol)
    sd = opt_n.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    opt2 = TorchOrderAdam(p_nat + [frozen], lr=1.0)
    opt2.load_state_dict(sd)
    assert opt2.state[p_nat[0]]["step"] == 4 and opt2.param_groups[0]["lr"] == 5e-4


@pytest.mark.parametrize("head_dtype", [torch.float32, torch.bfloat16])
def test_dino_head_reward_and_hinge_step_match_torch(ops, head_dtype):
    """A8b / A15 at the true DINOv2-B token shape: the native gather -> (L2 norm) -> Linear + GELU -> Linear(1) chain,
    the hybrid reward and the hinge loss with all four parameter gradients vs the same arithmetic written in torch fp32
    on the same bf16 tokens (adv_grpo/rewards.py:399-421, train_sd3_fast_dino_patch.py:186-219)."""
    from adv_grpo_b200.dinov2 import DINOHead, dino_hinge_d_loss, dino_patch_scores
    g = torch.Generator(device=DEV).manual_seed(3)
    B, T, D, n = 4, 1370, 768, 64
    feats = (torch.randn(2 * B, T, D, device=DEV, generator=g) * 2).bfloat16()
    idx = torch.randint(0, T - 1, (2 * B, n), device=DEV, generator=g)
    torch.manual_seed(0)
    head = DINOHead(in_dim=D).to(DEV).to(head_dtype)
    with torch.no_grad():
        head.layers[2].weight.mul_(8.0)                      # logits of O(1): both sides of the hinge are populated
    F = torch.nn.functional
    w1, b1, w2, b2 = (p.detach().float().clone().requires_grad_() for p in head.parameters())
    href = lambda t: F.linear(F.gelu(F.linear(t, w1, b1)), w2, b2).squeeze(-1)
    f32 = feats.float()
    gather = lambda f, i: torch.gather(f[:, 1:], 1, i.unsqueeze(-1).expand(-1, -1, D))
    # ---- reward path
    hybrid, cls_s, patch_s = dino_patch_scores(head, feats[:B], idx[:B], 0.7)
    nrm = lambda t: t / (t.norm(dim=-1, keepdim=True) + 1e-6)
    cls_r, patch_r = href(nrm(f32[:B, 0])), href(nrm(gather(f32[:B], idx[:B])))
    ref_h = 0.7 * cls_r + 0.3 * patch_r.mean(1)
    tol = 3e-2 if head_dtype == torch.bfloat16 else 6e-3      # bf16 head: the reference's bf16 rounding chain
    assert hybrid.shape == (B,) and patch_s.shape == (B, n) and hybrid.dtype == head_dtype
    assert (cls_s.float() - cls_r).abs().max().item() < tol * max(1.0, cls_r.abs().max().item())
    assert (patch_s.float() - patch_r).abs().max().item() < tol * max(1.0, patch_r.abs().max().item())
    assert (hybrid.float() - ref_h).abs().max().item() < tol * max(1.0, ref_h.abs().max().item())
    # ---- discriminator step
    loss, acc = dino_hinge_d_loss(head, feats[:B], feats[B:], idx[:B], idx[B:], 0.3)
    loss.backward()
    lr_, lf_ = href(f32[:B, 0]), href(f32[B:, 0])
    pr, pf = href(gather(f32[:B], idx[:B])), href(gather(f32[B:], idx[B:]))
    ref_loss = 0.5 * (F.relu(1 - lr_).mean() + F.relu(1 + lf_).mean()) + 0.3 * 0.5 * (F.relu(1 - pr).mean() + F.relu(1 + pf).mean())
    ref_loss.backward()
    ref_acc = 0.5 * ((lr_ > 0).float().mean() + (lf_ < 0).float().mean())
    assert abs(loss.item() - ref_loss.item()) < 1e-2 * max(1.0, abs(ref_loss.item())), (loss.item(), ref_loss.item())
    assert abs(acc.item() - ref_acc.item()) <= 0.26            # a CLS logit within rounding of zero may flip one of 8 votes
    for prm, ref in zip(head.parameters(), (w1, b1, w2, b2)):
        assert prm.grad is not None and prm.grad.dtype == head_dtype and prm.grad.shape == prm.shape
        cos = F.cosine_similarity(prm.grad.float().flatten(), ref.grad.flatten(), dim=0).item()
        assert cos > 0.995 and _rel(prm.grad, ref.grad) < 0.1, (cos, _rel(prm.grad, ref.grad))


def test_pickscore_head_kernel_both_arithmetics(ops):
    """Score tail of pickscore_scorer.py:44-51: fp32 mode vs fp64 math, bf16 mode vs the bf16 tensor expression."""
    g = torch.Generator(device=DEV).manual_seed(4)
    B, D = 16, 1024
    img = torch.randn(B, D, device=DEV, generator=g).bfloat16()
    txt = torch.randn(3, D, device=DEV, generator=g).bfloat16()
    index = torch.randint(0, 3, (B,), device=DEV, generator=g)
    ls = torch.tensor(4.6052, device=DEV)
    got = ops.pickscore_head(img, txt, index, ls, False)
    i64, t64 = img.double(), txt.double()[index]
    ref = ls.double().exp() * ((i64 / i64.norm(dim=-1, keepdim=True)) * (t64 / t64.norm(dim=-1, keepdim=True))).sum(-1) / 26
    assert (got.double() - ref).abs().max().item() < 1e-5
    ls16 = ls.bfloat16()
    got16 = ops.pickscore_head(img, txt, index, ls16, True)
    te, ie = txt[index], img
    want = (ls16.exp() * torch.bmm((te / te.norm(p=2, dim=-1, keepdim=True))[:, None, :],
                                   (ie / ie.norm(p=2, dim=-1, keepdim=True))[:, :, None]).reshape(-1)) / 26
    assert (got16 - want.float()).abs().max().item() <= 2.0 ** -7 * want.float().abs().max().item()
    one = ops.pickscore_head(img, txt[:1], None, ls, False)                         # one shared prompt, no index
    i0 = txt.double()[0]
    ref1 = ls.double().exp() * ((i64 / i64.norm(dim=-1, keepdim=True)) * (i0 / i0.norm())).sum(-1) / 26
    assert (one.double() - ref1).abs().max().item() < 1e-5


@pytest.mark.parametrize("rows,cols", [(64, 4096), (7, 260), (3, 16384)])
def test_row_softmax_f32(ops, rows, cols):
    g = torch.Generator(device=DEV).manual_seed(rows)
    x = torch.randn(rows, cols, device=DEV, generator=g) * 30
    ref = torch.softmax(x.double() * 0.044, dim=-1)
    got = ops.row_softmax_f32(x, scale=0.044)
    assert (got.double() - ref).abs().max().item() < 2e-6
    got_r = ops.row_softmax_f32(x.clone(), scale=0.044, round_tf32=True)
    assert (got_r.view(torch.int32) & 0x1FFF).abs().max().item() == 0             # TF32 values: low 13 mantissa bits clear
    assert (got_r.double() - ref).abs().max().item() <= 2.0 ** -11 * ref.max().item() + 2e-6
    y = x.clone()
    ops.row_softmax_f32(y, scale=0.044, out=y)                                     # in place
    assert torch.equal(y, got)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [(2, 64, 64, 32, 512, 3), (1, 64, 64, 128, 32, 3), (2, 9, 20, 32, 32, 1),
                                              (1, 4, 128, 256, 4096, 1), (1, 64, 64, 4096, 512, 1)])
@pytest.mark.parametrize("variant", [0, 2])
def test_conv_narrow_and_attention_shapes(ops, B, H, W, Cin, Cout, k, variant):
    """The shapes the all-native VAE decoder adds: conv_in (16 -> 32 padded input channels), conv_out (32-column tiles for
    the 3 -> 32 padded filters; variant 2 = the 128-column tile on the same problem), and the mid-block attention
    products as 1x1 convolutions (a weight matrix as the image; 4096 input channels)."""
    g = torch.Generator(device=DEV).manual_seed(B + H + W + Cin + Cout)
    x = torch.randn(B, Cin, H, W, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    w = torch.randn(Cout, Cin, k, k, device=DEV, generator=g) / (Cin * k * k) ** 0.5
    bias = torch.randn(Cout, device=DEV, generator=g)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = torch.nn.functional.conv2d(x.double(), w.double(), bias.double(), padding=k // 2).float()
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    from adv_grpo_b200 import _lib
    _lib.load().advgrpo_debug_set_conv_variant(variant)
    try:
        got = ops.conv2d_nhwc_tf32(x, ops.pack_conv_weight_tf32(w), bias, k)
    finally:
        _lib.load().advgrpo_debug_set_conv_variant(0)
    err = (got - ref).abs().max().item()
    assert err <= 4e-3 * ref.abs().max().item(), (err, ref.abs().max().item())


def test_vae_mid_attention_matches_fp64():
    """The decoder's single-head mid-block attention on the convolution kernel + row softmax vs fp64 attention."""
    from adv_grpo_b200 import weights
    from adv_grpo_b200.vae import AutoencoderKL
    vp = weights.init_vae_decoder(weights.VAE_SD3, seed=2, device="cpu")
    vae = AutoencoderKL(vp, weights.VAE_SD3, device=DEV)
    pre = "decoder.mid_block.attentions.0"
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(2, 512, 16, 16, device=DEV, generator=g).contiguous(memory_format=torch.channels_last)
    got = vae._mid_attn(pre, x)
    F = torch.nn.functional
    p = {k: v.double().to(DEV) for k, v in vp.items() if k.startswith(pre)}
    h = F.group_norm(x.double(), 32, p[pre + ".group_norm.weight"], p[pre + ".group_norm.bias"], eps=1e-6)
    h = h.permute(0, 2, 3, 1).reshape(2, 256, 512)
    q, k, v = (F.linear(h, p[f"{pre}.to_{n}.weight"], p[f"{pre}.to_{n}.bias"]) for n in "qkv")
    o = torch.softmax(q @ k.transpose(1, 2) * 512 ** -0.5, dim=-1) @ v
    o = F.linear(o, p[pre + ".to_out.0.weight"], p[pre + ".to_out.0.bias"])
    ref = x.double() + o.reshape(2, 16, 16, 512).permute(0, 3, 1, 2)
    err = (got.double() - ref).abs().max().item()
    assert err <= 4e-3 * ref.abs().max().item(), (err, ref.abs().max().item())


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("B,S,H,Dh", [(2, 257, 16, 80), (3, 50, 4, 64), (1, 512, 2, 80), (2, 33, 3, 128), (1, 1, 1, 16), (2, 77, 16, 64)])
def test_attention_small_fwd_bwd_matches_torch_fp32(ops, B, S, H, Dh, causal):
    """The attention core of the trainable CLIP-H blocks (A14): forward and all three input gradients vs fp32 torch SDPA
    on the same bf16 values; outputs are one bf16 rounding of fp32 results."""
    g = torch.Generator(device=DEV).manual_seed(S + Dh)
    q, k, v, do = (torch.randn(B, S, H, Dh, device=DEV, generator=g).bfloat16() for _ in range(4))
    qg, kg, vg = (t.clone().requires_grad_() for t in (q, k, v))
    o = ops.attention_small(qg, kg, vg, causal=causal)
    o.backward(do)
    qr, kr, vr = (t.float().transpose(1, 2).clone().requires_grad_() for t in (q, k, v))
    ref = torch.nn.functional.scaled_dot_product_attention(qr, kr, vr, is_causal=causal)
    ref.backward(do.float().transpose(1, 2))
    assert o.shape == (B, S, H, Dh) and o.dtype == torch.bfloat16
    assert _rel(o, ref.transpose(1, 2)) < 5e-3
    for got, r in ((qg.grad, qr.grad), (kg.grad, kr.grad), (vg.grad, vr.grad)):
        assert _rel(got, r.transpose(1, 2)) < 6e-3, _rel(got, r.transpose(1, 2))


def test_attention_small_rejects_what_does_not_fit(ops):
    from adv_grpo_b200 import _lib
    q = torch.zeros(1, 4096, 1, 64, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(_lib.AdvGrpoError):
        ops.attention_small(q, q, q)


@pytest.mark.parametrize("B,S,D", [(2, 77, 1536), (16, 1024, 1536), (3, 5, 256)])
def test_ln_modulate_full_grads_match_torch_fp32(ops, B, S, D):
    """LayerNorm-modulate with gradients to x AND to the per-sample shift / scale vectors (full fine-tuning): native forward,
    native dx, row-statistics + segmented column sums for d shift / d scale, vs fp32 torch autograd."""
    g = torch.Generator(device=DEV).manual_seed(B + S)
    x = (torch.randn(B, S, D, device=DEV, generator=g) * 2 + 0.3).bfloat16()
    mod = (0.3 * torch.randn(B, 6 * D, device=DEV, generator=g)).bfloat16()
    dy = torch.randn(B, S, D, device=DEV, generator=g).bfloat16()
    xg, mg = x.clone().requires_grad_(), mod.clone().requires_grad_()
    sh, sc = mg[:, D:2 * D], mg[:, 4 * D:5 * D]                          # row views of one [B, 6 D] matrix, like adaLN chunks
    y = ops.ln_modulate_full(xg, sh, sc)
    y.backward(dy)
    xr, mr = x.float().requires_grad_(), mod.float().requires_grad_()
    yr = torch.nn.functional.layer_norm(xr, (D,), eps=1e-6) * (1 + mr[:, None, 4 * D:5 * D]) + mr[:, None, D:2 * D]
    yr.backward(dy.float())
    assert _rel(y, yr) < 5e-3 and _rel(xg.grad, xr.grad) < 6e-3
    assert _rel(mg.grad[:, D:2 * D], mr.grad[:, D:2 * D]) < 6e-3           # d shift
    assert _rel(mg.grad[:, 4 * D:5 * D], mr.grad[:, 4 * D:5 * D]) < 6e-3   # d scale
    assert mg.grad[:, :D].abs().max().item() == 0

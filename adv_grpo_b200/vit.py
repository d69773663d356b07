"""Pre-LN ViT encoder blocks on the libadvgrpo_b200 kernels, shared by the PickScore CLIP-ViT-H/14
towers (`adv_grpo/pickscore_scorer.py:40-43`) and DINOv2-B/14 (`adv_grpo/rewards.py:397`).

Per block: LayerNorm (`ops.layer_norm`, the ln_modulate kernel in affine mode) -> fused QKV tcgen05 GEMM writing the token-major [B,S,3,H,Dp] buffer
the attention kernel reads through TMA -> flash attention -> out-projection GEMM with the residual
add (and DINOv2's LayerScale) fused as `x + gamma * (W o + b)` -> LayerNorm -> fc1 GEMM + GELU(erf)
epilogue -> fc2 GEMM + fused residual.  Heads narrower than the 64/128 the attention kernel supports
(CLIP-H: 80) are zero-padded INSIDE the packed weights, so no activation is ever padded or sliced.
"""
import torch
import torch.nn.functional as F

from . import ops


def _pad_heads_rows(w, b, heads, hd, hd_pad):
    """[heads*hd, K] -> [heads*hd_pad, K] with zero rows (and bias) for the padded head dims."""
    if hd == hd_pad:
        return w, b
    K = w.shape[1]
    wp = torch.zeros(heads, hd_pad, K, dtype=w.dtype, device=w.device)
    wp[:, :hd] = w.view(heads, hd, K)
    bp = torch.zeros(heads, hd_pad, dtype=b.dtype, device=b.device)
    bp[:, :hd] = b.view(heads, hd)
    return wp.reshape(heads * hd_pad, K), bp.reshape(-1)


def _pad_heads_cols(w, heads, hd, hd_pad):
    """[N, heads*hd] -> [N, heads*hd_pad] with zero columns."""
    if hd == hd_pad:
        return w
    N = w.shape[0]
    wp = torch.zeros(N, heads, hd_pad, dtype=w.dtype, device=w.device)
    wp[:, :, :hd] = w.view(N, heads, hd)
    return wp.reshape(N, heads * hd_pad)


class ViTBlock:
    def __init__(self, width, heads, wq, bq, wk, bk, wv, bv, wo, bo, ln1, ln2, fc1, fc2, eps, gamma1=None,
                 gamma2=None, act_epilogue=ops.EPI_GELU_ERF):
        hd = width // heads
        hd_pad = 64 if hd <= 64 else 128
        assert hd <= 128
        self.width, self.heads, self.hd, self.hd_pad, self.eps = width, heads, hd, hd_pad, eps
        parts = [_pad_heads_rows(w, b, heads, hd, hd_pad) for w, b in ((wq, bq), (wk, bk), (wv, bv))]
        self.w_qkv = torch.cat([p[0] for p in parts], 0).contiguous()
        self.b_qkv = torch.cat([p[1] for p in parts], 0).contiguous()
        self.w_o = _pad_heads_cols(wo, heads, hd, hd_pad).contiguous()
        self.b_o = bo
        self.ln1, self.ln2 = ln1, ln2
        self.w_fc1, self.b_fc1 = fc1
        self.w_fc2, self.b_fc2 = fc2
        ones = torch.ones(1, width, dtype=torch.bfloat16, device=wq.device)
        self.g1 = ones if gamma1 is None else gamma1.reshape(1, width).to(torch.bfloat16)
        self.g2 = ones if gamma2 is None else gamma2.reshape(1, width).to(torch.bfloat16)
        self.scale = hd ** -0.5
        self.act_epilogue = act_epilogue

    def __call__(self, x, causal=False):
        B, S, W = x.shape
        M = B * S
        h = ops.layer_norm(x, self.ln1[0], self.ln1[1], self.eps)
        qkv = ops.gemm(h, self.w_qkv, bias=self.b_qkv).view(B, S, 3, self.heads, self.hd_pad)
        o, _ = ops.attention_fwd(qkv, scale=self.scale, causal=causal, want_lse=False)
        x = ops.gemm(o.view(B, S, self.heads * self.hd_pad), self.w_o, bias=self.b_o,
                     epilogue=ops.EPI_GATE_RESIDUAL, residual=x, gate=self.g1, rows_per_gate=M)
        h = ops.layer_norm(x, self.ln2[0], self.ln2[1], self.eps)
        m = ops.gemm(h, self.w_fc1, bias=self.b_fc1, epilogue=self.act_epilogue)
        x = ops.gemm(m, self.w_fc2, bias=self.b_fc2, epilogue=ops.EPI_GATE_RESIDUAL, residual=x, gate=self.g2,
                     rows_per_gate=M)
        return x


def patch_embed(images, weight, bias, patch):
    """Conv2d(3, W, patch, stride=patch) as unfold + tcgen05 GEMM (K = 3*patch^2 zero-padded to 64k).
    `weight` is the pre-packed [W, Kpad] matrix from `pack_patch_weight`."""
    B, C, H, Wd = images.shape
    gh, gw = H // patch, Wd // patch
    x = images.to(torch.bfloat16).reshape(B, C, gh, patch, gw, patch).permute(0, 2, 4, 1, 3, 5)
    x = x.reshape(B * gh * gw, C * patch * patch)
    kpad = weight.shape[1]
    if kpad != x.shape[1]:
        x = F.pad(x, (0, kpad - x.shape[1]))
    return ops.gemm(x.contiguous(), weight, bias=bias).view(B, gh * gw, weight.shape[0])


def pack_patch_weight(conv_w):
    W = conv_w.shape[0]
    w = conv_w.reshape(W, -1)
    k = w.shape[1]
    kpad = (k + 63) // 64 * 64
    return F.pad(w, (0, kpad - k)).contiguous()

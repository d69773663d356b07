"""DINOv2 ViT-B/14 feature extractor + DINOHead on the B200 kernels: the `scorer` / `head` pair of
the reference's DINO adversarial reward (`scripts/train_sd3_fast_dino_patch.py:589-603`,
`adv_grpo/rewards.py:375-434`).  timm state-dict names (`vit_base_patch14_dinov2.lvd142m`)."""
import torch

from . import ops, vit
from .weights import DINOV2_B


class DinoV2:
    """Frozen backbone; `forward_features(images_bf16[B,3,518,518]) -> [B, 1+N, width]` (normed tokens)."""

    def __init__(self, params, cfg=DINOV2_B, device="cuda"):
        self.cfg = cfg
        self.num_features = cfg["width"]
        p = {k: v.to(device=device, dtype=torch.bfloat16) for k, v in params.items()}
        self.p = p
        w = cfg["width"]
        self.w_patch = vit.pack_patch_weight(p["patch_embed.proj.weight"])
        self.blocks = []
        for i in range(cfg["layers"]):
            b = f"blocks.{i}"
            wqkv, bqkv = p[b + ".attn.qkv.weight"], p[b + ".attn.qkv.bias"]
            self.blocks.append(vit.ViTBlock(
                w, cfg["heads"], wqkv[:w], bqkv[:w], wqkv[w:2 * w], bqkv[w:2 * w], wqkv[2 * w:], bqkv[2 * w:],
                p[b + ".attn.proj.weight"], p[b + ".attn.proj.bias"],
                (p[b + ".norm1.weight"], p[b + ".norm1.bias"]), (p[b + ".norm2.weight"], p[b + ".norm2.bias"]),
                (p[b + ".mlp.fc1.weight"], p[b + ".mlp.fc1.bias"]), (p[b + ".mlp.fc2.weight"], p[b + ".mlp.fc2.bias"]),
                1e-6, p[b + ".ls1.gamma"], p[b + ".ls2.gamma"]))

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    @torch.no_grad()
    def forward_features(self, images):
        p = self.p
        x = vit.patch_embed(images, self.w_patch, p["patch_embed.proj.bias"], self.cfg["patch"])
        x = torch.cat([p["cls_token"].expand(x.shape[0], -1, -1), x], 1) + p["pos_embed"]
        x = x.contiguous()
        for blk in self.blocks:
            x = blk(x)
        return ops.layer_norm(x, p["norm.weight"], p["norm.bias"], 1e-6)


class DINOHead(torch.nn.Module):
    """train_sd3_fast_dino_patch.py:592-603 (same module tree, so state dicts and DDP wrapping match)."""

    def __init__(self, in_dim=1024, hidden_dim=512):
        super().__init__()
        self.layers = torch.nn.Sequential(torch.nn.Linear(in_dim, hidden_dim), torch.nn.GELU(),
                                          torch.nn.Linear(hidden_dim, 1))

    def forward(self, x):
        return self.layers(x)


def _head_params(head):
    head = head.module if hasattr(head, "module") else head                 # DDP(head), train_sd3_fast_dino_patch.py:749
    l = head.layers
    return l[0].weight, l[0].bias, l[2].weight, l[2].bias


def dino_hinge_d_loss(head, feats_real, feats_fake, idx_real, idx_fake, patch_loss_weight=0.3):
    """Discriminator (head) loss of `train_sd3_fast_dino_patch.py:186-219`: hinge loss on the CLS token of real / fake
    images plus `patch_loss_weight` x the hinge loss on the sampled patch tokens (`idx_*` [B, n] = the torch.randint
    draws of :199-200; NO L2 normalisation in the D step, unlike the reward path).  Returns (loss, accuracy); everything
    -- token gather, Linear + GELU (tcgen05 GEMM), Linear(1), hinge, and the gradients of the four head parameters that
    `loss.backward()` deposits -- runs on the native kernels (`ops.dino_head_hinge_loss`, csrc/heads.cu)."""
    return ops.dino_head_hinge_loss(_head_params(head), feats_real, feats_fake, idx_real, idx_fake, patch_loss_weight)


def dino_patch_scores(head, feats, idx, cls_weight=0.7):
    """Reward path of `adv_grpo/rewards.py:399-421` behind `forward_features`: CLS + sampled patch tokens, L2-normalised,
    through the head, `cls_weight * cls + (1 - cls_weight) * mean(patch)`.  Returns (hybrid [B], cls [B], patch [B, n])."""
    w1, b1, w2, b2 = _head_params(head)
    B, n = idx.shape
    rows = ops.gather_rows_l2norm(feats, idx, l2norm=True, eps=1e-6)
    bf16_head = w1.dtype == torch.bfloat16
    logits, _, _ = ops.dino_head_forward((w1, b1, w2, b2), rows, bf16_head)
    hybrid = ops.dino_hybrid_score(logits, B, n, cls_weight, bf16_head)
    per = logits.view(B, 1 + n)
    out_dtype = w1.dtype
    return hybrid.to(out_dtype), per[:, 0].to(out_dtype), per[:, 1:].to(out_dtype)

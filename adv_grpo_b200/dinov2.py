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

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


def dino_hinge_d_loss(head, feats_real, feats_fake, idx_real, idx_fake, patch_loss_weight=0.3):
    """Discriminator (head) loss of `train_sd3_fast_dino_patch.py:186-219`: hinge loss on the CLS token of real / fake
    images plus `patch_loss_weight` x the hinge loss on the sampled patch tokens (`idx_*` [B, n] = the torch.randint
    draws of :199-200; NO L2 normalisation in the D step, unlike the reward path).  Returns (loss, accuracy)."""
    relu = torch.nn.functional.relu
    hp = next(head.parameters())
    fr, ff = feats_real.to(hp.dtype), feats_fake.to(hp.dtype)
    lr_, lf_ = head(fr[:, 0]).squeeze(-1), head(ff[:, 0]).squeeze(-1)
    image_loss = 0.5 * (relu(1.0 - lr_).mean() + relu(1.0 + lf_).mean())
    D = fr.shape[-1]
    sr = torch.gather(fr[:, 1:], 1, idx_real.unsqueeze(-1).expand(-1, -1, D))
    sf = torch.gather(ff[:, 1:], 1, idx_fake.unsqueeze(-1).expand(-1, -1, D))
    patch_loss = 0.5 * (relu(1.0 - head(sr).squeeze(-1)).mean() + relu(1.0 + head(sf).squeeze(-1)).mean())
    acc = 0.5 * ((lr_ > 0).float().mean() + (lf_ < 0).float().mean())
    return image_loss + patch_loss_weight * patch_loss, acc

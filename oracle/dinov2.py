"""Oracle: DINOv2 ViT-B/14 `forward_features` (timm `vit_base_patch14_dinov2.lvd142m`,
called at `adv_grpo/rewards.py:397` and `train_sd3_fast_dino_patch.py:182-184`), the
`DINOHead` (`train_sd3_fast_dino_patch.py:592-603`), the cls + random-patch hybrid
reward (`rewards.py:393-432`) and the discriminator hinge loss
(`train_sd3_fast_dino_patch.py:186-219`).  timm state-dict names.  timm is absent;
cross-checked against `transformers.Dinov2Model` (same architecture) in
tests/test_oracle_models.py.
Test infrastructure only (see oracle/__init__.py)."""
import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _ln(p, name, x):
    return F.layer_norm(x, (x.shape[-1],), p[name + ".weight"], p[name + ".bias"], 1e-6)


def forward_features(p, cfg, images):
    """timm `vit_base_patch14_dinov2.lvd142m` forward_features as called at rewards.py:397-399 / train_sd3_fast_dino_patch.py:589."""
    x = F.conv2d(images, p["patch_embed.proj.weight"], p["patch_embed.proj.bias"], stride=cfg["patch"])
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([p["cls_token"].expand(x.shape[0], -1, -1), x], 1) + p["pos_embed"]
    B, S, D = x.shape
    H = cfg["heads"]
    for i in range(cfg["layers"]):
        b = f"blocks.{i}"
        qkv = F.linear(_ln(p, b + ".norm1", x), p[b + ".attn.qkv.weight"], p[b + ".attn.qkv.bias"])
        q, k, v = qkv.view(B, S, 3, H, D // H).permute(2, 0, 3, 1, 4)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, S, D)
        x = x + p[b + ".ls1.gamma"] * F.linear(o, p[b + ".attn.proj.weight"], p[b + ".attn.proj.bias"])
        h = F.gelu(F.linear(_ln(p, b + ".norm2", x), p[b + ".mlp.fc1.weight"], p[b + ".mlp.fc1.bias"]))
        x = x + p[b + ".ls2.gamma"] * F.linear(h, p[b + ".mlp.fc2.weight"], p[b + ".mlp.fc2.bias"])
    return _ln(p, "norm", x)


def preprocess(images):
    """rewards.py:379-391: bicubic 518 (align_corners=False, no antialias) + ImageNet norm."""
    images = images.float()
    images = F.interpolate(images, size=(518, 518), mode="bicubic", align_corners=False)
    mean = torch.tensor(IMAGENET_MEAN)[None, :, None, None]
    std = torch.tensor(IMAGENET_STD)[None, :, None, None]
    return (images - mean) / std


def head(hp, x):
    """DINOHead: Linear(768,512) -> GELU -> Linear(512,1)."""
    h = F.gelu(F.linear(x, hp["layers.0.weight"], hp["layers.0.bias"]))
    return F.linear(h, hp["layers.2.weight"], hp["layers.2.bias"])


def patch_reward(hp, feats, idx, cls_weight=0.7):
    """rewards.py:399-421 given the sampled patch indices `idx` [B,n]."""
    cls_emb, patch_emb = feats[:, 0], feats[:, 1:]
    D = patch_emb.shape[-1]
    sp = torch.gather(patch_emb, 1, idx.unsqueeze(-1).expand(-1, -1, D))
    cls_emb = cls_emb / (cls_emb.norm(dim=-1, keepdim=True) + 1e-6)
    sp = sp / (sp.norm(dim=-1, keepdim=True) + 1e-6)
    cls_score = head(hp, cls_emb).squeeze(-1)
    patch_scores = head(hp, sp).squeeze(-1)
    return cls_weight * cls_score + (1 - cls_weight) * patch_scores.mean(1), cls_score, patch_scores


def hinge_d_loss(hp, feats_real, feats_fake, idx_real, idx_fake, patch_loss_weight=0.3):
    """train_sd3_fast_dino_patch.py:186-219 (note: NO L2 normalisation in the D step)."""
    cr, pr = feats_real[:, 0], feats_real[:, 1:]
    cf, pf = feats_fake[:, 0], feats_fake[:, 1:]
    lr, lf = head(hp, cr).squeeze(-1), head(hp, cf).squeeze(-1)
    image_loss = 0.5 * (F.relu(1.0 - lr).mean() + F.relu(1.0 + lf).mean())
    D = pr.shape[-1]
    sr = torch.gather(pr, 1, idx_real.unsqueeze(-1).expand(-1, -1, D))
    sf = torch.gather(pf, 1, idx_fake.unsqueeze(-1).expand(-1, -1, D))
    plr, plf = head(hp, sr).squeeze(-1), head(hp, sf).squeeze(-1)
    patch_loss = 0.5 * (F.relu(1.0 - plr).mean() + F.relu(1.0 + plf).mean())
    acc = 0.5 * ((lr > 0).float().mean() + (lf < 0).float().mean())
    return image_loss + patch_loss_weight * patch_loss, acc

"""CLIP (PickScore_v1 = CLIP-ViT-H/14) on the B200 kernels, exposing what the reference reaches through
`scorer.model`: `.get_image_features(pixel_values=)`, `.get_text_features(input_ids=)`, `.logit_scale`,
`.parameters()`, `.vision_model.encoder.layers[i:]` (`adv_grpo/pickscore_scorer.py:40-47`,
`adv_grpo/pick_score_training.py:101-102`, `scripts/train_sd3_fast_pickscore.py:1016-1020`).

The module tree and parameter names are those of transformers' CLIPModel, so a PickScore_v1 state
dict loads as is.  Frozen layers run on the tcgen05 kernels (packed weights, rebuilt when a
parameter's version counter changes); layers with requires_grad parameters under grad mode (the last
`tune_layer` vision blocks during the discriminator step) run through differentiable native ops: affine LayerNorm
(`ops.layer_norm`, native dx / d weight / d bias), every Linear on the tcgen05 GEMM with weight gradients on the
split-K TN kernel and bias gradients on the column-sum kernel (`ops.linear`), the MLP with the erf-GELU and its
derivative fused into GEMM epilogues (`ops.mlp_gelu`), and the 257-token, head_dim-80 softmax(QK^T)V core with its backward
on the shared-memory-resident short-sequence kernel (`ops.attention_small`, csrc/attn_small.cu: the D = 64 tcgen05
attention-backward kernel's TMEM layout has no room for a 128-wide head).  No torch SDPA / nn.Linear / LayerNorm call is
left on the discriminator step.
"""
import torch
from torch import nn

from . import ops, vit
from .weights import CLIP_H


class _Attn(nn.Module):
    def __init__(self, w):
        super().__init__()
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = (nn.Linear(w, w) for _ in range(4))


class _MLP(nn.Module):
    def __init__(self, w, m):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(w, m), nn.Linear(m, w)


class CLIPEncoderLayer(nn.Module):
    def __init__(self, width, heads, mlp):
        super().__init__()
        self.width, self.heads = width, heads
        self.layer_norm1 = nn.LayerNorm(width, eps=1e-5)
        self.self_attn = _Attn(width)
        self.layer_norm2 = nn.LayerNorm(width, eps=1e-5)
        self.mlp = _MLP(width, mlp)
        self._packed, self._packed_key = None, None

    def _fast(self):
        key = tuple(p._version for p in self.parameters()) + (next(self.parameters()).data_ptr(),)
        if self._packed is None or key != self._packed_key:
            a, m = self.self_attn, self.mlp
            d = lambda t: t.detach()
            self._packed = vit.ViTBlock(
                self.width, self.heads, d(a.q_proj.weight), d(a.q_proj.bias), d(a.k_proj.weight), d(a.k_proj.bias),
                d(a.v_proj.weight), d(a.v_proj.bias), d(a.out_proj.weight), d(a.out_proj.bias),
                (d(self.layer_norm1.weight), d(self.layer_norm1.bias)),
                (d(self.layer_norm2.weight), d(self.layer_norm2.bias)),
                (d(m.fc1.weight), d(m.fc1.bias)), (d(m.fc2.weight), d(m.fc2.bias)), 1e-5)
            self._packed_key = key
        return self._packed

    def forward(self, x, causal=False):
        needs_autograd = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if not needs_autograd:
            return self._fast()(x, causal)
        B, S, W = x.shape
        hd = W // self.heads
        ln1, ln2, a, m = self.layer_norm1, self.layer_norm2, self.self_attn, self.mlp
        h = ops.layer_norm(x, ln1.weight, ln1.bias, ln1.eps)
        q, k, v = (ops.linear(h, l.weight, l.bias).view(B, S, self.heads, hd) for l in (a.q_proj, a.k_proj, a.v_proj))
        o = ops.attention_small(q, k, v, scale=hd ** -0.5, causal=causal).reshape(B, S, W)
        x = x + ops.linear(o, a.out_proj.weight, a.out_proj.bias)
        return x + ops.mlp_gelu(ops.layer_norm(x, ln2.weight, ln2.bias, ln2.eps), m.fc1.weight, m.fc1.bias,
                                m.fc2.weight, m.fc2.bias)


def _ln(mod, x):
    """Affine LayerNorm on the native kernel (forward and, under autograd, backward: A14)."""
    return ops.layer_norm(x.contiguous(), mod.weight, mod.bias, mod.eps)


def _proj(lin, x):
    """Bias-free projection of the pooled token on the tcgen05 GEMM (differentiable through `ops.linear`)."""
    if torch.is_grad_enabled() and (x.requires_grad or lin.weight.requires_grad):
        return ops.linear(x.contiguous(), lin.weight)
    return ops.gemm(x.contiguous(), lin.weight.detach())


class _Encoder(nn.Module):
    def __init__(self, width, heads, mlp, layers):
        super().__init__()
        self.layers = nn.ModuleList([CLIPEncoderLayer(width, heads, mlp) for _ in range(layers)])


class _VisionEmbeddings(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        w, p = cfg["v_width"], cfg["patch"]
        self.class_embedding = nn.Parameter(torch.zeros(w))
        self.patch_embedding = nn.Conv2d(3, w, p, stride=p, bias=False)
        self.position_embedding = nn.Embedding((cfg["image"] // p) ** 2 + 1, w)


class _VisionModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        w = cfg["v_width"]
        self.embeddings = _VisionEmbeddings(cfg)
        self.pre_layrnorm = nn.LayerNorm(w, eps=1e-5)
        self.encoder = _Encoder(w, cfg["v_heads"], cfg["v_mlp"], cfg["v_layers"])
        self.post_layernorm = nn.LayerNorm(w, eps=1e-5)
        self._w_patch = None

    def forward(self, pixel_values):
        e = self.embeddings
        with torch.no_grad():
            if self._w_patch is None or self._w_patch.device != e.patch_embedding.weight.device:
                self._w_patch = vit.pack_patch_weight(e.patch_embedding.weight.detach())
            x = vit.patch_embed(pixel_values, self._w_patch, None, self.cfg["patch"])
            cls = e.class_embedding.detach().to(x.dtype).expand(x.shape[0], 1, -1)
            x = torch.cat([cls, x], 1) + e.position_embedding.weight.detach().to(x.dtype)[None]
            x = ops.layer_norm(x, self.pre_layrnorm.weight.detach(), self.pre_layrnorm.bias.detach(), 1e-5).contiguous()
        for layer in self.encoder.layers:
            x = layer(x, causal=False)
        return _ln(self.post_layernorm, x[:, 0])


class _TextEmbeddings(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.token_embedding = nn.Embedding(cfg["vocab"], cfg["t_width"])
        self.position_embedding = nn.Embedding(cfg["ctx"], cfg["t_width"])


class _TextModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        w = cfg["t_width"]
        self.embeddings = _TextEmbeddings(cfg)
        self.encoder = _Encoder(w, cfg["t_heads"], cfg["t_mlp"], cfg["t_layers"])
        self.final_layer_norm = nn.LayerNorm(w, eps=1e-5)

    def forward(self, input_ids):
        e = self.embeddings
        S = input_ids.shape[1]
        x = (e.token_embedding(input_ids) + e.position_embedding.weight[:S][None]).contiguous()
        for layer in self.encoder.layers:
            x = layer(x, causal=True)
        x = _ln(self.final_layer_norm, x)
        return x[torch.arange(x.shape[0], device=x.device), input_ids.argmax(-1)]   # EOS (highest id) pooling


class CLIPModel(nn.Module):
    def __init__(self, cfg=CLIP_H):
        super().__init__()
        self.cfg = cfg
        self.vision_model = _VisionModel(cfg)
        self.text_model = _TextModel(cfg)
        self.visual_projection = nn.Linear(cfg["v_width"], cfg["proj"], bias=False)
        self.text_projection = nn.Linear(cfg["t_width"], cfg["proj"], bias=False)
        self.logit_scale = nn.Parameter(torch.tensor(2.6592))

    @classmethod
    def from_params(cls, params, cfg=CLIP_H, device="cuda", dtype=torch.bfloat16):
        m = cls(cfg)
        m.load_state_dict({k: v for k, v in params.items()}, strict=True)
        return m.to(device=device, dtype=dtype).eval()

    def get_image_features(self, pixel_values=None, **_):
        return _proj(self.visual_projection, self.vision_model(pixel_values))

    def get_text_features(self, input_ids=None, attention_mask=None, **_):
        return _proj(self.text_projection, self.text_model(input_ids))

"""Oracle: SD3 / SD3.5-medium MMDiT(-X) forward, restated from the diffusers
0.33.1 architecture that the reference calls at `fast.py:630-637` and
`train_sd3_fast_pickscore.py:235-255` (`SD3Transformer2DModel`,
`JointTransformerBlock`, `JointAttnProcessor2_0`, `AdaLayerNormZero`,
`SD35AdaLayerNormZeroX`, `AdaLayerNormContinuous`, `PatchEmbed`,
`CombinedTimestepTextProjEmbeddings`) plus the peft LoRA wrapper the scripts
put on 8 attention projections (`train_sd3_fast_pickscore.py:488-505`).

Pure functions of a flat parameter dict that uses the diffusers state-dict
names.  diffusers is not installed here: **parity unpinned** for this body
(SURVEY.md section 8c); it is the architecture of `transformer/config.json`
of stabilityai/stable-diffusion-3.5-medium.

Test infrastructure only (see oracle/__init__.py).
"""
import math
import torch
import torch.nn.functional as F

LORA_TARGETS = ("attn.to_q", "attn.to_k", "attn.to_v", "attn.to_out.0",
                "attn.add_q_proj", "attn.add_k_proj", "attn.add_v_proj", "attn.to_add_out")


def sincos_1d(dim, pos):
    """diffusers get_1d_sincos_pos_embed_from_grid ([sin, cos] halves) behind PatchEmbed (call site fast.py:630-637)."""
    omega = torch.arange(dim // 2, dtype=torch.float64) / (dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = pos.reshape(-1).to(torch.float64)[:, None] * omega[None]
    return torch.cat([out.sin(), out.cos()], dim=1)


def cropped_pos_embed(dim, h, w, max_size=384, base_size=64):
    """diffusers PatchEmbed.cropped_pos_embed over get_2d_sincos_pos_embed(dim, max_size,
    base_size=base_size, interpolation_scale=1): centre crop of the fixed 2-D table."""
    top, left = (max_size - h) // 2, (max_size - w) // 2
    gh = torch.arange(top, top + h, dtype=torch.float64) / (max_size / base_size)
    gw = torch.arange(left, left + w, dtype=torch.float64) / (max_size / base_size)
    # grid = meshgrid(grid_w, grid_h): grid[0] varies along w, grid[1] along h
    g0 = gw[None, :].expand(h, w)
    g1 = gh[:, None].expand(h, w)
    emb = torch.cat([sincos_1d(dim // 2, g0), sincos_1d(dim // 2, g1)], dim=1)
    return emb.to(torch.float32).reshape(1, h * w, dim)


def timestep_embedding(t, dim=256, max_period=10000):
    """diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0) of CombinedTimestepTextProjEmbeddings (call site fast.py:630-637)."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32) / half
    emb = t.float()[:, None] * exponent.exp()[None]
    return torch.cat([emb.cos(), emb.sin()], dim=-1)       # flip_sin_to_cos=True


def layer_norm(x, eps=1e-6):
    """nn.LayerNorm(elementwise_affine=False, eps=1e-6) of AdaLayerNormZero / AdaLayerNormContinuous (call site fast.py:630-637)."""
    return F.layer_norm(x, (x.shape[-1],), eps=eps)


def rms_norm(x, w, eps=1e-6):
    """diffusers RMSNorm(head_dim, eps=1e-6) on q / k inside JointAttnProcessor2_0 (call site fast.py:630-637)."""
    var = x.float().pow(2).mean(-1, keepdim=True)
    return (x * torch.rsqrt(var + eps)).to(w.dtype) * w


def ada_layer_norm_zero(x, e):
    """diffusers AdaLayerNormZero (call site fast.py:630-637): e = linear(silu(temb)) [B, 6 d] chunks as (shift_msa, scale_msa,
    gate_msa, shift_mlp, scale_mlp, gate_mlp); returns the modulated input of the attention and the remaining four vectors."""
    sh, sc, g, sh_m, sc_m, g_m = e.chunk(6, dim=1)
    return layer_norm(x) * (1 + sc[:, None]) + sh[:, None], g, sh_m, sc_m, g_m


def ada_layer_norm_continuous(x, e):
    """diffusers AdaLayerNormContinuous (norm_out, and norm1_context of the context_pre_only last block): e = linear(silu(temb))
    [B, 2 d] chunks as (SCALE, SHIFT) -- the opposite order of AdaLayerNormZero."""
    sc, sh = e.chunk(2, dim=1)
    return layer_norm(x) * (1 + sc)[:, None] + sh[:, None]


class MMDiTOracle:
    def __init__(self, params, cfg, lora=None, lora_scale=2.0, dtype=torch.float32):
        """params: diffusers-named dict. cfg: dict(num_layers, heads, head_dim, dual_layers,
        qk_norm, patch_size, in_channels, pos_embed_max_size, base_size).
        lora: optional dict name -> (A[r,in], B[out,r]) for LORA_TARGETS of each block."""
        self.p = {k: v.to(dtype) for k, v in params.items()}
        self.cfg = cfg
        self.lora = {k: (a.to(dtype), b.to(dtype)) for k, (a, b) in (lora or {}).items()}
        self.lora_scale = lora_scale
        self.dtype = dtype

    def lin(self, name, x):
        """nn.Linear, or peft lora.Linear on the 8 targets of train_sd3_fast_pickscore.py:488-505: y + scale * B(A x)."""
        y = F.linear(x, self.p[name + ".weight"], self.p.get(name + ".bias"))
        if name in self.lora:                                   # peft lora.Linear: y + scale * B(A x)
            a, b = self.lora[name]
            y = y + self.lora_scale * F.linear(F.linear(x, a), b)
        return y

    def _heads(self, x):
        B, S, _ = x.shape
        return x.view(B, S, self.cfg["heads"], self.cfg["head_dim"]).transpose(1, 2)

    def _attn(self, pre, x, ctx=None, ctx_out=True):
        q, k, v = (self._heads(self.lin(f"{pre}.to_{n}", x)) for n in "qkv")
        if self.cfg["qk_norm"]:
            q = rms_norm(q, self.p[f"{pre}.norm_q.weight"])
            k = rms_norm(k, self.p[f"{pre}.norm_k.weight"])
        if ctx is not None:
            cq, ck, cv = (self._heads(self.lin(f"{pre}.add_{n}_proj", ctx)) for n in "qkv")
            if self.cfg["qk_norm"]:
                cq = rms_norm(cq, self.p[f"{pre}.norm_added_q.weight"])
                ck = rms_norm(ck, self.p[f"{pre}.norm_added_k.weight"])
            q, k, v = torch.cat([q, cq], 2), torch.cat([k, ck], 2), torch.cat([v, cv], 2)   # [image, text]
        o = F.scaled_dot_product_attention(q, k, v)
        B, H, S, D = o.shape
        o = o.transpose(1, 2).reshape(B, S, H * D)
        if ctx is None:
            return self.lin(f"{pre}.to_out.0", o), None
        n = x.shape[1]
        o, co = o[:, :n], o[:, n:]
        co = self.lin(f"{pre}.to_add_out", co) if ctx_out else None
        return self.lin(f"{pre}.to_out.0", o), co

    def _ff(self, pre, x):
        return self.lin(f"{pre}.net.2", F.gelu(self.lin(f"{pre}.net.0.proj", x), approximate="tanh"))

    def block(self, i, x, c, temb):
        """diffusers JointTransformerBlock.forward (incl. the SD3.5 dual-attention and the context_pre_only last block), call site fast.py:630-637."""
        pre = f"transformer_blocks.{i}"
        last = i == self.cfg["num_layers"] - 1
        dual = i in self.cfg["dual_layers"]
        e = self.lin(f"{pre}.norm1.linear", F.silu(temb))
        nx = layer_norm(x)
        if dual:                                               # SD35AdaLayerNormZeroX: the six AdaLayerNormZero chunks + three for attn2
            x1, g, sh_m, sc_m, g_m = ada_layer_norm_zero(x, e[:, :6 * x.shape[-1]])
            sh2, sc2, g2 = e[:, 6 * x.shape[-1]:].chunk(3, dim=1)
        else:
            x1, g, sh_m, sc_m, g_m = ada_layer_norm_zero(x, e)
        ce = self.lin(f"{pre}.norm1_context.linear", F.silu(temb))
        if last:                                               # AdaLayerNormContinuous: scale, shift
            c1 = ada_layer_norm_continuous(c, ce)
        else:
            c1, cg, csh_m, csc_m, cg_m = ada_layer_norm_zero(c, ce)
        a, ca = self._attn(f"{pre}.attn", x1, c1, ctx_out=not last)
        x = x + g[:, None] * a
        if dual:
            x2 = nx * (1 + sc2[:, None]) + sh2[:, None]
            a2, _ = self._attn(f"{pre}.attn2", x2)
            x = x + g2[:, None] * a2
        x = x + g_m[:, None] * self._ff(f"{pre}.ff", layer_norm(x) * (1 + sc_m[:, None]) + sh_m[:, None])
        if last:
            return x, None
        c = c + cg[:, None] * ca
        c = c + cg_m[:, None] * self._ff(f"{pre}.ff_context",
                                         layer_norm(c) * (1 + csc_m[:, None]) + csh_m[:, None])
        return x, c

    def forward(self, hidden_states, timestep, encoder_hidden_states, pooled_projections, upto=None):
        """diffusers SD3Transformer2DModel.forward as called at fast.py:630-637 and train_sd3_fast_pickscore.py:235-255."""
        cfg, dt = self.cfg, self.dtype
        ps = cfg["patch_size"]
        B, C, H, W = hidden_states.shape
        h, w = H // ps, W // ps
        x = F.conv2d(hidden_states.to(dt), self.p["pos_embed.proj.weight"], self.p["pos_embed.proj.bias"], stride=ps)
        x = x.flatten(2).transpose(1, 2)
        d = x.shape[-1]
        x = x + cropped_pos_embed(d, h, w, cfg["pos_embed_max_size"], cfg["base_size"]).to(dt)
        te = timestep_embedding(timestep).to(dt)
        temb = self.lin("time_text_embed.timestep_embedder.linear_2",
                        F.silu(self.lin("time_text_embed.timestep_embedder.linear_1", te)))
        temb = temb + self.lin("time_text_embed.text_embedder.linear_2",
                               F.silu(self.lin("time_text_embed.text_embedder.linear_1", pooled_projections.to(dt))))
        c = self.lin("context_embedder", encoder_hidden_states.to(dt))
        for i in range(cfg["num_layers"] if upto is None else upto):
            x, c = self.block(i, x, c, temb)
        if upto is not None:
            return x, c
        x = ada_layer_norm_continuous(x, self.lin("norm_out.linear", F.silu(temb)))
        x = self.lin("proj_out", x)
        x = x.reshape(B, h, w, ps, ps, cfg["in_channels"])
        x = torch.einsum("nhwpqc->nchpwq", x).reshape(B, cfg["in_channels"], h * ps, w * ps)
        return x

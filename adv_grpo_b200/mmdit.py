"""SD3 / SD3.5-medium MMDiT(-X) on the libadvgrpo_b200 kernels: the `pipeline.transformer` object of
the reference's rollout (`fast.py:630-637`) and replay (`train_sd3_fast_pickscore.py:233-255`),
i.e. diffusers' `SD3Transformer2DModel` wrapped by the peft LoRA of
`train_sd3_fast_pickscore.py:488-505`, re-designed for B200:

  * token-major bf16 activations end to end, no head transposes: the fused QKV projection writes
    [B, S, 3*H*D]; one kernel applies the per-head q/k RMSNorm and concatenates image + text rows
    into the joint buffer the attention kernel reads through TMA; the attention kernel writes the
    image / text rows of O into two contiguous buffers for the two out-projections;
  * every linear is the tcgen05 GEMM with its epilogue fused (bias, GELU-tanh, gate * y + residual)
    and the LoRA update accumulated into the same TMEM accumulator as a second product;
  * all adaLN modulation vectors of all 24 blocks come from ONE weight-bandwidth-bound GEMM per
    forward (the conditioning vector is shared by every block);
  * rollout (no_grad) and training replay run the SAME fused forward arithmetic, so
    ratio = exp(logp - logp_old) is exactly 1 for unchanged weights and identical inputs.

Gradients flow to the LoRA A/B matrices only (everything else is frozen); the backward of each
linear is dX = dY W (one tcgen05 GEMM against the pre-transposed weight, plus the LoRA term as a
second product); the LoRA weight gradients dt^T x / dy^T t are the skinny split-K TN kernel (ops.gemm_tn_skinny).

State-dict names follow diffusers / peft so released checkpoints map one to one.
"""
import math

import torch
import torch.nn.functional as F

from . import ops
from .weights import LORA_TARGETS


# ------------------------------------------------------------------------------ helpers
def _sincos_1d(dim, pos):
    omega = torch.arange(dim // 2, dtype=torch.float64, device=pos.device) / (dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = pos.reshape(-1).to(torch.float64)[:, None] * omega[None]
    return torch.cat([out.sin(), out.cos()], dim=1)


def cropped_pos_embed(dim, h, w, max_size, base_size, device):
    """Centre crop of diffusers' fixed 2-D sincos table (PatchEmbed.cropped_pos_embed), built
    directly for the crop instead of materialising the 384 x 384 x dim buffer."""
    top, left = (max_size - h) // 2, (max_size - w) // 2
    gh = torch.arange(top, top + h, dtype=torch.float64, device=device) / (max_size / base_size)
    gw = torch.arange(left, left + w, dtype=torch.float64, device=device) / (max_size / base_size)
    g0 = gw[None, :].expand(h, w)
    g1 = gh[:, None].expand(h, w)
    emb = torch.cat([_sincos_1d(dim // 2, g0), _sincos_1d(dim // 2, g1)], dim=1)
    return emb.to(torch.float32).reshape(1, h * w, dim)


def timestep_embedding(t, dim=256, max_period=10000):
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t.float()[:, None] * exponent.exp()[None]
    return torch.cat([emb.cos(), emb.sin()], dim=-1)


class _LinearFn(torch.autograd.Function):
    """y = epi(x W^T + (x A^T)(sB)^T + b) on the tcgen05 GEMM; backward w.r.t. x and the LoRA factors."""

    @staticmethod
    def forward(ctx, x, w, wt_holder, bias, lora_a, lora_w2, epilogue, residual, gate, rows_per_gate):
        M = x.numel() // x.shape[-1]
        t = None
        if lora_a is not None:
            t = ops.gemm(x, lora_a)                                   # [.., r_pad]
        need_grad = torch.is_grad_enabled() or any(ctx.needs_input_grad)
        preact = None
        if epilogue in (ops.EPI_GELU_TANH, ops.EPI_GELU_ERF) and any(ctx.needs_input_grad):
            preact = torch.empty((M, w.shape[0]), dtype=torch.bfloat16, device=x.device)
        y = ops.gemm(x, w, bias=bias, a2=t, w2=lora_w2, epilogue=epilogue, residual=residual, gate=gate,
                     rows_per_gate=rows_per_gate, preact_out=preact)
        ctx.save_for_backward(x, t, lora_a, lora_w2, gate, preact)
        ctx.meta = (wt_holder, epilogue, rows_per_gate)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, t, lora_a, lora_w2, gate, preact = ctx.saved_tensors
        wt_holder, epilogue, rows_per_gate = ctx.meta
        dy = dy.contiguous()
        lead = dy.shape[:-1]
        N = dy.shape[-1]
        dy2 = dy.reshape(-1, N)
        d_res = None
        if epilogue == ops.EPI_GATE_RESIDUAL:
            d_res = dy
            dy2 = ops.row_gate_mul(dy2, gate, rows_per_gate)
        elif epilogue == ops.EPI_GELU_TANH:
            dy2 = torch.ops.aten.gelu_backward(dy2, preact, approximate="tanh")
        elif epilogue == ops.EPI_GELU_ERF:
            dy2 = torch.ops.aten.gelu_backward(dy2, preact, approximate="none")
        dx = da = dw2 = None
        dt = None
        if lora_a is not None:
            dt = ops.gemm(dy2, lora_w2.t().contiguous())               # dY (sB) -> [M, r_pad]
            x2 = x.reshape(-1, x.shape[-1])
            if ctx.needs_input_grad[4]:
                da = ops.gemm_tn_skinny(dt, x2)                        # dt^T x -> [r_pad, K]   (tcgen05 split-K TN kernel)
            if ctx.needs_input_grad[5]:
                dw2 = ops.gemm_tn_skinny(t.reshape(-1, t.shape[-1]), dy2, transpose_out=True)   # dy^T t -> [N, r_pad]
        if ctx.needs_input_grad[0]:
            wt = wt_holder()
            if dt is not None:
                dx = ops.gemm(dy2, wt, a2=dt, w2=lora_a.t().contiguous())
            else:
                dx = ops.gemm(dy2, wt)
            dx = dx.reshape(*lead, wt.shape[0])
        return dx, None, None, None, da, dw2, None, d_res, None, None


class _DualLinearFn(torch.autograd.Function):
    """The image-stream and text-stream versions of one projection (different weights, same N / K / epilogue) as
    ONE persistent dual-problem GEMM launch, forward and backward.  Argument layout: the two problems' tensors
    interleaved (x0, x1, w0, w1, wt0, wt1, b0, b1, a0, a1, w2_0, w2_1, epilogue, r0, r1, g0, g1, rows0, rows1)."""

    @staticmethod
    def forward(ctx, x0, x1, w0, w1, wt0, wt1, b0, b1, la0, la1, lw0, lw1, epilogue, r0, r1, g0, g1, rows0, rows1):
        has_lora = la0 is not None
        t0 = t1 = None
        if has_lora:
            t0, t1 = ops.gemm_dual((x0, x1), (la0, la1))
        pre = (None, None)
        if epilogue in (ops.EPI_GELU_TANH, ops.EPI_GELU_ERF) and any(ctx.needs_input_grad):
            pre = tuple(torch.empty((x.numel() // x.shape[-1], w0.shape[0]), dtype=torch.bfloat16, device=x.device)
                        for x in (x0, x1))
        y0, y1 = ops.gemm_dual((x0, x1), (w0, w1), bias=(b0, b1), a2=(t0, t1), w2=(lw0, lw1), epilogue=epilogue,
                               residual=(r0, r1), gate=(g0, g1), rows_per_gate=(rows0, rows1), preact_out=pre)
        ctx.save_for_backward(x0, x1, t0, t1, la0, la1, lw0, lw1, g0, g1, pre[0], pre[1])
        ctx.meta = (wt0, wt1, epilogue, rows0, rows1)
        return y0, y1

    @staticmethod
    def backward(ctx, dy0, dy1):
        x0, x1, t0, t1, la0, la1, lw0, lw1, g0, g1, pre0, pre1 = ctx.saved_tensors
        wt0, wt1, epilogue, rows0, rows1 = ctx.meta
        xs, ts, las, lws, gs, pres, rows = (x0, x1), (t0, t1), (la0, la1), (lw0, lw1), (g0, g1), (pre0, pre1), (rows0, rows1)
        dys, d_res, dy2 = [dy0.contiguous(), dy1.contiguous()], [None, None], [None, None]
        for i in range(2):
            N = dys[i].shape[-1]
            d = dys[i].reshape(-1, N)
            if epilogue == ops.EPI_GATE_RESIDUAL:
                d_res[i] = dys[i]
                d = ops.row_gate_mul(d, gs[i], rows[i])
            elif epilogue == ops.EPI_GELU_TANH:
                d = torch.ops.aten.gelu_backward(d, pres[i], approximate="tanh")
            elif epilogue == ops.EPI_GELU_ERF:
                d = torch.ops.aten.gelu_backward(d, pres[i], approximate="none")
            dy2[i] = d
        das, dws, dts = [None, None], [None, None], (None, None)
        has_lora = la0 is not None
        if has_lora:
            dts = ops.gemm_dual(tuple(dy2), (lw0.t().contiguous(), lw1.t().contiguous()))
            for i in range(2):
                x2 = xs[i].reshape(-1, xs[i].shape[-1])
                if ctx.needs_input_grad[8 + i]:
                    das[i] = ops.gemm_tn_skinny(dts[i], x2)
                if ctx.needs_input_grad[10 + i]:
                    dws[i] = ops.gemm_tn_skinny(ts[i].reshape(-1, ts[i].shape[-1]), dy2[i], transpose_out=True)
        dx0 = dx1 = None
        need0, need1 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if need0 and need1:
            if has_lora:
                dx0, dx1 = ops.gemm_dual(tuple(dy2), (wt0(), wt1()), a2=dts, w2=(la0.t().contiguous(), la1.t().contiguous()))
            else:
                dx0, dx1 = ops.gemm_dual(tuple(dy2), (wt0(), wt1()))
            dx0 = dx0.reshape(*dys[0].shape[:-1], dx0.shape[-1])
            dx1 = dx1.reshape(*dys[1].shape[:-1], dx1.shape[-1])
        else:
            for i, need in enumerate((need0, need1)):
                if not need:
                    continue
                wt = (wt0, wt1)[i]()
                if has_lora:
                    dx = ops.gemm(dy2[i], wt, a2=dts[i], w2=las[i].t().contiguous())
                else:
                    dx = ops.gemm(dy2[i], wt)
                dx = dx.reshape(*dys[i].shape[:-1], wt.shape[0])
                if i == 0:
                    dx0 = dx
                else:
                    dx1 = dx
        return (dx0, dx1, None, None, None, None, None, None, das[0], das[1], dws[0], dws[1], None, d_res[0], d_res[1],
                None, None, None, None)


class _DualFeedForwardFn(torch.autograd.Function):
    """Both streams' feed-forward of one MMDiT block, x + gate * (GELU_tanh(x_mod W1^T + b1) W2^T + b2), as two
    dual-problem GEMM launches forward and two backward (the weights are frozen: no LoRA on the FF layers,
    train_sd3_fast_pickscore.py:490-499).  Backward: dz = ((dy * gate) W2) * gelu'(z) in the epilogue of the first
    GEMM (z = stored bf16 pre-activation), dx_mod = dz W1; the hidden activation itself is never stored."""

    @staticmethod
    def forward(ctx, x0, x1, w1_0, w1_1, w1t_0, w1t_1, b1_0, b1_1, w2_0, w2_1, w2t_0, w2t_1, b2_0, b2_1, r0, r1, g0, g1,
                rows0, rows1):
        pre = (None, None)
        if any(ctx.needs_input_grad):
            pre = tuple(torch.empty((x.numel() // x.shape[-1], w1_0.shape[0]), dtype=torch.bfloat16, device=x.device)
                        for x in (x0, x1))
        h0, h1 = ops.gemm_dual((x0, x1), (w1_0, w1_1), bias=(b1_0, b1_1), epilogue=ops.EPI_GELU_TANH, preact_out=pre)
        y0, y1 = ops.gemm_dual((h0, h1), (w2_0, w2_1), bias=(b2_0, b2_1), epilogue=ops.EPI_GATE_RESIDUAL,
                               residual=(r0, r1), gate=(g0, g1), rows_per_gate=(rows0, rows1))
        ctx.save_for_backward(g0, g1, pre[0], pre[1])
        ctx.meta = (w1t_0, w1t_1, w2t_0, w2t_1, rows0, rows1)
        return y0, y1

    @staticmethod
    def backward(ctx, dy0, dy1):
        g0, g1, pre0, pre1 = ctx.saved_tensors
        w1t_0, w1t_1, w2t_0, w2t_1, rows0, rows1 = ctx.meta
        dy0, dy1 = dy0.contiguous(), dy1.contiguous()
        d0 = ops.row_gate_mul(dy0.reshape(-1, dy0.shape[-1]), g0, rows0)
        d1 = ops.row_gate_mul(dy1.reshape(-1, dy1.shape[-1]), g1, rows1)
        dz0, dz1 = ops.gemm_dual((d0, d1), (w2t_0(), w2t_1()), epilogue=ops.EPI_GELU_TANH_GRAD, residual=(pre0, pre1))
        dx0, dx1 = ops.gemm_dual((dz0, dz1), (w1t_0(), w1t_1()))
        dx0 = dx0.reshape(*dy0.shape[:-1], dx0.shape[-1])
        dx1 = dx1.reshape(*dy1.shape[:-1], dx1.shape[-1])
        return (dx0, dx1) + (None,) * 12 + (dy0, dy1, None, None, None, None)


class _LoraPackFn(torch.autograd.Function):
    """All LoRA GEMM operands of the model from the ONE flat fp32 master buffer in three kernels: scale (alpha / r on
    the B factors), gather through a precomputed index (zero padding = index of an appended zero), cast to bf16.
    Outputs are the per-linear (A_pad, sB_pad) views of the packed buffer; backward concatenates their gradients and
    gathers them back through the inverse index.  Replaces ~800 cat / pad / mul launches per forward (and as many in
    the backward and after every optimizer step)."""

    @staticmethod
    def forward(ctx, flat, layout):
        ext = torch.cat([flat.detach() * layout["scale"], layout["zero"]])
        packed = ext.index_select(0, layout["idx"]).to(torch.bfloat16)
        ctx.layout = layout
        return tuple(packed[o:o + n].view(shape) for o, n, shape in layout["views"])

    @staticmethod
    def backward(ctx, *grads):
        layout = ctx.layout
        parts = [(g.reshape(-1).to(torch.bfloat16) if g is not None else torch.zeros(n, dtype=torch.bfloat16, device=layout["idx"].device))
                 for g, (o, n, shape) in zip(grads, layout["views"])]
        gp = torch.cat(parts)
        return gp.index_select(0, layout["inv"]).float() * layout["scale"], None


class _MasterUnpack(torch.autograd.Function):
    """Full fine-tuning (`config.use_lora = False`, train_sd3_fast_pickscore.py:488): the ONE flat fp32 master parameter
    -> the model's bf16 working weights as autograd leaves-by-proxy.  Forward hands out aliases of the working copies
    (no copy: they were refreshed from the master after the last optimizer step); backward adds every weight's bf16
    gradient into its fp32 view of `master.grad` in place (a view-based unpack would materialise a zero tensor of the
    whole 2.2 B-element buffer per weight)."""

    @staticmethod
    def forward(ctx, master, model):
        ctx.model = model
        return tuple(model.p[n].detach() for n in model._full_names)

    @staticmethod
    def backward(ctx, *grads):
        for gv, g in zip(ctx.model._full_grad_views, grads):
            if g is not None:
                gv.add_(g.reshape(gv.shape))
        return None, None


class _WT:
    """Lazily materialised transposed copy of a frozen weight (only layers that back-propagate
    to their input ever build one)."""

    def __init__(self, w):
        self.w, self.wt = w, None

    def __call__(self):
        if self.wt is None:
            self.wt = self.w.t().contiguous()
        return self.wt


class SD3Transformer2DModel(torch.nn.Module):
    def __init__(self, cfg, params, lora_rank=32, lora_alpha=64, lora=None, device="cuda"):
        super().__init__()
        self.cfg = dict(cfg)
        self.config = type("Config", (), dict(in_channels=cfg["in_channels"], patch_size=cfg["patch_size"],
                                              sample_size=cfg["base_size"] * cfg["patch_size"],
                                              joint_attention_dim=cfg["joint_dim"]))()
        self.device_ = torch.device(device)
        d = cfg["heads"] * cfg["head_dim"]
        self.d = d
        p = {k: v.to(device=self.device_, dtype=torch.bfloat16) for k, v in params.items()}
        self.p = p
        L = cfg["num_layers"]
        # ---- pack per-block weights (the originals stay reachable as views: state_dict()) ----
        self.blocks = []
        ada_w, ada_b, self.ada_off = [], [], []
        off = 0
        for i in range(L):
            b = f"transformer_blocks.{i}"
            last, dual = i == L - 1, i in cfg["dual_layers"]
            blk = dict(last=last, dual=dual, idx=i)

            def cat(names, suffix):
                return torch.cat([p[f"{b}.{n}.{suffix}"] for n in names], 0).contiguous()

            blk["w_qkv"], blk["b_qkv"] = cat(("attn.to_q", "attn.to_k", "attn.to_v"), "weight"), cat(("attn.to_q", "attn.to_k", "attn.to_v"), "bias")
            blk["w_cqkv"] = cat(("attn.add_q_proj", "attn.add_k_proj", "attn.add_v_proj"), "weight")
            blk["b_cqkv"] = cat(("attn.add_q_proj", "attn.add_k_proj", "attn.add_v_proj"), "bias")
            if dual:
                blk["w_qkv2"], blk["b_qkv2"] = cat(("attn2.to_q", "attn2.to_k", "attn2.to_v"), "weight"), cat(("attn2.to_q", "attn2.to_k", "attn2.to_v"), "bias")
            for key, name in (("out", "attn.to_out.0"), ("cout", "attn.to_add_out"), ("out2", "attn2.to_out.0"),
                              ("ff1", "ff.net.0.proj"), ("ff2", "ff.net.2"), ("cff1", "ff_context.net.0.proj"),
                              ("cff2", "ff_context.net.2")):
                if f"{b}.{name}.weight" in p:
                    blk["w_" + key], blk["b_" + key] = p[f"{b}.{name}.weight"], p[f"{b}.{name}.bias"]
            for key in [k for k in blk if k.startswith("w_")]:
                blk["wt_" + key[2:]] = _WT(blk[key])
            if cfg["qk_norm"]:
                for key, name in (("nq", "attn.norm_q"), ("nk", "attn.norm_k"), ("ncq", "attn.norm_added_q"),
                                  ("nck", "attn.norm_added_k"), ("nq2", "attn2.norm_q"), ("nk2", "attn2.norm_k")):
                    if f"{b}.{name}.weight" in p:
                        blk[key] = p[f"{b}.{name}.weight"]
            nx = 9 if dual else 6
            nc = 2 if last else 6
            ada_w += [p[f"{b}.norm1.linear.weight"], p[f"{b}.norm1_context.linear.weight"]]
            ada_b += [p[f"{b}.norm1.linear.bias"], p[f"{b}.norm1_context.linear.bias"]]
            self.ada_off.append((off, off + nx * d))
            off += (nx + nc) * d
            self.blocks.append(blk)
        ada_w += [p["norm_out.linear.weight"]]
        ada_b += [p["norm_out.linear.bias"]]
        self.ada_out_off = off
        self.ada_w = torch.cat(ada_w, 0).contiguous()
        self.ada_b = torch.cat(ada_b, 0).contiguous()
        self.w_patch = p["pos_embed.proj.weight"].reshape(d, -1).contiguous()
        self._pos_cache = {}
        # ---- LoRA: ONE flat fp32 master parameter (peft-layout A [r, in] / B [out, r] factors are views of it) ----
        self.lora_rank, self.lora_scale = lora_rank, (lora_alpha / lora_rank if lora_rank else 0.0)
        self.lora_A, self.lora_B = {}, {}            # name -> view of the flat buffer (shares storage; .grad = view of flat.grad)
        self._lora_names = []
        init, sizes = [], []
        if lora_rank:
            for i in range(L):
                for t in LORA_TARGETS:
                    name = f"transformer_blocks.{i}.{t}"
                    if name + ".weight" not in p:
                        continue
                    if lora is not None:
                        a, bb = lora[name]
                    else:
                        a = torch.randn(lora_rank, d) / lora_rank
                        bb = torch.zeros(d, lora_rank)
                    init.append((name, a.to(torch.float32), bb.to(torch.float32)))
                    self._lora_names.append(name)
        flat = torch.cat([t.reshape(-1) for _, a, bb in init for t in (a, bb)]) if init else torch.zeros(0)
        self.lora_flat = torch.nn.Parameter(flat.to(self.device_, torch.float32).contiguous())
        self.lora_flat.grad = torch.zeros_like(self.lora_flat)
        off = 0
        self._lora_off = {}
        for name, a, bb in init:
            key = name.replace(".", "_")
            for dct, t in ((self.lora_A, a), (self.lora_B, bb)):
                v = self.lora_flat.data[off:off + t.numel()].view(t.shape)
                v.grad = self.lora_flat.grad[off:off + t.numel()].view(t.shape)
                dct[key] = v
                self._lora_off[(key, "A" if dct is self.lora_A else "B")] = off
                off += t.numel()
        self._lora_layout = self._make_lora_layout() if init else None
        self.dual_gemm = True            # image + text projections of a block as one dual-problem GEMM launch
        self.fused_qkv_norm = True       # no-grad forward: q/k RMSNorm + concat inside the QKV GEMM epilogue
        self.fused_ff = True             # feed-forward pair as one autograd node (GELU backward in a GEMM epilogue)
        self._lora_cache = None
        self._lora_dirty = False
        self._lora_enabled = True
        self.full_finetune = False       # config.use_lora = False: enable_full_finetune()

    # ------------------------------------------------------------------ peft-like surface
    def to(self, *args, **kwargs):
        """The weights live on the device given at construction (packed bf16 operands); `.to(accelerator.device)` /
        `.to(dtype)` of the scripts is accepted and ignored."""
        return self

    def set_adapter(self, name="default"):
        self.lora_flat.requires_grad_(True)
        return self

    def trainable_parameters(self):
        """The single flat LoRA parameter (optimizer / clipping / EMA / all-reduce work on one tensor); the per-layer
        factors are `lora_A[key]` / `lora_B[key]` views of it.  Under full fine-tuning: the flat fp32 master of every weight."""
        if self.full_finetune:
            return [self.full_master]
        return [self.lora_flat]

    def lora_state_dict(self):
        out = {}
        for name in self._lora_names:
            key = name.replace(".", "_")
            out[f"base_model.model.{name}.lora_A.weight"] = self.lora_A[key]
            out[f"base_model.model.{name}.lora_B.weight"] = self.lora_B[key]
        return out

    def save_pretrained(self, path):
        """peft adapter directory (`adapter_config.json` + `adapter_model.safetensors`), as `save_ckpt` of
        `train_sd3_fast_pickscore.py:389-398` writes it."""
        if self.full_finetune:
            from .checkpoint import save_full_transformer
            return save_full_transformer(self, path)
        from .checkpoint import save_lora
        save_lora(self, path)

    def load_adapter(self, path, strict=True):
        """`PeftModel.from_pretrained(transformer, config.train.lora_path)` (`train_pick:506-509`)."""
        from .checkpoint import load_lora
        return load_lora(self, path, strict=strict)

    def invalidate_lora_cache(self):
        """Call after an optimizer step / EMA swap changed the LoRA parameters."""
        self._lora_dirty = True

    class _Disable:
        def __init__(self, m):
            self.m = m

        def __enter__(self):
            self.prev = self.m._lora_enabled
            self.m._lora_enabled = False

        def __exit__(self, *a):
            self.m._lora_enabled = self.prev

    def disable_adapter(self):
        return SD3Transformer2DModel._Disable(self)

    # ------------------------------------------------------------------ LoRA operand packing
    def _make_lora_layout(self):
        """Index maps between the flat master buffer and the packed bf16 GEMM operands: per fused linear
        (A_pad [r_pad, K], sB_pad [N, r_pad]) with r_pad a multiple of 64 (zero padded; block-diagonal for fused QKV)."""
        r, d, dev = self.lora_rank, self.d, self.device_
        n_flat = self.lora_flat.numel()
        idx_parts, views, keys = [], [], []
        scale = torch.ones(n_flat, dtype=torch.float32)
        pos = 0
        for blk in self.blocks:
            i = blk["idx"]
            entry = {}
            for pk_key, names in (("qkv", ("attn.to_q", "attn.to_k", "attn.to_v")),
                                  ("cqkv", ("attn.add_q_proj", "attn.add_k_proj", "attn.add_v_proj")),
                                  ("out", ("attn.to_out.0",)), ("cout", ("attn.to_add_out",))):
                ks = [f"transformer_blocks.{i}.{t}".replace(".", "_") for t in names]
                if ks[0] not in self.lora_A:
                    continue
                n = len(ks)
                rp = ((n * r + 63) // 64) * 64
                ia = torch.full((rp, d), n_flat, dtype=torch.int64)
                iw = torch.full((n * d, rp), n_flat, dtype=torch.int64)
                for j, k in enumerate(ks):
                    oa, ob = self._lora_off[(k, "A")], self._lora_off[(k, "B")]
                    ia[j * r:(j + 1) * r] = oa + torch.arange(r * d).view(r, d)
                    iw[j * d:(j + 1) * d, j * r:(j + 1) * r] = ob + torch.arange(d * r).view(d, r)
                    scale[ob:ob + d * r] = self.lora_scale
                for t in (ia, iw):
                    idx_parts.append(t.reshape(-1))
                    views.append((pos, t.numel(), tuple(t.shape)))
                    pos += t.numel()
                entry[pk_key] = (len(views) - 2, len(views) - 1)
            keys.append(entry)
        idx = torch.cat(idx_parts)
        inv = torch.empty(n_flat, dtype=torch.int64)
        valid = idx < n_flat
        inv[idx[valid]] = torch.arange(idx.numel())[valid]
        return dict(idx=idx.to(dev, torch.int32), inv=inv.to(dev, torch.int32), scale=scale.to(dev),
                    zero=torch.zeros(1, device=dev), views=views, keys=keys)

    def _packs_from(self, tensors):
        return [{k: (tensors[ia], tensors[iw]) for k, (ia, iw) in entry.items()} for entry in self._lora_layout["keys"]]

    def _pack_lora(self):
        if self.full_finetune:
            if self._lora_dirty:
                self._refresh_from_master()
            return [dict() for _ in self.blocks]
        if not self.lora_rank or not self._lora_enabled or self._lora_layout is None:
            return [dict() for _ in self.blocks]
        if torch.is_grad_enabled() and self.lora_flat.requires_grad:
            return self._packs_from(_LoraPackFn.apply(self.lora_flat, self._lora_layout))
        if self._lora_cache is not None and not self._lora_dirty:
            return self._lora_cache
        with torch.no_grad():
            tensors = _LoraPackFn.apply(self.lora_flat, self._lora_layout)
            if self._lora_cache is None:
                self._lora_cache_tensors = [t.clone() for t in tensors]
                self._lora_cache = self._packs_from(self._lora_cache_tensors)
            else:                               # in place: captured CUDA graphs keep pointing at these buffers
                torch._foreach_copy_(self._lora_cache_tensors, list(tensors))
        self._lora_dirty = False
        return self._lora_cache

    # ------------------------------------------------------------------ full (non-LoRA) fine-tuning, SURVEY.md section 8f-4
    def enable_full_finetune(self):
        """`config.use_lora = False` (train_sd3_fast_pickscore.py:488: no peft wrapper, every transformer parameter is
        trained).  Mixed precision the B200 way: ONE flat fp32 master buffer holds every weight (the optimizer, gradient
        clipping, EMA and the gradient all-reduce work on that one tensor, exactly as they do on the flat LoRA
        parameter), the bf16 working copies the kernels read are refreshed from it in place after every optimizer step
        (so captured CUDA graphs stay valid).  Rollout and replay both run `_forward_full` -- the same kernels and the
        same rounding points in both directions, which `ratio = exp(logp - logp_old) = 1` at `clip_range = 1e-5` needs."""
        if self.full_finetune:
            return self
        names = sorted(self.p)
        self._full_names = names
        self._lora_enabled = False
        self.lora_flat.requires_grad_(False)
        flat = torch.cat([self.p[n].reshape(-1).float() for n in names])
        self.full_master = torch.nn.Parameter(flat.contiguous())
        self.full_master.grad = torch.zeros_like(self.full_master)
        self._full_views, self._full_grad_views = [], []
        off = 0
        for n in names:
            k, shape = self.p[n].numel(), self.p[n].shape
            self._full_views.append(self.full_master.data[off:off + k].view(shape))
            self._full_grad_views.append(self.full_master.grad[off:off + k].view(shape))
            off += k
        self.full_finetune = True
        self._lora_dirty = False
        return self

    def full_state_dict(self):
        """diffusers-named fp32 weights (views of the master buffer)."""
        return dict(zip(self._full_names, self._full_views))

    @torch.no_grad()
    def _refresh_from_master(self):
        """master (fp32) -> bf16 working copies -> the packed operands of the fused no-grad path, all IN PLACE."""
        torch._foreach_copy_([self.p[n] for n in self._full_names], self._full_views)
        p = self.p
        for blk in self.blocks:
            b = f"transformer_blocks.{blk['idx']}"
            for key, names in (("qkv", ("attn.to_q", "attn.to_k", "attn.to_v")),
                               ("cqkv", ("attn.add_q_proj", "attn.add_k_proj", "attn.add_v_proj")),
                               ("qkv2", ("attn2.to_q", "attn2.to_k", "attn2.to_v"))):
                if "w_" + key in blk:
                    torch.cat([p[f"{b}.{n}.weight"] for n in names], 0, out=blk["w_" + key])
                    torch.cat([p[f"{b}.{n}.bias"] for n in names], 0, out=blk["b_" + key])
            for key in [k for k in blk if k.startswith("wt_")]:
                if blk[key].wt is not None:
                    blk[key].wt.copy_(blk[key].w.t())
        L = len(self.blocks)
        ada_w = [p[f"transformer_blocks.{i}.{n}.linear.weight"] for i in range(L) for n in ("norm1", "norm1_context")]
        ada_b = [p[f"transformer_blocks.{i}.{n}.linear.bias"] for i in range(L) for n in ("norm1", "norm1_context")]
        torch.cat(ada_w + [p["norm_out.linear.weight"]], 0, out=self.ada_w)
        torch.cat(ada_b + [p["norm_out.linear.bias"]], 0, out=self.ada_b)
        if self.w_patch.data_ptr() != p["pos_embed.proj.weight"].data_ptr():
            self.w_patch.copy_(p["pos_embed.proj.weight"].reshape(self.d, -1))
        if hasattr(self, "_proj_wt_holder") and self._proj_wt_holder.wt is not None:
            self._proj_wt_holder.wt.copy_(p["proj_out.weight"].t())
        self._lora_dirty = False

    def _forward_full(self, hidden_states, timestep, encoder_hidden_states, pooled_projections, upto=None):
        """The same network as `forward`, written for gradients to EVERY parameter: each Linear is `ops.linear`
        (tcgen05 GEMM forward, dX on the same kernel, dW on the split-K TN kernel, db on the column-sum kernel), the
        feed-forwards are `ops.mlp_gelu` (GELU and its derivative in GEMM epilogues), the joint attention is
        `ops.attention` (tcgen05 forward and backward), LayerNorm-modulate and per-head RMSNorm + concat run on their native
        forward / activation-gradient kernels and add the gradients of their parameters (shift / scale vectors, RMS
        weights) by torch reductions in the backward; only the gated residual adds are torch elementwise ops."""
        if self._lora_dirty:
            self._refresh_from_master()
        cfg, d, bf = self.cfg, self.d, torch.bfloat16
        H, D, ps = cfg["heads"], cfg["head_dim"], cfg["patch_size"]
        if torch.is_grad_enabled() and self.full_master.requires_grad:
            W = dict(zip(self._full_names, _MasterUnpack.apply(self.full_master, self)))
        else:
            W = self.p
        lin = lambda name, v: ops.linear(v, W[name + ".weight"], W.get(name + ".bias"))
        ln_mod = ops.ln_modulate_full                          # native forward / dx; shift / scale gradients by reduction

        def qkv_of(pre, names, v, wt):
            """One GEMM for the three projections (their weights concatenated: autograd splits the gradient back)."""
            wcat = torch.cat([W[f"{pre}.{n}.weight"] for n in names], 0)
            bcat = torch.cat([W[f"{pre}.{n}.bias"] for n in names], 0)
            return ops.linear(v, wcat, bcat, wt=wt)

        def attn(blk, pre, xq, cq=None, ctx_out=True):
            # wt_*: the blocks' lazily built, in-place refreshed transposes (the dX operands), shared with the fused path
            two = pre.endswith("attn2")
            qkv_x = qkv_of(pre, ("to_q", "to_k", "to_v"), xq, blk["wt_qkv2" if two else "wt_qkv"])
            qkv_c = None if cq is None else qkv_of(pre, ("add_q_proj", "add_k_proj", "add_v_proj"), cq, blk["wt_cqkv"])
            nw = lambda n: W.get(f"{pre}.{n}.weight") if cfg["qk_norm"] else None
            if cfg["qk_norm"]:
                joint = ops.qk_norm_concat_full(qkv_x, qkv_c, nw("norm_q"), nw("norm_k"),
                                                None if cq is None else nw("norm_added_q"),
                                                None if cq is None else nw("norm_added_k"), H, D)
            else:
                joint = ops.qk_norm_concat(qkv_x, qkv_c, None, None, None, None, H, D)
            o = ops.attention(joint)                                                              # [B, S, H, D]
            o = o.reshape(o.shape[0], o.shape[1], d)
            lin_t = lambda name, v, key: ops.linear(v, W[name + ".weight"], W.get(name + ".bias"), wt=blk[key])
            if cq is None:
                return lin_t(f"{pre}.to_out.0", o, "wt_out2" if two else "wt_out"), None
            n = xq.shape[1]
            return (lin_t(f"{pre}.to_out.0", o[:, :n], "wt_out"),
                    lin_t(f"{pre}.to_add_out", o[:, n:], "wt_cout") if ctx_out else None)

        def ff(blk, pre, v):
            c_ = "c" if pre.endswith("ff_context") else ""
            return ops.mlp_gelu(v, W[f"{pre}.net.0.proj.weight"], W[f"{pre}.net.0.proj.bias"], W[f"{pre}.net.2.weight"],
                                W[f"{pre}.net.2.bias"], approximate="tanh", w1t=blk[f"wt_{c_}ff1"], w2t=blk[f"wt_{c_}ff2"])

        gated = lambda res, gate, branch: torch.addcmul(res, gate[:, None], branch)               # res + gate * branch

        B, C, Hh, Ww = hidden_states.shape
        h, w = Hh // ps, Ww // ps
        xp = hidden_states.to(bf).reshape(B, C, h, ps, w, ps).permute(0, 2, 4, 1, 3, 5).reshape(B * h * w, C * ps * ps)
        key = (h, w)
        if key not in self._pos_cache:
            self._pos_cache[key] = cropped_pos_embed(d, h, w, cfg["pos_embed_max_size"], cfg["base_size"], self.device_).to(bf)
        x = ops.linear(xp.contiguous(), W["pos_embed.proj.weight"].reshape(d, -1), W["pos_embed.proj.bias"]).reshape(B, h * w, d)
        x = x + self._pos_cache[key]
        te = timestep_embedding(timestep.to(self.device_)).to(bf)
        temb = lin("time_text_embed.timestep_embedder.linear_2", F.silu(lin("time_text_embed.timestep_embedder.linear_1", te)))
        temb = temb + lin("time_text_embed.text_embedder.linear_2",
                          F.silu(lin("time_text_embed.text_embedder.linear_1", pooled_projections.to(bf))))
        st = F.silu(temb)
        c = lin("context_embedder", encoder_hidden_states.to(bf).contiguous())
        L = cfg["num_layers"]
        for i in range(L if upto is None else upto):
            pre, last, dual = f"transformer_blocks.{i}", i == L - 1, i in cfg["dual_layers"]
            blk = self.blocks[i]
            e = lin(f"{pre}.norm1.linear", st)
            if dual:
                sh, sc, g, sh_m, sc_m, g_m, sh2, sc2, g2 = e.chunk(9, dim=1)
            else:
                sh, sc, g, sh_m, sc_m, g_m = e.chunk(6, dim=1)
            x1 = ln_mod(x, sh, sc)
            ce = lin(f"{pre}.norm1_context.linear", st)
            if last:                                           # AdaLayerNormContinuous: (scale, shift)
                csc, csh = ce.chunk(2, dim=1)
            else:
                csh, csc, cg, csh_m, csc_m, cg_m = ce.chunk(6, dim=1)
            c1 = ln_mod(c, csh, csc)
            a, ca = attn(blk, f"{pre}.attn", x1, c1, ctx_out=not last)
            x_in = x
            x = gated(x, g, a)
            if dual:
                a2, _ = attn(blk, f"{pre}.attn2", ln_mod(x_in, sh2, sc2))
                x = gated(x, g2, a2)
            x = gated(x, g_m, ff(blk, f"{pre}.ff", ln_mod(x, sh_m, sc_m)))
            if last:
                c = None
                continue
            c = gated(c, cg, ca)
            c = gated(c, cg_m, ff(blk, f"{pre}.ff_context", ln_mod(c, csh_m, csc_m)))
        if upto is not None:
            return x, c
        sc, sh = lin("norm_out.linear", st).chunk(2, dim=1)
        x = lin("proj_out", ln_mod(x, sh, sc))
        cin = cfg["in_channels"]
        x = x.reshape(B, h, w, ps, ps, cin).permute(0, 5, 1, 3, 2, 4).reshape(B, cin, h * ps, w * ps)
        return (x,)

    # ------------------------------------------------------------------ building blocks
    def _lin(self, x, blk, key, lora_pack=None, epilogue=ops.EPI_NONE, residual=None, gate=None, rows=1):
        a = w2 = None
        if lora_pack is not None and key in lora_pack:
            a, w2 = lora_pack[key]
        return _LinearFn.apply(x, blk["w_" + key], blk["wt_" + key], blk["b_" + key], a, w2, epilogue, residual,
                               gate, rows)

    def _lin2(self, x, c, blk, key, ckey, lora_pack=None, epilogue=ops.EPI_NONE, res=(None, None), gate=(None, None),
              rows=(1, 1)):
        """Image-stream and text-stream projection `key` / `ckey` of one block in a single dual-problem launch."""
        a0 = w20 = a1 = w21 = None
        if lora_pack is not None and key in lora_pack and ckey in lora_pack:
            (a0, w20), (a1, w21) = lora_pack[key], lora_pack[ckey]
        return _DualLinearFn.apply(x, c, blk["w_" + key], blk["w_" + ckey], blk["wt_" + key], blk["wt_" + ckey],
                                   blk["b_" + key], blk["b_" + ckey], a0, a1, w20, w21, epilogue, res[0], res[1],
                                   gate[0], gate[1], rows[0], rows[1])

    def _qkv_norm(self, blk, x1, c1, pk):
        """no-grad path: fused QKV projection (+LoRA) + q/k RMSNorm + concat, one launch for both streams."""
        H, D = self.cfg["heads"], self.cfg["head_dim"]
        a2 = w2 = (None, None)
        if pk is not None and "qkv" in pk and "cqkv" in pk:
            (a0, w20), (a1, w21) = pk["qkv"], pk["cqkv"]
            a2, w2 = ops.gemm_dual((x1, c1), (a0, a1)), (w20, w21)
        return ops.gemm_qkv_norm(x1, c1, (blk["w_qkv"], blk["w_cqkv"]), (blk["b_qkv"], blk["b_cqkv"]),
                                 (blk.get("nq"), blk.get("ncq")), (blk.get("nk"), blk.get("nck")), H, D, a2=a2, w2=w2)

    def _attention(self, joint, split):
        if torch.is_grad_enabled() and joint.requires_grad:
            if split:
                oi, ot = ops.attention_split(joint, split)
                return oi.reshape(oi.shape[0], -1, self.d), ot.reshape(ot.shape[0], -1, self.d)
            o = ops.attention(joint)                               # [B, S, H, D]
            return o.reshape(o.shape[0], o.shape[1], self.d), None
        o, _ = ops.attention_fwd(joint, want_lse=False, split=split)
        if split:
            return o[0].reshape(o[0].shape[0], -1, self.d), o[1].reshape(o[1].shape[0], -1, self.d)
        return o.reshape(o.shape[0], -1, self.d), None

    def _block(self, blk, x, c, emb, pk):
        d, H, D = self.d, self.cfg["heads"], self.cfg["head_dim"]
        B, N, _ = x.shape
        Nc = c.shape[1]
        o0, o1 = self.ada_off[blk["idx"]]
        ex = emb[:, o0:o1]
        ec = emb[:, o1:o1 + (2 if blk["last"] else 6) * d]
        ch = lambda e, k: e[:, k * d:(k + 1) * d]
        # image stream: shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp (, shift2, scale2, gate2)
        if blk["dual"]:
            x1, x2 = ops.ln_modulate(x, ch(ex, 0), ch(ex, 1), ch(ex, 6), ch(ex, 7))
        else:
            x1 = ops.ln_modulate(x, ch(ex, 0), ch(ex, 1))
        if blk["last"]:
            c1 = ops.ln_modulate(c, ch(ec, 1), ch(ec, 0))           # AdaLayerNormContinuous: (scale, shift)
        else:
            c1 = ops.ln_modulate(c, ch(ec, 0), ch(ec, 1))
        fused_qkv = self.fused_qkv_norm and not torch.is_grad_enabled() and D == 64
        if fused_qkv:
            joint = self._qkv_norm(blk, x1, c1, pk)
        else:
            if self.dual_gemm:
                qkv_x, qkv_c = self._lin2(x1, c1, blk, "qkv", "cqkv", pk)
            else:
                qkv_x = self._lin(x1, blk, "qkv", pk)
                qkv_c = self._lin(c1, blk, "cqkv", pk)
            joint = ops.qk_norm_concat(qkv_x, qkv_c, blk.get("nq"), blk.get("nk"), blk.get("ncq"), blk.get("nck"), H, D)
        ox, oc = self._attention(joint, N)
        fused_out = self.dual_gemm and not blk["last"]
        if fused_out:
            x, c = self._lin2(ox, oc, blk, "out", "cout", pk, ops.EPI_GATE_RESIDUAL, (x, c), (ch(ex, 2), ch(ec, 2)), (N, Nc))
        else:
            x = self._lin(ox, blk, "out", pk, ops.EPI_GATE_RESIDUAL, x, ch(ex, 2), N)
        if blk["dual"] and fused_qkv:
            j2 = ops.gemm_qkv_norm(x2, None, (blk["w_qkv2"],), (blk["b_qkv2"],), (blk.get("nq2"),), (blk.get("nk2"),), H, D)
            o2, _ = self._attention(j2, 0)
            x = self._lin(o2, blk, "out2", None, ops.EPI_GATE_RESIDUAL, x, ch(ex, 8), N)
        elif blk["dual"]:
            qkv2 = self._lin(x2, blk, "qkv2")
            j2 = ops.qk_norm_concat(qkv2, None, blk.get("nq2"), blk.get("nk2"), None, None, H, D)
            o2, _ = self._attention(j2, 0)
            x = self._lin(o2, blk, "out2", None, ops.EPI_GATE_RESIDUAL, x, ch(ex, 8), N)
        xm = ops.ln_modulate(x, ch(ex, 3), ch(ex, 4))
        if blk["last"]:
            hmid = self._lin(xm, blk, "ff1", None, ops.EPI_GELU_TANH)
            x = self._lin(hmid, blk, "ff2", None, ops.EPI_GATE_RESIDUAL, x, ch(ex, 5), N)
            return x, None
        if not fused_out:
            c = self._lin(oc, blk, "cout", pk, ops.EPI_GATE_RESIDUAL, c, ch(ec, 2), Nc)
        cm = ops.ln_modulate(c, ch(ec, 3), ch(ec, 4))
        if self.dual_gemm and self.fused_ff:
            x, c = _DualFeedForwardFn.apply(xm, cm, blk["w_ff1"], blk["w_cff1"], blk["wt_ff1"], blk["wt_cff1"], blk["b_ff1"],
                                            blk["b_cff1"], blk["w_ff2"], blk["w_cff2"], blk["wt_ff2"], blk["wt_cff2"],
                                            blk["b_ff2"], blk["b_cff2"], x, c, ch(ex, 5), ch(ec, 5), N, Nc)
        elif self.dual_gemm:
            hx, hc = self._lin2(xm, cm, blk, "ff1", "cff1", None, ops.EPI_GELU_TANH)
            x, c = self._lin2(hx, hc, blk, "ff2", "cff2", None, ops.EPI_GATE_RESIDUAL, (x, c), (ch(ex, 5), ch(ec, 5)), (N, Nc))
        else:
            hmid = self._lin(xm, blk, "ff1", None, ops.EPI_GELU_TANH)
            x = self._lin(hmid, blk, "ff2", None, ops.EPI_GATE_RESIDUAL, x, ch(ex, 5), N)
            hmid = self._lin(cm, blk, "cff1", None, ops.EPI_GELU_TANH)
            c = self._lin(hmid, blk, "cff2", None, ops.EPI_GATE_RESIDUAL, c, ch(ec, 5), Nc)
        return x, c

    # ------------------------------------------------------------------ forward
    def forward(self, hidden_states, timestep, encoder_hidden_states, pooled_projections,
                joint_attention_kwargs=None, return_dict=False, upto=None):
        if self.full_finetune:
            return self._forward_full(hidden_states, timestep, encoder_hidden_states, pooled_projections, upto=upto)
        cfg, p, d = self.cfg, self.p, self.d
        ps = cfg["patch_size"]
        B, C, Hh, Ww = hidden_states.shape
        h, w = Hh // ps, Ww // ps
        bf = torch.bfloat16
        # patch embed = GEMM over unfolded 2x2 patches (+ fixed sincos positions)
        xp = hidden_states.to(bf).reshape(B, C, h, ps, w, ps).permute(0, 2, 4, 1, 3, 5).reshape(B * h * w, C * ps * ps)
        key = (h, w)
        if key not in self._pos_cache:
            self._pos_cache[key] = cropped_pos_embed(d, h, w, cfg["pos_embed_max_size"], cfg["base_size"],
                                                     self.device_).to(bf)
        x = ops.gemm(xp.contiguous(), self.w_patch, bias=p["pos_embed.proj.bias"]).reshape(B, h * w, d)
        x = x + self._pos_cache[key]
        # conditioning vector and every block's adaLN modulation in one GEMM
        te = timestep_embedding(timestep.to(self.device_)).to(bf)
        lin = lambda n, v: ops.gemm(v.contiguous(), p[n + ".weight"], bias=p[n + ".bias"])   # B-row conditioning MLPs
        temb = lin("time_text_embed.timestep_embedder.linear_2", F.silu(lin("time_text_embed.timestep_embedder.linear_1", te)))
        temb = temb + lin("time_text_embed.text_embedder.linear_2",
                          F.silu(lin("time_text_embed.text_embedder.linear_1", pooled_projections.to(bf))))
        emb = ops.gemm(F.silu(temb).contiguous(), self.ada_w, bias=self.ada_b)            # [B, sum]
        c = ops.gemm(encoder_hidden_states.to(bf).contiguous(), p["context_embedder.weight"],
                     bias=p["context_embedder.bias"])
        pk = self._pack_lora()
        nblk = len(self.blocks) if upto is None else upto
        for blk, lp in zip(self.blocks[:nblk], pk):
            x, c = self._block(blk, x, c, emb, lp)
        if upto is not None:
            return x, c
        eo = emb[:, self.ada_out_off:self.ada_out_off + 2 * d]
        x = ops.ln_modulate(x, eo[:, d:], eo[:, :d])                                       # (scale, shift) order
        x = _LinearFn.apply(x, p["proj_out.weight"], self._proj_wt(), p["proj_out.bias"], None, None, ops.EPI_NONE,
                            None, None, 1)
        cin = cfg["in_channels"]
        x = x.reshape(B, h, w, ps, ps, cin).permute(0, 5, 1, 3, 2, 4).reshape(B, cin, h * ps, w * ps)
        return (x,)

    def _proj_wt(self):
        if not hasattr(self, "_proj_wt_holder"):
            self._proj_wt_holder = _WT(self.p["proj_out.weight"])
        return self._proj_wt_holder

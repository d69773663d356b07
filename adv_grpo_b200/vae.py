"""SD3 VAE decoder (`pipeline.vae` of the reference: `fast.py:667-669`, kept in fp32 like
`train_sd3_fast_pickscore.py:481`) + the `VaeImageProcessor.postprocess(output_type='pt')` step.

Channels-last fp32 end to end.  EVERY convolution runs on the tcgen05 TF32 implicit-GEMM kernel of csrc/conv.cu (TF32
tensor cores with fp32 accumulation, like the reference's fp32 VAE under cudnn.allow_tf32); conv_in (16 input channels)
and conv_out (3 output channels) are zero-padded to the kernel's 32-channel granule (conv_out on 32-column tiles).
Everything between the convolutions is fused into our streaming kernels so that no tensor makes an extra HBM round trip:
  * GroupNorm + SiLU in two passes, with the PRECEDING convolution's bias folded into the statistics/apply
    (the convolutions run bias-free: no separate bias pass, no NCHW<->NHWC copies around group_norm);
  * residual add + conv2 bias (+ shortcut bias) in one pass;
  * nearest 2x upsampling as one vectorised pass; GroupNorm-apply and upsample hand their outputs over as TF32 values
    (round to nearest) so the tensor core's operand truncation is exact;
  * the single-head mid-block attention on the same kernel: the q / k / v / out projections are 1x1 convolutions, the
    scores Q K^T are a 1x1 convolution whose "weights" are the sample's K matrix, V^T comes out of a 1x1 convolution of
    the V weight matrix against the tokens, softmax is one fp32 row kernel (csrc/heads.cu) that hands P over as TF32
    values, and P V is a 1x1 convolution with V^T as the weights -- no library GEMM, no fp32 SIMT fmha.
diffusers state-dict names."""
import torch

from . import ops
from .weights import VAE_SD3


class AutoencoderKL:
    def __init__(self, params, cfg=VAE_SD3, device="cuda", dtype=torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        self.config = type("Config", (), dict(scaling_factor=1.5305, shift_factor=0.0609))()
        self.p = {}
        for k, v in params.items():
            v = v.to(device=device, dtype=dtype)
            if v.dim() == 4:
                v = v.contiguous(memory_format=torch.channels_last)
            self.p[k] = v
        if dtype != torch.float32:
            raise ValueError(f"AutoencoderKL is an fp32 model (train_pick:481 keeps the VAE in float32), got {dtype}")
        self.fused = True            # every op below is a native kernel; decode() on a non-CUDA tensor raises (ops._need_cuda)
        self.native_conv = True       # 3x3 / 1x1 convolutions on the tcgen05 TF32 implicit-GEMM kernel (csrc/conv.cu)
        self.max_latent_pixels = 8 * 64 * 64    # decode() splits larger batches (8 images at 512 px, 2 at 1024 px per call)

    def to(self, *a, **k):
        return self

    def requires_grad_(self, flag=True):
        return self

    def eval(self):
        return self

    def state_dict(self):
        return {k: v for k, v in self.p.items() if not k.endswith((".packed_tf32", ".fused_bias"))}

    # ---- building blocks -------------------------------------------------------------------------------
    def _fusable(self, C):
        return self.fused and C % 128 == 0 and 256 % (C // 4) == 0

    def _gn(self, name, x, silu=False, in_bias=None):
        C = x.shape[1]
        if not self._fusable(C):
            raise ValueError(f"GroupNorm over {C} channels is not supported by gn_stats / gn_apply (channels must be a "
                             "multiple of 128 with 256 % (C / 4) == 0: 128, 256, 512, 1024); no PyTorch fallback")
        return ops.group_norm_silu_nhwc(x, self.p[name + ".weight"], self.p[name + ".bias"], 32, 1e-6, silu,
                                        in_bias=in_bias, round_tf32=self.native_conv)

    def _conv(self, name, x, pad=1, bias=True):
        w = self.p[name + ".weight"]
        Cout, Cin, k, _ = w.shape
        if not (self.native_conv and self.fused and k in (1, 3) and pad == k // 2):
            raise ValueError(f"convolution {name} ({Cin} -> {Cout}, k={k}, pad={pad}) is not supported by conv_tf32_kernel "
                             "(k in {1, 3}, pad = k // 2); there is no PyTorch fallback")
        cin_p, cout_p = -(-Cin // 32) * 32, -(-Cout // 32) * 32
        key = name + ".packed_tf32"
        if key not in self.p:
            if (cin_p, cout_p) != (Cin, Cout):          # conv_in: 16 -> 32 zero input channels; conv_out: 3 -> 32 zero filters
                wp = torch.zeros(cout_p, cin_p, k, k, dtype=w.dtype, device=w.device)
                wp[:Cout, :Cin] = w
                w = wp
                bp = torch.zeros(cout_p, dtype=w.dtype, device=w.device)
                bp[:Cout] = self.p[name + ".bias"]
                self.p[name + ".fused_bias"] = bp
            self.p[key] = ops.pack_conv_weight_tf32(w)
        if cin_p != Cin:
            xp = torch.empty((x.shape[0], cin_p, x.shape[2], x.shape[3]), dtype=x.dtype, device=x.device,
                             memory_format=torch.channels_last).zero_()
            xp[:, :Cin] = x
            x = xp
        b = None
        if bias:
            b = self.p[name + ".fused_bias"] if cout_p != Cout or cin_p != Cin else self.p[name + ".bias"]
        y = ops.conv2d_nhwc_tf32(x, self.p[key], b, k)
        return y if cout_p == Cout else y[:, :Cout]

    def _resnet(self, pre, x):
        p = self.p
        C_out = p[pre + ".conv1.weight"].shape[0]
        fuse = self._fusable(C_out)
        h = self._conv(pre + ".conv1", self._gn(pre + ".norm1", x, silu=True), bias=not fuse)
        h = self._gn(pre + ".norm2", h, silu=True, in_bias=p[pre + ".conv1.bias"] if fuse else None)
        h = self._conv(pre + ".conv2", h, bias=not fuse)
        has_sc = pre + ".conv_shortcut.weight" in p
        if has_sc:
            x = self._conv(pre + ".conv_shortcut", x, pad=0, bias=not fuse)
        if not fuse:
            return x + h
        bias = p[pre + ".conv2.bias"]
        if has_sc:
            key = pre + ".fused_bias"
            if key not in p:
                p[key] = (p[pre + ".conv2.bias"] + p[pre + ".conv_shortcut.bias"]).contiguous()
            bias = p[key]
        return ops.add_bias_nhwc(x, h, bias)

    def _lin_w(self, name):
        """nn.Linear weight [Cout, Cin] as a packed 1x1-convolution weight (TF32-rounded)."""
        key = name + ".packed_tf32"
        if key not in self.p:
            w = self.p[name + ".weight"]
            self.p[key] = ops.pack_conv_weight_tf32(w.reshape(w.shape[0], w.shape[1], 1, 1))
        return self.p[key]

    def _mid_attn(self, pre, x):
        p = self.p
        B, C, H, W = x.shape
        S = H * W
        h = self._gn(pre + ".group_norm", x)                                   # TF32-rounded tokens, NHWC = [B, S, C]
        q = ops.conv2d_nhwc_tf32(h, self._lin_w(pre + ".to_q"), p[pre + ".to_q.bias"], 1)
        k = ops.conv2d_nhwc_tf32(h, self._lin_w(pre + ".to_k"), p[pre + ".to_k.bias"], 1)
        o = torch.empty_like(q, memory_format=torch.channels_last)
        o_rows = o.permute(0, 2, 3, 1)                                         # [B, H, W, C] view of the NHWC storage
        k_rows = k.permute(0, 2, 3, 1).reshape(B, S, C)
        h_rows = h.permute(0, 2, 3, 1).reshape(B, S, C)
        # the V weight matrix as a "C-pixel image" [1, C, C / cw, cw]: a 1x1 convolution of it against the tokens
        # (weights = h[b], [S, C]) yields V^T [C, S] directly, the K-major operand the P V product needs
        cw = min(C, 128)
        wv_img = self._lin_w(pre + ".to_v").view(1, C // cw, cw, C).permute(0, 3, 1, 2)
        for b in range(B):                                                     # single head, per sample: K / V are weights
            s = ops.conv2d_nhwc_tf32(q[b:b + 1], k_rows[b], None, 1)           # [1, S(keys), H, W]: scores[query, key]
            s_rows = s.permute(0, 2, 3, 1).reshape(S, S)
            ops.row_softmax_f32(s_rows, scale=C ** -0.5, round_tf32=True, out=s_rows)
            vt = ops.conv2d_nhwc_tf32(wv_img, h_rows[b], None, 1)              # [1, S, C / cw, cw] = V^T (bias-free)
            vt = vt.permute(0, 2, 3, 1).reshape(C, S)
            # rows of P sum to one, so the V bias passes through the attention unchanged: it is this product's bias
            ob = ops.conv2d_nhwc_tf32(s, vt, p[pre + ".to_v.bias"], 1)         # [1, C, H, W]
            o_rows[b].copy_(ob.permute(0, 2, 3, 1)[0])
        o = ops.conv2d_nhwc_tf32(o, self._lin_w(pre + ".to_out.0"), p[pre + ".to_out.0.bias"], 1)
        return ops.add_bias_nhwc(x, o, None)

    def _upsample(self, x):
        return ops.upsample_nearest2x_nhwc(x)

    @torch.no_grad()
    def decode(self, z, return_dict=False):
        B, _, h, w = z.shape
        per_call = max(1, self.max_latent_pixels // (h * w))
        if B > per_call:                               # bound the activation footprint (fp32 NHWC maps) of large batches
            return (torch.cat([self._decode(z[i:i + per_call]) for i in range(0, B, per_call)]),)
        return (self._decode(z),)

    def _decode(self, z):
        ops._need_cuda(z)
        x = z.to(self.dtype).contiguous(memory_format=torch.channels_last)
        x = self._conv("decoder.conv_in", x)
        x = self._resnet("decoder.mid_block.resnets.0", x)
        x = self._mid_attn("decoder.mid_block.attentions.0", x)
        x = self._resnet("decoder.mid_block.resnets.1", x)
        n_up = len(self.cfg["block_out"])
        for i in range(n_up):
            for j in range(self.cfg["layers_per_block"] + 1):
                x = self._resnet(f"decoder.up_blocks.{i}.resnets.{j}", x)
            if i < n_up - 1:
                x = self._conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", self._upsample(x))
        x = self._gn("decoder.conv_norm_out", x, silu=True)
        return self._conv("decoder.conv_out", x)


class VaeImageProcessor:
    """Only what the hot path uses: postprocess(image, output_type='pt') = denormalise to [0,1]."""

    def postprocess(self, image, output_type="pt"):
        if output_type != "pt":
            raise NotImplementedError("only output_type='pt' is on the hot path (fast.py:670)")
        return (image / 2 + 0.5).clamp(0, 1)

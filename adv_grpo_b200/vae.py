"""SD3 VAE decoder (`pipeline.vae` of the reference: `fast.py:667-669`, kept in fp32 like
`train_sd3_fast_pickscore.py:481`) + the `VaeImageProcessor.postprocess(output_type='pt')` step.
The convolutions are plain library calls (cuDNN, TF32 like the reference's allow_tf32) in
channels_last; SURVEY.md section 8(f) ranks a tcgen05 implicit-GEMM decoder as "next".
diffusers state-dict names."""
import torch
import torch.nn.functional as F

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

    def to(self, *a, **k):
        return self

    def _gn(self, name, x, silu=False):
        C = x.shape[1]
        if x.is_cuda and x.dtype == torch.float32 and C % 128 == 0 and 256 % (C // 4) == 0:
            return ops.group_norm_silu_nhwc(x, self.p[name + ".weight"], self.p[name + ".bias"], 32, 1e-6, silu)
        y = F.group_norm(x, 32, self.p[name + ".weight"], self.p[name + ".bias"], eps=1e-6)   # tiny test configs
        return F.silu(y) if silu else y

    def _conv(self, name, x, pad=1):
        return F.conv2d(x, self.p[name + ".weight"], self.p[name + ".bias"], padding=pad)

    def _resnet(self, pre, x):
        h = self._conv(pre + ".conv1", self._gn(pre + ".norm1", x, silu=True))
        h = self._conv(pre + ".conv2", self._gn(pre + ".norm2", h, silu=True))
        if pre + ".conv_shortcut.weight" in self.p:
            x = self._conv(pre + ".conv_shortcut", x, pad=0)
        return x + h

    def _mid_attn(self, pre, x):
        p = self.p
        B, C, H, W = x.shape
        h = self._gn(pre + ".group_norm", x).reshape(B, C, H * W).transpose(1, 2)
        q, k, v = (F.linear(h, p[f"{pre}.to_{n}.weight"], p[f"{pre}.to_{n}.bias"]) for n in "qkv")
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = F.linear(o, p[pre + ".to_out.0.weight"], p[pre + ".to_out.0.bias"])
        return x + o.transpose(1, 2).reshape(B, C, H, W)

    @torch.no_grad()
    def decode(self, z, return_dict=False):
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
                x = F.interpolate(x, scale_factor=2.0, mode="nearest")
                x = self._conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", x)
        x = self._gn("decoder.conv_norm_out", x, silu=True)
        return (self._conv("decoder.conv_out", x),)


class VaeImageProcessor:
    """Only what the hot path uses: postprocess(image, output_type='pt') = denormalise to [0,1]."""

    def postprocess(self, image, output_type="pt"):
        if output_type != "pt":
            raise NotImplementedError("only output_type='pt' is on the hot path (fast.py:670)")
        return (image / 2 + 0.5).clamp(0, 1)

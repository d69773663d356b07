"""Oracle: SD3 VAE decoder + postprocess (`fast.py:667-670`: latents/scaling+shift ->
`AutoencoderKL.decode` in fp32 -> `VaeImageProcessor.postprocess(output_type='pt')`),
restated from the diffusers 0.33.1 `AutoencoderKL` decoder with the SD3 config
(latent_channels 16, block_out_channels [128,256,512,512], layers_per_block 2,
norm_num_groups 32, no post_quant_conv, scaling 1.5305, shift 0.0609).
diffusers is absent here: **parity unpinned**; topology cross-read against
torchtitan/experiments/flux/model/autoencoder.py (same decoder family).
Test infrastructure only (see oracle/__init__.py)."""
import torch
import torch.nn.functional as F

SCALING_FACTOR = 1.5305
SHIFT_FACTOR = 0.0609


def _gn(p, name, x):
    return F.group_norm(x, 32, p[name + ".weight"], p[name + ".bias"], eps=1e-6)


def _conv(p, name, x, pad=1):
    return F.conv2d(x, p[name + ".weight"], p[name + ".bias"], padding=pad)


def _resnet(p, pre, x):
    h = _conv(p, pre + ".conv1", F.silu(_gn(p, pre + ".norm1", x)))
    h = _conv(p, pre + ".conv2", F.silu(_gn(p, pre + ".norm2", h)))
    if pre + ".conv_shortcut.weight" in p:
        x = _conv(p, pre + ".conv_shortcut", x, pad=0)
    return x + h


def _mid_attn(p, pre, x):
    B, C, H, W = x.shape
    h = _gn(p, pre + ".group_norm", x).view(B, C, H * W).transpose(1, 2)
    q = F.linear(h, p[pre + ".to_q.weight"], p[pre + ".to_q.bias"])
    k = F.linear(h, p[pre + ".to_k.weight"], p[pre + ".to_k.bias"])
    v = F.linear(h, p[pre + ".to_v.weight"], p[pre + ".to_v.bias"])
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
    o = F.linear(o, p[pre + ".to_out.0.weight"], p[pre + ".to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(B, C, H, W)


def vae_decode(p, z):
    """p: diffusers-named decoder params ("decoder.*"), z: [B,16,h,w] fp32 (already de-scaled)."""
    p = {k: v.float() for k, v in p.items()}
    x = _conv(p, "decoder.conv_in", z.float())
    x = _resnet(p, "decoder.mid_block.resnets.0", x)
    x = _mid_attn(p, "decoder.mid_block.attentions.0", x)
    x = _resnet(p, "decoder.mid_block.resnets.1", x)
    for i in range(4):
        for j in range(3):
            x = _resnet(p, f"decoder.up_blocks.{i}.resnets.{j}", x)
        if i < 3:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = _conv(p, f"decoder.up_blocks.{i}.upsamplers.0.conv", x)
    x = F.silu(_gn(p, "decoder.conv_norm_out", x))
    return _conv(p, "decoder.conv_out", x)


def decode_latents_to_image(p, latents):
    """`latents / scaling_factor + shift_factor` -> vae.decode -> postprocess('pt') (fast.py:667-670)."""
    z = latents.float() / SCALING_FACTOR + SHIFT_FACTOR          # fast.py:667
    img = vae_decode(p, z)                                       # fast.py:669
    return (img / 2 + 0.5).clamp(0, 1)                           # postprocess('pt') -> denormalize

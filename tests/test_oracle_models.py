"""Cross-check the oracle's restated model bodies against same-architecture implementations that ARE
installed here (transformers CLIPModel / Dinov2Model, Pillow) -- SURVEY.md section 8c.  Tiny configs, CPU."""
import numpy as np
import torch

from oracle import clip as clip_o
from oracle import dinov2 as dino_o
from oracle import preprocess as pre_o


def test_clip_oracle_matches_transformers():
    from transformers import CLIPConfig, CLIPModel
    cfg = CLIPConfig(
        text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                         vocab_size=300, max_position_embeddings=77, hidden_act="gelu", eos_token_id=2, bos_token_id=0,
                         pad_token_id=1, projection_dim=32),
        vision_config=dict(hidden_size=96, intermediate_size=192, num_hidden_layers=2, num_attention_heads=4,
                           image_size=56, patch_size=14, hidden_act="gelu", projection_dim=32),
        projection_dim=32)
    torch.manual_seed(0)
    m = CLIPModel(cfg).eval()
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ocfg = dict(patch=14, v_layers=2, v_heads=4, t_layers=2, t_heads=4)
    pix = torch.randn(3, 3, 56, 56)
    ids = torch.randint(3, 299, (3, 20))
    ids[:, -1] = 299                                     # highest id last = EOS position for argmax pooling
    with torch.no_grad():
        ref_i = m.get_image_features(pixel_values=pix)
        ref_t = m.get_text_features(input_ids=ids)
        ref_i = getattr(ref_i, "pooler_output", ref_i)   # transformers 5.x returns an output object
        ref_t = getattr(ref_t, "pooler_output", ref_t)
        got_i = clip_o.image_features(p, ocfg, pix)
        got_t = clip_o.text_features(p, ocfg, ids)
    assert torch.allclose(got_i, ref_i, atol=2e-5, rtol=1e-4)
    assert torch.allclose(got_t, ref_t, atol=2e-5, rtol=1e-4)


def test_dinov2_oracle_matches_transformers():
    from transformers import Dinov2Config, Dinov2Model
    cfg = Dinov2Config(hidden_size=64, num_hidden_layers=2, num_attention_heads=4, mlp_ratio=2, image_size=56,
                       patch_size=14, layerscale_value=0.3, hidden_act="gelu", qkv_bias=True, layer_norm_eps=1e-6)
    torch.manual_seed(0)
    m = Dinov2Model(cfg).eval()
    sd = m.state_dict()
    p = {"patch_embed.proj.weight": sd["embeddings.patch_embeddings.projection.weight"],
         "patch_embed.proj.bias": sd["embeddings.patch_embeddings.projection.bias"],
         "cls_token": sd["embeddings.cls_token"], "pos_embed": sd["embeddings.position_embeddings"],
         "norm.weight": sd["layernorm.weight"], "norm.bias": sd["layernorm.bias"]}
    for i in range(2):
        h, b = f"encoder.layer.{i}", f"blocks.{i}"
        p[b + ".norm1.weight"], p[b + ".norm1.bias"] = sd[h + ".norm1.weight"], sd[h + ".norm1.bias"]
        p[b + ".norm2.weight"], p[b + ".norm2.bias"] = sd[h + ".norm2.weight"], sd[h + ".norm2.bias"]
        p[b + ".attn.qkv.weight"] = torch.cat([sd[f"{h}.attention.attention.{n}.weight"] for n in ("query", "key", "value")])
        p[b + ".attn.qkv.bias"] = torch.cat([sd[f"{h}.attention.attention.{n}.bias"] for n in ("query", "key", "value")])
        p[b + ".attn.proj.weight"], p[b + ".attn.proj.bias"] = sd[h + ".attention.output.dense.weight"], sd[h + ".attention.output.dense.bias"]
        p[b + ".ls1.gamma"], p[b + ".ls2.gamma"] = sd[h + ".layer_scale1.lambda1"], sd[h + ".layer_scale2.lambda1"]
        p[b + ".mlp.fc1.weight"], p[b + ".mlp.fc1.bias"] = sd[h + ".mlp.fc1.weight"], sd[h + ".mlp.fc1.bias"]
        p[b + ".mlp.fc2.weight"], p[b + ".mlp.fc2.bias"] = sd[h + ".mlp.fc2.weight"], sd[h + ".mlp.fc2.bias"]
    x = torch.randn(2, 3, 56, 56)
    with torch.no_grad():
        ref = m(pixel_values=x).last_hidden_state
        got = dino_o.forward_features(p, dict(patch=14, heads=4, layers=2), x)
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-4)


def test_pil_resize_restatement_is_bit_exact_with_pillow():
    from PIL import Image
    rng = np.random.RandomState(0)
    for H, out in ((512, 224), (300, 224), (128, 224)):
        img = rng.randint(0, 256, size=(H, H, 3), dtype=np.uint8)
        ref = np.array(Image.fromarray(img).resize((out, out), resample=Image.BICUBIC))
        got = pre_o.pil_bicubic_resize_u8(img.transpose(2, 0, 1), out).transpose(1, 2, 0)
        assert np.array_equal(ref, got), (H, out)


def test_dino_head_reward_and_hinge_loss_shapes():
    torch.manual_seed(0)
    hp = {"layers.0.weight": torch.randn(16, 32) * 0.1, "layers.0.bias": torch.zeros(16),
          "layers.2.weight": torch.randn(1, 16) * 0.1, "layers.2.bias": torch.zeros(1)}
    feats = torch.randn(3, 10, 32)
    idx = torch.randint(0, 9, (3, 4))
    hybrid, cls_s, patch_s = dino_o.patch_reward(hp, feats, idx)
    assert hybrid.shape == (3,) and patch_s.shape == (3, 4)
    assert torch.allclose(hybrid, 0.7 * cls_s + 0.3 * patch_s.mean(1))
    loss, acc = dino_o.hinge_d_loss(hp, feats, feats.flip(0), idx, idx)
    assert loss.dim() == 0 and 0 <= acc <= 1


def test_vae_decoder_oracle_matches_flux_autoencoder():
    """torchtitan's FLUX autoencoder (installed here) is the same LDM decoder family as
    the SD3 VAE (GroupNorm(32, eps 1e-6) + swish resnets, single-head mid attention, nearest-2x upsample).
    Map the oracle's diffusers-named weights onto it (the standard LDM <-> diffusers key conversion) and
    compare the decoder outputs."""
    import pytest
    ae = pytest.importorskip("torchtitan.experiments.flux.model.autoencoder")
    from adv_grpo_b200 import weights
    from oracle import vae as vae_o
    cfg = dict(latent_channels=16, block_out=(32, 64, 128, 128), layers_per_block=2)
    p = weights.init_vae_decoder(cfg, seed=3, device="cpu")
    dec = ae.Decoder(ch=32, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2, in_channels=3, resolution=64,
                     z_channels=16).eval()
    sd = {}

    def put(dst, src, conv1x1=False):
        for s in ("weight", "bias"):
            v = p[f"decoder.{src}.{s}"]
            sd[f"{dst}.{s}"] = v[:, :, None, None] if (conv1x1 and s == "weight" and v.dim() == 2) else v

    def resnet(dst, src):
        for n in ("norm1", "conv1", "norm2", "conv2"):
            put(f"{dst}.{n}", f"{src}.{n}")
        if f"decoder.{src}.conv_shortcut.weight" in p:
            put(f"{dst}.nin_shortcut", f"{src}.conv_shortcut")

    put("conv_in", "conv_in")
    resnet("mid.block_1", "mid_block.resnets.0")
    resnet("mid.block_2", "mid_block.resnets.1")
    put("mid.attn_1.norm", "mid_block.attentions.0.group_norm")
    for a, b in (("q", "to_q"), ("k", "to_k"), ("v", "to_v"), ("proj_out", "to_out.0")):
        put(f"mid.attn_1.{a}", f"mid_block.attentions.0.{b}", conv1x1=True)
    for i in range(4):                       # diffusers up_blocks.i (coarse -> fine) = LDM up[3 - i]
        for j in range(3):
            resnet(f"up.{3 - i}.block.{j}", f"up_blocks.{i}.resnets.{j}")
        if i < 3:
            put(f"up.{3 - i}.upsample.conv", f"up_blocks.{i}.upsamplers.0.conv")
    put("norm_out", "conv_norm_out")
    put("conv_out", "conv_out")
    missing, unexpected = dec.load_state_dict(sd, strict=True)
    z = torch.randn(2, 16, 8, 8, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        ref = dec(z)
        got = vae_o.vae_decode(p, z)
    assert got.shape == ref.shape == (2, 3, 64, 64)
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-4), (got - ref).abs().max()


def test_clip_text_with_projection_oracle_matches_transformers():
    """encode_prompt's CLIP branch (train_dreambooth_lora_sd3.py:57-93): hidden_states[-2] and text_embeds, for both
    activation flavours (CLIP-L quick_gelu, CLIP-G gelu)."""
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection
    from oracle import text_encoders as te_o
    for act in ("quick_gelu", "gelu"):
        cfg = CLIPTextConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=4,
                             vocab_size=300, max_position_embeddings=77, hidden_act=act, eos_token_id=2, bos_token_id=0,
                             pad_token_id=1, projection_dim=48)
        torch.manual_seed(0)
        m = CLIPTextModelWithProjection(cfg).eval()
        p = {k: v.detach().clone() for k, v in m.state_dict().items()}
        ids = torch.randint(3, 298, (2, 77))
        ids[0, 10:] = 299                                   # EOS (highest id) then EOS padding, as the CLIP tokenizers pad
        ids[1, 30:] = 299
        with torch.no_grad():
            ref = m(ids, output_hidden_states=True)
            got_te, got_h = te_o.clip_text_with_projection(p, dict(layers=3, heads=4, act=act, eos_id=2), ids)
        assert len(got_h) == len(ref.hidden_states) == 4
        assert torch.allclose(got_te, ref[0], atol=2e-5, rtol=1e-4)
        assert torch.allclose(got_te, ref.text_embeds, atol=2e-5, rtol=1e-4)
        assert torch.allclose(got_h[-2], ref.hidden_states[-2], atol=2e-5, rtol=1e-4)


def test_t5_encoder_oracle_matches_transformers():
    """encode_prompt's T5 branch (train_dreambooth_lora_sd3.py:13-55): T5 v1.1 encoder (gated-gelu, RMS norm,
    relative-position bias, no mask) incl. the bucket function."""
    from transformers import T5Config, T5EncoderModel
    from oracle import text_encoders as te_o
    cfg = T5Config(vocab_size=200, d_model=64, d_kv=16, d_ff=96, num_layers=3, num_heads=4,
                   relative_attention_num_buckets=32, relative_attention_max_distance=128,
                   feed_forward_proj="gated-gelu", dropout_rate=0.0, is_encoder_decoder=False, use_cache=False)
    torch.manual_seed(0)
    m = T5EncoderModel(cfg).eval()
    with torch.no_grad():
        m.encoder.block[0].layer[0].SelfAttention.relative_attention_bias.weight.normal_(0, 0.5)
        for blk in m.encoder.block:                      # non-trivial norm weights
            blk.layer[0].layer_norm.weight.normal_(1, 0.1)
            blk.layer[1].layer_norm.weight.normal_(1, 0.1)
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ids = torch.randint(2, 199, (2, 200))                # S = 200 > max_distance: exercises the log buckets + clamp
    ids[:, 150:] = 0
    with torch.no_grad():
        ref = m(ids)[0]
        got = te_o.t5_encoder(p, dict(layers=3, heads=4, d_kv=16), ids)
    assert torch.allclose(got, ref, atol=3e-5, rtol=1e-4), (got - ref).abs().max()


def test_mmdit_joint_block_oracle_matches_flux_double_stream_block():
    """torchtitan's FLUX `DoubleStreamBlock` (installed here) is structurally the MMDiT
    joint block (adaLN-Zero with 6 chunks in the order shift/scale/gate x2 from `Linear(SiLU(vec))`, affine-free
    LayerNorm eps 1e-6, per-head RMS q/k norm, ONE softmax over the concatenated text+image tokens, gated
    residuals, GELU(tanh) feed-forward).  Differences: FLUX concatenates [text, image] (attention without a mask
    is invariant to the key order, so the per-stream outputs are the same) and applies RoPE (identity at
    position 0).  Map the oracle's diffusers-named weights of a non-dual, non-last block onto it."""
    import pytest
    layers = pytest.importorskip("torchtitan.experiments.flux.model.layers")
    from adv_grpo_b200 import weights
    from oracle.mmdit import MMDiTOracle, timestep_embedding

    cfg = dict(weights.MMDIT_TINY, num_layers=3, dual_layers=())
    p = weights.init_mmdit(cfg, seed=11, device="cpu", dtype=torch.float32)
    H, D = cfg["heads"], cfg["head_dim"]
    d = H * D
    blk = layers.DoubleStreamBlock(d, H, mlp_ratio=4.0, qkv_bias=True).eval()
    for n in (blk.img_attn.norm, blk.txt_attn.norm):
        n.query_norm.eps = n.key_norm.eps = 1e-6          # diffusers RMSNorm(eps=1e-6); nn.RMSNorm defaults to finfo.eps
    i = 1
    pre = f"transformer_blocks.{i}"
    sd = {}
    for s in ("weight", "bias"):
        sd[f"img_mod.lin.{s}"] = p[f"{pre}.norm1.linear.{s}"]
        sd[f"txt_mod.lin.{s}"] = p[f"{pre}.norm1_context.linear.{s}"]
        sd[f"img_attn.qkv.{s}"] = torch.cat([p[f"{pre}.attn.to_{n}.{s}"] for n in "qkv"])
        sd[f"txt_attn.qkv.{s}"] = torch.cat([p[f"{pre}.attn.add_{n}_proj.{s}"] for n in "qkv"])
        sd[f"img_attn.proj.{s}"] = p[f"{pre}.attn.to_out.0.{s}"]
        sd[f"txt_attn.proj.{s}"] = p[f"{pre}.attn.to_add_out.{s}"]
        sd[f"img_mlp.0.{s}"] = p[f"{pre}.ff.net.0.proj.{s}"]
        sd[f"img_mlp.2.{s}"] = p[f"{pre}.ff.net.2.{s}"]
        sd[f"txt_mlp.0.{s}"] = p[f"{pre}.ff_context.net.0.proj.{s}"]
        sd[f"txt_mlp.2.{s}"] = p[f"{pre}.ff_context.net.2.{s}"]
    sd["img_attn.norm.query_norm.weight"] = p[f"{pre}.attn.norm_q.weight"]
    sd["img_attn.norm.key_norm.weight"] = p[f"{pre}.attn.norm_k.weight"]
    sd["txt_attn.norm.query_norm.weight"] = p[f"{pre}.attn.norm_added_q.weight"]
    sd["txt_attn.norm.key_norm.weight"] = p[f"{pre}.attn.norm_added_k.weight"]
    blk.load_state_dict(sd, strict=True)

    g = torch.Generator().manual_seed(12)
    B, N, T = 2, 36, 13
    x, c, temb = torch.randn(B, N, d, generator=g), torch.randn(B, T, d, generator=g), torch.randn(B, d, generator=g)
    pe = torch.eye(2).expand(1, 1, T + N, D // 2, 2, 2)   # RoPE at position 0
    oracle = MMDiTOracle(p, dict(cfg, dual_layers=set()))
    with torch.no_grad():
        ref_x, ref_c = blk(x, c, temb, pe)
        got_x, got_c = oracle.block(i, x, c, temb)
    assert torch.allclose(got_x, ref_x, atol=2e-5, rtol=1e-4), (got_x - ref_x).abs().max()
    assert torch.allclose(got_c, ref_c, atol=2e-5, rtol=1e-4), (got_c - ref_c).abs().max()
    # the sinusoidal timestep features (diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0))
    t = torch.tensor([1000.0, 464.876, 8.9286])
    assert torch.allclose(timestep_embedding(t), layers.timestep_embedding(t, 256, time_factor=1.0), atol=1e-6)


def test_mmdit_pos_embed_matches_mae_sincos_tables():
    """diffusers `get_2d_sincos_pos_embed` (behind PatchEmbed.cropped_pos_embed, the MMDiT's additive position table)
    is the MAE table with the grid divided by `grid_size / base_size`; transformers ships the MAE original.  Pins the
    axis order (w first), the [sin, cos] halves and the half-h / half-w split of the oracle; the centre crop is then
    the same table evaluated on the cropped, scaled coordinates."""
    import pytest
    mae = pytest.importorskip("transformers.models.vit_mae.modeling_vit_mae")
    from oracle.mmdit import cropped_pos_embed
    dim, n = 64, 12
    full = cropped_pos_embed(dim, n, n, max_size=n, base_size=n)[0].numpy()          # no crop, scale 1
    np.testing.assert_allclose(full, mae.get_2d_sincos_pos_embed(dim, n), atol=1e-6)
    h, w, max_size, base = 6, 10, 24, 8                                                # SD3.5: 384 / 64, crop to the latent grid
    top, left = (max_size - h) // 2, (max_size - w) // 2
    gh = np.arange(top, top + h, dtype=np.float64) / (max_size / base)
    gw = np.arange(left, left + w, dtype=np.float64) / (max_size / base)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, h, w)
    want = mae.get_2d_sincos_pos_embed_from_grid(dim, grid)
    got = cropped_pos_embed(dim, h, w, max_size=max_size, base_size=base)[0].numpy()
    np.testing.assert_allclose(got, want, atol=1e-6)
    # and it IS a crop of the full max_size table
    table = cropped_pos_embed(dim, max_size, max_size, max_size=max_size, base_size=base)[0].reshape(max_size, max_size, dim)
    np.testing.assert_allclose(got.reshape(h, w, dim), table[top:top + h, left:left + w].numpy(), atol=1e-6)


def test_mmdit_dual_attention_branch_matches_flux_self_attention():
    """The image-only `attn2` of the SD3.5 dual-attention blocks (JointAttnProcessor2_0 without a context: to_q/k/v,
    per-head RMS q/k norm, SDPA, to_out) against torchtitan's FLUX `SelfAttention` (fused qkv + QKNorm + attention +
    proj; RoPE at position 0 = identity) with the oracle's diffusers-named weights mapped onto it."""
    import pytest
    layers = pytest.importorskip("torchtitan.experiments.flux.model.layers")
    from adv_grpo_b200 import weights
    from oracle.mmdit import MMDiTOracle
    cfg = dict(weights.MMDIT_TINY)
    p = weights.init_mmdit(cfg, seed=21, device="cpu", dtype=torch.float32)
    H, D = cfg["heads"], cfg["head_dim"]
    d = H * D
    pre = "transformer_blocks.0.attn2"
    sa = layers.SelfAttention(d, num_heads=H, qkv_bias=True).eval()
    sa.norm.query_norm.eps = sa.norm.key_norm.eps = 1e-6
    sd = {"norm.query_norm.weight": p[f"{pre}.norm_q.weight"], "norm.key_norm.weight": p[f"{pre}.norm_k.weight"]}
    for s in ("weight", "bias"):
        sd[f"qkv.{s}"] = torch.cat([p[f"{pre}.to_{n}.{s}"] for n in "qkv"])
        sd[f"proj.{s}"] = p[f"{pre}.to_out.0.{s}"]
    sa.load_state_dict(sd, strict=True)
    x = torch.randn(2, 36, d, generator=torch.Generator().manual_seed(22))
    pe = torch.eye(2).expand(1, 1, 36, D // 2, 2, 2)
    oracle = MMDiTOracle(p, dict(cfg, dual_layers=set(cfg["dual_layers"])))
    with torch.no_grad():
        ref = sa(x, pe)
        got, none = oracle._attn(pre, x)
    assert none is None
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-4), (got - ref).abs().max()


def test_scheduler_shift_is_the_flux_time_shift():
    """The static shift of FlowMatchEulerDiscreteScheduler, sigma = shift * s / (1 + (shift - 1) * s), is FLUX's
    `time_shift(mu = log(shift), sigma = 1, s)` (torchtitan ships the FLUX sampler).  Pins the functional form of the
    oracle's schedule; the diffusers-specific parts (the shift applied to the training table AND again in
    set_timesteps, the linspace end points) are restated from diffusers 0.33.1 and stay unpinned."""
    import math
    import pytest
    sampling = pytest.importorskip("torchtitan.experiments.flux.sampling")
    from oracle.scheduler import FlowMatchEulerOracle
    sch = FlowMatchEulerOracle(shift=3.0)
    sch.set_timesteps(10)
    s = torch.linspace(sch.sigma_max, sch.sigma_min, 10, dtype=torch.float64)          # "t / 1000" space
    want = sampling.time_shift(math.log(3.0), 1.0, s)
    assert torch.allclose(sch.sigmas[:-1].double(), want, atol=1e-6)
    assert sch.sigmas[-1] == 0 and torch.allclose(sch.timesteps, sch.sigmas[:-1] * 1000)


def test_mmdit_adaln_chunk_orders_match_transformers_dit_modules():
    """Independent pin of the two adaLN conventions the oracle restates from diffusers: transformers ships, inside the
    Qwen2.5-Omni token2wav DiT, module-for-module descendants of diffusers' `AdaLayerNormZero` (chunks = shift_msa,
    scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp; LayerNorm without affine, eps 1e-6; `linear(silu(emb))`) and of
    `AdaLayerNormContinuous` as used for the final layer (chunks = SCALE, SHIFT -- the opposite order).  Same weights, same
    inputs: `oracle.mmdit.ada_layer_norm_zero` / `ada_layer_norm_continuous` (used by every block and by norm_out /
    the last block's norm1_context) must reproduce them."""
    import pytest
    qo = pytest.importorskip("transformers.models.qwen2_5_omni.modeling_qwen2_5_omni")
    from oracle import mmdit as mm
    torch.manual_seed(0)
    d, B, S = 64, 3, 7
    x, temb = torch.randn(B, S, d), torch.randn(B, d)
    zero = qo.Qwen2_5_OmniAdaLayerNormZero(d)
    with torch.no_grad():
        ref = zero(x, emb=temb)
        e = torch.nn.functional.linear(torch.nn.functional.silu(temb), zero.linear.weight, zero.linear.bias)
        got = mm.ada_layer_norm_zero(x, e)
    assert len(ref) == len(got) == 5
    for r, g in zip(ref, got):                                   # modulated x, gate_msa, shift_mlp, scale_mlp, gate_mlp
        assert torch.allclose(r, g, atol=1e-6)
    final = qo.Qwen2_5_OmniAdaLayerNormZero_Final(d)
    with torch.no_grad():
        ref_f = final(x, temb)
        e2 = torch.nn.functional.linear(torch.nn.functional.silu(temb), final.linear.weight, final.linear.bias)
        got_f = mm.ada_layer_norm_continuous(x, e2)
        swapped = mm.layer_norm(x) * (1 + e2.chunk(2, dim=1)[1])[:, None] + e2.chunk(2, dim=1)[0][:, None]
    assert torch.allclose(ref_f, got_f, atol=1e-6)
    assert not torch.allclose(ref_f, swapped, atol=1e-3)         # the order matters: (shift, scale) would be caught


from jpeg_util import _cmyk_jpeg, _jpeg_bytes  # noqa: E402


JPEG_CASES = [((64, 64), dict(quality=90, subsampling=0)), ((48, 80), dict(quality=75, subsampling=2)),
              ((37, 53), dict(quality=85, subsampling=2)), ((33, 47), dict(quality=60, subsampling=1)),
              ((17, 9), dict(quality=80, subsampling=2)), ((40, 40), dict(quality=50, subsampling=2, restart_marker_blocks=2)),
              ((1, 1), dict(quality=90, subsampling=2)), ((8, 2), dict(quality=90, subsampling=1)),
              ((24, 40), dict(quality=100, subsampling=0)), ((30, 45), dict(quality=80, gray=True)),
              # progressive (SOF2): DC / AC first + refinement scans, end-of-band runs
              ((48, 80), dict(quality=75, subsampling=2, progressive=True)), ((33, 47), dict(quality=60, subsampling=1, progressive=True)),
              ((37, 53), dict(quality=95, subsampling=0, progressive=True)), ((30, 45), dict(quality=80, gray=True, progressive=True)),
              ((40, 40), dict(quality=50, subsampling=2, progressive=True, restart_marker_blocks=2)),
              # narrow planes: plain replication instead of the triangle filter when ceil(W / 2) <= 2 (jinit_upsampler)
              ((47, 3), dict(quality=100, subsampling=2)), ((16, 4), dict(quality=90, subsampling=1)), ((9, 5), dict(quality=90, subsampling=2))]


def test_jpeg_oracle_matches_pillow():
    """oracle/jpeg.py (numpy restatement of libjpeg's default baseline decode: Huffman, islow IDCT, fancy upsampling, YCbCr
    tables) returns exactly Pillow's pixels -- the pin of the oracle the GPU decoder is tested against -- for 4:4:4, 4:2:2,
    4:2:0, grayscale, odd sizes down to 1x1, restart intervals, sequential and progressive files; CMYK is reported unsupported."""
    import io
    import pytest
    from PIL import Image
    from oracle import jpeg as jpeg_o
    for (h, w), kw in JPEG_CASES:
        data = _jpeg_bytes(h, w, seed=h + w, **kw)
        ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        assert np.array_equal(jpeg_o.decode_rgb(data), ref), ((h, w), kw)
    with pytest.raises(jpeg_o.JpegUnsupported):
        jpeg_o.decode_rgb(_cmyk_jpeg())


def test_jpeg_host_entropy_decoder_matches_oracle():
    """The library's host Huffman decoder (csrc/jpeg.cu, plain C++; runs without a GPU) yields the oracle's coefficient
    blocks and quantisation tables bit for bit, and classifies unsupported files without raising."""
    from adv_grpo_b200 import jpeg as jpeg_b
    from oracle import jpeg as jpeg_o
    for (h, w), kw in JPEG_CASES:
        data = _jpeg_bytes(h, w, seed=h + w, **kw)
        info = jpeg_o.parse(data)
        ref = jpeg_o.entropy_decode(data, info)
        got, qt, gi = jpeg_b.coefficients_as_numpy(data)
        assert gi.supported == 1 and (gi.height, gi.width) == (h, w) and gi.ncomp == len(ref)
        for c in range(gi.ncomp):
            assert np.array_equal(got[c], ref[c]), ((h, w), kw, c)
            assert np.array_equal(qt[c], info["q"][info["frame"]["comps"][c]["tq"]])
    assert jpeg_b.jpeg_info(_jpeg_bytes(16, 16, progressive=True)).progressive == 1
    assert jpeg_b.jpeg_info(_cmyk_jpeg()).supported == 0
    assert jpeg_b.coefficients_as_numpy(_cmyk_jpeg())[0] is None
    import pytest
    from adv_grpo_b200 import _lib
    with pytest.raises(_lib.AdvGrpoError):
        jpeg_b.jpeg_info(b"not a jpeg at all")


def test_jpeg_host_decoder_fuzz_against_pillow():
    """Randomised files (noise / flat / patterned content, quality 1..100, every sampling mode, sequential and progressive,
    restart intervals, optimised Huffman tables): the library's host entropy decoder + the oracle's numpy back end (the
    arithmetic the GPU kernels implement) reproduce Pillow's pixels exactly."""
    import io
    from PIL import Image
    from adv_grpo_b200 import jpeg as jpeg_b
    from oracle import jpeg as jpeg_o
    rng = np.random.default_rng(123)
    for _ in range(24):
        h, w = int(rng.integers(1, 70)), int(rng.integers(1, 70))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        elif kind == 1:
            img = np.full((h, w, 3), int(rng.integers(0, 256)), dtype=np.uint8)
        else:
            yy, xx = np.mgrid[0:h, 0:w]
            img = np.stack([(xx * 3 + yy) % 256, (yy * 5) % 256, (xx * yy) % 256], -1).astype(np.uint8)
        kw = dict(quality=int(rng.choice([1, 10, 35, 75, 90, 100])), subsampling=int(rng.integers(0, 3)),
                  progressive=bool(rng.integers(0, 2)))
        if rng.integers(0, 3) == 0:
            kw["restart_marker_blocks"] = int(rng.integers(1, 5))
        if rng.integers(0, 4) == 0:
            kw["optimize"] = True
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, format="JPEG", **kw)
        data = buf.getvalue()
        ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        coefs, _, info = jpeg_b.coefficients_as_numpy(data)
        assert coefs is not None and (info.height, info.width) == (h, w), kw
        oinfo = jpeg_o.parse(data)
        got = jpeg_o.assemble_rgb(oinfo, coefs)
        assert np.array_equal(got, ref), ((h, w), kw)


def test_png_oracle_and_host_inflate_match_pillow_and_zlib():
    """oracle/png.py (numpy restatement: chunk walk, unfiltering incl. Average and Paeth, RGB conversion) returns exactly
    Pillow's pixels, and the library's host inflate (csrc/png.cu, plain C++; runs without a GPU) returns exactly zlib's bytes --
    Pillow-written files of every colour type and hand-assembled files with random filter types, stored / fixed / dynamic
    deflate blocks and split IDAT chunks, every supported bit depth, Adam7-interlaced files."""
    import io
    import pytest
    from PIL import Image
    from adv_grpo_b200 import png as png_b
    from oracle import png as png_o
    from png_util import handmade_png, pillow_png
    files = [pillow_png(40, 56, m, seed=i) for i, m in enumerate(("RGB", "RGBA", "L", "LA", "P"))]
    files += [pillow_png(1, 1, "RGB"), pillow_png(17, 3, "RGBA"), pillow_png(33, 70, "RGB", compress_level=0)]
    raws = [None] * len(files)
    for t in range(16):
        f, raw = handmade_png(int(1 + t * 5 % 37), int(1 + t * 7 % 41), [0, 2, 3, 4, 6][t % 5], seed=t, level=[0, 1, 6, 9][t % 4],
                              split=[None, 5, 100][t % 3], kind=t % 3)
        files.append(f)
        raws.append(raw)
    for data, raw in zip(files, raws):
        ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        assert np.array_equal(png_o.decode_rgb(data), ref)
        got_raw, _, info = png_b.inflate(data)
        assert info.supported == 1 and (info.height, info.width) == ref.shape[:2]
        import zlib
        assert got_raw.numpy().tobytes() == zlib.decompress(png_o.parse(data)["idat"])
        if raw is not None:
            assert got_raw.numpy().tobytes() == raw
    # every bit depth Pillow maps onto 8-bit RGB: sub-byte greyscale / palette, 16-bit greyscale (clipped) and truecolour
    for t, (ct, bd) in enumerate(((0, 1), (0, 2), (0, 4), (0, 16), (2, 16), (6, 16), (3, 1), (3, 2), (3, 4), (4, 16))):
        data, raw = handmade_png(9 + t, 21 - t, ct, seed=100 + t, bd=bd, kind=t % 2)
        ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        assert np.array_equal(png_o.decode_rgb(data), ref), (ct, bd)
        got_raw, _, info = png_b.inflate(data)
        assert info.supported == 1 and got_raw.numpy().tobytes() == raw
    # Adam7-interlaced files: seven reduced images with their own scan lines
    for t, (ct, bd) in enumerate(((2, 8), (6, 8), (0, 4), (3, 2), (2, 16), (0, 8), (4, 16))):
        for h, w in ((19 + t, 23 - t), (3, 2), (1, 1), (8, 9)):
            data, raw = handmade_png(h, w, ct, seed=200 + t + h, bd=bd, interlace=1)
            ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
            assert np.array_equal(png_o.decode_rgb(data), ref), (ct, bd, h, w)
            got_raw, _, info = png_b.inflate(data)
            assert info.supported == 1 and info.interlace == 1 and got_raw.numpy().tobytes() == raw
    # a depth / colour-type pair PNG does not define (4-bit truecolour): classified, not decoded
    data = handmade_png(8, 8, 2, bd=4)[0]
    assert png_b.png_info(data).supported == 0 and png_b.inflate(data)[0] is None
    with pytest.raises(png_o.PngUnsupported):
        png_o.decode_rgb(data)
    from adv_grpo_b200 import _lib
    with pytest.raises(_lib.AdvGrpoError):
        png_b.png_info(b"definitely not a png")

"""GPU parity of the text-encoding step (`compute_text_embeddings` / `encode_prompt`,
train_sd3_fast_pickscore.py:186-193, train_dreambooth_lora_sd3.py:13-144) against the CPU oracle
(oracle/text_encoders.py, itself pinned to transformers' CLIPTextModelWithProjection / T5EncoderModel in
tests/test_oracle_models.py), plus the two kernel features it adds: the attention score bias and the quick-GELU epilogue."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(got, ref):
    return ((got.float().cpu() - ref.float()).abs().max() / ref.float().abs().max()).item()


def _clip_ids(B, vocab, g):
    ids = torch.randint(3, vocab - 2, (B, 77), generator=g)
    for b in range(B):
        n = int(torch.randint(5, 70, (1,), generator=g))
        ids[b, n:] = vocab - 1                      # EOS = highest id, also the padding token (CLIP tokenizers)
    return ids


@pytest.mark.parametrize("S,H", [(128, 4), (200, 2), (77, 64), (333, 3)])
def test_attention_with_score_bias_matches_torch(S, H):
    from adv_grpo_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(S + H)
    B, D = 2, 64
    qkv = torch.randn(B, S, 3, H, D, device=DEV, generator=g).bfloat16()
    bias = 2.0 * torch.randn(H, S, S, device=DEV, generator=g)
    out, lse = ops.attention_fwd_bias(qkv, bias, scale=0.3, want_lse=True)
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3).float() for i in range(3))
    sc = 0.3 * (q @ k.transpose(-1, -2)) + bias[None]
    ref = (sc.softmax(-1) @ v).permute(0, 2, 1, 3)
    assert _rel(out.cpu(), ref.cpu()) < 1.5e-2
    assert torch.allclose(lse, torch.logsumexp(sc, -1), atol=2e-3, rtol=1e-4)


def test_quick_gelu_epilogue_matches_torch():
    from adv_grpo_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(3)
    M, N, K = 300, 1544, 128
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g).bfloat16()
    z = (a.float() @ w.float().T + bias.float()).bfloat16().float()
    got = ops.gemm(a, w, bias=bias, epilogue=ops.EPI_QUICK_GELU)
    assert _rel(got.cpu(), (z * torch.sigmoid(1.702 * z)).cpu()) < 6e-3


@pytest.mark.parametrize("which", ["l", "g"])
def test_clip_text_with_projection_matches_oracle(which):
    from adv_grpo_b200 import weights
    from adv_grpo_b200.text_encoders import CLIPTextModelWithProjection
    from oracle import text_encoders as te_o
    cfg = weights.CLIP_L_TEXT_TINY if which == "l" else weights.CLIP_G_TEXT_TINY
    p = weights.init_clip_text(cfg, seed=5, device="cpu", dtype=torch.bfloat16)
    enc = CLIPTextModelWithProjection(p, cfg, device=DEV)
    ids = _clip_ids(3, cfg["vocab"], torch.Generator().manual_seed(1))
    out = enc(ids.to(DEV), output_hidden_states=True)
    te, hidden = te_o.clip_text_with_projection({k: v.float() for k, v in p.items()}, cfg, ids)
    assert len(out.hidden_states) == cfg["layers"] + 1 and out[0] is out.text_embeds
    assert _rel(out.hidden_states[-2], hidden[-2]) < 3e-2
    assert _rel(out[0], te) < 3e-2
    assert enc(ids.to(DEV)).hidden_states is None


def test_t5_encoder_matches_oracle_tiny_and_true_width():
    from adv_grpo_b200 import weights
    from adv_grpo_b200.text_encoders import T5EncoderModel
    from oracle import text_encoders as te_o
    # tiny stack, and two layers at the true T5-XXL widths (d_model 4096, 64 heads x 64, d_ff 10240, 128 tokens)
    for cfg, B, S in ((weights.T5_TINY, 2, 200), (dict(weights.T5_XXL, layers=2, vocab=512), 1, 128)):
        p = weights.init_t5_encoder(cfg, seed=6, device="cpu", dtype=torch.bfloat16)
        enc = T5EncoderModel(p, cfg, device=DEV)
        ids = torch.randint(2, cfg["vocab"], (B, S), generator=torch.Generator().manual_seed(2))
        ids[:, S - 20:] = 0
        got = enc(ids.to(DEV))[0]
        ref = te_o.t5_encoder({k: v.float() for k, v in p.items()}, cfg, ids)
        assert got.shape == ref.shape == (B, S, cfg["d_model"]) and got.dtype == torch.bfloat16
        assert _rel(got, ref) < 3e-2, (cfg["d_model"], _rel(got, ref))
        # the bias table itself is exact (integer bucket arithmetic + a gather)
        assert torch.equal(enc.position_bias(S).cpu(), te_o.t5_position_bias({k: v.float() for k, v in p.items()}, S,
                                                                             cfg["num_buckets"], cfg["max_distance"]))


def test_encode_prompt_matches_oracle_and_reference_layout():
    from adv_grpo_b200 import install_as_adv_grpo, weights
    from adv_grpo_b200.text_encoders import CLIPTextModelWithProjection, T5EncoderModel
    from oracle import text_encoders as te_o
    install_as_adv_grpo()
    from adv_grpo.diffusers_patch.train_dreambooth_lora_sd3 import compute_text_embeddings, encode_prompt
    cl, cg, ct = weights.CLIP_L_TEXT_TINY, weights.CLIP_G_TEXT_TINY, weights.T5_TINY
    pl, pg, pt = weights.init_clip_text(cl, 5), weights.init_clip_text(cg, 7), weights.init_t5_encoder(ct, 6)
    encs = [CLIPTextModelWithProjection(pl, cl, device=DEV), CLIPTextModelWithProjection(pg, cg, device=DEV),
            T5EncoderModel(pt, ct, device=DEV)]
    g = torch.Generator().manual_seed(4)
    ids_l, ids_g = _clip_ids(2, cl["vocab"], g), _clip_ids(2, cg["vocab"], g)
    ids_t = torch.randint(2, ct["vocab"], (2, 128), generator=g)
    pe, pooled = encode_prompt(encs, [None, None, None], ["a", "b"], 128, text_input_ids_list=[ids_l, ids_g, ids_t])
    f = lambda p: {k: v.float() for k, v in p.items()}
    pe_o, pooled_o = te_o.encode_prompt((f(pl), cl), (f(pg), cg), (f(pt), ct), ids_l, ids_g, ids_t)
    assert pe.shape == pe_o.shape == (2, 77 + 128, ct["d_model"]) and pooled.shape == (2, cl["proj"] + cg["proj"])
    assert _rel(pe, pe_o) < 3e-2 and _rel(pooled, pooled_o) < 3e-2
    assert torch.count_nonzero(pe[:, :77, cl["width"] + cg["width"]:]) == 0          # zero padding of the CLIP rows
    # a tokenizer object with the transformers call convention works as in the reference
    class Tok:
        def __init__(self, ids):
            self.ids = ids

        def __call__(self, prompt, padding, max_length, truncation, return_tensors, **kw):
            assert padding == "max_length" and truncation and return_tensors == "pt" and self.ids.shape[1] == max_length
            return type("Enc", (), {"input_ids": self.ids[:len(prompt)]})()
    pe2, pooled2 = compute_text_embeddings(["a", "b"], encs, [Tok(ids_l), Tok(ids_g), Tok(ids_t)], 128, DEV)
    assert torch.equal(pe2, pe) and torch.equal(pooled2, pooled)
    pe3, _ = encode_prompt(encs, [None] * 3, "a", 128, num_images_per_prompt=3,
                           text_input_ids_list=[ids_l[:1], ids_g[:1], ids_t[:1]])
    assert pe3.shape == (3, 205, ct["d_model"]) and torch.equal(pe3[2], pe[0])

"""The three text encoders of `compute_text_embeddings` (`scripts/train_sd3_fast_pickscore.py:186-193`) on the
libadvgrpo_b200 kernels: CLIP-L and CLIP-G text towers with projection (transformers
`CLIPTextModelWithProjection`: `encoder(ids, output_hidden_states=True)` -> `[0]` = text_embeds,
`.hidden_states[-2]`) and the T5-XXL encoder (`T5EncoderModel`: `encoder(ids)[0]`), called by `encode_prompt`
(`adv_grpo/diffusers_patch/train_dreambooth_lora_sd3.py:13-144`, mirrored in
`adv_grpo_b200/diffusers_patch/train_dreambooth_lora_sd3.py`).  SURVEY.md section 8f, rank 1.

Every linear is the tcgen05 GEMM with its epilogue fused (bias, GELU-erf / quick-GELU / GELU-tanh, residual add);
attention is the flash-attention kernel (causal for CLIP; with T5's additive relative-position bias, un-scaled
scores, for T5).  One prompt is 77 + 128 tokens, so the step is bound by streaming the 11 GB of frozen bf16
weights once (T5-XXL: 9.4 GB), not by FLOPs.  Parameters keep their transformers state-dict names.
"""
import math

import torch
import torch.nn.functional as F

from . import ops, vit


class _ClipTextOutput(tuple):
    """Indexable like transformers' CLIPTextModelOutput: [0] = text_embeds, [1] = last_hidden_state."""

    def __new__(cls, text_embeds, last_hidden_state, hidden_states):
        o = super().__new__(cls, (text_embeds, last_hidden_state))
        o.text_embeds, o.last_hidden_state, o.hidden_states = text_embeds, last_hidden_state, hidden_states
        return o


class CLIPTextModelWithProjection:
    def requires_grad_(self, flag=True):
        return self

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def __init__(self, params, cfg, device="cuda", dtype=torch.bfloat16):
        self.cfg, self.device, self.dtype = dict(cfg), torch.device(device), dtype
        p = {k: v.to(device=self.device, dtype=dtype) for k, v in params.items()}
        self.p = p
        w = cfg["width"]
        act = ops.EPI_QUICK_GELU if cfg["act"] == "quick_gelu" else ops.EPI_GELU_ERF
        self.blocks = []
        for i in range(cfg["layers"]):
            l = f"text_model.encoder.layers.{i}"
            g = lambda n: (p[f"{l}.{n}.weight"], p[f"{l}.{n}.bias"])
            (wq, bq), (wk, bk), (wv, bv), (wo, bo) = (g(f"self_attn.{n}_proj") for n in ("q", "k", "v", "out"))
            self.blocks.append(vit.ViTBlock(w, cfg["heads"], wq, bq, wk, bk, wv, bv, wo, bo, g("layer_norm1"),
                                            g("layer_norm2"), g("mlp.fc1"), g("mlp.fc2"), 1e-5, act_epilogue=act))

    def state_dict(self):
        return self.p

    @torch.no_grad()
    def __call__(self, input_ids, output_hidden_states=False, **_):
        p, cfg = self.p, self.cfg
        input_ids = input_ids.to(self.device)
        B, S = input_ids.shape
        x = (p["text_model.embeddings.token_embedding.weight"][input_ids]
             + p["text_model.embeddings.position_embedding.weight"][:S][None]).contiguous()
        hidden = [x]
        for blk in self.blocks:
            x = blk(x, causal=True)
            hidden.append(x)
        last = ops.layer_norm(x.contiguous(), p["text_model.final_layer_norm.weight"],
                              p["text_model.final_layer_norm.bias"], 1e-5)
        if cfg.get("eos_id", 2) == 2:
            pos = input_ids.argmax(-1)
        else:
            pos = (input_ids == cfg["eos_id"]).int().argmax(-1)
        pooled = last[torch.arange(B, device=self.device), pos]
        text_embeds = ops.gemm(pooled.contiguous(), p["text_projection.weight"])
        return _ClipTextOutput(text_embeds, last, tuple(hidden) if output_hidden_states else None)


def _relative_position_bucket(relative_position, num_buckets=32, max_distance=128):
    """T5Attention._relative_position_bucket (bidirectional), evaluated once per sequence length on the device."""
    num_buckets //= 2
    ret = (relative_position > 0).long() * num_buckets
    n = relative_position.abs()
    max_exact = num_buckets // 2
    large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).long()
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return ret + torch.where(n < max_exact, n, large)


def _t5_layer_norm(x, w, eps=1e-6):
    """T5LayerNorm: RMS statistics in fp32, cast back to the weight dtype, then scale (no mean, no bias)."""
    var = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps)).to(w.dtype)


class T5EncoderModel:
    def requires_grad_(self, flag=True):
        return self

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def __init__(self, params, cfg, device="cuda", dtype=torch.bfloat16):
        self.cfg, self.device, self.dtype = dict(cfg), torch.device(device), dtype
        p = {k: v.to(device=self.device, dtype=dtype) for k, v in params.items()}
        self.p = p
        d = cfg["d_model"]
        self.ones = torch.ones(1, d, dtype=dtype, device=self.device)
        self.blocks = []
        for i in range(cfg["layers"]):
            b = f"encoder.block.{i}"
            a = b + ".layer.0.SelfAttention"
            f = b + ".layer.1.DenseReluDense"
            self.blocks.append(dict(
                w_qkv=torch.cat([p[f"{a}.{n}.weight"] for n in "qkv"], 0).contiguous(), w_o=p[a + ".o.weight"],
                ln1=p[b + ".layer.0.layer_norm.weight"], ln2=p[b + ".layer.1.layer_norm.weight"],
                wi_0=p[f + ".wi_0.weight"], wi_1=p[f + ".wi_1.weight"], wo=p[f + ".wo.weight"]))
        self._bias = {}

    def state_dict(self):
        return self.p

    def position_bias(self, S):
        """[H, S, S] f32 additive score bias (block 0's relative_attention_bias, shared by every layer)."""
        if S not in self._bias:
            cfg = self.cfg
            ctx = torch.arange(S, device=self.device)[:, None]
            mem = torch.arange(S, device=self.device)[None, :]
            bucket = _relative_position_bucket(mem - ctx, cfg.get("num_buckets", 32), cfg.get("max_distance", 128))
            w = self.p["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
            self._bias[S] = w[bucket].permute(2, 0, 1).float().contiguous()
        return self._bias[S]

    @torch.no_grad()
    def __call__(self, input_ids, **_):
        cfg = self.cfg
        input_ids = input_ids.to(self.device)
        B, S = input_ids.shape
        H, dk = cfg["heads"], cfg["d_kv"]
        M = B * S
        x = self.p["shared.weight"][input_ids].contiguous()
        bias = self.position_bias(S)
        for blk in self.blocks:
            h = _t5_layer_norm(x, blk["ln1"])
            qkv = ops.gemm(h, blk["w_qkv"]).view(B, S, 3, H, dk)
            o, _ = ops.attention_fwd_bias(qkv, bias, scale=1.0)                       # T5 does not scale the scores
            x = ops.gemm(o.view(B, S, H * dk), blk["w_o"], epilogue=ops.EPI_GATE_RESIDUAL, residual=x, gate=self.ones,
                         rows_per_gate=M)
            h = _t5_layer_norm(x, blk["ln2"])
            g = ops.gemm(h, blk["wi_0"], epilogue=ops.EPI_GELU_TANH) * ops.gemm(h, blk["wi_1"])   # gated gelu_new
            x = ops.gemm(g, blk["wo"], epilogue=ops.EPI_GATE_RESIDUAL, residual=x, gate=self.ones, rows_per_gate=M)
        return (_t5_layer_norm(x, self.p["encoder.final_layer_norm.weight"]),)


class GraphedPromptEncoder:
    """`compute_text_embeddings` for ONE prompt (`train_sd3_fast_pickscore.py:186-193`) as a CUDA graph over static
    token-id buffers: the eager step is ~900 small launches (77 / 128-token GEMMs that stream 10.7 GB of frozen
    weights) and is CPU-dispatch-bound; results are cached per prompt (the encoders are frozen), which the
    reference does not do (it re-encodes the prompt of every batch)."""

    def __init__(self, text_encoders, max_sequence_length=128, device="cuda", cache=True):
        self.encoders, self.S, self.device = list(text_encoders), int(max_sequence_length), torch.device(device)
        self.static_ids = [torch.zeros(1, 77, dtype=torch.long, device=self.device),
                           torch.zeros(1, 77, dtype=torch.long, device=self.device),
                           torch.zeros(1, self.S, dtype=torch.long, device=self.device)]
        self.graph, self.out, self.n_kernels = None, None, 0
        self.cache = {} if cache else None

    def _run(self):
        from .diffusers_patch.train_dreambooth_lora_sd3 import encode_prompt
        return encode_prompt(self.encoders, [None, None, None], "prompt", self.S, text_input_ids_list=self.static_ids)

    @torch.no_grad()
    def __call__(self, ids_clip_l, ids_clip_g, ids_t5, key=None):
        """token ids [1, 77], [1, 77], [1, max_sequence_length] -> (prompt_embeds [1, 77 + S, d], pooled [1, P])."""
        from . import _lib
        if self.cache is not None and key is not None and key in self.cache:
            return self.cache[key]
        for dst, src in zip(self.static_ids, (ids_clip_l, ids_clip_g, ids_t5)):
            dst.copy_(src.to(self.device).reshape(dst.shape))
        if self.graph is None:
            for _ in range(2):                       # warm-up (workspace allocation, lazily built bias tables)
                self._run()
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(self.graph):
                self.out = self._run()
            self.n_kernels = _lib.launch_count() - n0
        self.graph.replay()
        _lib.add_launches(self.n_kernels)
        res = tuple(t.clone() for t in self.out)
        if self.cache is not None and key is not None:
            self.cache[key] = res
        return res

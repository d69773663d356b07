"""Text-encoding step at the true sizes of stable-diffusion-3.5-medium's encoders (CLIP-L, CLIP-G, T5-XXL; random
init, synthetic token ids): CUDA-event time of encode_prompt for one prompt and the weight-streaming roofline."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adv_grpo_b200 import weights
from adv_grpo_b200.diffusers_patch.train_dreambooth_lora_sd3 import encode_prompt
from adv_grpo_b200.text_encoders import CLIPTextModelWithProjection, T5EncoderModel

dev = "cuda"
encs, nbytes = [], 0
for cfg, init in ((weights.CLIP_L_TEXT, weights.init_clip_text), (weights.CLIP_G_TEXT, weights.init_clip_text),
                  (weights.T5_XXL, weights.init_t5_encoder)):
    p = init(cfg, device=dev)
    emb = [k for k in p if "embedding" in k or k == "shared.weight"]
    nbytes += sum(v.numel() * 2 for k, v in p.items() if k not in emb)          # the embedding tables are gathered, not streamed
    encs.append((CLIPTextModelWithProjection if "act" in cfg else T5EncoderModel)(p, cfg, device=dev))
    del p
g = torch.Generator().manual_seed(0)
ids_c = torch.randint(3, 49406, (1, 77), generator=g); ids_c[:, 20:] = 49407
ids_t = torch.randint(2, 32000, (1, 128), generator=g); ids_t[:, 30:] = 0
ids = [ids_c.to(dev), ids_c.to(dev), ids_t.to(dev)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    pe, pooled = encode_prompt(encs, [None] * 3, "p", 128, text_input_ids_list=ids)
ts = []
for _ in range(7):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pe, pooled = encode_prompt(encs, [None] * 3, "p", 128, text_input_ids_list=ids); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort(); ms = ts[len(ts) // 2]
from adv_grpo_b200.text_encoders import GraphedPromptEncoder
ge = GraphedPromptEncoder(encs, 128, device=dev, cache=False)
for _ in range(3):
    pe_g, pooled_g = ge(*ids)
assert torch.equal(pe_g, pe) and torch.equal(pooled_g, pooled), "graphed encode differs from eager"
tg = []
for _ in range(7):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ge(*ids); e1.record()
    torch.cuda.synchronize(); tg.append(e0.elapsed_time(e1))
tg.sort(); ms_g = tg[len(tg) // 2]
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6551.0}
gbs = nbytes / ms / 1e6
print(json.dumps({"what": "encode_prompt, 1 prompt (77 CLIP-L + 77 CLIP-G + 128 T5-XXL tokens), eager launches",
                  "ms": round(ms, 3), "ms_cuda_graph": round(ms_g, 3), "graph_GBps": round(nbytes / ms_g / 1e6, 1), "prompt_embeds": list(pe.shape), "pooled": list(pooled.shape),
                  "weight_bytes_streamed": nbytes, "achieved_GBps": round(gbs, 1), "hbm_peak_GBps": peaks["hbm_gbs"],
                  "frac": round(gbs / peaks["hbm_gbs"], 3)}))

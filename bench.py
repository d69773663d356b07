#!/usr/bin/env python
"""Benchmark of the Adv-GRPO hot path (BASELINE.json metric: GRPO samples/sec, rollout + score + update,
SD3.5-medium, 512x512, 10 denoise steps, G = 8, PickScore reward, LoRA r=32).

    python bench.py --gpus N --steps K --warmup W           # our arm (one rank per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on the host CPU cores

One "step" = one GRPO epoch slice per rank: NB = 2 prompt groups of G = 8 images are rolled out
(10 MMDiT steps at CFG batch 16, fused CFG/SDE/log-prob, VAE decode), scored (PickScore on generated and
reference images), turned into group-relative advantages, and trained on (2 SDE-window timesteps per
group, forward + backward through the LoRA MMDiT, clipped GRPO loss, all-reduce, clip + AdamW + EMA):
2 optimizer steps per step, like the reference epoch.  value = samples (images) through that whole
slice per second, summed over ranks.  Synthetic prompts / seeded random weights at the true shapes (no
datasets or checkpoints exist on the box).  `e2e` repeats the measurement with every step's inputs
(prompt embeddings, reference images) copied from pinned host memory and the step's metrics read back.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# `metric` is BASELINE.json's own string (the workload detail -- SD3.5-medium LoRA r32, 512x512, G = 8, 10 steps,
# PickScore -- is spelled out in `config.workload`)
METRIC = "GRPO samples/sec (rollout+score+update) SD3-m G=8 10-step @1/2/4/8 B200"
try:
    with open(os.path.join(ROOT, "BASELINE.json")) as _f:
        METRIC = json.load(_f).get("metric", METRIC)
except (OSError, ValueError):
    pass
UNIT = "samples/s"
NB, G, T_STEPS, T_TRAIN, RES = 2, 8, 10, 2, 512          # BASELINE config 2 (the default; the CPU arm is quoted on it)

# BASELINE.json configs 2-5 (config 1 is the CPU plumbing case: tests/, not a bench line).  `groups` = prompt groups per
# rank per step in weak scaling; `--scaling strong` fixes 16 groups per step over all ranks instead.
CONFIGS = {
    2: dict(preset="pickscore_cotrain_sd3_fast", res=512, steps=10, G=8, groups=2, train_d=False,
            workload="SD3.5-medium LoRA r32 512x512, 10 denoise steps (2 trained), G=8, CFG 4.5, PickScore (CLIP-ViT-H/14) "
                     "reward on generated+reference images; {nb} groups + 2 optimizer steps per rank per step (BASELINE config 2)"),
    3: dict(preset="dino_patch_cotrain_sd3_fast", res=512, steps=10, G=8, groups=2, train_d=True,
            workload="SD3.5-medium LoRA r32 512x512, 10 denoise steps (2 trained), G=8, DINOv2-B/14 patch adversarial reward "
                     "(CLS + 64 patches, DINOHead) on generated+reference images; {nb} groups per rank per step; the step "
                     "alternates generator GRPO updates with hinge-loss head (D) updates as train_sd3_fast_dino_patch.py:1097 "
                     "(d_times = 10: 9 of 10 epochs are D steps) (BASELINE config 3)"),
    4: dict(preset="pickscore_sd3_fast", res=1024, steps=20, G=16, groups=1, train_d=False,
            workload="SD3.5-medium LoRA r32 1024x1024 (4301 joint tokens), 20 denoise steps (2 trained), G=16, multi-reward "
                     "{{pickscore: 0.5, ocr: 0.5}} (rewards.py registry; frozen fp32 PickScore + the host OCR plugin with a "
                     "deterministic recogniser stub, SURVEY 8d); {nb} group(s) + 2 optimizer steps per rank per step, the "
                     "replay of a group runs in micro-batches of 8 samples (BASELINE config 4)"),
    5: dict(preset="pickscore_cotrain_sd3_fast", res=512, steps=10, G=8, groups=2, train_d=True,
            workload="SD3.5-medium LoRA r32 512x512, 10 steps, G=8, adversarial co-update: steps alternate a generator GRPO "
                     "update and a PickScore discriminator update (CLIPCriterion on generated vs reference images, last "
                     "vision block trainable, train_sd3_fast_pickscore.py:151-183,1003-1037); {nb} groups per rank per step "
                     "(BASELINE config 5)"),
}
STRONG_GROUPS = 16


def config_dict(config_id, nb, world, full_finetune=False):
    """The `config` object of the JSON line; both arms (ours and `--impl reference`) print the SAME object."""
    spec = CONFIGS[config_id]
    return {"workload": spec["workload"].format(nb=nb) + (" -- FULL fine-tuning (use_lora=False), not LoRA"
                                                           if full_finetune else ""),
            "baseline_config": config_id,
            "groups_per_step_all_ranks": nb * world,
            "l2": "per-step working set (4.4 GB weights + activations) exceeds the 126 MB L2; no flush needed",
            "parallelism": f"dp{world} (prompt groups sharded, LoRA-grad all-reduce)"}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        load = [x for x in sm if mx and x > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_sample(steps=1, warmup=0, verbose=False, budget_s=None):
    """Times the oracle (the CPU restatement of the reference algorithm, oracle/) on the host cores on a
    bounded sample of the config-2 workload and extrapolates by the exact step counts:
      per sample = 10 x MMDiT fwd (CFG pair, B=2) + 2 x (fwd + bwd, B=2) + 1 VAE decode + 2 PickScore image fwd
    Returns (samples_per_s list per step, cores, description)."""
    import torch
    from adv_grpo_b200 import weights
    from oracle import clip as clip_o
    from oracle import vae as vae_o
    from oracle.mmdit import MMDiTOracle
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    cfg = weights.SD35_MEDIUM
    t0 = time.perf_counter()
    params = weights.init_mmdit(cfg, seed=0, device="cpu", dtype=torch.bfloat16)
    lora = weights.init_lora(cfg, rank=32, seed=1, perturb_b=0.01)
    lora = {k: (a.requires_grad_(True), b.requires_grad_(True)) for k, (a, b) in lora.items()}
    oracle = MMDiTOracle(params, dict(cfg, dual_layers=set(cfg["dual_layers"])), lora=lora, lora_scale=2.0)
    del params
    vparams = weights.init_vae_decoder(weights.VAE_SD3, seed=2, device="cpu")
    cparams = {k: v.float() for k, v in weights.init_clip(weights.CLIP_H, seed=3, device="cpu").items()}
    ccfg = dict(patch=14, v_layers=32, v_heads=16, t_layers=24, t_heads=16)
    if verbose:
        print(f"[cpu baseline] init {time.perf_counter() - t0:.1f}s, {cores} threads", file=sys.stderr)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 16, RES // 8, RES // 8, generator=g)
    t = torch.full((2,), 500.0)
    ctx = torch.randn(2, 205, 4096, generator=g)
    pooled = torch.randn(2, 2048, generator=g)
    results, step_wall = [], []
    t_start = time.perf_counter()
    for it in range(warmup + steps):
        # bounded run: the CPU timings repeat to ~2 %, so once the wall budget is spent the remaining steps
        # would only repeat the same sample; the executed count is reported in the description
        if budget_s is not None and results and time.perf_counter() - t_start > budget_s:
            break
        with torch.no_grad():
            a = time.perf_counter()
            oracle.forward(x, t, ctx, pooled)
            t_fwd = time.perf_counter() - a
        a = time.perf_counter()
        out = oracle.forward(x, t, ctx, pooled)
        out.square().mean().backward()
        t_fb = time.perf_counter() - a
        with torch.no_grad():
            a = time.perf_counter()
            vae_o.decode_latents_to_image(vparams, x[:1])
            t_vae = time.perf_counter() - a
            a = time.perf_counter()
            clip_o.image_features(cparams, ccfg, torch.randn(1, 3, 224, 224, generator=g))
            t_clip = time.perf_counter() - a
        per_sample = T_STEPS * t_fwd + T_TRAIN * t_fb + t_vae + 2 * t_clip
        if verbose:
            print(f"[cpu baseline] fwd(B=2) {t_fwd:.2f}s  fwd+bwd(B=2) {t_fb:.2f}s  vae {t_vae:.2f}s  clip {t_clip:.2f}s "
                  f"-> {per_sample:.1f}s/sample", file=sys.stderr)
        if it >= warmup:
            results.append(1.0 / per_sample)
            step_wall.append(t_fwd + t_fb + t_vae + t_clip)        # what this bounded step actually took on the host
    desc = ("oracle (torch fp32 restatement of the reference path) on host cores: timed 1 MMDiT fwd at B=2 (one CFG "
            "pair), 1 fwd+bwd at B=2, 1 VAE decode, 1 PickScore image fwd at SD3.5-M/512px shapes; per-sample time = "
            "10*fwd + 2*(fwd+bwd) + vae + 2*clip (extrapolated by step counts)")
    if len(results) < steps:
        desc += f"; {len(results)} of {steps} requested steps executed within the {budget_s:.0f} s wall budget"
    cpu_reference_sample.last_step_wall_s = step_wall
    return results, cores, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config != 2:
        print(json.dumps({"impl": "reference", "unavailable": "the CPU arm is bounded to BASELINE config 2 (config "
                          f"{args.config} needs hours of host time per sample)"}))
        return
    vals, cores, desc = cpu_reference_sample(steps=args.steps, warmup=min(args.warmup, 1), verbose=True,
                                             budget_s=args.cpu_budget)
    v = statistics.mean(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "steps_executed": len(vals),
            # ms_per_step = the measured wall time of one bounded step (the timed sample); `value` extrapolates that sample to
            # a whole GRPO sample by the step counts, ms_per_sample_extrapolated = 1000 / value
            "warmup": args.warmup, "ms_per_step": 1000.0 * statistics.mean(cpu_reference_sample.last_step_wall_s),
            "ms_per_sample_extrapolated": 1000.0 / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # the same config object as the GPU arm prints for this launch (the CPU arm is one process on rank 0's host cores;
            # its per-sample time does not depend on how many groups a step holds)
            "config": config_dict(2, CONFIGS[2]["groups"], max(int(os.environ.get("WORLD_SIZE", "1")), 1)),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def _ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel, extracted from the committed
    `ncu --set full` captures by scripts/ncu_traffic.py into profiles/ncu_traffic.json (None when no capture exists)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            ent = json.load(f).get(key)
        return (ent["dram_bytes"], ent["source"]) if ent else (None, None)
    except (OSError, ValueError, KeyError):
        return None, None


def _ocr_stub(img):
    """Deterministic recogniser for the host OCR plugin (SURVEY 8d allows a stub: PaddleOCR is not installable here):
    the recognised text is a function of the image content, so rewards differ between samples of a group."""
    v = int(img[::64, ::64].astype("int64").sum()) % 7
    return [("text" + "x" * v, 0.9)]


def run_ours(args):
    import torch
    import torch.distributed as dist
    from adv_grpo_b200 import _lib, ops, rewards, weights
    from adv_grpo_b200.config import load_config
    from adv_grpo_b200.pipeline import StableDiffusion3Pipeline
    from adv_grpo_b200.trainer import GRPOTrainer, SyntheticTextEmbedder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    _lib.call("advgrpo_device_check", local)
    torch.backends.cuda.matmul.allow_tf32 = True        # reference: train_sd3_fast_pickscore.py:537-538
    torch.backends.cudnn.allow_tf32 = True

    spec = CONFIGS[args.config]
    res, t_steps, gsz = spec["res"], spec["steps"], spec["G"]
    nb = spec["groups"]
    if args.scaling == "strong":
        total = STRONG_GROUPS if args.config != 4 else 8
        if total % world:
            raise SystemExit(f"--scaling strong needs a world size dividing {total}")
        nb = total // world

    pipe = StableDiffusion3Pipeline.from_seed(weights.SD35_MEDIUM, weights.VAE_SD3, device=dev, seed=0)
    cfg = load_config(spec["preset"])
    cfg.resolution = res
    cfg.sample.num_steps = t_steps
    cfg.sample.mini_num_image_per_prompt = gsz
    cfg.sample.num_image_per_prompt = gsz                 # north_star: every rank rolls out its own G-sample groups
    cfg.sample.shard_groups_across_ranks = False
    cfg.sample.num_batches_per_epoch = nb
    cfg.train.gradient_accumulation_steps = max(nb // 2, 1)   # 2 optimizer steps per step, like the reference epoch
    cfg.train_d = spec["train_d"]
    if args.full_finetune:
        cfg.use_lora = False
    scorer = head = None
    if args.config in (2, 5):
        from adv_grpo_b200.pickscore_scorer import PickScoreScorer
        scorer = PickScoreScorer(device=dev, dtype=torch.bfloat16)
    elif args.config == 3:
        from adv_grpo_b200.dinov2 import DINOHead, DinoV2
        scorer = DinoV2(weights.init_dinov2(weights.DINOV2_B, seed=4, device=dev, dtype=torch.bfloat16), weights.DINOV2_B,
                        device=dev)
        head = DINOHead(in_dim=scorer.num_features).to(dev)
        cfg.d_times = 2                                   # alternate G / D epochs so a short run times both
    elif args.config == 4:
        rewards.OCR_KWARGS["recognizer"] = _ocr_stub
        cfg.train.micro_batch = 8
        cfg.train.gradient_accumulation_steps = 1
    if args.config == 4:
        prompts = [f'a storefront sign that says "text{i % 5}" number {i}' for i in range(99)]
    else:
        prompts = [f"synthetic prompt {i}" for i in range(99)]

    class HostStagedEmbedder(SyntheticTextEmbedder):
        """e2e leg: every call copies the prompt embeddings from pinned host memory."""

        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.cache, self.h2d = {}, 0

        def __call__(self, idx):
            if idx not in self.cache:
                g = torch.Generator().manual_seed(1000 + int(idx))
                e = torch.randn(1, self.n_tokens, self.joint_dim, generator=g).bfloat16().pin_memory()
                p = torch.randn(1, self.pooled_dim, generator=g).bfloat16().pin_memory()
                self.cache[idx] = (e, p)
            e, p = self.cache[idx]
            self.h2d += e.numel() * 2 + p.numel() * 2
            return e.to(self.device, non_blocking=True), p.to(self.device, non_blocking=True)

    class DeviceEmbedder(SyntheticTextEmbedder):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.cache = {}

        def __call__(self, idx):
            if idx not in self.cache:
                self.cache[idx] = super().__call__(idx)
            return self.cache[idx]

    host_refs, dev_refs = {}, {}
    counters = {"h2d": 0}

    def ref_host(idx, n, size):
        if idx not in host_refs:
            g = torch.Generator().manual_seed(11 + int(idx))
            host_refs[idx] = torch.rand(n, 3, size, size, generator=g).pin_memory()
        counters["h2d"] += host_refs[idx].numel() * 4
        return host_refs[idx].to(dev, non_blocking=True)

    def ref_dev(idx, n, size):
        if idx not in dev_refs:
            g = torch.Generator().manual_seed(11 + int(idx))
            dev_refs[idx] = torch.rand(n, 3, size, size, generator=g).to(dev)
        return dev_refs[idx]

    emb_dev = DeviceEmbedder(device=dev)
    emb_host = HostStagedEmbedder(device=dev)
    trainer = GRPOTrainer(cfg, pipe, prompts, scorer=scorer, head=head, embedder=emb_dev, device=dev,
                          reference_image_fn=ref_dev)
    if spec["train_d"]:
        trainer.d_schedule = lambda epoch: epoch % 2 == 1          # G, D, G, D, ... (both updates inside a short run)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k, read_back):
        d2h = 0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        for _ in range(k):
            info = trainer.run_epoch()
            if read_back:
                keys = [kk for kk in ("loss", "reward_mean", "approx_kl", "d_loss") if kk in info]
                vals = torch.stack([info[kk].float().reshape(()) for kk in keys]).cpu()
                d2h += vals.numel() * 4
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), _lib.launch_count() - n0, d2h

    # ---- warm-up (also captures the CUDA graphs and pre-stages every prompt this run will touch) ----
    n_warm = max(args.warmup, 3)
    if spec["train_d"]:
        n_warm += n_warm % 2                       # keep the G / D alternation aligned with the timed region
    trainer.embedder, trainer.reference_image_fn = emb_host, ref_host
    timed(1, True)
    trainer.embedder, trainer.reference_image_fn = emb_dev, ref_dev
    timed(n_warm - 1, False)
    # pre-stage the device-resident inputs of the timed region (value leg: inputs already in HBM)
    for ep in range(trainer.epoch, trainer.epoch + 2 * args.steps + 2):
        for i in range(nb):
            idx = trainer.sampler.indices_for_epoch(ep * nb + i)[rank][0]
            emb_dev(idx), ref_dev(idx, gsz, res), emb_host(idx), ref_host(idx, gsz, res)
    emb_host.h2d, counters["h2d"] = 0, 0

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_dev, launches, _ = timed(args.steps, False)
    trainer.embedder, trainer.reference_image_fn = emb_host, ref_host
    ms_e2e, _, d2h = timed(args.steps, True)
    clk = clocks.stop() if rank == 0 else None
    # phase split of ONE further step (device-resident inputs, CUDA events between the phases; explains `value`,
    # is not part of it)
    trainer.embedder, trainer.reference_image_fn = emb_dev, ref_dev
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    barrier()
    ev[0].record()
    smp = trainer.sample_epoch()
    ev[1].record()
    adv = trainer.compute_advantages(smp)
    ev[2].record()
    trainer.train_generator(smp, adv)
    ev[3].record()
    if spec["train_d"]:
        trainer.discriminator_step(smp)
    ev[4].record()
    trainer.epoch += 1
    barrier()
    phases = {"rollout_decode_score": ev[0].elapsed_time(ev[1]), "advantages": ev[1].elapsed_time(ev[2]),
              "update": ev[2].elapsed_time(ev[3])}
    if spec["train_d"]:
        phases["discriminator_update"] = ev[3].elapsed_time(ev[4])
        if "d_step_ms" in trainer.last_info:                       # ADVGRPO_TRACE_DSTEP=1: backward / grad sync / optimizer
            phases["discriminator_update_split"] = trainer.last_info["d_step_ms"]
    del smp, adv
    samples = world * nb * gsz * args.steps
    value = samples / (ms_dev / 1e3)
    e2e_value = samples / (ms_e2e / 1e3)
    h2d_per_step = (emb_host.h2d + counters["h2d"]) / args.steps

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (tcgen05 GEMM) and of the attention kernels, measured live ----
    peak_burst, peak_sustained, hbm, src = _peaks()

    def time_kernel(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    n_img = (res // 16) ** 2
    bsz = 2 * min(gsz, 8)                            # CFG batch of one launch (config 4 replays in micro-batches of 8)
    M, K, N, K2 = bsz * n_img, 1536, 4608, 128
    A = torch.randn(M, K, device=dev).bfloat16()
    W = torch.randn(N, K, device=dev).bfloat16()
    A2 = torch.randn(M, K2, device=dev).bfloat16()
    W2 = torch.randn(N, K2, device=dev).bfloat16()
    bias = torch.randn(N, device=dev).bfloat16()
    ms_gemm = time_kernel(lambda: ops.gemm(A, W, bias=bias, a2=A2, w2=W2))
    gemm_flops = 2.0 * M * N * (K + K2)
    S = n_img + 205
    qkv = torch.randn(bsz, S, 3, 24, 64, device=dev).bfloat16()
    ms_attn = time_kernel(lambda: ops.attention_fwd(qkv, want_lse=False, split=n_img))
    attn_flops = 4.0 * bsz * 24 * S * S * 64
    out, lse = ops.attention_fwd(qkv)
    dout = torch.randn_like(out)
    ms_attn_bwd = time_kernel(lambda: ops.attention_bwd(qkv, out, dout, lse), iters=10)
    achieved = gemm_flops / ms_gemm / 1e9
    attn_alg_bytes = 2.0 * bsz * S * 24 * 64 * 4          # q, k, v read + o written, bf16
    tr_gemm, tr_gemm_src = _ncu_traffic("gemm_qkv_lora")
    tr_af, tr_af_src = _ncu_traffic("attn_fwd")
    tr_ab, tr_ab_src = _ncu_traffic("attn_bwd")
    roofline = {"bound": "tensor", "kernel": "gemm_kernel<256> (fused QKV projection + LoRA second product, "
                f"M={M} N={N} K={K}+{K2})", "achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s",
                "frac": achieved / peak_burst,
                # dram bytes per launch from the committed ncu --set full capture (profiles/ncu_traffic.json; captured at
                # the config-2 shape); algorithmic bytes = A + W + A2 + W2 + C in bf16
                "traffic": tr_gemm if args.config in (2, 3, 5) else None, "traffic_source": tr_gemm_src,
                "algorithmic_bytes": 2.0 * (M * K + N * K + M * K2 + N * K2 + M * N),
                "peak_source": f"{src} burst bf16 (kernel timed alone)",
                # north_star's graded contraction: the joint text-image attention of one MMDiT block
                "attention_fwd": {"bound": "tensor", "kernel": f"attn_fwd_pair_kernel (B={bsz} H=24 S={S} D=64, image/text "
                                  "output split)", "achieved": attn_flops / ms_attn / 1e9, "peak": peak_burst,
                                  "unit": "TFLOP/s", "frac": attn_flops / ms_attn / 1e9 / peak_burst,
                                  "traffic": tr_af if args.config in (2, 3, 5) else None, "traffic_source": tr_af_src,
                                  "algorithmic_bytes": attn_alg_bytes},
                "attention_bwd": {"bound": "tensor", "kernel": f"attn_bwd_kernel (B={bsz} H=24 S={S} D=64)",
                                  "achieved": 2.5 * attn_flops / ms_attn_bwd / 1e9, "peak": peak_burst, "unit": "TFLOP/s",
                                  "frac": 2.5 * attn_flops / ms_attn_bwd / 1e9 / peak_burst,
                                  "traffic": tr_ab if args.config in (2, 3, 5) else None, "traffic_source": tr_ab_src,
                                  "algorithmic_bytes": 2.0 * attn_alg_bytes + 2.0 * bsz * S * 24 * 64 * 3}}
    # FLOPs per GRPO sample: SURVEY 8d convention (2 T F + 2 T_train 3 F) and the executed count (LoRA-only backward:
    # dX GEMMs = the forward GEMM share, attention backward = 2.5 x attention forward)
    f_fwd = {512: 2.224, 1024: 10.92}.get(res, 2.224)
    attn_share = {512: 0.138, 1024: 0.373}.get(res, 0.138)
    conv_flops = 2 * t_steps * f_fwd + 2 * T_TRAIN * 3 * f_fwd
    exec_flops = 2 * t_steps * f_fwd + 2 * T_TRAIN * (f_fwd + (1 - attn_share) * f_fwd + 2.5 * attn_share * f_fwd)
    kernels = {
        "attn_shape": f"B={bsz} H=24 S={S} D=64",
        "tflop_per_sample_convention": conv_flops, "tflop_per_sample_executed": exec_flops,
        "step_tflops": conv_flops * value / world, "step_frac_of_sustained": conv_flops * value / world / peak_sustained,
        "step_tflops_executed": exec_flops * value / world,
        "step_frac_of_sustained_executed": exec_flops * value / world / peak_sustained,
    }
    del A, W, A2, W2, qkv, out, dout
    torch.cuda.empty_cache()

    cpu = None
    if world == 1 and not args.no_cpu_baseline and args.config == 2:
        vals, cores, desc = cpu_reference_sample(steps=1, warmup=0)
        cpu = {"value": vals[0], "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_dict(args.config, nb, world, args.full_finetune),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_per_step,
                    "d2h_bytes_per_step": d2h / args.steps},
            "gpu_launches": launches, "clocks": clk, "roofline": roofline, "kernels": kernels,
            "phases_ms": phases}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json config: 2 PickScore GRPO (default, the headline), 3 DINOv2-patch adversarial, "
                         "4 1024x1024 / 20 steps / G=16 PickScore+OCR, 5 generator + PickScore discriminator co-update")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: fixed groups per rank (default); strong: 16 groups per step split over the ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--full-finetune", action="store_true",
                    help="config.use_lora = False (SURVEY 8f-4): every transformer weight trains; not a BASELINE config")
    ap.add_argument("--cpu-budget", type=float, default=150.0,
                    help="--impl reference: wall-clock bound (s) of the timed CPU samples (init excluded)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

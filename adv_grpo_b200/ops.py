"""Torch-facing wrappers of the libadvgrpo_b200 C ABI.

PyTorch is plumbing here: device memory, streams and autograd bookkeeping.  Every
numerical result on the hot path comes from the sm_100a kernels; there is no
PyTorch/CPU fallback -- non-CUDA tensors raise.
"""
import math
import os
import threading

import torch

from . import _lib

_ws_cache = {}
_ws_lock = threading.Lock()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.AdvGrpoError("adv_grpo_b200 ops run on CUDA tensors only (no CPU fallback)")


def _workspace(tag, nbytes, device):
    """Per (op, device, stream, thread) scratch buffer: concurrent streams never share one, and neither do two host
    threads enqueueing multi-kernel entry points on the same stream (ctypes releases the GIL during the call)."""
    key = (tag, device.index, _stream(), threading.get_ident())
    with _ws_lock:
        buf = _ws_cache.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            _ws_cache[key] = buf
    return buf


def _bf16c(t):
    if t.dtype != torch.bfloat16:
        raise _lib.AdvGrpoError(f"expected bfloat16 tensor, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


# --------------------------------------------------------------------------- A4 + A5
SDE_FLOW_CPS, SDE_FLOW_SDE = 0, 1      # sde_step_with_logprob_new (sde.py:77-139) / sde_step_with_logprob (sde.py:13-73)


def cfg_sde_step_logprob(v_uncond, v_text, x, timesteps, sched_timesteps, sigmas, guidance_scale,
                         noise_level, prev_sample=None, noise=None, seed=0, offset=0,
                         want_mean=False, want_prev=True, variant=SDE_FLOW_CPS):
    """Fused CFG + Flow-CPS (or Flow-SDE) step + log-prob. Returns (prev_sample_bf16|None, log_prob f32[B],
    prev_sample_mean f32|None, std_dev_t f32[B])."""
    _need_cuda(v_text, x)
    v_text, x = _bf16c(v_text), _bf16c(x)
    v_uncond = None if v_uncond is None else _bf16c(v_uncond)
    B = x.shape[0]
    n = x.numel() // B
    dev = x.device
    timesteps = timesteps.to(device=dev, dtype=torch.float32).contiguous().reshape(-1)
    sched_timesteps = sched_timesteps.to(device=dev, dtype=torch.float32).contiguous()
    sigmas = sigmas.to(device=dev, dtype=torch.float32).contiguous()
    T = sched_timesteps.numel()
    if sigmas.numel() != T + 1:
        raise _lib.AdvGrpoError("sigmas must have len(sched_timesteps) + 1 entries")
    prev_in = None if prev_sample is None else _bf16c(prev_sample)
    if noise is not None:
        noise = noise.to(device=dev, dtype=torch.float32).contiguous()
    prev_out = torch.empty_like(x) if (prev_in is None and want_prev) else None
    mean_out = torch.empty(x.shape, dtype=torch.float32, device=dev) if want_mean else None
    logp = torch.empty(B, dtype=torch.float32, device=dev)
    std = torch.empty(B, dtype=torch.float32, device=dev)
    ws_bytes = _lib.query("advgrpo_sde_step_workspace_bytes", B, n)
    ws = _workspace("sde", ws_bytes, dev)
    _lib.call("advgrpo_cfg_sde_step_logprob_variant", _ptr(v_uncond), _ptr(v_text), _ptr(x), _ptr(prev_in),
              _ptr(noise), _ptr(timesteps), timesteps.numel(), _ptr(sched_timesteps), _ptr(sigmas), T,
              _ptr(prev_out), _ptr(mean_out), _ptr(logp), _ptr(std), B, n, float(guidance_scale),
              float(noise_level), int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1), _ptr(ws),
              ws.numel(), int(variant), _stream())
    return prev_out, logp, mean_out, std


class _SdeLogProbReplay(torch.autograd.Function):
    """log_prob of a stored transition under the current model output; differentiable w.r.t.
    the (CFG-batched) transformer output (train_sd3_fast_pickscore.py:233-267).  With `mean_ref`
    (prev_sample_mean of the adapter-disabled forward) it also returns the per-sample KL regulariser
    kl[b] = mean((mu - mu_ref)^2) of train_sd3_fast_pickscore.py:1124-1128, differentiable as well."""

    @staticmethod
    def forward(ctx, noise_pred, x, prev_sample, timesteps, sched_timesteps, sigmas, guidance_scale,
                noise_level, cfg, want_mean, mean_ref, variant=SDE_FLOW_CPS):
        noise_pred = _bf16c(noise_pred)
        if cfg:
            vu, vt = noise_pred.chunk(2)
        else:
            vu, vt = None, noise_pred
        _, logp, mean, std = cfg_sde_step_logprob(vu, vt, x, timesteps, sched_timesteps, sigmas,
                                                  guidance_scale, noise_level, prev_sample=prev_sample,
                                                  want_mean=want_mean or mean_ref is not None, variant=variant)
        kl = torch.empty(0, device=x.device)
        if mean_ref is not None:
            mean_ref = mean_ref.to(torch.float32).contiguous()
            kl = ((mean - mean_ref) ** 2).reshape(x.shape[0], -1).mean(1)
        ctx.save_for_backward(noise_pred, x, prev_sample, timesteps, sched_timesteps, sigmas, mean_ref)
        ctx.meta = (float(guidance_scale), float(noise_level), bool(cfg), int(variant))
        ctx.mark_non_differentiable(std)
        if mean is None:
            mean = torch.empty(0, device=x.device)
        ctx.mark_non_differentiable(mean)
        return logp, mean, std, kl

    @staticmethod
    def backward(ctx, g_logp, _gm, _gs, g_kl):
        noise_pred, x, prev, timesteps, sched_t, sigmas, mean_ref = ctx.saved_tensors
        gs, nl, cfg, variant = ctx.meta
        B = x.shape[0]
        n = x.numel() // B
        dev = x.device
        grad = torch.empty_like(noise_pred)
        if cfg:
            vu, vt = noise_pred.chunk(2)
            gvu, gvt = grad[:B], grad[B:]
        else:
            vu, vt, gvu, gvt = None, noise_pred, None, grad
        x, prev = _bf16c(x), _bf16c(prev)
        timesteps = timesteps.to(device=dev, dtype=torch.float32).contiguous().reshape(-1)
        g_logp = (torch.zeros(B, device=dev) if g_logp is None else g_logp).to(torch.float32).contiguous()
        if mean_ref is not None:
            g_kl = (torch.zeros(B, device=dev) if g_kl is None else g_kl).to(torch.float32).contiguous()
        else:
            g_kl = None
        _lib.call("advgrpo_cfg_sde_logprob_bwd_variant", _ptr(vu), _ptr(vt), _ptr(x), _ptr(prev), _ptr(timesteps),
                  timesteps.numel(), _ptr(sched_t), _ptr(sigmas), sched_t.numel(), _ptr(g_logp), _ptr(g_kl),
                  _ptr(mean_ref), _ptr(gvu), _ptr(gvt), B, n, gs, nl, variant, _stream())
        return grad, None, None, None, None, None, None, None, None, None, None, None


def sde_logprob_replay(noise_pred, x, prev_sample, timesteps, sched_timesteps, sigmas, guidance_scale,
                       noise_level, cfg=True, want_mean=False, mean_ref=None, variant=SDE_FLOW_CPS):
    """Returns (log_prob, prev_sample_mean or None, std_dev_t), plus the per-sample KL term when `mean_ref` is given."""
    dev = x.device
    sched_timesteps = sched_timesteps.to(device=dev, dtype=torch.float32).contiguous()
    sigmas = sigmas.to(device=dev, dtype=torch.float32).contiguous()
    logp, mean, std, kl = _SdeLogProbReplay.apply(noise_pred, _bf16c(x), _bf16c(prev_sample), timesteps,
                                                  sched_timesteps, sigmas, guidance_scale, noise_level, cfg,
                                                  want_mean, mean_ref, variant)
    if mean_ref is not None:
        return logp, (mean if want_mean else None), std, kl
    return logp, (mean if want_mean else None), std


# --------------------------------------------------------------------------- A9
ADV_MODES = {"grpo": 0, "rwr": 1, "sft": 2, "dpo": 3}


def group_advantage(rewards, group_keys, global_std=True, want_stats=True, mode="grpo"):
    """rewards f32 [N] or [N,T] (CUDA); group_keys int64 [N] or [N,L] (CUDA); mode = the `type` of
    PerPromptStatTracker.update.  Returns (advantages f64 same shape as rewards, stats f64[4] or None)."""
    _need_cuda(rewards, group_keys)
    squeeze = rewards.dim() == 1
    r = rewards.to(torch.float32).reshape(rewards.shape[0], -1).contiguous()
    N, T = r.shape
    keys = group_keys.to(torch.int64).reshape(N, -1).contiguous()
    adv = torch.empty((N, T), dtype=torch.float64, device=r.device)
    stats = torch.zeros(4, dtype=torch.float64, device=r.device) if want_stats else None
    ws_bytes = _lib.query("advgrpo_group_advantage_workspace_bytes", N, T)
    ws = _workspace("adv", ws_bytes, r.device)
    _lib.call("advgrpo_group_advantage_mode", _ptr(r), _ptr(keys), keys.shape[1], N, T, int(bool(global_std)),
              ADV_MODES[mode], _ptr(adv), _ptr(stats), _ptr(ws), ws.numel(), _stream())
    return (adv[:, 0] if squeeze else adv), stats


# --------------------------------------------------------------------------- A12
def clip_adamw(param, grad, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
               max_grad_norm=1.0, zero_grad=True, want_norm=True):
    """In-place global-norm clip + AdamW step (+ gradient clear) on one flat fp32 CUDA tensor
    (train_sd3_fast_pickscore.py:1165-1171).  Returns the pre-clip gradient norm (f32 [1], device) or None."""
    _need_cuda(param, grad, exp_avg, exp_avg_sq)
    n = param.numel()
    for t in (param, grad, exp_avg, exp_avg_sq):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n:
            raise ValueError("clip_adamw: param / grad / exp_avg / exp_avg_sq must be contiguous fp32 of one size")
    norm = torch.zeros(1, dtype=torch.float32, device=param.device) if (want_norm and max_grad_norm > 0) else None
    ws_bytes = _lib.query("advgrpo_clip_adamw_workspace_bytes", n)
    ws = _workspace("adamw", ws_bytes, param.device)
    _lib.call("advgrpo_clip_adamw", _ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), n, float(lr),
              float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step), float(max_grad_norm),
              int(bool(zero_grad)), _ptr(norm), _ptr(ws), ws.numel(), _stream())
    return norm


# --------------------------------------------------------------------------- A11
class _GrpoClipLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_prob, old_log_prob, advantages, clip_range, adv_clip_max, grad_scale):
        _need_cuda(log_prob, old_log_prob, advantages)
        lp = log_prob.to(torch.float32).contiguous()
        lpo = old_log_prob.to(torch.float32).contiguous()
        adv = advantages.to(torch.float64)
        B = lp.numel()
        out = torch.empty(6, dtype=torch.float64, device=lp.device)
        g = torch.empty(B, dtype=torch.float32, device=lp.device)
        stride = adv.stride(0) if adv.dim() == 1 else 1
        if adv.dim() != 1:
            adv = adv.reshape(-1).contiguous()
        _lib.call("advgrpo_grpo_clip_loss", _ptr(lp), _ptr(lpo), _ptr(adv), stride, B, float(clip_range),
                  float(adv_clip_max), float(grad_scale), _ptr(out), _ptr(g), _stream())
        ctx.save_for_backward(g)
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, g_loss, _g_out):
        (g,) = ctx.saved_tensors
        return (g * g_loss.to(torch.float32)), None, None, None, None, None


def grpo_clip_loss(log_prob, old_log_prob, advantages, clip_range, adv_clip_max, grad_scale=1.0):
    """Returns (loss f64 scalar, stats f64[6] = loss, approx_kl, clipfrac, clipfrac_gt_one,
    clipfrac_lt_one, policy_loss).  `grad_scale` folds the gradient-accumulation divisor in."""
    return _GrpoClipLoss.apply(log_prob, old_log_prob, advantages, clip_range, adv_clip_max, grad_scale)


# --------------------------------------------------------------------------- adaLN LayerNorm-modulate
class _LnModulate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, shift, scale, shift2, scale2, eps):
        _need_cuda(x)
        x = _bf16c(x)
        B, S, D = x.shape
        stride = shift.stride(0)
        for m in (shift, scale, shift2, scale2):
            if m is not None and (m.stride(-1) != 1 or m.stride(0) != stride or m.dtype != torch.bfloat16):
                raise _lib.AdvGrpoError("modulation chunks must be bf16 row views of one [B, k*D] matrix")
        y = torch.empty_like(x)
        y2 = torch.empty_like(x) if shift2 is not None else None
        _lib.call("advgrpo_ln_modulate_fwd", _ptr(x), _ptr(shift), _ptr(scale), _ptr(shift2), _ptr(scale2),
                  stride, _ptr(y), _ptr(y2), B, S, D, float(eps), _stream())
        ctx.save_for_backward(x, scale, scale2)
        ctx.eps = float(eps)
        if y2 is None:
            return y
        return y, y2

    @staticmethod
    def backward(ctx, dy, dy2=None):
        x, scale, scale2 = ctx.saved_tensors
        B, S, D = x.shape
        dy = _bf16c(dy)
        dy2 = None if dy2 is None else _bf16c(dy2)
        dx = torch.empty_like(x)
        _lib.call("advgrpo_ln_modulate_bwd", _ptr(x), _ptr(scale), _ptr(scale2), scale.stride(0), _ptr(dy),
                  _ptr(dy2), _ptr(dx), 0, B, S, D, ctx.eps, _stream())
        return dx, None, None, None, None, None


def ln_modulate(x, shift, scale, shift2=None, scale2=None, eps=1e-6):
    """LayerNorm(no affine)(x) * (1 + scale[:, None]) + shift[:, None]; optional second modulation.
    Gradient flows to x only (the modulation comes from frozen adaLN weights)."""
    return _LnModulate.apply(x, shift, scale, shift2, scale2, eps)


def _layer_norm_fwd(x, weight, bias, eps):
    D = x.shape[-1]
    _need_cuda(x, weight, bias)
    if D % 256 or D > 2048:
        raise ValueError(f"layer_norm: width {D} is not supported by the native kernel (multiple of 256, <= 2048); "
                         "there is no PyTorch fallback")
    x, weight, bias = _bf16c(x), _bf16c(weight), _bf16c(bias)
    y = torch.empty_like(x)
    _lib.call("advgrpo_layer_norm_affine", _ptr(x), _ptr(weight), _ptr(bias), _ptr(y), x.numel() // D, D, float(eps), _stream())
    return y


class _LayerNormAffine(torch.autograd.Function):
    """Affine LayerNorm with a native backward (dx, and d weight / d bias when they are trainable): the LayerNorms of the
    CLIP vision blocks the PickScore discriminator step trains (train_sd3_fast_pickscore.py:1016-1029)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x = _bf16c(x)
        ctx.save_for_backward(x, weight)
        ctx.eps = float(eps)
        ctx.param_dtype = weight.dtype
        return _layer_norm_fwd(x, weight.detach().to(torch.bfloat16), bias.detach().to(torch.bfloat16), eps)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        D = x.shape[-1]
        rows = x.numel() // D
        dy = _bf16c(dy)
        want_p = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx = torch.empty_like(x)
        dwb = torch.empty((2, D), dtype=torch.float32, device=x.device) if want_p else None
        ws = None
        if want_p:
            ws = _workspace("ln_bwd", _lib.query("advgrpo_layer_norm_affine_bwd_workspace_bytes", rows, D), x.device)
        _lib.call("advgrpo_layer_norm_affine_bwd", _ptr(x), _ptr(_bf16c(weight.detach().to(torch.bfloat16))), _ptr(dy), _ptr(dx),
                  _ptr(dwb), None if dwb is None else dwb[1].data_ptr(), rows, D, ctx.eps, _ptr(ws),
                  0 if ws is None else ws.numel(), _stream())
        dw = dwb[0].to(ctx.param_dtype) if ctx.needs_input_grad[1] else None
        db = dwb[1].to(ctx.param_dtype) if ctx.needs_input_grad[2] else None
        return (dx if ctx.needs_input_grad[0] else None), dw, db, None


class _LnModulateFull(torch.autograd.Function):
    """`ln_modulate` for full fine-tuning: the same forward kernel, and a backward that also returns the gradients of the
    shift / scale vectors (they come from TRAINABLE adaLN linears there): dx on the native kernel, d shift = sum_t dy and
    d scale = sum_t dy * xhat per sample on the row-statistics + segmented column-sum kernels (csrc/heads.cu)."""

    @staticmethod
    def forward(ctx, x, shift, scale, eps):
        x = _bf16c(x)
        B, S, D = x.shape
        y = torch.empty_like(x)
        _lib.call("advgrpo_ln_modulate_fwd", _ptr(x), _ptr(shift), _ptr(scale), None, None, shift.stride(0), _ptr(y), None,
                  B, S, D, float(eps), _stream())
        ctx.save_for_backward(x, scale)
        ctx.eps = float(eps)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, scale = ctx.saved_tensors
        B, S, D = x.shape
        dy = _bf16c(dy)
        dx = torch.empty_like(x)
        _lib.call("advgrpo_ln_modulate_bwd", _ptr(x), _ptr(scale), None, scale.stride(0), _ptr(dy), None, _ptr(dx), 0, B, S, D,
                  ctx.eps, _stream())
        grads = torch.empty((2, B, D), dtype=torch.float32, device=x.device)          # [d shift | d scale]
        ws = _workspace("ln_mod_grads", _lib.query("advgrpo_ln_modulation_grads_workspace_bytes", B, S, D), x.device)
        _lib.call("advgrpo_ln_modulation_grads", _ptr(x), _ptr(dy), _ptr(grads), grads[1].data_ptr(), B, S, D, ctx.eps,
                  _ptr(ws), ws.numel(), _stream())
        return dx, grads[0].to(torch.bfloat16), grads[1].to(torch.bfloat16), None


def ln_modulate_full(x, shift, scale, eps=1e-6):
    """LayerNorm(no affine)(x) * (1 + scale[:, None]) + shift[:, None] with gradients to x, shift AND scale."""
    for m in (shift, scale):
        if m.stride(-1) != 1 or m.dtype != torch.bfloat16 or m.stride(0) != shift.stride(0):
            raise _lib.AdvGrpoError("modulation chunks must be bf16 row views of one [B, k*D] matrix")
    return _LnModulateFull.apply(x, shift, scale, eps)


class _QkNormConcatFull(torch.autograd.Function):
    """`qk_norm_concat` for full fine-tuning: native forward and activation gradients; the gradients of the four per-head
    RMSNorm weight vectors (d w[c] = sum over tokens and heads of dy * xhat) as torch reductions."""

    @staticmethod
    def forward(ctx, qkv_img, qkv_txt, wq_img, wk_img, wq_txt, wk_txt, H, D, eps):
        out = _QkNormConcat.forward(ctx, qkv_img, qkv_txt, wq_img, wk_img, wq_txt, wk_txt, H, D, eps)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv_img, qkv_txt, wq_img, wk_img, wq_txt, wk_txt = ctx.saved_tensors
        H, D, eps = ctx.meta
        d_img, d_txt = _QkNormConcat.backward(ctx, dout)[:2]
        B, S_img, _ = qkv_img.shape

        def wgrad(src, dy_rows, sec):
            x = src.view(src.shape[0], src.shape[1], 3, H, D)[:, :, sec].float()
            xhat = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
            return (dy_rows[:, :, sec].float() * xhat).sum((0, 1, 2)).to(torch.bfloat16)

        dq_i, dk_i = wgrad(qkv_img, dout[:, :S_img], 0), wgrad(qkv_img, dout[:, :S_img], 1)
        dq_t = dk_t = None
        if qkv_txt is not None:
            dq_t, dk_t = wgrad(qkv_txt, dout[:, S_img:], 0), wgrad(qkv_txt, dout[:, S_img:], 1)
        return d_img, d_txt, dq_i, dk_i, dq_t, dk_t, None, None, None


def qk_norm_concat_full(qkv_img, qkv_txt, wq_img, wk_img, wq_txt, wk_txt, H, D=64, eps=1e-6):
    return _QkNormConcatFull.apply(qkv_img, qkv_txt, wq_img, wk_img, wq_txt, wk_txt, H, D, eps)


def layer_norm(x, weight, bias, eps):
    """Affine LayerNorm over the last dim of the reward towers' blocks (bf16, width a multiple of 256) on the
    `ln_modulate` kernel; differentiable (native backward) when the input or the parameters require a gradient."""
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or bias.requires_grad):
        return _LayerNormAffine.apply(x, weight, bias, eps)
    return _layer_norm_fwd(x.contiguous(), weight.detach(), bias.detach(), eps)


# --------------------------------------------------------------------------- q/k RMSNorm + concat
class _QkNormConcat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv_img, qkv_txt, wq_img, wk_img, wq_txt, wk_txt, H, D, eps):
        _need_cuda(qkv_img)
        qkv_img = _bf16c(qkv_img)
        qkv_txt = None if qkv_txt is None else _bf16c(qkv_txt)
        B, S_img, _ = qkv_img.shape
        S_txt = 0 if qkv_txt is None else qkv_txt.shape[1]
        out = torch.empty((B, S_img + S_txt, 3, H, D), dtype=torch.bfloat16, device=qkv_img.device)
        _lib.call("advgrpo_qk_norm_concat_fwd", _ptr(qkv_img), _ptr(qkv_txt), _ptr(wq_img), _ptr(wk_img),
                  _ptr(wq_txt), _ptr(wk_txt), _ptr(out), B, S_img, S_txt, H, D, float(eps), _stream())
        ctx.save_for_backward(qkv_img, qkv_txt, wq_img, wk_img, wq_txt, wk_txt)
        ctx.meta = (H, D, float(eps))
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv_img, qkv_txt, wq_img, wk_img, wq_txt, wk_txt = ctx.saved_tensors
        H, D, eps = ctx.meta
        B, S_img, _ = qkv_img.shape
        S_txt = 0 if qkv_txt is None else qkv_txt.shape[1]
        dout = _bf16c(dout)
        d_img = torch.empty_like(qkv_img)
        d_txt = None if qkv_txt is None else torch.empty_like(qkv_txt)
        _lib.call("advgrpo_qk_norm_concat_bwd", _ptr(qkv_img), _ptr(qkv_txt), _ptr(wq_img), _ptr(wk_img),
                  _ptr(wq_txt), _ptr(wk_txt), _ptr(dout), _ptr(d_img), _ptr(d_txt), B, S_img, S_txt, H, D,
                  eps, _stream())
        return d_img, d_txt, None, None, None, None, None, None, None


def qk_norm_concat(qkv_img, qkv_txt, wq_img, wk_img, wq_txt, wk_txt, H, D=64, eps=1e-6):
    """[B,S_img,3HD] (+ [B,S_txt,3HD]) -> joint token-major [B,S,3,H,D] with per-head RMSNorm on q,k."""
    return _QkNormConcat.apply(qkv_img, qkv_txt, wq_img, wk_img, wq_txt, wk_txt, H, D, eps)


# --------------------------------------------------------------------------- attention
def attention_fwd(qkv, scale=None, causal=False, want_lse=True, variant=0, split=0):
    """Returns (out, lse); with 0 < split < S, out is the pair (out[:, :split], out[:, split:]) written
    as two contiguous tensors (image rows / text rows of the joint sequence)."""
    _need_cuda(qkv)
    qkv = _bf16c(qkv)
    B, S, three, H, D = qkv.shape
    assert three == 3
    scale = (1.0 / math.sqrt(D)) if scale is None else scale
    out2 = None
    if split:
        out = torch.empty((B, split, H, D), dtype=torch.bfloat16, device=qkv.device)
        out2 = torch.empty((B, S - split, H, D), dtype=torch.bfloat16, device=qkv.device)
    else:
        out = torch.empty((B, S, H, D), dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty((B, H, S), dtype=torch.float32, device=qkv.device) if want_lse else None
    if variant:
        _lib.call("advgrpo_attn_fwd_variant", _ptr(qkv), _ptr(out), _ptr(lse), B, S, H, D, float(scale),
                  int(causal), int(variant), _stream())
    else:
        _lib.call("advgrpo_attn_fwd", _ptr(qkv), _ptr(out), _ptr(out2), int(split), _ptr(lse), B, S, H, D,
                  float(scale), int(causal), _stream())
    return ((out, out2) if split else out), lse


def attention_fwd_bias(qkv, bias, scale=1.0, want_lse=False):
    """softmax(scale * Q K^T + bias[h]) V with bias f32 [H, S, S] (T5 relative-position attention)."""
    _need_cuda(qkv, bias)
    qkv = _bf16c(qkv)
    B, S, three, H, D = qkv.shape
    bias = bias.to(torch.float32).contiguous()
    assert three == 3 and tuple(bias.shape) == (H, S, S)
    out = torch.empty((B, S, H, D), dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty((B, H, S), dtype=torch.float32, device=qkv.device) if want_lse else None
    _lib.call("advgrpo_attn_fwd_bias", _ptr(qkv), _ptr(bias), _ptr(out), _ptr(lse), B, S, H, D, float(scale), _stream())
    return out, lse


def attention_bwd(qkv, out, dout, lse, scale=None, causal=False):
    qkv, out, dout = _bf16c(qkv), _bf16c(out), _bf16c(dout)
    B, S, _, H, D = qkv.shape
    scale = (1.0 / math.sqrt(D)) if scale is None else scale
    dqkv = torch.empty_like(qkv)
    ws_bytes = _lib.query("advgrpo_attn_bwd_workspace_bytes", B, S, H, D)
    ws = _workspace("attn_bwd", ws_bytes, qkv.device)
    _lib.call("advgrpo_attn_bwd", _ptr(qkv), _ptr(out), _ptr(dout), _ptr(lse), _ptr(dqkv), B, S, H, D,
              float(scale), int(causal), _ptr(ws), ws.numel(), _stream())
    return dqkv


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, scale, causal):
        out, lse = attention_fwd(qkv, scale, causal, want_lse=True)
        ctx.save_for_backward(qkv, out, lse)
        ctx.meta = (scale, causal)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        scale, causal = ctx.meta
        return attention_bwd(qkv, out, dout, lse, scale, causal), None, None


class _AttentionSplit(torch.autograd.Function):
    """Joint attention whose output leaves as the (image rows, text rows) pair; the backward re-joins the
    two incoming gradients with ONE concatenation instead of autograd's zero-fill + slice-copy + add."""

    @staticmethod
    def forward(ctx, qkv, scale, causal, split):
        out, lse = attention_fwd(qkv, scale, causal, want_lse=True)
        ctx.save_for_backward(qkv, out, lse)
        ctx.meta = (scale, causal)
        return out[:, :split].contiguous(), out[:, split:].contiguous()

    @staticmethod
    def backward(ctx, d_img, d_txt):
        qkv, out, lse = ctx.saved_tensors
        scale, causal = ctx.meta
        dout = torch.cat([d_img, d_txt], dim=1)
        return attention_bwd(qkv, out, dout, lse, scale, causal), None, None, None


def attention_split(qkv, split, scale=None, causal=False):
    """Differentiable joint attention returning (out[:, :split], out[:, split:]) as contiguous tensors."""
    return _AttentionSplit.apply(qkv, scale, causal, split)


def attention(qkv, scale=None, causal=False):
    """qkv bf16 [B,S,3,H,D] -> out bf16 [B,S,H,D]; differentiable."""
    if torch.is_grad_enabled() and qkv.requires_grad:
        return _Attention.apply(qkv, scale, causal)
    return attention_fwd(qkv, scale, causal, want_lse=False)[0]


class _AttentionSmall(torch.autograd.Function):
    """softmax(scale q k^T) v for short sequences / odd head sizes with a native backward (csrc/attn_small.cu): the
    attention core of the trainable CLIP-H vision blocks (257 tokens, head_dim 80) in the PickScore discriminator step."""

    @staticmethod
    def forward(ctx, q, k, v, scale, causal):
        q, k, v = _bf16c(q), _bf16c(k), _bf16c(v)
        B, S, H, Dh = q.shape
        o = torch.empty_like(q)
        lse = torch.empty((B, H, S), dtype=torch.float32, device=q.device)
        _lib.call("advgrpo_attn_small_fwd", _ptr(q), _ptr(k), _ptr(v), _ptr(o), _ptr(lse), B, S, H, Dh, float(scale),
                  int(bool(causal)), _stream())
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.meta = (float(scale), int(bool(causal)))
        return o

    @staticmethod
    def backward(ctx, dout):
        q, k, v, o, lse = ctx.saved_tensors
        B, S, H, Dh = q.shape
        dout = _bf16c(dout)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        delta = torch.empty_like(lse)
        _lib.call("advgrpo_attn_small_bwd", _ptr(q), _ptr(k), _ptr(v), _ptr(o), _ptr(dout), _ptr(lse), _ptr(delta), _ptr(dq),
                  _ptr(dk), _ptr(dv), B, S, H, Dh, ctx.meta[0], ctx.meta[1], _stream())
        return dq, dk, dv, None, None


def attention_small(q, k, v, scale=None, causal=False):
    """q, k, v bf16 [B, S, H, Dh] (Dh even, <= 128; S * Dh bounded by shared memory) -> o bf16 [B, S, H, Dh]; differentiable."""
    _need_cuda(q, k, v)
    scale = (1.0 / math.sqrt(q.shape[-1])) if scale is None else scale
    return _AttentionSmall.apply(q, k, v, scale, causal)


# --------------------------------------------------------------------------- GEMM
def set_gemm_variant(v):
    """Test/bench hook: 0 = auto (CTA pairs when the shape allows), 1 = single-CTA tiles only."""
    _lib.load().advgrpo_debug_set_gemm_variant(int(v))


EPI_NONE, EPI_GELU_TANH, EPI_GELU_ERF, EPI_GATE_RESIDUAL, EPI_QUICK_GELU, EPI_GELU_TANH_GRAD = 0, 1, 2, 3, 5, 6
EPI_GELU_ERF_GRAD = 7


def gemm(a, w, bias=None, a2=None, w2=None, epilogue=EPI_NONE, residual=None, gate=None,
         rows_per_gate=1, out=None, preact_out=None):
    """C = epi(a @ w.T (+ a2 @ w2.T) + bias).  a [M,K], w [N,K] bf16 row-major (strided rows ok)."""
    _need_cuda(a, w)
    lead = a.shape[:-1]
    a2d = a.reshape(-1, a.shape[-1])
    if a2d.stride(-1) != 1:
        a2d = a2d.contiguous()
    M, K = a2d.shape
    N = w.shape[0]
    if w.stride(-1) != 1:
        w = w.contiguous()
    c = out if out is not None else torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    c2d = c.reshape(M, N)
    K2 = 0
    if a2 is not None:
        a2 = a2.reshape(M, -1)
        K2 = a2.shape[1]
    r2d = None if residual is None else residual.reshape(M, N)
    _lib.call("advgrpo_gemm_bf16", _ptr(a2d), a2d.stride(0), _ptr(w), w.stride(0), _ptr(a2),
              0 if a2 is None else a2.stride(0), _ptr(w2), 0 if w2 is None else w2.stride(0), K2, _ptr(bias),
              _ptr(c2d), c2d.stride(0), M, N, K, int(epilogue), _ptr(r2d), 0 if r2d is None else r2d.stride(0),
              _ptr(gate), 0 if gate is None else gate.stride(0), int(rows_per_gate), _ptr(preact_out), _stream())
    return c2d.reshape(*lead, N)


def gemm_tn_skinny(a, b, transpose_out=False):
    """a^T @ b over the leading (token) axis: a bf16 [Kt, Ms] (Ms <= 256), b bf16 [Kt, Nb] -> bf16 [Ms, Nb] (or [Nb, Ms]).
    The LoRA weight-gradient products dt^T x and (dy^T t = (t^T dy)^T) on the tcgen05 split-K kernel."""
    _need_cuda(a, b)
    a, b = _bf16c(a), _bf16c(b)
    Kt, Ms = a.shape
    Kt2, Nb = b.shape
    if Kt != Kt2:
        raise _lib.AdvGrpoError(f"gemm_tn_skinny: row counts differ ({Kt} vs {Kt2})")
    out = torch.empty((Nb, Ms) if transpose_out else (Ms, Nb), dtype=torch.bfloat16, device=a.device)
    ws_bytes = _lib.query("advgrpo_gemm_tn_skinny_workspace_bytes", Kt, Ms, Nb)
    ws = _workspace("gemm_tn", ws_bytes, a.device)
    _lib.call("advgrpo_gemm_tn_skinny", _ptr(a), _ptr(b), _ptr(out), Kt, Ms, Nb, int(bool(transpose_out)), _ptr(ws),
              ws.numel(), _stream())
    return out


def gemm_tn(a, b, out=None):
    """a^T @ b over the leading (row) axis for ANY widths: a bf16 [R, M], b bf16 [R, N] -> bf16 [M, N] -- the weight
    gradient dy^T x of a trainable Linear (discriminator step) on the same tcgen05 split-K kernel as the LoRA products."""
    _need_cuda(a, b)
    a, b = _bf16c(a), _bf16c(b)
    R, M = a.shape
    R2, N = b.shape
    if R != R2:
        raise _lib.AdvGrpoError(f"gemm_tn: row counts differ ({R} vs {R2})")
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
    elif out.shape != (M, N) or out.dtype != torch.bfloat16 or not out.is_contiguous():
        raise _lib.AdvGrpoError("gemm_tn: out must be a contiguous bf16 [M, N] tensor")
    _lib.call("advgrpo_gemm_tn_skinny", _ptr(a), _ptr(b), _ptr(out), R, M, N, 0, None, 0, _stream())
    return out


def col_sum(a, b=None, row_scale=None):
    """f32 [C] = sum_r row_scale[r] * a[r, c] * b[r, c] (b / row_scale optional, not both); a, b bf16 [R, C]."""
    _need_cuda(a)
    a = _bf16c(a)
    R, C = a.shape
    b = None if b is None else _bf16c(b)
    if row_scale is not None:
        row_scale = row_scale.to(torch.float32).contiguous()
    out = torch.empty(C, dtype=torch.float32, device=a.device)
    ws = _workspace("col_sum", _lib.query("advgrpo_col_sum_workspace_bytes", R, C), a.device)
    _lib.call("advgrpo_col_sum", _ptr(a), C, _ptr(b), C, _ptr(row_scale), _ptr(out), R, C, _ptr(ws), ws.numel(), _stream())
    return out


def _w16(t):
    """bf16 view of a parameter for the tensor-core kernels (fp32 modules are cast per call; bf16 ones are used as is)."""
    t = t.detach()
    return (t if t.dtype == torch.bfloat16 else t.to(torch.bfloat16)).contiguous()


class _Linear(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 GEMM with a native backward: dx = dy W (the same kernel on the transposed weight),
    dW = dy^T x (`gemm_tn`), db = column sums of dy (`col_sum`)."""

    @staticmethod
    def forward(ctx, x, weight, bias, wt=None):
        x2 = _bf16c(x).reshape(-1, x.shape[-1])
        w = _w16(weight)
        ctx.save_for_backward(x2, w)
        ctx.meta = (x.shape, weight.dtype, None if bias is None else bias.dtype)
        ctx.wt = wt                                      # optional callable -> cached contiguous W^T (the dX operand)
        y = gemm(x2, w, bias=None if bias is None else _w16(bias))
        return y.reshape(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        shape, wdt, bdt = ctx.meta
        dy2 = _bf16c(dy).reshape(-1, w.shape[0])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm(dy2, ctx.wt() if ctx.wt is not None else w.t().contiguous()).reshape(shape)
        if ctx.needs_input_grad[1]:
            dw = gemm_tn(dy2, x2).to(wdt)
        if bdt is not None and ctx.needs_input_grad[2]:
            db = col_sum(dy2).to(bdt)
        return dx, dw, db, None


def linear(x, weight, bias=None, wt=None):
    """Differentiable nn.Linear on the native kernels (bf16 compute, fp32 accumulation).  in / out features multiples of 64.
    `wt`: optional callable returning a cached contiguous transpose of the weight for the input-gradient GEMM."""
    return _Linear.apply(x, weight, bias, wt)


class _MlpGelu(torch.autograd.Function):
    """fc2(GELU_erf(fc1(x))) of a ViT block with the activation fused into the fc1 epilogue and its derivative into the
    epilogue of the fc2 input-gradient GEMM (no elementwise pass, the hidden activation's gradient never exists in HBM)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, tanh, w1t=None, w2t=None):
        x2 = _bf16c(x).reshape(-1, x.shape[-1])
        w1h, w2h = _w16(w1), _w16(w2)
        need = any(ctx.needs_input_grad)
        z = torch.empty((x2.shape[0], w1h.shape[0]), dtype=torch.bfloat16, device=x2.device) if need else None
        a = gemm(x2, w1h, bias=_w16(b1), epilogue=EPI_GELU_TANH if tanh else EPI_GELU_ERF, preact_out=z)
        y = gemm(a, w2h, bias=_w16(b2))
        if need:
            ctx.save_for_backward(x2, w1h, w2h, z, a)
        ctx.meta = (x.shape, w1.dtype, b1.dtype, w2.dtype, b2.dtype, bool(tanh))
        ctx.wts = (w1t, w2t)
        return y.reshape(*x.shape[:-1], w2h.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w1h, w2h, z, a = ctx.saved_tensors
        shape, w1dt, b1dt, w2dt, b2dt, tanh = ctx.meta
        dy2 = _bf16c(dy).reshape(-1, w2h.shape[0])
        w1t, w2t = ctx.wts
        dz = gemm(dy2, w2t() if w2t is not None else w2h.t().contiguous(),
                  epilogue=EPI_GELU_TANH_GRAD if tanh else EPI_GELU_ERF_GRAD, residual=z)
        dx = gemm(dz, w1t() if w1t is not None else w1h.t().contiguous()).reshape(shape) if ctx.needs_input_grad[0] else None
        dw1 = gemm_tn(dz, x2).to(w1dt) if ctx.needs_input_grad[1] else None
        db1 = col_sum(dz).to(b1dt) if ctx.needs_input_grad[2] else None
        dw2 = gemm_tn(dy2, a).to(w2dt) if ctx.needs_input_grad[3] else None
        db2 = col_sum(dy2).to(b2dt) if ctx.needs_input_grad[4] else None
        return dx, dw1, db1, dw2, db2, None, None, None


def mlp_gelu(x, w1, b1, w2, b2, approximate="none", w1t=None, w2t=None):
    """fc2(GELU(fc1(x))) with a native forward and backward (all four parameter gradients); approximate = "none" (erf:
    CLIP / DINOv2) or "tanh" (the MMDiT feed-forward).  w1t / w2t: optional callables returning cached transposes."""
    return _MlpGelu.apply(x, w1, b1, w2, b2, approximate == "tanh", w1t, w2t)


# --------------------------------------------------------------------------- score heads / discriminator step
def gather_rows_l2norm(feats, idx, l2norm, eps=1e-6):
    """feats bf16 [B, T, D], idx int64 [B, n] -> bf16 [B * (1 + n), D]: CLS row then the n patch rows 1 + idx per image,
    optionally L2-normalised (`x / (|x| + eps)` with the bf16 rounding order of rewards.py:411-412)."""
    _need_cuda(feats, idx)
    feats = _bf16c(feats)
    B, T, D = feats.shape
    idx = idx.to(torch.int64).contiguous()
    n = idx.shape[1]
    out = torch.empty((B * (1 + n), D), dtype=torch.bfloat16, device=feats.device)
    _lib.call("advgrpo_gather_rows_l2norm", _ptr(feats), _ptr(idx), _ptr(out), B, T, n, D, int(bool(l2norm)), float(eps), _stream())
    return out


def head_logits(a, w2, b2, round_bf16):
    a = _bf16c(a)
    R, Hd = a.shape
    logits = torch.empty(R, dtype=torch.float32, device=a.device)
    _lib.call("advgrpo_head_logits", _ptr(a), _ptr(w2), _ptr(b2), _ptr(logits), R, Hd, int(bool(round_bf16)), _stream())
    return logits


def dino_hybrid_score(logits, B, n, cls_weight, round_bf16):
    hybrid = torch.empty(B, dtype=torch.float32, device=logits.device)
    _lib.call("advgrpo_dino_hybrid_score", _ptr(logits), _ptr(hybrid), B, n, float(cls_weight), int(bool(round_bf16)), _stream())
    return hybrid


def dino_head_forward(head_params, rows, round_bf16, want_preact=False):
    """DINOHead (Linear -> GELU -> Linear(1), train_sd3_fast_dino_patch.py:592-603) on the kernels: tcgen05 GEMM with the
    erf-GELU epilogue, then the row-dot kernel.  head_params = (w1, b1, w2, b2) in any float dtype.  Returns
    (logits f32 [R], a, z) with z = the pre-activation (None unless want_preact)."""
    w1, b1, w2, b2 = (_w16(t) for t in head_params)
    z = torch.empty((rows.shape[0], w1.shape[0]), dtype=torch.bfloat16, device=rows.device) if want_preact else None
    a = gemm(rows, w1, bias=b1, epilogue=EPI_GELU_ERF, preact_out=z)
    return head_logits(a, w2.reshape(-1), b2.reshape(-1), round_bf16), a, z


class _DinoHeadHinge(torch.autograd.Function):
    """Discriminator loss of train_dino (train_sd3_fast_dino_patch.py:186-219) with the gradients of the four head
    parameters computed natively in the forward pass (the loss is closed-form in the logits): gather -> GEMM + GELU ->
    row-dot -> hinge -> dz -> (gemm_tn, col_sum).  Returns (loss, accuracy)."""

    @staticmethod
    def forward(ctx, w1, b1, w2, b2, feats_real, feats_fake, idx_real, idx_fake, patch_loss_weight):
        feats = torch.cat([feats_real, feats_fake]).to(torch.bfloat16)
        idx = torch.cat([idx_real, idx_fake])
        Br, Bf, n = feats_real.shape[0], feats_fake.shape[0], idx.shape[1]
        rows = gather_rows_l2norm(feats, idx, l2norm=False)                 # the D step does not normalise the tokens
        bf16_head = w1.dtype == torch.bfloat16
        logits, a, z = dino_head_forward((w1, b1, w2, b2), rows, bf16_head, want_preact=True)
        dl = torch.empty_like(logits)
        out3 = torch.empty(3, dtype=torch.float32, device=logits.device)
        _lib.call("advgrpo_dino_hinge_loss", _ptr(logits), _ptr(dl), _ptr(out3), Br, Bf, n, float(patch_loss_weight), _stream())
        w2h = _w16(w2).reshape(-1)
        dz = torch.empty_like(z)
        _lib.call("advgrpo_head_dz", _ptr(dl), _ptr(w2h), _ptr(z), _ptr(dz), z.shape[0], z.shape[1], _stream())
        dw1 = gemm_tn(dz, rows).to(w1.dtype)
        db1 = col_sum(dz).to(b1.dtype)
        dw2 = col_sum(a, row_scale=dl).reshape(w2.shape).to(w2.dtype)
        db2 = out3[2].reshape(b2.shape).to(b2.dtype)
        ctx.save_for_backward(dw1, db1, dw2, db2)
        ctx.mark_non_differentiable(out3)
        return out3[0].clone(), out3

    @staticmethod
    def backward(ctx, g_loss, _g):
        dw1, db1, dw2, db2 = ctx.saved_tensors
        g = g_loss.to(torch.float32)
        sc = lambda t: (t.float() * g).to(t.dtype)
        return sc(dw1), sc(db1), sc(dw2), sc(db2), None, None, None, None, None


def dino_head_hinge_loss(head_params, feats_real, feats_fake, idx_real, idx_fake, patch_loss_weight=0.3):
    """Returns (loss, accuracy); `loss.backward()` deposits the native gradients in the head parameters."""
    loss, out3 = _DinoHeadHinge.apply(*head_params, feats_real, feats_fake, idx_real, idx_fake, patch_loss_weight)
    return loss, out3[1]


def pickscore_head(img_feat, txt_feat, txt_index, logit_scale, bf16_arithmetic):
    """scores f32 [B] = exp(logit_scale) * cos(img[b], txt[txt_index[b]]) / 26 (pickscore_scorer.py:44-51)."""
    _need_cuda(img_feat, txt_feat)
    img_feat, txt_feat = _bf16c(img_feat), _bf16c(txt_feat)
    B, D = img_feat.shape
    ls = logit_scale.detach()
    if ls.dtype not in (torch.bfloat16, torch.float32):
        ls = ls.float()
    scores = torch.empty(B, dtype=torch.float32, device=img_feat.device)
    ti = None if txt_index is None else txt_index.to(torch.int64).contiguous()
    _lib.call("advgrpo_pickscore_head", _ptr(img_feat), _ptr(txt_feat), _ptr(ti), _ptr(ls), int(ls.dtype == torch.bfloat16),
              _ptr(scores), B, txt_feat.shape[0], D, int(bool(bf16_arithmetic)), _stream())
    return scores


def adam_torch_order(param, grad, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8, zero_grad=False):
    """In-place torch.optim.Adam update of one tensor (bf16 or fp32 parameter with moments of the same dtype) in torch's
    multi-tensor op order, rounding to the parameter dtype after every op (csrc/heads.cu)."""
    _need_cuda(param, grad, exp_avg, exp_avg_sq)
    n = param.numel()
    for t in (param, grad, exp_avg, exp_avg_sq):
        if not t.is_contiguous() or t.numel() != n or t.dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("adam_torch_order: contiguous bf16 / fp32 tensors of one size expected")
    if exp_avg.dtype != param.dtype or exp_avg_sq.dtype != param.dtype:
        raise ValueError("adam_torch_order: the moments carry the parameter's dtype (as torch.optim.Adam creates them)")
    _lib.call("advgrpo_adam_torch_order", _ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), n,
              int(param.dtype == torch.bfloat16), int(grad.dtype == torch.bfloat16), float(lr), float(betas[0]),
              float(betas[1]), float(eps), int(step), int(bool(zero_grad)), _stream())


def row_softmax_f32(x, scale=1.0, round_tf32=False, out=None):
    """softmax(scale * x) over the last dim of an fp32 matrix (in place when out is x)."""
    _need_cuda(x)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise _lib.AdvGrpoError("row_softmax_f32 expects a contiguous float32 tensor")
    cols = x.shape[-1]
    out = torch.empty_like(x) if out is None else out
    _lib.call("advgrpo_row_softmax_f32", _ptr(x), _ptr(out), x.numel() // cols, cols, float(scale), int(bool(round_tf32)), _stream())
    return out


def row_gate_mul(x, gate, rows_per_gate):
    """x [M, N] bf16 (any leading shape) * gate[m // rows_per_gate] (gate [G, N], row-strided) in one pass."""
    _need_cuda(x, gate)
    x = _bf16c(x)
    N = x.shape[-1]
    M = x.numel() // N
    out = torch.empty_like(x)
    _lib.call("advgrpo_row_gate_mul", _ptr(x), _ptr(gate), gate.stride(0), int(rows_per_gate), _ptr(out), M, N, _stream())
    return out


def _arr_p(vals):
    import ctypes
    return (ctypes.c_void_p * 2)(*[None if v is None else v for v in vals])


def _arr_i(vals):
    import ctypes
    return (ctypes.c_int64 * 2)(*[int(v) for v in vals])


# Experimental (off by default, ADVGRPO_GEMM_TAIL_SPLIT=1): when the 256x256 pair tiles of a
# dual launch fill k full waves plus a small remainder, run the rows behind the remainder as a second launch (which
# picks 128x128 tiles: one partial wave of quarter-cost tiles instead of a whole extra wave of pair tiles).
GEMM_TAIL_SPLIT = os.environ.get("ADVGRPO_GEMM_TAIL_SPLIT", "0") == "1"


def _tail_split_rows(M0, M1, N, rows_per_gate1, sms):
    """Rows of problem 1 to keep in the dual launch (the rest goes to the tail launch), or None."""
    pairs, tn = sms // 2, (N + 255) // 256
    rt0, rt1 = (M0 + 255) // 256, (M1 + 255) // 256
    tiles = (rt0 + rt1) * tn
    full, rem = divmod(tiles, pairs)
    if full < 2 or rem == 0 or 2 * rem > pairs or (full * pairs) % tn:
        return None
    keep_tiles = full * pairs // tn - rt0            # row tiles of problem 1 that complete the last full wave
    if keep_tiles <= 0:
        return None
    split = (keep_tiles * 256 // rows_per_gate1) * rows_per_gate1      # whole gate groups only
    if split <= 0 or split >= M1 or (split + 255) // 256 != keep_tiles:
        return None
    tail_tiles = ((M1 - split + 127) // 128) * ((N + 127) // 128)
    return split if tail_tiles <= sms else None


def gemm_dual(a, w, bias=(None, None), a2=(None, None), w2=(None, None), epilogue=EPI_NONE, residual=(None, None),
              gate=(None, None), rows_per_gate=(1, 1), preact_out=(None, None)):
    """Two problems (same N, K, K2, epilogue; different operands / row counts) in one persistent launch.
    Every argument is a pair; returns the pair of outputs."""
    _need_cuda(a[0], a[1], w[0], w[1])
    a2d, lead, cs = [], [], []
    for i in range(2):
        lead.append(a[i].shape[:-1])
        t = a[i].reshape(-1, a[i].shape[-1])
        a2d.append(t if t.stride(-1) == 1 else t.contiguous())
    K, N = a2d[0].shape[1], w[0].shape[0]
    assert a2d[1].shape[1] == K and w[1].shape[0] == N and w[0].shape[1] == K and w[1].shape[1] == K
    w = [x if x.stride(-1) == 1 else x.contiguous() for x in w]
    M = [t.shape[0] for t in a2d]
    for i in range(2):
        cs.append(torch.empty((M[i], N), dtype=torch.bfloat16, device=a2d[i].device))
    has2 = a2[0] is not None
    a2r = [None if x is None else x.reshape(M[i], -1) for i, x in enumerate(a2)]
    K2 = a2r[0].shape[1] if has2 else 0
    r2d = [None if x is None else x.reshape(M[i], N) for i, x in enumerate(residual)]
    st = lambda xs: [0 if x is None else x.stride(0) for x in xs]
    split = None
    if GEMM_TAIL_SPLIT:
        split = _tail_split_rows(M[0], M[1], N, int(rows_per_gate[1]) if gate[1] is not None else 1,
                                 torch.cuda.get_device_properties(a2d[0].device).multi_processor_count)
    M_main = M if split is None else [M[0], split]
    _lib.call("advgrpo_gemm_bf16_dual", _arr_p([_ptr(t) for t in a2d]), _arr_i(st(a2d)), _arr_p([_ptr(t) for t in w]),
              _arr_i(st(w)), _arr_p([_ptr(t) for t in a2r]), _arr_i(st(a2r)), _arr_p([_ptr(t) for t in w2]),
              _arr_i(st(w2)), K2, _arr_p([_ptr(t) for t in bias]), _arr_p([_ptr(t) for t in cs]), _arr_i(st(cs)),
              _arr_i(M_main), N, K, int(epilogue), _arr_p([_ptr(t) for t in r2d]), _arr_i(st(r2d)),
              _arr_p([_ptr(t) for t in gate]), _arr_i(st(gate)), _arr_i(rows_per_gate),
              _arr_p([_ptr(t) for t in preact_out]), _stream())
    if split is not None:                                  # the rows behind the last full wave, as their own launch
        rows = slice(split, None)
        cut = lambda t: None if t is None else t.reshape(M[1], -1)[rows]
        gemm(a2d[1][rows], w[1], bias=bias[1], a2=cut(a2r[1]), w2=w2[1], epilogue=epilogue, residual=cut(r2d[1]),
             gate=None if gate[1] is None else gate[1][split // int(rows_per_gate[1]):],
             rows_per_gate=int(rows_per_gate[1]), out=cs[1][rows], preact_out=cut(preact_out[1]))
    return cs[0].reshape(*lead[0], N), cs[1].reshape(*lead[1], N)


def gemm_qkv_norm(x, c, w, bias, norm_q, norm_k, H, D=64, a2=(None, None), w2=(None, None), eps=1e-6,
                  prenorm_out=(None, None)):
    """Fused QKV projection + per-head q/k RMSNorm + [image, text] concat in one persistent dual-problem launch
    (`advgrpo_gemm_qkv_norm`): x [B,S_img,K], c [B,S_txt,K] or None -> joint [B,S_img+S_txt,3,H,D].
    w / bias / norm_q / norm_k / a2 / w2 are (image, text) pairs.  No autograd (rollout path); bit-identical to
    gemm_dual + qk_norm_concat."""
    _need_cuda(x, w[0])
    B, S_img, K = x.shape
    S_txt = 0 if c is None else c.shape[1]
    n = 2 if c is not None else 1
    xs = [x.reshape(-1, K)] + ([c.reshape(-1, K)] if c is not None else [])
    xs = [t if t.stride(-1) == 1 else t.contiguous() for t in xs]
    ws = [t if t.stride(-1) == 1 else t.contiguous() for t in w[:n]]
    pad = lambda seq: list(seq[:n]) + [None] * (2 - n)
    xs2, ws2 = pad(xs), pad(ws)
    a2r = [None if t is None else t.reshape(xs[i].shape[0], -1) for i, t in enumerate(pad(a2)) ]
    K2 = a2r[0].shape[1] if a2r[0] is not None else 0
    joint = torch.empty((B, S_img + S_txt, 3, H, D), dtype=torch.bfloat16, device=x.device)
    st = lambda ts: [0 if t is None else t.stride(0) for t in ts]
    _lib.call("advgrpo_gemm_qkv_norm", _arr_p([_ptr(t) for t in xs2]), _arr_i(st(xs2)), _arr_p([_ptr(t) for t in ws2]),
              _arr_i(st(ws2)), _arr_p([_ptr(t) for t in a2r]), _arr_i(st(a2r)), _arr_p([_ptr(t) for t in pad(w2)]),
              _arr_i(st(pad(w2))), K2, _arr_p([_ptr(t) for t in pad(bias)]), _arr_p([_ptr(t) for t in pad(norm_q)]),
              _arr_p([_ptr(t) for t in pad(norm_k)]), _ptr(joint), _arr_p([_ptr(t) for t in pad(prenorm_out)]), B,
              _arr_i([S_img, S_txt]), H, D, K, float(eps), _stream())
    return joint


# --------------------------------------------------------------------------- reward preprocessing
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
_const_cache = {}


def _const3(vals, device):
    key = (vals, device.index)
    t = _const_cache.get(key)
    if t is None:
        t = torch.tensor(vals, dtype=torch.float32, device=device)
        _const_cache[key] = t
    return t


def clip_preprocess(images, out_size=224, dtype=torch.bfloat16, want_u8=False):
    """bf16 [B,3,H,W] in [0,1] -> CLIPProcessor pixel_values [B,3,out,out] (PIL-exact 8-bit resize)."""
    _need_cuda(images)
    is_u8 = images.dtype == torch.uint8
    images = images.contiguous() if is_u8 else _bf16c(images)
    B, C, H, W = images.shape
    assert C == 3
    dev = images.device
    pix = torch.empty((B, 3, out_size, out_size), dtype=dtype, device=dev)
    u8 = torch.empty((B, 3, out_size, out_size), dtype=torch.uint8, device=dev) if want_u8 else None
    ws_bytes = _lib.query("advgrpo_clip_preprocess_workspace_bytes", B, H, W, out_size)
    ws = _workspace("clip_pre", ws_bytes, dev)
    _lib.call("advgrpo_clip_preprocess", _ptr(images), int(is_u8), B, H, W, out_size, _ptr(_const3(CLIP_MEAN, dev)),
              _ptr(_const3(CLIP_STD, dev)), _ptr(pix), int(dtype == torch.float32), _ptr(u8), _ptr(ws),
              ws.numel(), _stream())
    return (pix, u8) if want_u8 else pix


def pil_resize_bilinear(img_hwc_u8, out_h, out_w, want_u8=False):
    """uint8 [H, W, 3] (a decoded PIL RGB image, on the device) -> float32 [3, out_h, out_w] in [0, 1]: Pillow's
    antialiased BILINEAR `resize` + torchvision `ToTensor()` bit for bit (train_sd3_fast_pickscore.py:791-797)."""
    _need_cuda(img_hwc_u8)
    if img_hwc_u8.dtype != torch.uint8 or img_hwc_u8.dim() != 3 or img_hwc_u8.shape[2] != 3:
        raise _lib.AdvGrpoError("pil_resize_bilinear expects a uint8 [H, W, 3] tensor")
    img = img_hwc_u8.contiguous()
    H, W, _ = img.shape
    dev = img.device
    out = torch.empty((3, out_h, out_w), dtype=torch.float32, device=dev)
    u8 = torch.empty((3, out_h, out_w), dtype=torch.uint8, device=dev) if want_u8 else None
    ws = _workspace("pil_resize", _lib.query("advgrpo_pil_resize_bilinear_workspace_bytes", H, W, out_h, out_w), dev)
    _lib.call("advgrpo_pil_resize_bilinear_u8", _ptr(img), H, W, out_h, out_w, _ptr(out), _ptr(u8), _ptr(ws), ws.numel(), _stream())
    return (out, u8) if want_u8 else out


def dino_preprocess(images, out_size=518):
    """[B,3,H,W] (bf16 or f32, [0,1]) -> bicubic out x out, ImageNet-normalised, bf16."""
    _need_cuda(images)
    if images.dtype not in (torch.bfloat16, torch.float32):
        images = images.float()
    images = images.contiguous()
    B, C, H, W = images.shape
    dev = images.device
    pix = torch.empty((B, 3, out_size, out_size), dtype=torch.bfloat16, device=dev)
    _lib.call("advgrpo_dino_preprocess", _ptr(images), int(images.dtype == torch.float32), B, H, W, out_size,
              _ptr(_const3(IMAGENET_MEAN, dev)), _ptr(_const3(IMAGENET_STD, dev)), _ptr(pix), _stream())
    return pix


# --------------------------------------------------------------------------- VAE GroupNorm (+SiLU)
def group_norm_silu_nhwc(x, gamma, beta, groups=32, eps=1e-6, silu=True, in_bias=None, round_tf32=False):
    """x: f32 [B, C, H, W] stored channels_last (or [B, H, W, C] contiguous).  Returns the same logical shape,
    channels_last."""
    _need_cuda(x)
    if x.dim() != 4 or x.dtype != torch.float32:
        raise _lib.AdvGrpoError("group_norm_silu_nhwc expects a 4-D float32 tensor")
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    B, C, H, W = x.shape
    y = torch.empty_like(x, memory_format=torch.channels_last)
    ws_bytes = _lib.query("advgrpo_group_norm_workspace_bytes", B, groups)
    ws = _workspace("gn", ws_bytes, x.device)
    _lib.call("advgrpo_group_norm_silu_nhwc", _ptr(x), _ptr(in_bias), _ptr(gamma), _ptr(beta), _ptr(y), B, H * W, C, groups,
              float(eps), int(bool(silu)) | (2 if round_tf32 else 0), _ptr(ws), ws.numel(), _stream())
    return y


def pack_conv_weight_tf32(w):
    """torch conv weight [Cout, Cin, kh, kw] -> tap-major [Cout, kh*kw*Cin] fp32, rounded to the nearest TF32 value
    (10-bit mantissa) so the tensor core's truncation of the low 13 mantissa bits is exact for the weights."""
    w = w.detach().float().permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()
    bits = w.view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32).contiguous()


def conv2d_nhwc_tf32(x, w_packed, bias, ksize):
    """x: f32 [B, Cin, H, W] stored channels_last; w_packed from `pack_conv_weight_tf32`.  3x3 / pad 1 or 1x1
    convolution on the tcgen05 TF32 implicit-GEMM kernel.  Returns f32 [B, Cout, H, W], channels_last."""
    _need_cuda(x, w_packed)
    if x.dim() != 4 or x.dtype != torch.float32:
        raise _lib.AdvGrpoError("conv2d_nhwc_tf32 expects a 4-D float32 tensor")
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    B, Cin, H, W = x.shape
    Cout = w_packed.shape[0]
    assert w_packed.shape[1] == ksize * ksize * Cin
    y = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    _lib.call("advgrpo_conv2d_nhwc_tf32", _ptr(x), _ptr(w_packed), _ptr(bias), _ptr(y), B, H, W, Cin, Cout, int(ksize), _stream())
    return y


def add_bias_nhwc(a, b, bias=None):
    """a + b + bias[c] for channels_last fp32 4-D tensors (fused residual add of the VAE resnet blocks)."""
    _need_cuda(a, b)
    a = a if a.is_contiguous(memory_format=torch.channels_last) else a.contiguous(memory_format=torch.channels_last)
    b = b if b.is_contiguous(memory_format=torch.channels_last) else b.contiguous(memory_format=torch.channels_last)
    B, C, H, W = a.shape
    out = torch.empty_like(a, memory_format=torch.channels_last)
    _lib.call("advgrpo_add_bias_nhwc", _ptr(a), _ptr(b), _ptr(bias), _ptr(out), B * H * W, C, _stream())
    return out


def upsample_nearest2x_nhwc(x):
    _need_cuda(x)
    x = x if x.is_contiguous(memory_format=torch.channels_last) else x.contiguous(memory_format=torch.channels_last)
    B, C, H, W = x.shape
    y = torch.empty((B, C, 2 * H, 2 * W), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    _lib.call("advgrpo_upsample_nearest2x_nhwc", _ptr(x), _ptr(y), B, H, W, C, _stream())
    return y

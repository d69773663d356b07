/* libadvgrpo_b200 -- C ABI of the B200-native (sm_100a) kernels behind the Adv-GRPO
 * rollout -> score -> advantage -> update hot path.
 *
 * The reference (showlab/Adv-GRPO @ 8287f90) is pure Python over
 * diffusers/transformers/peft and has no FFI of its own; each entry point below
 * names the reference code (file:line under /root/reference) whose arithmetic it
 * replaces.  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless noted;
 *   - the caller owns every buffer including workspaces; nothing here allocates;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no entry
 *     point synchronises with the host;
 *   - return value: 0 on success, a negative ADVGRPO_ERR_* code otherwise, with a
 *     message retrievable (per calling thread) from advgrpo_last_error();
 *   - re-entrant: no global mutable state besides the per-thread error string, so
 *     the reward thread pool (train_sd3_fast_pickscore.py:668,816-817) and the main
 *     thread may call concurrently on different streams.
 *   - bf16 tensors are raw uint16 storage (torch.bfloat16); "f32"/"f64" are IEEE.
 */
#ifndef ADVGRPO_B200_H_
#define ADVGRPO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADVGRPO_ABI_VERSION 1

#define ADVGRPO_OK 0
#define ADVGRPO_ERR_BAD_ARG (-1)
#define ADVGRPO_ERR_CUDA (-2)
#define ADVGRPO_ERR_UNSUPPORTED (-3)
#define ADVGRPO_ERR_WORKSPACE (-4)

typedef void* advgrpo_stream_t; /* cudaStream_t */

int advgrpo_abi_version(void);
/* Message of the last failing call made by the calling thread ("" if none). */
const char* advgrpo_last_error(void);
/* 0 if device `dev` is an sm_100 part this library can run on, ADVGRPO_ERR_UNSUPPORTED otherwise. */
int advgrpo_device_check(int dev);

/* ------------------------------------------------------------------------------------
 * A4 + A5: classifier-free-guidance combine + Flow-CPS SDE step + per-sample log-prob.
 * Replaces adv_grpo/diffusers_patch/sd3_sde_with_logprob.py:100-139
 * (sde_step_with_logprob_new) and the CFG combine at
 * adv_grpo/diffusers_patch/sd3_pipeline_with_logprob_fast.py:640-642 /
 * scripts/train_sd3_fast_pickscore.py:242-247, including the host-synchronising
 * index_for_timestep lookups (sde.py:106,110), which happen on the device here.
 *
 *   v = v_uncond + guidance * (v_text - v_uncond)     (each op rounded to bf16, as the
 *                                                      reference's bf16 tensor ops do)
 *   sigma = sigmas[idx(t)], sigma' = sigmas[idx(t)+1], std = sigma' * sin(noise_level*pi/2)
 *   mu = (x - sigma v)(1 - sigma') + (x + (1 - sigma) v) sqrt(sigma'^2 - std^2)
 *   rollout: prev = mu + std * eps ; replay: prev given
 *   log_prob[b] = -mean_n (prev - mu)^2
 *
 * v_uncond may be NULL (no guidance: v = v_text).  v_*: bf16 [B, n]; x: bf16 [B, n].
 * timesteps: f32 [B] or [1] (t_count = B or 1; 1 broadcasts as in the rollout,
 * fast.py:649).  sched_timesteps: f32 [T]; sigmas: f32 [T+1].
 * noise: f32 [B, n] injected noise, or NULL to draw it in-kernel from Philox4x32-10
 * keyed by (seed, offset).  prev_in: bf16 [B, n] (replay) or NULL (rollout).
 * Outputs: prev_out bf16 [B, n] (rollout only; the bf16-rounded next latents that the
 * reference stores, fast.py:654-655; may be NULL in replay), prev_mean_out f32 [B, n]
 * (optional, NULL to skip), log_prob f32 [B], std_out f32 [B] (optional).
 * workspace: advgrpo_sde_step_workspace_bytes(B, n) bytes.
 */
size_t advgrpo_sde_step_workspace_bytes(int64_t B, int64_t n);
int advgrpo_cfg_sde_step_logprob(const void* v_uncond, const void* v_text, const void* x,
                                 const void* prev_in, const float* noise, const float* timesteps,
                                 int64_t t_count, const float* sched_timesteps, const float* sigmas,
                                 int64_t T, void* prev_out, float* prev_mean_out, float* log_prob,
                                 float* std_out, int64_t B, int64_t n, float guidance,
                                 float noise_level, uint64_t seed, uint64_t offset, void* workspace,
                                 size_t workspace_bytes, advgrpo_stream_t stream);
/* Backward of the replay form w.r.t. the transformer output (what autograd derives at
 * train_sd3_fast_pickscore.py:1165 through :258-267 and :242-247):
 *   d log_prob[b] / d v = (2/n) (prev - mu) * ((1 - sigma) sqrt(sigma'^2 - std^2) - sigma (1 - sigma'))
 * chained through the bf16 CFG combine.  grad_log_prob: f32 [B].  Writes grad_v_uncond
 * (bf16 [B, n], may be NULL when v_uncond is NULL) and grad_v_text (bf16 [B, n]). */
int advgrpo_cfg_sde_logprob_bwd(const void* v_uncond, const void* v_text, const void* x,
                                const void* prev_in, const float* timesteps, int64_t t_count,
                                const float* sched_timesteps, const float* sigmas, int64_t T,
                                const float* grad_log_prob, void* grad_v_uncond, void* grad_v_text,
                                int64_t B, int64_t n, float guidance, float noise_level,
                                advgrpo_stream_t stream);

/* The same backward with the KL regulariser of train_sd3_fast_pickscore.py:1105-1108,1124-1128 (train.beta > 0):
 *   kl[b] = mean_{CHW}((mu - mu_ref)^2), mu_ref = prev_sample_mean of the adapter-disabled forward (f32 [B, n]);
 *   d kl[b] / d v = (2/n) (mu - mu_ref) d mu / d v is added with weight grad_kl[b] (f32 [B]).
 * grad_kl == mean_ref == NULL reduces to advgrpo_cfg_sde_logprob_bwd. */
int advgrpo_cfg_sde_logprob_kl_bwd(const void* v_uncond, const void* v_text, const void* x,
                                   const void* prev_in, const float* timesteps, int64_t t_count,
                                   const float* sched_timesteps, const float* sigmas, int64_t T,
                                   const float* grad_log_prob, const float* grad_kl, const float* mean_ref,
                                   void* grad_v_uncond, void* grad_v_text, int64_t B, int64_t n, float guidance,
                                   float noise_level, advgrpo_stream_t stream);

/* The same two entry points with the step variant selectable: 0 = Flow-CPS (`sde_step_with_logprob_new`, sde.py:77-139,
 * the one both training scripts import), 1 = Flow-SDE (`sde_step_with_logprob`, sde.py:13-73):
 *   std = sqrt(sigma / (1 - (sigma == 1 ? sigmas[1] : sigma))) * noise_level, dt = sigma' - sigma,
 *   mu = x (1 + std^2 / (2 sigma) dt) + v (1 + std^2 (1 - sigma) / (2 sigma)) dt, prev = mu + std sqrt(-dt) eps,
 *   log_prob = mean_{CHW}( -(prev - mu)^2 / (2 (std sqrt(-dt))^2) - log(std sqrt(-dt)) - log sqrt(2 pi) ). */
int advgrpo_cfg_sde_step_logprob_variant(const void* v_uncond, const void* v_text, const void* x,
                                         const void* prev_in, const float* noise, const float* timesteps,
                                         int64_t t_count, const float* sched_timesteps, const float* sigmas,
                                         int64_t T, void* prev_out, float* prev_mean_out, float* log_prob,
                                         float* std_out, int64_t B, int64_t n, float guidance,
                                         float noise_level, uint64_t seed, uint64_t offset, void* workspace,
                                         size_t workspace_bytes, int variant, advgrpo_stream_t stream);
int advgrpo_cfg_sde_logprob_bwd_variant(const void* v_uncond, const void* v_text, const void* x,
                                        const void* prev_in, const float* timesteps, int64_t t_count,
                                        const float* sched_timesteps, const float* sigmas, int64_t T,
                                        const float* grad_log_prob, const float* grad_kl, const float* mean_ref,
                                        void* grad_v_uncond, void* grad_v_text, int64_t B, int64_t n, float guidance,
                                        float noise_level, int variant, advgrpo_stream_t stream);

/* ------------------------------------------------------------------------------------
 * A9: group-relative advantage.  Replaces adv_grpo/stat_tracking.py:18-47
 * (PerPromptStatTracker.update, type='grpo') together with the prompt-identity
 * round trip of scripts/train_sd3_fast_pickscore.py:962-970 and the statistics of
 * calculate_zero_std_ratio (:195-229).
 *   adv[i,t] = (r[i,t] - mean_{j in group(i)} r[j,t]) / (std + 1e-4),
 *   std = population std over all N rewards of column t (global_std != 0) or over the group.
 * rewards: f32 [N, T].  Group identity: either group_keys int64 [N] (key_len = 1) or the
 * tokenised prompt rows int64 [N, key_len] (the reference's `prompt_ids`, key_len = 256);
 * rows with identical keys form a group.  advantages: f64 [N, T] (the reference returns
 * float64, quirk Q5).  stats (optional, f64 [4]): {n_groups, mean group size,
 * zero_std_ratio, mean per-group std of column 0}.
 */
/* `type` of PerPromptStatTracker.update (stat_tracking.py:46-70).  GRPO is the only one the training scripts
 * use; the others are per-group selections of the same [N, T] rewards:
 *   RWR: the rewards themselves;  SFT: 1.0 where the reward equals the maximum over the group's whole [n, T]
 *   block;  DPO (T = 1 only): +1 at the group's first arg-max, -1 at its first arg-min, members in array order,
 *   an all-equal group gets -1 / +1 on its first / second member. */
#define ADVGRPO_ADV_GRPO 0
#define ADVGRPO_ADV_RWR 1
#define ADVGRPO_ADV_SFT 2
#define ADVGRPO_ADV_DPO 3
size_t advgrpo_group_advantage_workspace_bytes(int64_t N, int64_t T);
int advgrpo_group_advantage_mode(const float* rewards, const int64_t* group_keys, int64_t key_len,
                                 int64_t N, int64_t T, int global_std, int mode, double* advantages,
                                 double* stats, void* workspace, size_t workspace_bytes,
                                 advgrpo_stream_t stream);
/* mode = ADVGRPO_ADV_GRPO */
int advgrpo_group_advantage(const float* rewards, const int64_t* group_keys, int64_t key_len,
                            int64_t N, int64_t T, int global_std, double* advantages,
                            double* stats, void* workspace, size_t workspace_bytes,
                            advgrpo_stream_t stream);

/* ------------------------------------------------------------------------------------
 * A11: GRPO clipped policy-gradient loss, its logged statistics and the backward seed.
 * Replaces scripts/train_sd3_fast_pickscore.py:1111-1162 (float64 arithmetic, Q5).
 *   A = clamp(adv, +-adv_clip_max); rho = exp(lp - lp_old);
 *   loss = mean(max(-A rho, -A clamp(rho, 1 - clip, 1 + clip)))
 * log_prob, old_log_prob: f32 [B]; advantages: f64 [B] with element stride adv_stride
 * (so a column of the [N, T] advantage matrix can be passed in place).
 * out: f64 [6] = {loss, approx_kl, clipfrac, clipfrac_gt_one, clipfrac_lt_one, policy_loss}.
 * grad_log_prob: f32 [B] = grad_scale * d loss / d log_prob (may be NULL).
 */
int advgrpo_grpo_clip_loss(const float* log_prob, const float* old_log_prob,
                           const double* advantages, int64_t adv_stride, int64_t B,
                           double clip_range, double adv_clip_max, double grad_scale, double* out,
                           float* grad_log_prob, advgrpo_stream_t stream);

/* ------------------------------------------------------------------------------------
 * A12: global-norm gradient clipping + AdamW + gradient clear on the flat fp32 LoRA master
 * parameter.  Replaces scripts/train_sd3_fast_pickscore.py:1165-1171
 * (accelerator.clip_grad_norm_(params, max_grad_norm); optimizer.step(); optimizer.zero_grad(),
 * optimizer = torch.optim.AdamW, :515-521).  The all-reduce of the gradient (NCCL) runs before
 * this call.  Two kernels, no host read of the norm:
 *   norm = ||g||_2;  g <- g * min(max_grad_norm / (norm + 1e-6), 1)      (skipped if max_grad_norm <= 0)
 *   p <- p - lr wd p;  m <- m + (1 - b1)(g - m);  v <- b2 v + (1 - b2) g^2;
 *   p <- p - lr / (1 - b1^step) * m / (sqrt(v) / sqrt(1 - b2^step) + eps)
 * param, grad, exp_avg, exp_avg_sq: f32 [n], 16-byte aligned.  step >= 1 is the 1-based count of
 * this update.  zero_grad != 0 clears grad in the same pass (else it is left clipped, as torch
 * does in place).  grad_norm_out (optional, f32 [1]) receives the pre-clip norm.
 */
size_t advgrpo_clip_adamw_workspace_bytes(int64_t n);
int advgrpo_clip_adamw(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                       double lr, double beta1, double beta2, double eps, double weight_decay,
                       int64_t step, double max_grad_norm, int zero_grad, float* grad_norm_out,
                       void* workspace, size_t workspace_bytes, advgrpo_stream_t stream);

/* ------------------------------------------------------------------------------------
 * A3 glue (MMDiT block, diffusers AdaLayerNormZero / SD35AdaLayerNormZeroX /
 * AdaLayerNormContinuous as called from SD3Transformer2DModel, reference call site
 * fast.py:630-637): y = LayerNorm(x; no affine, eps) * (1 + scale[b]) + shift[b].
 * x: bf16 [B, S, D]; shift/scale: bf16 rows of an adaLN embedding, row stride
 * mod_stride elements ([B, k*D] matrices: pass the chunk base pointers).  y2 (with
 * shift2/scale2) is the second modulation of the dual-attention blocks; pass NULL to skip.
 * D must be a multiple of 256 and <= 2048.
 */
int advgrpo_ln_modulate_fwd(const void* x, const void* shift, const void* scale, const void* shift2,
                            const void* scale2, int64_t mod_stride, void* y, void* y2, int64_t B,
                            int64_t S, int64_t D, float eps, advgrpo_stream_t stream);
/* dx (+)= d/dx [ y , y2 ] given dy (and dy2, may be NULL).  accumulate != 0 adds into dx. */
int advgrpo_ln_modulate_bwd(const void* x, const void* scale, const void* scale2,
                            int64_t mod_stride, const void* dy, const void* dy2, void* dx,
                            int accumulate, int64_t B, int64_t S, int64_t D, float eps,
                            advgrpo_stream_t stream);

/* A8a/A8b glue: affine LayerNorm of the reward towers' pre-LN blocks (transformers CLIPEncoderLayer
 * layer_norm1/2 behind adv_grpo/pickscore_scorer.py:40-43; timm Block norm1/2 behind
 * adv_grpo/rewards.py:397-399): y = (x - mean) * rstd * weight + bias, fp32 statistics, bf16 in/out.
 * x, y: bf16 [rows, D]; weight, bias: bf16 [D].  D must be a multiple of 256 and <= 2048.
 */
int advgrpo_layer_norm_affine(const void* x, const void* weight, const void* bias, void* y, int64_t rows,
                              int64_t D, float eps, advgrpo_stream_t stream);

/* Per-head RMSNorm of q and k (diffusers RMSNorm(head_dim, eps) inside
 * JointAttnProcessor2_0) fused with the [image, text] sequence concat: builds the joint
 * token-major buffer qkv_joint bf16 [B, S_img + S_txt, 3, H, D] that the attention
 * kernel reads through TMA.  qkv_img: bf16 [B, S_img, 3*H*D] (fused to_q/to_k/to_v
 * output); qkv_txt: bf16 [B, S_txt, 3*H*D] or NULL (S_txt = 0, dual attention attn2).
 * w_*: bf16 [D] RMSNorm weights, or NULL for no normalisation (SD3-medium). D = 64. */
int advgrpo_qk_norm_concat_fwd(const void* qkv_img, const void* qkv_txt, const void* wq_img,
                               const void* wk_img, const void* wq_txt, const void* wk_txt,
                               void* qkv_joint, int64_t B, int64_t S_img, int64_t S_txt, int64_t H,
                               int64_t D, float eps, advgrpo_stream_t stream);
int advgrpo_qk_norm_concat_bwd(const void* qkv_img, const void* qkv_txt, const void* wq_img,
                               const void* wk_img, const void* wq_txt, const void* wk_txt,
                               const void* dqkv_joint, void* dqkv_img, void* dqkv_txt, int64_t B,
                               int64_t S_img, int64_t S_txt, int64_t H, int64_t D, float eps,
                               advgrpo_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Attention (tcgen05 + TMA).  Replaces F.scaled_dot_product_attention inside
 * diffusers JointAttnProcessor2_0 (MMDiT joint text-image attention and the SD3.5
 * image-only attn2; reference call site fast.py:630-637, train_sd3_fast_pickscore.py:
 * 235-255), transformers CLIPAttention (PickScore ViT-H/14 towers,
 * adv_grpo/pickscore_scorer.py:40-43) and timm Attention (DINOv2-B/14,
 * adv_grpo/rewards.py:397).
 * qkv: bf16 [B, S, 3, H, D] token-major; out: bf16 [B, S, H, D].  If out2 != NULL the output
 * rows are split like the joint sequence (JointAttnProcessor2_0: image rows, then text rows):
 * rows [0, S_split) -> out [B, S_split, H, D], rows [S_split, S) -> out2 [B, S - S_split, H, D].
 * lse: f32 [B, H, S] natural-log-sum-exp of the scaled scores (NULL to skip; needed by bwd).
 * D in {64, 128}; any S >= 1 (ragged tail masked); causal != 0 applies the lower-
 * triangular mask (CLIP text tower).
 */
int advgrpo_attn_fwd(const void* qkv, void* out, void* out2, int64_t S_split, float* lse, int64_t B,
                     int64_t S, int64_t H, int64_t D, float scale, int causal,
                     advgrpo_stream_t stream);
/* The same forward with an additive score bias shared by all samples: P = softmax(scale * Q K^T + bias[h]),
 * bias f32 [H, S, S] (query-major).  This is T5's relative-position attention (scale = 1) in the text-encoding step
 * compute_text_embeddings / encode_prompt (train_sd3_fast_pickscore.py:186-193,
 * diffusers_patch/train_dreambooth_lora_sd3.py:98-144). */
int advgrpo_attn_fwd_bias(const void* qkv, const float* bias, void* out, float* lse, int64_t B, int64_t S, int64_t H,
                          int64_t D, float scale, advgrpo_stream_t stream);
/* dqkv: bf16 [B, S, 3, H, D].  workspace: advgrpo_attn_bwd_workspace_bytes(...) bytes. */
size_t advgrpo_attn_bwd_workspace_bytes(int64_t B, int64_t S, int64_t H, int64_t D);
int advgrpo_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                     void* dqkv, int64_t B, int64_t S, int64_t H, int64_t D, float scale,
                     int causal, void* workspace, size_t workspace_bytes,
                     advgrpo_stream_t stream);

/* out[m, :] = x[m, :] * gate[m / rows_per_gate, :] (bf16; gate rows of stride gate_stride): the backward of the adaLN
 * gate in `x + gate * f(.)` (d f = dy * gate) as one HBM pass.  N multiple of 8. */
int advgrpo_row_gate_mul(const void* x, const void* gate, int64_t gate_stride, int64_t rows_per_gate, void* out,
                         int64_t M, int64_t N, advgrpo_stream_t stream);

/* Skinny "TN" contraction over the token axis (tcgen05 + TMA, deterministic split-K): the LoRA weight-gradient products
 * of the replay backward -- peft lora.Linear's autograd through `loss.backward()` at train_sd3_fast_pickscore.py:1165:
 *   C[m, n] = sum_k A[k, m] * B[k, n],   A bf16 [Kt, Ms] (Ms <= 256, multiple of 8), B bf16 [Kt, Nb] (Nb multiple of 8)
 *   grad lora_A = dt^T x  (A = dt [tokens, r], B = x [tokens, K]);   grad lora_B^T = t^T dy, written transposed
 * out: bf16 [Ms, Nb], or [Nb, Ms] when transpose_out != 0.  fp32 accumulation; split-K over an 8-CTA cluster whose partial
 * tiles are summed through distributed shared memory in a fixed order (bit-reproducible).  workspace: none needed today
 * (advgrpo_gemm_tn_skinny_workspace_bytes returns 0; the arguments stay in the ABI). */
size_t advgrpo_gemm_tn_skinny_workspace_bytes(int64_t Kt, int64_t Ms, int64_t Nb);
int advgrpo_gemm_tn_skinny(const void* a, const void* b, void* out, int64_t Kt, int64_t Ms, int64_t Nb, int transpose_out,
                           void* workspace, size_t workspace_bytes, advgrpo_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Dense contraction with fused epilogue (tcgen05 + TMA), the nn.Linear / peft
 * lora.Linear layers of SD3Transformer2DModel and of the reward ViTs:
 *   C[M, N] = epilogue( A[M, K] @ W[N, K]^T  (+ A2[M, K2] @ W2[N, K2]^T)  + bias[N] )
 * A, W, A2, W2, C, residual: bf16 row-major with the given leading dimensions (elements);
 * bias: bf16 [N] or NULL.  The optional second product is the LoRA update
 * (A2 = x A_lora^T, W2 = scale * B_lora; train_sd3_fast_pickscore.py:488-505).
 * epilogue: ADVGRPO_EPI_*.  GATE_RESIDUAL: C = residual + gate[row / rows_per_gate] * (.)
 * with gate bf16 rows of stride gate_stride (the adaLN gate chunk).
 * GELU_*: applied to the bf16-rounded pre-activation z = bf16(acc + bias); if preact_out != NULL,
 * z is also stored there (bf16 [M, N], leading dimension ldc) for the backward pass.
 * K, K2 multiples of 64; N multiple of 8; M arbitrary.
 */
#define ADVGRPO_EPI_NONE 0
#define ADVGRPO_EPI_GELU_TANH 1
#define ADVGRPO_EPI_GELU_ERF 2
#define ADVGRPO_EPI_GATE_RESIDUAL 3
#define ADVGRPO_EPI_QKNORM 4 /* only through advgrpo_gemm_qkv_norm */
#define ADVGRPO_EPI_QUICK_GELU 5 /* x * sigmoid(1.702 x): the CLIP-L text encoder of encode_prompt */
#define ADVGRPO_EPI_GELU_TANH_GRAD 6 /* C = acc * gelu_tanh'(z), z = bf16 [M, N] passed as `residual` (ld ldr): the backward
                                        of a GELU feed-forward, dz = (dy W2) * gelu'(z), without a separate elementwise pass */
#define ADVGRPO_EPI_GELU_ERF_GRAD 7  /* the same with the erf GELU's derivative (CLIP-H / DINOv2 MLPs: the discriminator step's
                                        backward through the trainable vision blocks, train_sd3_fast_pickscore.py:1016-1029) */
int advgrpo_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const void* A2,
                      int64_t lda2, const void* W2, int64_t ldw2, int64_t K2, const void* bias,
                      void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue,
                      const void* residual, int64_t ldr, const void* gate, int64_t gate_stride,
                      int64_t rows_per_gate, void* preact_out, advgrpo_stream_t stream);

/* Two problems in ONE persistent launch: same N, K, K2 and epilogue, different operands and row counts
 * (every argument array has 2 entries; optional arrays may be NULL).  The MMDiT block runs its image-stream
 * and text-stream projections (different weights, 1024 vs 205 tokens per sample) this way: the short text
 * problem fills the tail wave of the image problem instead of paying its own launch and wave quantisation. */
int advgrpo_gemm_bf16_dual(const void* const* A, const int64_t* lda, const void* const* W, const int64_t* ldw,
                           const void* const* A2, const int64_t* lda2, const void* const* W2, const int64_t* ldw2,
                           int64_t K2, const void* const* bias, void* const* C, const int64_t* ldc, const int64_t* M,
                           int64_t N, int64_t K, int epilogue, const void* const* residual, const int64_t* ldr,
                           const void* const* gate, const int64_t* gate_stride, const int64_t* rows_per_gate,
                           void* const* preact_out, advgrpo_stream_t stream);

/* Fused QKV projection of one MMDiT attention (JointAttnProcessor2_0: to_q/to_k/to_v and add_q/k/v_proj, per-head
 * RMSNorm norm_q / norm_k / norm_added_q / norm_added_k, torch.cat([image, text], dim=1); reference call site
 * sd3_pipeline_with_logprob_fast.py:630-637) as ONE persistent launch: problem 0 = image stream, problem 1 = text
 * stream (S[1] == 0: image-only attention, the attn2 of the SD3.5 dual blocks).  A[i]: bf16 [B * S[i], K];
 * W[i]: bf16 [3 * H * D, K] (q | k | v rows); optional LoRA second product as in advgrpo_gemm_bf16; bias[i] bf16 [3HD].
 * Epilogue: z = bf16(acc + bias); q and k heads are RMS-normalised over D = 64 with norm_q[i] / norm_k[i] (bf16 [64],
 * both NULL = no normalisation) exactly as advgrpo_qk_norm_concat_fwd does; the result is TMA-stored straight into
 * qkv_joint bf16 [B, S[0] + S[1], 3, H, D] (image tokens first).  prenorm_out[i] (optional): z as flat [B * S[i], 3HD]
 * for the backward pass. */
int advgrpo_gemm_qkv_norm(const void* const* A, const int64_t* lda, const void* const* W, const int64_t* ldw,
                          const void* const* A2, const int64_t* lda2, const void* const* W2, const int64_t* ldw2,
                          int64_t K2, const void* const* bias, const void* const* norm_q, const void* const* norm_k,
                          void* qkv_joint, void* const* prenorm_out, int64_t B, const int64_t* S, int64_t H,
                          int64_t D, int64_t K, float eps, advgrpo_stream_t stream);

/* ------------------------------------------------------------------------------------
 * A8a preprocessing: the reward image path of adv_grpo/rewards.py:581-584 +
 * adv_grpo/pickscore_scorer.py:21-28 (CLIPProcessor) without the host round trip:
 *   u8 = clamp(round(bf16(img) * 255), 0, 255)           (bf16 arithmetic, quirk Q6)
 *   PIL-compatible antialiased bicubic resize to out_size x out_size (two passes, 8-bit
 *   intermediate, Pillow's 22-bit fixed-point coefficients) -> /255 -> (x - mean) / std.
 * images: bf16 [B, 3, H, W] in [0,1], or (images_u8 != 0) uint8 [B, 3, H, W] already-quantised
 * planes (the PIL inputs of train_pickscore, train_sd3_fast_pickscore.py:162-163);
 * pixels: bf16 or f32 [B, 3, out, out];
 * u8_out (optional): uint8 [B, 3, out, out] resized bytes for bit-exact checks.
 * workspace: advgrpo_clip_preprocess_workspace_bytes(B, H, W, out) bytes.
 */
size_t advgrpo_clip_preprocess_workspace_bytes(int64_t B, int64_t H, int64_t W, int64_t out);
int advgrpo_clip_preprocess(const void* images, int images_u8, int64_t B, int64_t H, int64_t W, int64_t out,
                            const float* mean3, const float* std3, void* pixels, int pixels_f32,
                            uint8_t* u8_out, void* workspace, size_t workspace_bytes,
                            advgrpo_stream_t stream);
/* Reference ("real") images of the adversarial loop: `transforms.Resize((S, S))` on the decoded PIL
 * image + `ToTensor()` (scripts/train_sd3_fast_pickscore.py:791-797) on the device.  img_hwc: the interleaved RGB bytes
 * `Image.open(path).convert("RGB")` yields, uint8 [H, W, 3] of ANY size; out_chw: f32 [3, out_h, out_w] = byte / 255.
 * Pillow's antialiased BILINEAR resample bit for bit (ImagingResample: triangle filter of support max(1, in / out),
 * horizontal then vertical pass with an 8-bit intermediate, 22-bit fixed-point coefficients).  u8_out_chw (optional): the
 * resized bytes, uint8 [3, out_h, out_w].  The entropy (Huffman / inflate) decode of the file stays a host step. */
size_t advgrpo_pil_resize_bilinear_workspace_bytes(int64_t H, int64_t W, int64_t out_h, int64_t out_w);
int advgrpo_pil_resize_bilinear_u8(const uint8_t* img_hwc, int64_t H, int64_t W, int64_t out_h, int64_t out_w, float* out_chw,
                                   uint8_t* u8_out_chw, void* workspace, size_t workspace_bytes, advgrpo_stream_t stream);
/* Baseline JPEG decode of the reference images (`Image.open(fpath).convert("RGB")`,
 * scripts/train_sd3_fast_pickscore.py:773-786), hybrid: marker parse + sequential Huffman entropy decode on the HOST (plain C++
 * inside this library), dequantisation + islow integer IDCT + fancy chroma upsampling + YCbCr -> RGB on the DEVICE.
 * Bit-exact with libjpeg(-turbo)'s default settings, i.e. with Pillow, for baseline / extended-sequential AND progressive
 * 8-bit Huffman files, grayscale or YCbCr with 4:4:4, 4:2:2 (2x1) or 4:2:0 (2x2) sampling, with or without restart intervals.
 * advgrpo_jpeg_parse fills `info` (host call); info->supported == 0 marks a valid file outside that subset (arithmetic,
 * lossless, 12-bit, CMYK / RGB-coded, multi-scan sequential, progressive scripts that stop short of full precision): the caller
 * keeps its host decoder for it -- the return value is still 0.  The host functions treat the file as untrusted and take only streams
 * every decoder reads the same way: a scan segment that does not end on its last block, a restart marker out of sequence, a missing
 * EOI, a Huffman table libjpeg refuses, an unknown marker, or coefficients beyond the range of 8-bit samples is an error
 * (ADVGRPO_ERR_BAD_ARG / _UNSUPPORTED), never a guess.
 * advgrpo_jpeg_entropy_decode (host call): coefs_host int16 [advgrpo_jpeg_coef_count(info)] = per component
 * [blocks_h, blocks_w, 64] quantised coefficients in natural order; qtabs_host uint16 [3 * 64] natural-order tables.
 * advgrpo_jpeg_idct_to_rgb: device pointers of the same two arrays -> rgb_hwc_dev uint8 [height, width, 3]. */
typedef struct {
  int32_t width, height, ncomp;
  int32_t h[3], v[3], tq[3];          /* sampling factors and quantisation-table ids per component */
  int32_t blocks_w[3], blocks_h[3];   /* block grid per component, padded to whole MCUs */
  int32_t restart_interval;
  int32_t supported;
  int32_t progressive;                /* SOF2: several scans accumulate into the coefficient blocks */
} advgrpo_jpeg_info;
int advgrpo_jpeg_parse(const uint8_t* file, size_t nbytes, advgrpo_jpeg_info* info);
size_t advgrpo_jpeg_coef_count(const advgrpo_jpeg_info* info);
int advgrpo_jpeg_entropy_decode(const uint8_t* file, size_t nbytes, int16_t* coefs_host, uint16_t* qtabs_host);
size_t advgrpo_jpeg_workspace_bytes(const advgrpo_jpeg_info* info);
int advgrpo_jpeg_idct_to_rgb(const int16_t* coefs_dev, const uint16_t* qtabs_dev, const advgrpo_jpeg_info* info,
                             uint8_t* rgb_hwc_dev, void* workspace, size_t workspace_bytes, advgrpo_stream_t stream);
/* PNG decode of the reference images (the adversarial loop's reference images are PNG files, README.md:114-128 of the
 * reference; `Image.open(fpath).convert("RGB")`, scripts/train_sd3_fast_pickscore.py:773-786), hybrid: chunk walk + zlib inflate
 * of the IDAT stream on the HOST (plain C++ in this library: stored / fixed / dynamic Huffman blocks, LZ77 copies), scan-line
 * unfiltering (None / Sub / Up / Average / Paeth) as an anti-diagonal wavefront + conversion to interleaved RGB on the
 * DEVICE.  Byte-exact with Pillow for non-interlaced and Adam7-interlaced files of every colour type (truecolour, truecolour + alpha: alpha
 * dropped, greyscale (+ alpha): replicated, palette: looked up) at every bit depth Pillow maps onto 8-bit RGB (1 / 2 / 4 / 8 /
 * 16-bit greyscale, 8 / 16-bit greyscale + alpha and truecolour (+ alpha), 1..8-bit palette).  advgrpo_png_parse (host) fills `info`;
 * supported == 0 marks a file this decoder does not take (a depth / colour-type pair PNG does not define, a palette image without PLTE).
 * The host functions treat the file as untrusted: every chunk checksum, the chunk order, zlib's code-completeness rules, the Adler-32
 * trailer and the filter types are verified, and any violation is ADVGRPO_ERR_BAD_ARG (the caller's host decoder then decides).
 * advgrpo_png_inflate (host): raw_host uint8 [advgrpo_png_raw_bytes(info)] = height x (1 filter byte + rowbytes) filtered scan
 * lines (interlaced: the same for each of the seven reduced images, concatenated), palette_host uint8 [768].  advgrpo_png_unfilter_to_rgb: device copies of both -> rgb_hwc_dev uint8 [height, width, 3]
 * (workspace: advgrpo_png_workspace_bytes, unused for truecolour). */
typedef struct {
  int32_t width, height, bit_depth, color_type, interlace;
  int32_t channels, rowbytes, palette_entries, supported;
} advgrpo_png_info;
int advgrpo_png_parse(const uint8_t* file, size_t nbytes, advgrpo_png_info* info);
size_t advgrpo_png_raw_bytes(const advgrpo_png_info* info);
int advgrpo_png_inflate(const uint8_t* file, size_t nbytes, uint8_t* raw_host, uint8_t* palette_host);
size_t advgrpo_png_workspace_bytes(const advgrpo_png_info* info);
int advgrpo_png_unfilter_to_rgb(const uint8_t* raw_dev, const uint8_t* palette_dev, const advgrpo_png_info* info,
                                uint8_t* rgb_hwc_dev, void* workspace, size_t workspace_bytes, advgrpo_stream_t stream);
/* A8b preprocessing (adv_grpo/rewards.py:379-391): bicubic (A = -0.75, align_corners =
 * False, no antialias) resize to out x out, ImageNet normalisation, cast to bf16. */
int advgrpo_dino_preprocess(const void* images, int images_f32, int64_t B, int64_t H, int64_t W,
                            int64_t out, const float* mean3, const float* std3, void* pixels,
                            advgrpo_stream_t stream);

/* ------------------------------------------------------------------------------------
 * A6 glue: GroupNorm(groups, affine, eps) + optional SiLU on NHWC fp32 activations -- the normalisation
 * between the convolutions of the SD3 VAE decoder (diffusers AutoencoderKL.decode, reference call site
 * adv_grpo/diffusers_patch/sd3_pipeline_with_logprob_fast.py:669).  x, y: f32 [B, HW, C] (channels last);
 * gamma, beta: f32 [C].  C / groups must be a multiple of 4.  In place (y == x) is allowed.
 * in_bias (f32 [C], may be NULL) is added to x first: the bias of the preceding convolution, so the
 * convolution itself runs bias-free and no separate bias pass exists.
 * silu: bit 0 = apply SiLU; bit 1 = round the output to the nearest TF32 value (the output feeds
 * advgrpo_conv2d_nhwc_tf32, whose tensor-core operands would otherwise be truncated).
 */
size_t advgrpo_group_norm_workspace_bytes(int64_t B, int64_t groups);
int advgrpo_group_norm_silu_nhwc(const float* x, const float* in_bias, const float* gamma, const float* beta,
                                 float* y, int64_t B, int64_t HW, int64_t C, int64_t groups, float eps, int silu,
                                 void* workspace, size_t workspace_bytes, advgrpo_stream_t stream);
/* out = a + b + bias[c] on NHWC fp32 [rows, C] (ResnetBlock2D residual add with the conv2 / shortcut biases
 * folded in; bias may be NULL), and nearest-neighbour 2x upsampling x [B,H,W,C] -> y [B,2H,2W,C] (Upsample2D; the
 * values are rounded to TF32: the result is the input of the upsampler's convolution). */
int advgrpo_add_bias_nhwc(const float* a, const float* b, const float* bias, float* out, int64_t rows, int64_t C,
                          advgrpo_stream_t stream);
int advgrpo_upsample_nearest2x_nhwc(const float* x, float* y, int64_t B, int64_t H, int64_t W, int64_t C,
                                    advgrpo_stream_t stream);

/* ------------------------------------------------------------------------------------
 * A6: the decoder's convolutions as an implicit GEMM on tcgen05 in TF32 with fp32 accumulation (the reference runs
 * `pipeline.vae.decode` in fp32, i.e. TF32 tensor-core convolutions under PyTorch's default cudnn.allow_tf32;
 * adv_grpo/diffusers_patch/sd3_pipeline_with_logprob_fast.py:667-670, train_sd3_fast_pickscore.py:481).
 * x: f32 [B, H, W, Cin] (channels last); w: f32 [Cout, ksize*ksize, Cin] (tap-major, i.e. the torch weight
 * [Cout, Cin, kh, kw] permuted to [Cout, kh, kw, Cin]); bias: f32 [Cout] or NULL; y: f32 [B, H, W, Cout].
 * ksize 3 (stride 1, zero padding 1) or 1.  Cin, Cout multiples of 32.  No workspace: the zero padding comes
 * from the TMA unit's out-of-bounds fill, there is no im2col buffer.
 */
int advgrpo_conv2d_nhwc_tf32(const float* x, const float* w, const float* bias, float* y, int64_t B, int64_t H, int64_t W,
                             int64_t Cin, int64_t Cout, int ksize, advgrpo_stream_t stream);

/* ------------------------------------------------------------------------------------
 * A8a / A8b score heads and the A14 / A15 discriminator step (`score_head_{pickscore,dino_patch}`).
 *
 * advgrpo_gather_rows_l2norm: feats bf16 [B, T, D] (DINOv2 `forward_features`: token 0 = CLS) and idx int64 [B, n]
 *   (the torch.randint draw of adv_grpo/rewards.py:406 / train_sd3_fast_dino_patch.py:199-200) -> out bf16 [B (1 + n), D]:
 *   row b (1 + n) is the CLS token, the next n rows are patch tokens 1 + idx[b, :].  l2norm != 0: every row is divided by
 *   (|row| + eps) with the bf16 rounding order of `x / (x.norm(dim=-1, keepdim=True) + 1e-6)` on bf16 tensors
 *   (rewards.py:411-412; the D step of train_dino does not normalise: l2norm = 0).
 * advgrpo_head_logits: logits f32 [R] = a[R, Hd] . w2[Hd] + b2 -- the Linear(hidden, 1) of DINOHead
 *   (train_sd3_fast_dino_patch.py:592-603); `a` = GELU(Linear(in, hidden)) from advgrpo_gemm_bf16 (GELU_ERF epilogue).
 *   round_bf16: round the logit to bf16 (the reference head is a bf16 module, train_sd3_fast_dino_patch.py:745).
 * advgrpo_dino_hybrid_score: hybrid[b] = cls_weight * logit[b, 0] + (1 - cls_weight) * mean_j logit[b, 1 + j]
 *   (rewards.py:414-419).
 * advgrpo_dino_hinge_loss: the discriminator loss of train_dino (train_sd3_fast_dino_patch.py:186-219) on logits
 *   [(B_real + B_fake), 1 + n] (real images first): out3 = {loss, accuracy on the CLS logits, sum of dlogits};
 *   dlogits f32 [R] = d loss / d logit.
 * advgrpo_head_dz: dz bf16 [R, Hd] = (dlogits[r] * w2[:]) * gelu_erf'(z[r, :]): backward through Linear(hidden, 1) + GELU.
 * advgrpo_col_sum: out f32 [C] = sum_r scale[r] * a[r, c] * b[r, c] (b, scale optional, not both): bias gradients and the
 *   head's d w2 = sum_r dlogits[r] a[r, :]; deterministic two-stage reduction.  a, b bf16, leading dimensions lda / ldb.
 */
int advgrpo_gather_rows_l2norm(const void* feats, const int64_t* idx, void* out, int64_t B, int64_t T, int64_t n, int64_t D,
                               int l2norm, float eps, advgrpo_stream_t stream);
int advgrpo_head_logits(const void* a, const void* w2, const void* b2, float* logits, int64_t R, int64_t Hd, int round_bf16,
                        advgrpo_stream_t stream);
int advgrpo_dino_hybrid_score(const float* logits, float* hybrid, int64_t B, int64_t n, float cls_weight, int round_bf16,
                              advgrpo_stream_t stream);
int advgrpo_dino_hinge_loss(const float* logits, float* dlogits, float* out3, int64_t B_real, int64_t B_fake, int64_t n,
                            float patch_loss_weight, advgrpo_stream_t stream);
int advgrpo_head_dz(const float* dlogits, const void* w2, const void* z, void* dz, int64_t R, int64_t Hd,
                    advgrpo_stream_t stream);
size_t advgrpo_col_sum_workspace_bytes(int64_t rows, int64_t C);
int advgrpo_col_sum(const void* a, int64_t lda, const void* b, int64_t ldb, const float* row_scale, float* out, int64_t rows,
                    int64_t C, void* workspace, size_t workspace_bytes, advgrpo_stream_t stream);
/* PickScore score tail (adv_grpo/pickscore_scorer.py:44-51): scores[b] = exp(logit_scale) * <img[b] / |img[b]|,
 * txt[i] / |txt[i]|> / 26 with i = txt_index[b] (NULL: b % n_txt).  img bf16 [B, D], txt bf16 [n_txt, D]; logit_scale is a
 * DEVICE scalar (bf16 or f32: no host read).  bf16_arithmetic != 0 reproduces the reference's bf16 model: norms, quotients,
 * the dot product, the scaling and the division by 26 are each rounded to bf16 (quirk Q10); 0 = fp32 tail. */
int advgrpo_pickscore_head(const void* img_feat, const void* txt_feat, const int64_t* txt_index, const void* logit_scale,
                           int logit_scale_is_bf16, float* scores, int64_t B, int64_t n_txt, int64_t D, int bf16_arithmetic,
                           advgrpo_stream_t stream);
/* Backward of the affine LayerNorm of a trainable CLIP block (train_sd3_fast_pickscore.py:1016-1029): dx bf16 [rows, D];
 * dweight / dbias f32 [D] (both NULL when the LayerNorm parameters are frozen).  D a multiple of 8, <= 2048. */
size_t advgrpo_layer_norm_affine_bwd_workspace_bytes(int64_t rows, int64_t D);
int advgrpo_layer_norm_affine_bwd(const void* x, const void* weight, const void* dy, void* dx, float* dweight, float* dbias,
                                  int64_t rows, int64_t D, float eps, void* workspace, size_t workspace_bytes,
                                  advgrpo_stream_t stream);
/* Gradients of the adaLN modulation vectors of one LayerNorm-modulate under full fine-tuning (config.use_lora = False,
 * train_sd3_fast_pickscore.py:488: the adaLN linears train): for each of nseg samples (rows_per_seg tokens each)
 * dshift[s, :] = sum_t dy[s, t, :] and dscale[s, :] = sum_t dy[s, t, :] * xhat[s, t, :], xhat = LayerNorm(no affine, eps)(x).
 * x, dy: bf16 [nseg * rows_per_seg, D]; outputs f32 [nseg, D]. */
size_t advgrpo_ln_modulation_grads_workspace_bytes(int64_t nseg, int64_t rows_per_seg, int64_t D);
int advgrpo_ln_modulation_grads(const void* x, const void* dy, float* dshift, float* dscale, int64_t nseg, int64_t rows_per_seg,
                                int64_t D, float eps, void* workspace, size_t workspace_bytes, advgrpo_stream_t stream);
/* The discriminator optimizers (`torch.optim.Adam(params, lr=config.d_lr, betas=(0.5, 0.999))`, train_sd3_fast_pickscore.py:658,
 * train_sd3_fast_dino_patch.py:750) as one pass per tensor in torch's multi-tensor op order (lerp_, mul_, addcmul_, sqrt,
 * div_, add_, addcdiv_), rounding to the parameter dtype after every op: bf16 parameters with bf16 moments follow the
 * reference's arithmetic.  param / exp_avg / exp_avg_sq: bf16 or f32 [n] (param_is_bf16); grad bf16 or f32 (grad_is_bf16);
 * step >= 1 is the step count AFTER this update; zero_grad != 0 clears the gradient in the same pass. */
int advgrpo_adam_torch_order(void* param, void* grad, void* exp_avg, void* exp_avg_sq, int64_t n, int param_is_bf16,
                             int grad_is_bf16, double lr, double beta1, double beta2, double eps, int64_t step, int zero_grad,
                             advgrpo_stream_t stream);
/* y = softmax(scale * x) over fp32 rows (in place allowed), optionally rounded to TF32: the probability matrix of the VAE
 * decoder's single-head mid-block attention (diffusers AutoencoderKL, sd3_pipeline_with_logprob_fast.py:669) between the two
 * TF32 tensor-core products advgrpo_conv2d_nhwc_tf32 computes (scores = 1x1 convolution with K as the weights). */
int advgrpo_row_softmax_f32(const float* x, float* y, int64_t rows, int64_t cols, float scale, int round_tf32,
                            advgrpo_stream_t stream);

/* Differentiable attention for short sequences and head sizes outside the tcgen05 backward kernel (CLIP-ViT-H/14: 257
 * tokens, head_dim 80): the softmax(Q K^T) V core of the vision blocks the PickScore discriminator step trains
 * (train_sd3_fast_pickscore.py:1016-1029; F.scaled_dot_product_attention inside transformers' CLIPAttention and its
 * autograd).  fp32 arithmetic on the CUDA cores with K / V (backward: also Q / dO) of one (sample, head) resident in
 * shared memory; q, k, v, o, dout, dq, dk, dv: bf16 [B, S, H, Dh] (the projections' [B, S, H * Dh] outputs viewed per
 * head); lse, delta: f32 [B, H, S] (delta is scratch written by the backward).  Dh even, <= 128; S * Dh bounded by the
 * 227 KB of shared memory (ADVGRPO_ERR_UNSUPPORTED beyond: S <= 544 at Dh = 80).  causal != 0: key j <= query i. */
int advgrpo_attn_small_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int64_t B, int64_t S, int64_t H,
                           int64_t Dh, float scale, int causal, advgrpo_stream_t stream);
int advgrpo_attn_small_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                           float* delta, void* dq, void* dk, void* dv, int64_t B, int64_t S, int64_t H, int64_t Dh, float scale,
                           int causal, advgrpo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ADVGRPO_B200_H_ */

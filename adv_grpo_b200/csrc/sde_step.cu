// Fused classifier-free-guidance combine + Flow-CPS SDE step + per-sample log-prob
// (forward for rollout and replay, backward for replay).  HBM-bound streaming kernel:
// 16-byte vector loads, fp32 math in the reference's op order (no FMA contraction so mu
// is bit-identical to the torch-fp32 evaluation), warp-shuffle + block reduction of
// sum (prev - mu)^2, deterministic cross-block finish by the last arriving block.
//
// Reference: adv_grpo/diffusers_patch/sd3_sde_with_logprob.py:100-139,
//            adv_grpo/diffusers_patch/sd3_pipeline_with_logprob_fast.py:640-655,
//            scripts/train_sd3_fast_pickscore.py:242-267.
#include <math.h>

#include "common.cuh"

namespace advgrpo {
namespace {

constexpr int kThreads = 256;
constexpr int kVec = 8;  // bf16 elements per 16-byte load

struct StepCoef {
  float sigma, sigma_prev, std, c, one_m_sigma, one_m_sigma_prev;
  // Flow-SDE variant (sde.py:13-73): mu = x * ax + (v * cv) * dt, noise scale s_noise = std * sqrt(-dt)
  float ax, cv, dt, s_noise;
  int mode;                 // 0 = Flow-CPS (sde_step_with_logprob_new), 1 = Flow-SDE (sde_step_with_logprob)
  bool valid;
};

// sde.py:106-122 scalar part; index_for_timestep done on the device (no .item()).
__device__ __forceinline__ StepCoef step_coef(const float* timesteps, int64_t t_count, int b,
                                              const float* sched_t, const float* sigmas, int T,
                                              float sin_level, int mode = 0, float noise_level = 0.f) {
  const float t = timesteps[t_count == 1 ? 0 : b];
  // diffusers index_for_timestep: the 2nd match if the timestep occurs more than once
  int first = -1, second = -1;
  StepCoef k;
  float sigma_max = 0.f;
  if (T <= 31) {
    // the whole schedule fits one warp: lane i holds sched_t[i] and sigmas[i] (T + 1 entries), the match is a ballot and
    // the coefficients come by shuffle -- ONE global round trip instead of the three dependent ones (timestep -> scan ->
    // sigmas[idx]) that made up a third of this launch-latency-bound kernel at rollout sizes (profiles/r2_hbm_kernels)
    const int lane = threadIdx.x & 31;
    const float tl = lane < T ? sched_t[lane] : 0.f;
    const float sg = lane <= T ? sigmas[lane] : 0.f;
    unsigned m = __ballot_sync(0xffffffffu, lane < T && tl == t);
    if (m) {
      first = __ffs(m) - 1;
      m &= m - 1;
      if (m) second = __ffs(m) - 1;
    }
    int idx = second >= 0 ? second : first;
    k.valid = idx >= 0;
    if (!k.valid) idx = 0;
    k.sigma = __shfl_sync(0xffffffffu, sg, idx);
    k.sigma_prev = __shfl_sync(0xffffffffu, sg, idx + 1);
    sigma_max = __shfl_sync(0xffffffffu, sg, 1);
  } else {
    for (int i = 0; i < T; ++i) {
      if (sched_t[i] == t) {
        if (first < 0) first = i;
        else if (second < 0) second = i;
      }
    }
    int idx = second >= 0 ? second : first;
    k.valid = idx >= 0;
    if (!k.valid) idx = 0;
    k.sigma = sigmas[idx];
    k.sigma_prev = sigmas[idx + 1];
    sigma_max = mode == 1 ? sigmas[1] : 0.f;               // Flow-SDE only (the API requires T >= 2 there)
  }
  k.std = __fmul_rn(k.sigma_prev, sin_level);
  k.c = __fsqrt_rn(__fsub_rn(__fmul_rn(k.sigma_prev, k.sigma_prev), __fmul_rn(k.std, k.std)));
  k.one_m_sigma = __fsub_rn(1.0f, k.sigma);
  k.one_m_sigma_prev = __fsub_rn(1.0f, k.sigma_prev);
  k.mode = mode;
  k.s_noise = k.std;
  k.ax = k.cv = k.dt = 0.f;
  if (mode == 1) {
    // sde.py:46-53 in torch's fp32 op order (every scalar below is a [B,1,1,1] fp32 tensor in the reference)
    const float den = __fsub_rn(1.0f, k.sigma == 1.0f ? sigma_max : k.sigma);
    k.std = __fmul_rn(__fsqrt_rn(__fdiv_rn(k.sigma, den)), noise_level);
    k.dt = __fsub_rn(k.sigma_prev, k.sigma);
    const float std2 = __fmul_rn(k.std, k.std);
    const float two_sigma = __fmul_rn(2.0f, k.sigma);
    k.ax = __fadd_rn(1.0f, __fmul_rn(__fdiv_rn(std2, two_sigma), k.dt));
    k.cv = __fadd_rn(1.0f, __fdiv_rn(__fmul_rn(std2, k.one_m_sigma), two_sigma));
    k.s_noise = __fmul_rn(k.std, __fsqrt_rn(__fmul_rn(-1.0f, k.dt)));
  }
  return k;
}

__device__ __forceinline__ float cfg_bf16(float u, float t, float g) {
  // every op rounds to bf16, as torch does on bf16 tensors (fast.py:641-642)
  float d = bf16_round(__fsub_rn(t, u));
  float e = bf16_round(__fmul_rn(g, d));
  return bf16_round(__fadd_rn(u, e));
}

__device__ __forceinline__ float mean_of(float x, float v, const StepCoef& k) {
  if (k.mode == 1) return __fadd_rn(__fmul_rn(x, k.ax), __fmul_rn(__fmul_rn(v, k.cv), k.dt));   // sde.py:53
  float x0 = __fsub_rn(x, __fmul_rn(k.sigma, v));                 // sde.py:120
  float x1 = __fadd_rn(x, __fmul_rn(v, k.one_m_sigma));           // sde.py:121
  return __fadd_rn(__fmul_rn(x0, k.one_m_sigma_prev), __fmul_rn(x1, k.c));  // sde.py:122
}

// Philox4x32-10 (Salmon et al. 2011) -> 4 standard normals by Box-Muller.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void normal4(uint64_t seed, uint64_t ctr, float (&z)[4]) {
  uint32_t r[4];
  philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const float kInv = 2.3283064365386963e-10f;  // 2^-32
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float u1 = fmaf((float)r[2 * i], kInv, kInv * 0.5f);   // (0, 1]
    float u2 = fmaf((float)r[2 * i + 1], kInv, kInv * 0.5f);
    // fast-math Box-Muller (MUFU lg2 / sqrt / sin / cos, ~2^-21 absolute error on the normals): the IEEE logf / sqrtf /
    // sincospif sequence cost ~50 instructions per pair and made the rollout form of this kernel issue-bound
    float rad, s, c;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-2.0f * __logf(u1)));
    __sincosf(6.283185307179586f * u2, &s, &c);
    z[2 * i] = rad * c;
    z[2 * i + 1] = rad * s;
  }
}

struct FwdArgs {
  const __nv_bfloat16 *vu, *vt, *x, *prev_in;
  const float* noise;
  const float *timesteps, *sched_t, *sigmas;
  int64_t t_count;
  int T;
  __nv_bfloat16* prev_out;
  float *prev_mean_out, *log_prob, *std_out;
  int64_t n;
  float guidance, sin_level;
  uint64_t seed, offset;
  double* partial;
  unsigned int* tickets;
  int mode;
  float noise_level;
};

__global__ void __launch_bounds__(kThreads) sde_fwd_kernel(FwdArgs a) {
  pdl_trigger();   // programmatic dependent launch (common.cuh): no-ops unless the launch carries the attribute
  pdl_wait();
  __shared__ float scratch[32];
  __shared__ bool is_last;
  const int b = blockIdx.y;
  const StepCoef k = step_coef(a.timesteps, a.t_count, b, a.sched_t, a.sigmas, a.T, a.sin_level, a.mode, a.noise_level);
  const int64_t base = (int64_t)b * a.n;
  const int64_t nvec = a.n / kVec;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * kThreads) {
    const int64_t e = base + i * kVec;
    float vt[8], vu[8], x[8], pv[8], mu[8];
    unpack8(*reinterpret_cast<const bf16x8*>(a.vt + e), vt);
    unpack8(*reinterpret_cast<const bf16x8*>(a.x + e), x);
    if (a.vu) {
      unpack8(*reinterpret_cast<const bf16x8*>(a.vu + e), vu);
#pragma unroll
      for (int j = 0; j < 8; ++j) vt[j] = cfg_bf16(vu[j], vt[j], a.guidance);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) mu[j] = mean_of(x[j], vt[j], k);
    if (a.prev_in) {
      unpack8(*reinterpret_cast<const bf16x8*>(a.prev_in + e), pv);
    } else {
      float z[8];
      if (a.noise) {
        const float4* np = reinterpret_cast<const float4*>(a.noise + e);
        float4 z0 = np[0], z1 = np[1];
        z[0] = z0.x; z[1] = z0.y; z[2] = z0.z; z[3] = z0.w;
        z[4] = z1.x; z[5] = z1.y; z[6] = z1.z; z[7] = z1.w;
      } else {
        float z4[4];
        normal4(a.seed, a.offset + (uint64_t)(e >> 2), z4);
        z[0] = z4[0]; z[1] = z4[1]; z[2] = z4[2]; z[3] = z4[3];
        normal4(a.seed, a.offset + (uint64_t)(e >> 2) + 1, z4);
        z[4] = z4[0]; z[5] = z4[1]; z[6] = z4[2]; z[7] = z4[3];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) pv[j] = __fadd_rn(mu[j], __fmul_rn(k.s_noise, z[j]));  // sde.py:131 / :62
      if (a.prev_out) *reinterpret_cast<bf16x8*>(a.prev_out + e) = pack8(pv);        // fast.py:654-655
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float d = __fsub_rn(pv[j], mu[j]);
      acc = fmaf(d, d, acc);
    }
    if (a.prev_mean_out) {
      float4* mp = reinterpret_cast<float4*>(a.prev_mean_out + e);
      mp[0] = make_float4(mu[0], mu[1], mu[2], mu[3]);
      mp[1] = make_float4(mu[4], mu[5], mu[6], mu[7]);
    }
  }
  float tot = block_sum(acc, scratch);
  if (threadIdx.x == 0) {
    a.partial[(int64_t)b * gridDim.x + blockIdx.x] = (double)tot;
    __threadfence();
    unsigned int t = atomicAdd(&a.tickets[b], 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    __threadfence();
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) s += a.partial[(int64_t)b * gridDim.x + i];
    s = warp_sum(s);
    if (threadIdx.x == 0) {
      float lp = (float)(-(s / (double)a.n));                      // sde.py:134-137
      if (k.mode == 1) {                                           // sde.py:64-71 (Gaussian log-density, mean over CHW)
        const double sn = (double)k.s_noise;
        lp = (float)(-(s / (double)a.n) / (2.0 * sn * sn) - log(sn) - 0.91893853320467274178);   // log(sqrt(2 pi))
      }
      a.log_prob[b] = k.valid ? lp : __int_as_float(0x7fc00000);
      if (a.std_out) a.std_out[b] = k.std;
      a.tickets[b] = 0u;  // leave the workspace clean for the next call
    }
  }
}

struct BwdArgs {
  const __nv_bfloat16 *vu, *vt, *x, *prev_in;
  const float *timesteps, *sched_t, *sigmas, *grad_lp;
  const float *grad_kl, *mean_ref;     // KL term of train_sd3_fast_pickscore.py:1124-1128 (both NULL when beta == 0)
  int64_t t_count;
  int T;
  __nv_bfloat16 *gvu, *gvt;
  int64_t n;
  float guidance, sin_level;
  int mode;
  float noise_level;
};

__global__ void __launch_bounds__(kThreads) sde_bwd_kernel(BwdArgs a) {
  pdl_trigger();   // programmatic dependent launch (common.cuh): no-ops unless the launch carries the attribute
  pdl_wait();
  const int b = blockIdx.y;
  const StepCoef k = step_coef(a.timesteps, a.t_count, b, a.sched_t, a.sigmas, a.T, a.sin_level, a.mode, a.noise_level);
  // Flow-CPS: d mu / d v = (1 - sigma) c - sigma (1 - sigma'), d logp / d mu = (2/n) (prev - mu)
  // Flow-SDE: d mu / d v = cv dt,                              d logp / d mu = (prev - mu) / (n s^2)
  const float dmu_dv = k.mode == 1 ? k.cv * k.dt : k.one_m_sigma * k.c - k.sigma * k.one_m_sigma_prev;
  const float dlp = k.mode == 1 ? 1.0f / ((float)a.n * k.s_noise * k.s_noise) : 2.0f / (float)a.n;
  const float coef = a.grad_lp[b] * dlp * dmu_dv;
  // kl[b] = mean((mu - mu_ref)^2):  d kl[b] / d v = (2/n) (mu - mu_ref) d mu / d v
  const float coef_kl = a.mean_ref ? a.grad_kl[b] * (2.0f / (float)a.n) * dmu_dv : 0.f;
  const int64_t base = (int64_t)b * a.n;
  const int64_t nvec = a.n / kVec;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * kThreads) {
    const int64_t e = base + i * kVec;
    float vt[8], vu[8], x[8], pv[8], gt[8], gu[8];
    unpack8(*reinterpret_cast<const bf16x8*>(a.vt + e), vt);
    unpack8(*reinterpret_cast<const bf16x8*>(a.x + e), x);
    unpack8(*reinterpret_cast<const bf16x8*>(a.prev_in + e), pv);
    if (a.vu) {
      unpack8(*reinterpret_cast<const bf16x8*>(a.vu + e), vu);
#pragma unroll
      for (int j = 0; j < 8; ++j) vt[j] = cfg_bf16(vu[j], vt[j], a.guidance);
    }
    float mr[8];
    if (a.mean_ref) {
      const float4 m0 = *reinterpret_cast<const float4*>(a.mean_ref + e);
      const float4 m1 = *reinterpret_cast<const float4*>(a.mean_ref + e + 4);
      mr[0] = m0.x; mr[1] = m0.y; mr[2] = m0.z; mr[3] = m0.w;
      mr[4] = m1.x; mr[5] = m1.y; mr[6] = m1.z; mr[7] = m1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float mu = mean_of(x[j], vt[j], k);
      float gf = coef * (pv[j] - mu);
      if (a.mean_ref) gf = fmaf(coef_kl, mu - mr[j], gf);
      float g = bf16_round(gf);                     // grad at noise_pred (bf16 tensor)
      if (a.vu) {
        float gd = bf16_round(a.guidance * g);      // through e = g * d
        gt[j] = gd;                                 // d = t - u
        gu[j] = bf16_round(g - gd);                 // u + e  and  -d
      } else {
        gt[j] = g;
      }
    }
    *reinterpret_cast<bf16x8*>(a.gvt + e) = pack8(gt);
    if (a.vu) *reinterpret_cast<bf16x8*>(a.gvu + e) = pack8(gu);
  }
}

int blocks_per_sample(int64_t B, int64_t n) {
  int64_t need = (n / kVec + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)sm_count() * 8 / (B > 0 ? B : 1);
  if (cap < 1) cap = 1;
  int64_t g = need < cap ? need : cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace
}  // namespace advgrpo

using namespace advgrpo;

extern "C" {

size_t advgrpo_sde_step_workspace_bytes(int64_t B, int64_t n) {
  (void)n;
  // partial sums (worst case grid) + tickets
  return (size_t)B * 2048 * sizeof(double) + (size_t)B * sizeof(unsigned int) + 16;
}

int advgrpo_cfg_sde_step_logprob(const void* v_uncond, const void* v_text, const void* x,
                                 const void* prev_in, const float* noise, const float* timesteps,
                                 int64_t t_count, const float* sched_timesteps, const float* sigmas,
                                 int64_t T, void* prev_out, float* prev_mean_out, float* log_prob,
                                 float* std_out, int64_t B, int64_t n, float guidance,
                                 float noise_level, uint64_t seed, uint64_t offset, void* workspace,
                                 size_t workspace_bytes, advgrpo_stream_t stream) {
  return advgrpo_cfg_sde_step_logprob_variant(v_uncond, v_text, x, prev_in, noise, timesteps, t_count, sched_timesteps,
                                              sigmas, T, prev_out, prev_mean_out, log_prob, std_out, B, n, guidance,
                                              noise_level, seed, offset, workspace, workspace_bytes, 0, stream);
}

int advgrpo_cfg_sde_step_logprob_variant(const void* v_uncond, const void* v_text, const void* x,
                                         const void* prev_in, const float* noise, const float* timesteps,
                                         int64_t t_count, const float* sched_timesteps, const float* sigmas,
                                         int64_t T, void* prev_out, float* prev_mean_out, float* log_prob,
                                         float* std_out, int64_t B, int64_t n, float guidance,
                                         float noise_level, uint64_t seed, uint64_t offset, void* workspace,
                                         size_t workspace_bytes, int variant, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(variant == 0 || variant == 1, "cfg_sde_step_logprob: variant must be 0 (Flow-CPS) or 1 (Flow-SDE)");
  ADVGRPO_CHECK_ARG(variant == 0 || T >= 2, "cfg_sde_step_logprob: Flow-SDE needs at least two schedule steps (sigma_max = sigmas[1])");
  ADVGRPO_CHECK_ARG(v_text && x && timesteps && sched_timesteps && sigmas && log_prob,
                    "cfg_sde_step_logprob: null required pointer");
  ADVGRPO_CHECK_ARG(B > 0 && n > 0 && n % kVec == 0, "cfg_sde_step_logprob: n=%lld must be a positive multiple of 8",
                    (long long)n);
  ADVGRPO_CHECK_ARG(t_count == 1 || t_count == B, "cfg_sde_step_logprob: t_count must be 1 or B");
  ADVGRPO_CHECK_ARG(T >= 1 && T <= 4096, "cfg_sde_step_logprob: bad schedule length %lld", (long long)T);
  ADVGRPO_CHECK_ARG(prev_in || prev_out, "cfg_sde_step_logprob: rollout form needs prev_out");
  ADVGRPO_CHECK_ARG(aligned16(v_text) && aligned16(x) && (!v_uncond || aligned16(v_uncond)) &&
                        (!prev_in || aligned16(prev_in)) && (!prev_out || aligned16(prev_out)) &&
                        (!noise || aligned16(noise)) && (!prev_mean_out || aligned16(prev_mean_out)),
                    "cfg_sde_step_logprob: tensors must be 16-byte aligned");
  if (workspace_bytes < advgrpo_sde_step_workspace_bytes(B, n) || !workspace)
    return set_error(ADVGRPO_ERR_WORKSPACE, "cfg_sde_step_logprob: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int gx = blocks_per_sample(B, n);
  if (gx > 2048) gx = 2048;
  FwdArgs a;
  a.vu = (const __nv_bfloat16*)v_uncond; a.vt = (const __nv_bfloat16*)v_text;
  a.x = (const __nv_bfloat16*)x; a.prev_in = (const __nv_bfloat16*)prev_in;
  a.noise = noise; a.timesteps = timesteps; a.sched_t = sched_timesteps; a.sigmas = sigmas;
  a.t_count = t_count; a.T = (int)T; a.prev_out = (__nv_bfloat16*)prev_out;
  a.prev_mean_out = prev_mean_out; a.log_prob = log_prob; a.std_out = std_out; a.n = n;
  a.guidance = guidance;
  a.sin_level = (float)sin((double)noise_level * M_PI / 2.0);   // sde.py:119 (python double -> f32 scalar)
  a.seed = seed; a.offset = offset;
  a.mode = variant; a.noise_level = noise_level;
  a.partial = (double*)workspace;
  a.tickets = (unsigned int*)((char*)workspace + (size_t)B * 2048 * sizeof(double));
  ADVGRPO_CUDA_CALL(cudaMemsetAsync(a.tickets, 0, (size_t)B * sizeof(unsigned int), st));
  ADVGRPO_CUDA_CALL(launch_chain(sde_fwd_kernel, dim3(gx, (unsigned)B), dim3(kThreads), 0, st, 1, a));
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_cfg_sde_logprob_bwd(const void* v_uncond, const void* v_text, const void* x,
                                const void* prev_in, const float* timesteps, int64_t t_count,
                                const float* sched_timesteps, const float* sigmas, int64_t T,
                                const float* grad_log_prob, void* grad_v_uncond, void* grad_v_text,
                                int64_t B, int64_t n, float guidance, float noise_level,
                                advgrpo_stream_t stream) {
  return advgrpo_cfg_sde_logprob_bwd_variant(v_uncond, v_text, x, prev_in, timesteps, t_count, sched_timesteps, sigmas, T,
                                             grad_log_prob, nullptr, nullptr, grad_v_uncond, grad_v_text, B, n, guidance,
                                             noise_level, 0, stream);
}

int advgrpo_cfg_sde_logprob_kl_bwd(const void* v_uncond, const void* v_text, const void* x,
                                   const void* prev_in, const float* timesteps, int64_t t_count,
                                   const float* sched_timesteps, const float* sigmas, int64_t T,
                                   const float* grad_log_prob, const float* grad_kl, const float* mean_ref,
                                   void* grad_v_uncond, void* grad_v_text, int64_t B, int64_t n, float guidance,
                                   float noise_level, advgrpo_stream_t stream) {
  return advgrpo_cfg_sde_logprob_bwd_variant(v_uncond, v_text, x, prev_in, timesteps, t_count, sched_timesteps, sigmas, T,
                                             grad_log_prob, grad_kl, mean_ref, grad_v_uncond, grad_v_text, B, n, guidance,
                                             noise_level, 0, stream);
}

int advgrpo_cfg_sde_logprob_bwd_variant(const void* v_uncond, const void* v_text, const void* x,
                                        const void* prev_in, const float* timesteps, int64_t t_count,
                                        const float* sched_timesteps, const float* sigmas, int64_t T,
                                        const float* grad_log_prob, const float* grad_kl, const float* mean_ref,
                                        void* grad_v_uncond, void* grad_v_text, int64_t B, int64_t n, float guidance,
                                        float noise_level, int variant, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(variant == 0 || variant == 1, "cfg_sde_logprob_bwd: variant must be 0 (Flow-CPS) or 1 (Flow-SDE)");
  ADVGRPO_CHECK_ARG((grad_kl == nullptr) == (mean_ref == nullptr), "cfg_sde_logprob_kl_bwd: grad_kl and mean_ref go together");
  ADVGRPO_CHECK_ARG(!mean_ref || aligned16(mean_ref), "cfg_sde_logprob_kl_bwd: mean_ref must be 16-byte aligned");
  ADVGRPO_CHECK_ARG(v_text && x && prev_in && timesteps && sched_timesteps && sigmas &&
                        grad_log_prob && grad_v_text,
                    "cfg_sde_logprob_bwd: null required pointer");
  ADVGRPO_CHECK_ARG(!v_uncond || grad_v_uncond, "cfg_sde_logprob_bwd: grad_v_uncond required with v_uncond");
  ADVGRPO_CHECK_ARG(B > 0 && n > 0 && n % kVec == 0, "cfg_sde_logprob_bwd: n must be a positive multiple of 8");
  ADVGRPO_CHECK_ARG(t_count == 1 || t_count == B, "cfg_sde_logprob_bwd: t_count must be 1 or B");
  ADVGRPO_CHECK_ARG(aligned16(v_text) && aligned16(x) && aligned16(prev_in) && aligned16(grad_v_text) &&
                        (!v_uncond || (aligned16(v_uncond) && aligned16(grad_v_uncond))),
                    "cfg_sde_logprob_bwd: tensors must be 16-byte aligned");
  BwdArgs a;
  a.vu = (const __nv_bfloat16*)v_uncond; a.vt = (const __nv_bfloat16*)v_text;
  a.x = (const __nv_bfloat16*)x; a.prev_in = (const __nv_bfloat16*)prev_in;
  a.timesteps = timesteps; a.sched_t = sched_timesteps; a.sigmas = sigmas; a.grad_lp = grad_log_prob;
  a.grad_kl = grad_kl; a.mean_ref = mean_ref;
  a.t_count = t_count; a.T = (int)T; a.gvu = (__nv_bfloat16*)grad_v_uncond;
  a.gvt = (__nv_bfloat16*)grad_v_text; a.n = n; a.guidance = guidance;
  a.sin_level = (float)sin((double)noise_level * M_PI / 2.0);
  a.mode = variant; a.noise_level = noise_level;
  int gx = blocks_per_sample(B, n);
  ADVGRPO_CUDA_CALL(launch_chain(sde_bwd_kernel, dim3(gx, (unsigned)B), dim3(kThreads), 0, (cudaStream_t)stream, 1, a));
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

}  // extern "C"

// MMDiT glue kernels (HBM-bound, warp-shuffle reductions, 16-byte vector access):
//   * LayerNorm(no affine) + adaLN modulate, one or two modulations per pass (dual blocks)
//   * per-head q/k RMSNorm fused with the [image, text] sequence concat into the joint
//     token-major QKV buffer the attention kernel reads through TMA
// and their backward passes w.r.t. the activation stream (the modulation vectors and the
// RMSNorm weights are frozen in LoRA training: train_sd3_fast_pickscore.py:488-505).
//
// Reference semantics: diffusers AdaLayerNormZero / SD35AdaLayerNormZeroX /
// AdaLayerNormContinuous and JointAttnProcessor2_0 + RMSNorm, reached from
// adv_grpo/diffusers_patch/sd3_pipeline_with_logprob_fast.py:630-637.
#include "common.cuh"

namespace advgrpo {
namespace {

constexpr int kWarpsPerBlock = 8;

// AFFINE: plain LayerNorm with per-channel weight / bias (`scale` = weight, `shift` = bias, mod_stride 0):
// y = n * weight + bias instead of n * (1 + scale) + shift.
template <int NV, bool AFFINE = false>  // D = NV * 256
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
ln_modulate_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ shift,
                       const __nv_bfloat16* __restrict__ scale, const __nv_bfloat16* __restrict__ shift2,
                       const __nv_bfloat16* __restrict__ scale2, int64_t mod_stride,
                       __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ y2, int64_t rows,
                       int64_t S, float eps) {
  pdl_trigger();   // programmatic dependent launch (common.cuh): no-ops unless the launch carries the attribute
  pdl_wait();
  constexpr int D = NV * 256;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int64_t b = row / S;
  const __nv_bfloat16* xr = x + row * D;
  // The row stays PACKED in registers (NV x 16 bytes per lane) and is unpacked once per pass: 24 registers instead of 48
  // at width 1536, so four 8-warp blocks (64 registers per thread) instead of three (80) are resident per SM and a third
  // more row loads are in flight (ncu r2: 34 % of the warp slots active, 4.4 TB/s).  Same fp32 operations in the same
  // order as before: bit-identical output.
  bf16x8 raw[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) raw[i] = *reinterpret_cast<const bf16x8*>(xr + (i * 32 + lane) * 8);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float v[8];
    unpack8(raw[i], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float v[8];
    unpack8(raw[i], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[j] - mean;
      q = fmaf(d, d, q);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  const __nv_bfloat16* sh = shift + b * mod_stride;
  const __nv_bfloat16* sc = scale + b * mod_stride;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int off = (i * 32 + lane) * 8;
    float v[8], fsh[8], fsc[8], o[8];
    unpack8(raw[i], v);
    unpack8(*reinterpret_cast<const bf16x8*>(sh + off), fsh);
    unpack8(*reinterpret_cast<const bf16x8*>(sc + off), fsc);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] -= mean;
      o[j] = fmaf(v[j] * rstd, AFFINE ? fsc[j] : 1.0f + fsc[j], fsh[j]);
    }
    *reinterpret_cast<bf16x8*>(y + row * D + off) = pack8(o);
    if (!AFFINE && y2) {
      unpack8(*reinterpret_cast<const bf16x8*>(shift2 + b * mod_stride + off), fsh);
      unpack8(*reinterpret_cast<const bf16x8*>(scale2 + b * mod_stride + off), fsc);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(v[j] * rstd, 1.0f + fsc[j], fsh[j]);
      *reinterpret_cast<bf16x8*>(y2 + row * D + off) = pack8(o);
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ln_modulate_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ scale,
                       const __nv_bfloat16* __restrict__ scale2, int64_t mod_stride,
                       const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy2,
                       __nv_bfloat16* __restrict__ dx, int accumulate, int64_t rows, int64_t S,
                       float eps) {
  pdl_trigger();   // programmatic dependent launch (common.cuh): no-ops unless the launch carries the attribute
  pdl_wait();
  constexpr int D = NV * 256;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int64_t b = row / S;
  float v[NV][8], g[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    unpack8(*reinterpret_cast<const bf16x8*>(x + row * D + (i * 32 + lane) * 8), v[i]);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[i][j];
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[i][j] -= mean;
      q = fmaf(v[i][j], v[i][j], q);
    }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int off = (i * 32 + lane) * 8;
    float fsc[8], fd[8];
    unpack8(*reinterpret_cast<const bf16x8*>(scale + b * mod_stride + off), fsc);
    unpack8(*reinterpret_cast<const bf16x8*>(dy + row * D + off), fd);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[i][j] = fd[j] * (1.0f + fsc[j]);
    if (dy2) {
      unpack8(*reinterpret_cast<const bf16x8*>(scale2 + b * mod_stride + off), fsc);
      unpack8(*reinterpret_cast<const bf16x8*>(dy2 + row * D + off), fd);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[i][j] = fmaf(fd[j], 1.0f + fsc[j], g[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[i][j] *= rstd;  // xhat
      sg += g[i][j];
      sgx = fmaf(g[i][j], v[i][j], sgx);
    }
  }
  sg = warp_sum(sg) * (1.0f / D);
  sgx = warp_sum(sgx) * (1.0f / D);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int off = (i * 32 + lane) * 8;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - sg - v[i][j] * sgx);
    if (accumulate) {
      float prev[8];
      unpack8(*reinterpret_cast<const bf16x8*>(dx + row * D + off), prev);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += prev[j];
    }
    *reinterpret_cast<bf16x8*>(dx + row * D + off) = pack8(o);
  }
}

// One CTA per joint token; thread t owns 16-byte chunk t of each of the q, k, v sections
// (H*D/8 threads); 8 consecutive lanes hold one head (D = 64).
__global__ void qk_norm_concat_fwd_kernel(const __nv_bfloat16* __restrict__ qkv_img,
                                          const __nv_bfloat16* __restrict__ qkv_txt,
                                          const __nv_bfloat16* __restrict__ wq_img,
                                          const __nv_bfloat16* __restrict__ wk_img,
                                          const __nv_bfloat16* __restrict__ wq_txt,
                                          const __nv_bfloat16* __restrict__ wk_txt,
                                          __nv_bfloat16* __restrict__ out, int64_t S_img,
                                          int64_t S_txt, int HD, float eps) {
  pdl_trigger();   // programmatic dependent launch (common.cuh): no-ops unless the launch carries the attribute
  pdl_wait();
  const int64_t S = S_img + S_txt;
  const int64_t b = blockIdx.x / S, s = blockIdx.x % S;
  const bool is_img = s < S_img;
  const __nv_bfloat16* src = is_img ? qkv_img + (b * S_img + s) * 3 * HD
                                    : qkv_txt + (b * S_txt + (s - S_img)) * 3 * HD;
  __nv_bfloat16* dst = out + (b * S + s) * 3 * HD;
  const int t = threadIdx.x;
  const int chunk = (t & 7) * 8;
#pragma unroll
  for (int sec = 0; sec < 3; ++sec) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(src + sec * HD + t * 8), f);
    const __nv_bfloat16* w = sec == 0 ? (is_img ? wq_img : wq_txt) : (sec == 1 ? (is_img ? wk_img : wk_txt) : nullptr);
    if (w) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
      ss += __shfl_xor_sync(0xffffffffu, ss, 1);
      ss += __shfl_xor_sync(0xffffffffu, ss, 2);
      ss += __shfl_xor_sync(0xffffffffu, ss, 4);
      const float r = rsqrtf(ss * (1.0f / 64.0f) + eps);
      float fw[8];
      unpack8(*reinterpret_cast<const bf16x8*>(w + chunk), fw);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = f[j] * r * fw[j];
    }
    *reinterpret_cast<bf16x8*>(dst + sec * HD + t * 8) = pack8(f);
  }
}

__global__ void qk_norm_concat_bwd_kernel(const __nv_bfloat16* __restrict__ qkv_img,
                                          const __nv_bfloat16* __restrict__ qkv_txt,
                                          const __nv_bfloat16* __restrict__ wq_img,
                                          const __nv_bfloat16* __restrict__ wk_img,
                                          const __nv_bfloat16* __restrict__ wq_txt,
                                          const __nv_bfloat16* __restrict__ wk_txt,
                                          const __nv_bfloat16* __restrict__ dout,
                                          __nv_bfloat16* __restrict__ dq_img,
                                          __nv_bfloat16* __restrict__ dq_txt, int64_t S_img,
                                          int64_t S_txt, int HD, float eps) {
  pdl_trigger();   // programmatic dependent launch (common.cuh): no-ops unless the launch carries the attribute
  pdl_wait();
  const int64_t S = S_img + S_txt;
  const int64_t b = blockIdx.x / S, s = blockIdx.x % S;
  const bool is_img = s < S_img;
  const int64_t src_off = is_img ? (b * S_img + s) * 3 * HD : (b * S_txt + (s - S_img)) * 3 * HD;
  const __nv_bfloat16* src = (is_img ? qkv_img : qkv_txt) + src_off;
  __nv_bfloat16* dst = (is_img ? dq_img : dq_txt) + src_off;
  const __nv_bfloat16* g = dout + (b * S + s) * 3 * HD;
  const int t = threadIdx.x;
  const int chunk = (t & 7) * 8;
#pragma unroll
  for (int sec = 0; sec < 3; ++sec) {
    float dy[8];
    unpack8(*reinterpret_cast<const bf16x8*>(g + sec * HD + t * 8), dy);
    const __nv_bfloat16* w = sec == 0 ? (is_img ? wq_img : wq_txt) : (sec == 1 ? (is_img ? wk_img : wk_txt) : nullptr);
    if (w) {
      float f[8], fw[8];
      unpack8(*reinterpret_cast<const bf16x8*>(src + sec * HD + t * 8), f);
      unpack8(*reinterpret_cast<const bf16x8*>(w + chunk), fw);
      float ss = 0.f, dot = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dy[j] *= fw[j];
        ss = fmaf(f[j], f[j], ss);
        dot = fmaf(f[j], dy[j], dot);
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        dot += __shfl_xor_sync(0xffffffffu, dot, o);
      }
      const float r = rsqrtf(ss * (1.0f / 64.0f) + eps);
      const float c = r * r * r * dot * (1.0f / 64.0f);
#pragma unroll
      for (int j = 0; j < 8; ++j) dy[j] = r * dy[j] - f[j] * c;
    }
    *reinterpret_cast<bf16x8*>(dst + sec * HD + t * 8) = pack8(dy);
  }
}

__global__ void __launch_bounds__(256)
row_gate_mul_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gate, int64_t gate_stride,
                    int64_t rows_per_gate, __nv_bfloat16* __restrict__ out, int64_t M, int nvec) {
  pdl_trigger();   // programmatic dependent launch (common.cuh): no-ops unless the launch carries the attribute
  pdl_wait();
  const int64_t total = M * nvec;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t m = i / nvec;
    const int c = (int)(i - m * nvec) * 8;
    float a[8], g[8];
    unpack8(*reinterpret_cast<const bf16x8*>(x + i * 8), a);
    unpack8(*reinterpret_cast<const bf16x8*>(gate + (m / rows_per_gate) * gate_stride + c), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= g[j];
    *reinterpret_cast<bf16x8*>(out + i * 8) = pack8(a);
  }
}

}  // namespace
}  // namespace advgrpo

using namespace advgrpo;

#define DISPATCH_NV(NVV, ...)                          \
  switch (NVV) {                                       \
    case 1: { constexpr int NV = 1; __VA_ARGS__; break; } \
    case 2: { constexpr int NV = 2; __VA_ARGS__; break; } \
    case 3: { constexpr int NV = 3; __VA_ARGS__; break; } \
    case 4: { constexpr int NV = 4; __VA_ARGS__; break; } \
    case 5: { constexpr int NV = 5; __VA_ARGS__; break; } \
    case 6: { constexpr int NV = 6; __VA_ARGS__; break; } \
    case 8: { constexpr int NV = 8; __VA_ARGS__; break; } \
    default: return set_error(ADVGRPO_ERR_UNSUPPORTED, "ln_modulate: unsupported D=%lld", (long long)D); \
  }

extern "C" {

int advgrpo_ln_modulate_fwd(const void* x, const void* shift, const void* scale, const void* shift2,
                            const void* scale2, int64_t mod_stride, void* y, void* y2, int64_t B,
                            int64_t S, int64_t D, float eps, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(x && shift && scale && y, "ln_modulate_fwd: null pointer");
  ADVGRPO_CHECK_ARG(!y2 || (shift2 && scale2), "ln_modulate_fwd: y2 needs shift2/scale2");
  ADVGRPO_CHECK_ARG(D % 256 == 0 && D <= 2048 && mod_stride % 8 == 0, "ln_modulate_fwd: D=%lld must be a multiple of 256 (<= 2048), mod_stride a multiple of 8", (long long)D);
  ADVGRPO_CHECK_ARG(aligned16(x) && aligned16(shift) && aligned16(scale) && aligned16(y) &&
                        (!y2 || (aligned16(y2) && aligned16(shift2) && aligned16(scale2))),
                    "ln_modulate_fwd: tensors must be 16-byte aligned");
  const int64_t rows = B * S;
  if (rows == 0) return ADVGRPO_OK;
  const unsigned grid = (unsigned)((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
  cudaError_t le = cudaSuccess;
  DISPATCH_NV((int)(D / 256), le = launch_chain(ln_modulate_fwd_kernel<NV, false>, dim3(grid), dim3(kWarpsPerBlock * 32), 0,
      (cudaStream_t)stream, 1, (const __nv_bfloat16*)x, (const __nv_bfloat16*)shift, (const __nv_bfloat16*)scale,
      (const __nv_bfloat16*)shift2, (const __nv_bfloat16*)scale2, mod_stride, (__nv_bfloat16*)y,
      (__nv_bfloat16*)y2, rows, S, eps));
  ADVGRPO_CUDA_CALL(le);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_layer_norm_affine(const void* x, const void* weight, const void* bias, void* y, int64_t rows,
                              int64_t D, float eps, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(x && weight && bias && y, "layer_norm_affine: null pointer");
  ADVGRPO_CHECK_ARG(D % 256 == 0 && D <= 2048, "layer_norm_affine: D=%lld must be a multiple of 256 (<= 2048)", (long long)D);
  ADVGRPO_CHECK_ARG(aligned16(x) && aligned16(weight) && aligned16(bias) && aligned16(y),
                    "layer_norm_affine: tensors must be 16-byte aligned");
  if (rows <= 0) return ADVGRPO_OK;
  const unsigned grid = (unsigned)((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
  DISPATCH_NV((int)(D / 256), ln_modulate_fwd_kernel<NV, true><<<grid, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)bias, (const __nv_bfloat16*)weight, nullptr, nullptr, 0,
      (__nv_bfloat16*)y, nullptr, rows, rows, eps));
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_ln_modulate_bwd(const void* x, const void* scale, const void* scale2,
                            int64_t mod_stride, const void* dy, const void* dy2, void* dx,
                            int accumulate, int64_t B, int64_t S, int64_t D, float eps,
                            advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(x && scale && dy && dx, "ln_modulate_bwd: null pointer");
  ADVGRPO_CHECK_ARG(!dy2 || scale2, "ln_modulate_bwd: dy2 needs scale2");
  ADVGRPO_CHECK_ARG(D % 256 == 0 && D <= 2048 && mod_stride % 8 == 0, "ln_modulate_bwd: D=%lld must be a multiple of 256 (<= 2048)", (long long)D);
  ADVGRPO_CHECK_ARG(aligned16(x) && aligned16(scale) && aligned16(dy) && aligned16(dx) &&
                        (!dy2 || (aligned16(dy2) && aligned16(scale2))),
                    "ln_modulate_bwd: tensors must be 16-byte aligned");
  const int64_t rows = B * S;
  if (rows == 0) return ADVGRPO_OK;
  const unsigned grid = (unsigned)((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
  cudaError_t le = cudaSuccess;
  DISPATCH_NV((int)(D / 256), le = launch_chain(ln_modulate_bwd_kernel<NV>, dim3(grid), dim3(kWarpsPerBlock * 32), 0,
      (cudaStream_t)stream, 1, (const __nv_bfloat16*)x, (const __nv_bfloat16*)scale, (const __nv_bfloat16*)scale2, mod_stride,
      (const __nv_bfloat16*)dy, (const __nv_bfloat16*)dy2, (__nv_bfloat16*)dx, accumulate, rows, S, eps));
  ADVGRPO_CUDA_CALL(le);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_row_gate_mul(const void* x, const void* gate, int64_t gate_stride, int64_t rows_per_gate, void* out,
                         int64_t M, int64_t N, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(x && gate && out, "row_gate_mul: null pointer");
  ADVGRPO_CHECK_ARG(N % 8 == 0 && gate_stride % 8 == 0 && rows_per_gate >= 1, "row_gate_mul: N and gate_stride must be multiples of 8");
  ADVGRPO_CHECK_ARG(aligned16(x) && aligned16(gate) && aligned16(out), "row_gate_mul: 16-byte alignment");
  if (M == 0 || N == 0) return ADVGRPO_OK;
  const int64_t total = M * (N / 8);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  ADVGRPO_CUDA_CALL(launch_chain(row_gate_mul_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, 1,
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)gate, gate_stride, rows_per_gate, (__nv_bfloat16*)out, M, (int)(N / 8)));
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_qk_norm_concat_fwd(const void* qkv_img, const void* qkv_txt, const void* wq_img,
                               const void* wk_img, const void* wq_txt, const void* wk_txt,
                               void* qkv_joint, int64_t B, int64_t S_img, int64_t S_txt, int64_t H,
                               int64_t D, float eps, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(qkv_img && qkv_joint, "qk_norm_concat_fwd: null pointer");
  ADVGRPO_CHECK_ARG(D == 64, "qk_norm_concat_fwd: head_dim must be 64 (got %lld)", (long long)D);
  ADVGRPO_CHECK_ARG(H >= 1 && H * D / 8 <= 1024 && (H * D / 8) % 32 == 0, "qk_norm_concat_fwd: unsupported H=%lld", (long long)H);
  ADVGRPO_CHECK_ARG((S_txt == 0) == (qkv_txt == nullptr), "qk_norm_concat_fwd: qkv_txt must be given iff S_txt > 0");
  ADVGRPO_CHECK_ARG((wq_img == nullptr) == (wk_img == nullptr), "qk_norm_concat_fwd: wq_img/wk_img must both be set or both NULL");
  ADVGRPO_CHECK_ARG(aligned16(qkv_img) && aligned16(qkv_joint) && (!qkv_txt || aligned16(qkv_txt)), "qk_norm_concat_fwd: 16-byte alignment");
  const int64_t tokens = B * (S_img + S_txt);
  if (tokens == 0) return ADVGRPO_OK;
  ADVGRPO_CUDA_CALL(launch_chain(qk_norm_concat_fwd_kernel, dim3((unsigned)tokens), dim3((unsigned)(H * D / 8)), 0,
      (cudaStream_t)stream, 1, (const __nv_bfloat16*)qkv_img, (const __nv_bfloat16*)qkv_txt, (const __nv_bfloat16*)wq_img,
      (const __nv_bfloat16*)wk_img, (const __nv_bfloat16*)wq_txt, (const __nv_bfloat16*)wk_txt,
      (__nv_bfloat16*)qkv_joint, S_img, S_txt, (int)(H * D), eps));
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_qk_norm_concat_bwd(const void* qkv_img, const void* qkv_txt, const void* wq_img,
                               const void* wk_img, const void* wq_txt, const void* wk_txt,
                               const void* dqkv_joint, void* dqkv_img, void* dqkv_txt, int64_t B,
                               int64_t S_img, int64_t S_txt, int64_t H, int64_t D, float eps,
                               advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(qkv_img && dqkv_joint && dqkv_img, "qk_norm_concat_bwd: null pointer");
  ADVGRPO_CHECK_ARG(D == 64, "qk_norm_concat_bwd: head_dim must be 64 (got %lld)", (long long)D);
  ADVGRPO_CHECK_ARG(H >= 1 && H * D / 8 <= 1024 && (H * D / 8) % 32 == 0, "qk_norm_concat_bwd: unsupported H=%lld", (long long)H);
  ADVGRPO_CHECK_ARG((S_txt == 0) == (qkv_txt == nullptr) && (S_txt == 0) == (dqkv_txt == nullptr), "qk_norm_concat_bwd: text tensors must be given iff S_txt > 0");
  const int64_t tokens = B * (S_img + S_txt);
  if (tokens == 0) return ADVGRPO_OK;
  ADVGRPO_CUDA_CALL(launch_chain(qk_norm_concat_bwd_kernel, dim3((unsigned)tokens), dim3((unsigned)(H * D / 8)), 0,
      (cudaStream_t)stream, 1, (const __nv_bfloat16*)qkv_img, (const __nv_bfloat16*)qkv_txt, (const __nv_bfloat16*)wq_img,
      (const __nv_bfloat16*)wk_img, (const __nv_bfloat16*)wq_txt, (const __nv_bfloat16*)wk_txt,
      (const __nv_bfloat16*)dqkv_joint, (__nv_bfloat16*)dqkv_img, (__nv_bfloat16*)dqkv_txt, S_img, S_txt,
      (int)(H * D), eps));
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

}  // extern "C"

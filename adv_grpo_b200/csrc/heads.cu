// Reward-head and discriminator-step kernels (latency / HBM-bound, warp-shuffle reductions, 16-byte vector access):
//   * gather_rows_l2norm   : CLS + the sampled patch tokens of every image -> one [B (1 + n), D] row matrix, optionally
//                            L2-normalised with the reference's bf16 rounding order (adv_grpo/rewards.py:399-412)
//   * head_logits          : logit[r] = <a[r, :], w2> + b2 -- the Linear(hidden, 1) of DINOHead
//                            (scripts/train_sd3_fast_dino_patch.py:592-603) behind the tcgen05 GEMM + GELU of its first layer
//   * dino_hybrid          : 0.7 cls + 0.3 mean(patch scores) per image (rewards.py:414-419)
//   * dino_hinge           : hinge loss of train_dino (train_sd3_fast_dino_patch.py:186-219), accuracy, d loss / d logit
//   * head_dz              : dz = (dl w2) * gelu_erf'(z) -- backward through Linear(hidden, 1) and the GELU in one pass
//   * col_sum              : sum_r s[r] a[r, c] b[r, c] (bias gradients, dw2 of the head), deterministic two-stage reduction
//   * ln_modulation_grads  : per-sample d shift = sum_t dy, d scale = sum_t dy * xhat of an adaLN LayerNorm-modulate (full
//                            fine-tuning: the modulation vectors come from trainable linears), row statistics + segmented
//                            column sums
//   * layer_norm_affine_bwd: dx of an affine LayerNorm + d weight / d bias (the trainable CLIP blocks of the PickScore
//                            discriminator step, scripts/train_sd3_fast_pickscore.py:1016-1029)
//   * adam_torch_order     : torch.optim.Adam's multi-tensor update (the discriminator optimizers, train_pick:658,
//                            train_dino:750: Adam(lr = d_lr, betas = (0.5, 0.999))) as ONE pass per tensor, rounding to the
//                            parameter dtype after every torch._foreach_* step so bf16 parameters follow the reference bit
//                            for bit (lerp_, mul_, addcmul_, sqrt, div_, add_, addcdiv_)
//   * row_softmax_f32      : softmax over fp32 rows, output rounded to TF32 (the VAE mid-block attention's P matrix
//                            between two TF32 tensor-core products)
//   * pickscore_head       : logit_scale * <img / |img|, txt / |txt|> / 26 with the reference's bf16 rounding order
//                            (adv_grpo/pickscore_scorer.py:44-51)
#include <math.h>

#include "common.cuh"

namespace advgrpo {
namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// d/dz of the erf GELU, accurate erff / expf (tiny tensors: accuracy over speed)
__device__ __forceinline__ float gelu_erf_grad_acc(float z) {
  const float cdf = 0.5f * (1.0f + erff(z * 0.7071067811865476f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * z * z);
  return fmaf(z, pdf, cdf);
}

// ------------------------------------------------------------------------------------------------ gather + L2 norm
// One warp per output row.  row = b (1 + n) + j: j = 0 -> token 0 (CLS), j >= 1 -> token 1 + idx[b, j - 1].
__global__ void __launch_bounds__(kWarps * 32)
gather_rows_l2norm_kernel(const __nv_bfloat16* __restrict__ feats, const int64_t* __restrict__ idx,
                          __nv_bfloat16* __restrict__ out, int64_t B, int64_t T, int64_t n, int D, int l2norm,
                          float eps) {
  const int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= B * (1 + n)) return;
  const int lane = threadIdx.x & 31;
  const int64_t b = row / (1 + n), j = row % (1 + n);
  int64_t tok = 0;
  if (j > 0) {
    tok = idx[b * n + (j - 1)];
    tok = tok < 0 ? 0 : (tok > T - 2 ? T - 2 : tok);
    tok += 1;
  }
  const __nv_bfloat16* src = feats + (b * T + tok) * D;
  __nv_bfloat16* dst = out + row * D;
  const int nvec = D >> 3;
  float ss = 0.f;
  if (l2norm) {
    for (int v = lane; v < nvec; v += 32) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(src + v * 8), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) ss = fmaf(f[k], f[k], ss);
    }
    ss = warp_sum(ss);
  }
  // x / (bf16(|x|) + 1e-6) evaluated like the bf16 tensor expression: norm, sum and quotient each rounded to bf16
  const float denom = l2norm ? bf16_round(bf16_round(sqrtf(ss)) + eps) : 1.0f;
  for (int v = lane; v < nvec; v += 32) {
    bf16x8 raw = *reinterpret_cast<const bf16x8*>(src + v * 8);
    if (l2norm) {
      float f[8];
      unpack8(raw, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = __fdiv_rn(f[k], denom);
      raw = pack8(f);
    }
    *reinterpret_cast<bf16x8*>(dst + v * 8) = raw;
  }
}

// ------------------------------------------------------------------------------------------------ Linear(hidden, 1)
__global__ void __launch_bounds__(kWarps * 32)
head_logits_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ w2,
                   const __nv_bfloat16* __restrict__ b2, float* __restrict__ logits, int64_t R, int Hd, int round_bf16) {
  const int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= R) return;
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int v = lane; v < (Hd >> 3); v += 32) {
    float f[8], w[8];
    unpack8(*reinterpret_cast<const bf16x8*>(a + row * Hd + v * 8), f);
    unpack8(*reinterpret_cast<const bf16x8*>(w2 + v * 8), w);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fmaf(f[k], w[k], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    acc += b2 ? __bfloat162float(b2[0]) : 0.f;
    logits[row] = round_bf16 ? bf16_round(acc) : acc;     // a bf16 Linear's output is rounded once
  }
}

// hybrid[b] = cls_w cls + (1 - cls_w) mean_j patch, each step rounded to bf16 like the bf16 tensor expression
__global__ void dino_hybrid_kernel(const float* __restrict__ logits, float* __restrict__ hybrid, int64_t B, int64_t n,
                                   float cls_w, int round_bf16) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* l = logits + b * (1 + n);
  float s = 0.f;
  for (int64_t j = 1; j <= n; ++j) s += l[j];
  const float mean = n > 0 ? s / (float)n : 0.f;
  hybrid[b] = round_bf16 ? bf16_round(bf16_round(cls_w * l[0]) + bf16_round((1.0f - cls_w) * bf16_round(mean)))
                         : fmaf(cls_w, l[0], (1.0f - cls_w) * mean);
}

// Single block.  Images [0, B_real) are real, the rest fake.  loss = 0.5 (mean relu(1 - l_real_cls) + mean relu(1 + l_fake_cls))
// + w_patch 0.5 (the same over the patch logits).  out[0] = loss, out[1] = accuracy on the CLS logits, out[2] = sum of dl
// (the gradient of the output bias); dl[r] = d loss / d logit[r].
__global__ void __launch_bounds__(256)
dino_hinge_kernel(const float* __restrict__ logits, float* __restrict__ dl, float* __restrict__ out, int64_t B_real,
                  int64_t B_fake, int64_t n, float w_patch) {
  __shared__ float scratch[32];
  const int64_t R = (B_real + B_fake) * (1 + n);
  const float c_cls_r = 0.5f / (float)B_real, c_cls_f = 0.5f / (float)B_fake;
  const float c_p_r = n > 0 ? w_patch * 0.5f / (float)(B_real * n) : 0.f;
  const float c_p_f = n > 0 ? w_patch * 0.5f / (float)(B_fake * n) : 0.f;
  float loss = 0.f, acc = 0.f, sdl = 0.f;
  for (int64_t r = threadIdx.x; r < R; r += blockDim.x) {
    const int64_t img = r / (1 + n), j = r % (1 + n);
    const bool real = img < B_real;
    const float l = logits[r];
    const float c = j == 0 ? (real ? c_cls_r : c_cls_f) : (real ? c_p_r : c_p_f);
    const float margin = real ? 1.0f - l : 1.0f + l;
    float d = 0.f;
    if (margin > 0.f) {
      loss = fmaf(c, margin, loss);
      d = real ? -c : c;
    }
    dl[r] = d;
    sdl += d;
    if (j == 0) acc += (real ? (l > 0.f) : (l < 0.f)) ? (real ? c_cls_r : c_cls_f) : 0.f;
  }
  loss = block_sum(loss, scratch);
  acc = block_sum(acc, scratch);
  sdl = block_sum(sdl, scratch);
  if (threadIdx.x == 0) {
    out[0] = loss;
    out[1] = acc;
    out[2] = sdl;
  }
}

// dz[r, :] = bf16(bf16(dl[r] w2[:]) gelu'(z[r, :]))
__global__ void __launch_bounds__(kWarps * 32)
head_dz_kernel(const float* __restrict__ dl, const __nv_bfloat16* __restrict__ w2, const __nv_bfloat16* __restrict__ z,
               __nv_bfloat16* __restrict__ dz, int64_t R, int Hd) {
  const int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= R) return;
  const int lane = threadIdx.x & 31;
  const float d = dl[row];
  for (int v = lane; v < (Hd >> 3); v += 32) {
    float zz[8], w[8], o[8];
    unpack8(*reinterpret_cast<const bf16x8*>(z + row * Hd + v * 8), zz);
    unpack8(*reinterpret_cast<const bf16x8*>(w2 + v * 8), w);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = bf16_round(d * w[k]) * gelu_erf_grad_acc(zz[k]);
    *reinterpret_cast<bf16x8*>(dz + row * Hd + v * 8) = pack8(o);
  }
}

// ------------------------------------------------------------------------------------------------ column sums
// MODE 0: sum_r s[r] a[r, c] (s optional);  MODE 1: sum_r a[r, c] b[r, c];
// MODE 2 (LayerNorm parameter gradients): out0[c] = sum_r a[r, c] xhat[r, c], out1[c] = sum_r a[r, c] with
//        xhat = (b[r, c] - stats[r].mean) stats[r].rstd   (a = dy, b = x).
// Block = 8 warps x 256 columns; warp w of block (bx, by) walks rows by * rpb + w, + 8, ...; the eight warps are combined
// through shared memory in a fixed order, the row slabs by the finish kernel in a fixed order: bit-reproducible.
template <int MODE>
__global__ void __launch_bounds__(kWarps * 32)
col_sum_partial_kernel(const __nv_bfloat16* __restrict__ a, int64_t lda, const __nv_bfloat16* __restrict__ b, int64_t ldb,
                       const float* __restrict__ s, const float2* __restrict__ stats, float* __restrict__ partial,
                       int64_t rows, int64_t C, int64_t rows_per_block) {
  __shared__ float sm[(MODE == 2 ? 2 : 1) * kWarps * 256];
  // blockIdx.z = segment (rows [z * rows, (z + 1) * rows) of the matrices, its own slab of partials): per-sample sums
  {
    const int64_t seg = blockIdx.z;
    a += seg * rows * lda;
    if (b) b += seg * rows * ldb;
    if (s) s += seg * rows;
    if (stats) stats += seg * rows;
    partial += seg * (int64_t)(MODE == 2 ? 2 : 1) * gridDim.y * C;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t c0 = (int64_t)blockIdx.x * 256 + lane * 8;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r_end = min(rows, r_begin + rows_per_block);
  float acc[8], acc2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = acc2[k] = 0.f;
  if (c0 < C) {
    for (int64_t r = r_begin + warp; r < r_end; r += kWarps) {
      float fa[8];
      unpack8(*reinterpret_cast<const bf16x8*>(a + r * lda + c0), fa);
      if (MODE == 0) {
        const float w = s ? s[r] : 1.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(w, fa[k], acc[k]);
      } else {
        float fb[8];
        unpack8(*reinterpret_cast<const bf16x8*>(b + r * ldb + c0), fb);
        if (MODE == 1) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fmaf(fa[k], fb[k], acc[k]);
        } else {
          const float2 st = stats[r];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            acc[k] = fmaf(fa[k], (fb[k] - st.x) * st.y, acc[k]);
            acc2[k] += fa[k];
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sm[warp * 256 + lane * 8 + k] = acc[k];
    if (MODE == 2) sm[kWarps * 256 + warp * 256 + lane * 8 + k] = acc2[k];
  }
  __syncthreads();
  const int c = threadIdx.x;                     // 256 threads = 256 columns of the block
  const int64_t col = (int64_t)blockIdx.x * 256 + c;
  if (col < C) {
    float t = 0.f, t2 = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      t += sm[w * 256 + c];
      if (MODE == 2) t2 += sm[kWarps * 256 + w * 256 + c];
    }
    partial[(int64_t)blockIdx.y * C + col] = t;
    if (MODE == 2) partial[((int64_t)gridDim.y + blockIdx.y) * C + col] = t2;
  }
}

__global__ void col_sum_finish_kernel(const float* __restrict__ partial, float* __restrict__ out0, float* __restrict__ out1,
                                      int64_t C, int ny) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  partial += (int64_t)blockIdx.y * (out1 ? 2 : 1) * ny * C;      // blockIdx.y = segment
  out0 += (int64_t)blockIdx.y * C;
  if (out1) out1 += (int64_t)blockIdx.y * C;
  float t = 0.f, t2 = 0.f;
  for (int y = 0; y < ny; ++y) {
    t += partial[(int64_t)y * C + c];
    if (out1) t2 += partial[((int64_t)ny + y) * C + c];
  }
  out0[c] = t;
  if (out1) out1[c] = t2;
}

int col_sum_ny(int64_t rows) {
  int64_t ny = (rows + 63) / 64;
  return (int)(ny < 1 ? 1 : (ny > 64 ? 64 : ny));
}

// per-row LayerNorm statistics (mean, rstd) of a bf16 [rows, D] matrix, one warp per row
__global__ void __launch_bounds__(kWarps * 32)
row_stats_kernel(const __nv_bfloat16* __restrict__ x, float2* __restrict__ stats, int64_t rows, int D, float eps) {
  const int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  float s = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(x + row * D + v * 8), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += f[k];
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(x + row * D + v * 8), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float d = f[k] - mean;
      q = fmaf(d, d, q);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  if (lane == 0) stats[row] = make_float2(mean, rstd);
}

// ------------------------------------------------------------------------------------------------ affine LayerNorm bwd
// One warp per row, any D that is a multiple of 8 and <= 2048 (8 chunks of 256 per lane at most).
__global__ void __launch_bounds__(kWarps * 32)
ln_affine_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ weight,
                     const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, float2* __restrict__ stats,
                     int64_t rows, int D, float eps) {
  const int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nvec = D >> 3;
  float v[8][8], g[8][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = i * 32 + lane;
    if (c < nvec) {
      unpack8(*reinterpret_cast<const bf16x8*>(x + row * D + c * 8), v[i]);
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[i][k];
    }
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i * 32 + lane < nvec) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        v[i][k] -= mean;
        q = fmaf(v[i][k], v[i][k], q);
      }
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
  if (stats && lane == 0) stats[row] = make_float2(mean, rstd);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = i * 32 + lane;
    if (c < nvec) {
      float fw[8], fd[8];
      unpack8(*reinterpret_cast<const bf16x8*>(weight + c * 8), fw);
      unpack8(*reinterpret_cast<const bf16x8*>(dy + row * D + c * 8), fd);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        g[i][k] = fd[k] * fw[k];
        v[i][k] *= rstd;
        sg += g[i][k];
        sgx = fmaf(g[i][k], v[i][k], sgx);
      }
    }
  }
  sg = warp_sum(sg) / (float)D;
  sgx = warp_sum(sgx) / (float)D;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = i * 32 + lane;
    if (c < nvec) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = rstd * (g[i][k] - sg - v[i][k] * sgx);
      *reinterpret_cast<bf16x8*>(dx + row * D + c * 8) = pack8(o);
    }
  }
}

// ------------------------------------------------------------------------------------------------ Adam, torch op order
template <typename T> __device__ __forceinline__ float ld_as_float(const T* p, int64_t i);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p, int64_t i) { return p[i]; }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) {
  return __bfloat162float(p[i]);
}
template <typename T> __device__ __forceinline__ float rnd(float x);
template <> __device__ __forceinline__ float rnd<float>(float x) { return x; }
template <> __device__ __forceinline__ float rnd<__nv_bfloat16>(float x) { return bf16_round(x); }
__device__ __forceinline__ void st_from_float(float* p, int64_t i, float v) { p[i] = v; }
__device__ __forceinline__ void st_from_float(__nv_bfloat16* p, int64_t i, float v) { p[i] = __float2bfloat16_rn(v); }

struct AdamT {
  float w;          // 1 - beta1 (the lerp weight)
  float beta2, beta2_c, bc2_sqrt, eps, neg_step_size;
};

// torch/optim/adam.py::_multi_tensor_adam (weight_decay = 0, amsgrad = False, maximize = False, not capturable); every
// _foreach_ op computes in fp32 and stores in the tensor dtype, which R() reproduces:
//   exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2); exp_avg_sq.addcmul_(grad, grad, value = 1 - beta2);
//   d = exp_avg_sq.sqrt(); d.div_(sqrt(1 - beta2^t)); d.add_(eps); param.addcdiv_(exp_avg, d, value = -lr / (1 - beta1^t))
template <typename T, typename G>
__global__ void __launch_bounds__(256)
adam_torch_order_kernel(T* __restrict__ p, const G* __restrict__ g, T* __restrict__ m, T* __restrict__ v, int64_t n,
                        const AdamT a, int zero_grad, G* g_mut) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float gi = ld_as_float<G>(g, i);
    const float gt = rnd<T>(gi);                        // the optimizer sees the gradient in the parameter dtype
    float mi = ld_as_float<T>(m, i), vi = ld_as_float<T>(v, i), pi = ld_as_float<T>(p, i);
    const float diff = gt - mi;                        // ATen lerp: weight < 0.5 ? self + w diff : end - diff (1 - w)
    mi = rnd<T>(fabsf(a.w) < 0.5f ? fmaf(a.w, diff, mi) : fmaf(-diff, 1.0f - a.w, gt));
    vi = rnd<T>(vi * a.beta2);
    vi = rnd<T>(fmaf(a.beta2_c * gt, gt, vi));
    float d = rnd<T>(sqrtf(vi));
    d = rnd<T>(__fdiv_rn(d, a.bc2_sqrt));
    d = rnd<T>(d + a.eps);
    pi = rnd<T>(fmaf(a.neg_step_size, __fdiv_rn(mi, d), pi));
    st_from_float(p, i, pi);
    st_from_float(m, i, mi);
    st_from_float(v, i, vi);
    if (zero_grad) st_from_float(g_mut, i, 0.f);
  }
}

// ------------------------------------------------------------------------------------------------ fp32 row softmax
// One block per row: y = softmax(scale * x) over `cols` fp32 values, rounded to the nearest TF32 value when round_tf32
// (the row is the A operand of a TF32 tensor-core product whose truncation is then exact).  In place allowed.
__global__ void __launch_bounds__(256)
row_softmax_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t cols, float scale_log2, int round_tf32) {
  __shared__ float scratch[32];
  const float* xr = x + (int64_t)blockIdx.x * cols;
  float* yr = y + (int64_t)blockIdx.x * cols;
  float mx = -INFINITY;
  for (int64_t c = threadIdx.x * 4; c < cols; c += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  mx = warp_max(mx);
  {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) scratch[w] = mx;
    __syncthreads();
    mx = warp_max(l < 8 ? scratch[l] : -INFINITY);
    __syncthreads();
  }
  const float off = mx * scale_log2;
  float s = 0.f;
  for (int64_t c = threadIdx.x * 4; c < cols; c += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    s += (ex2f(fmaf(v.x, scale_log2, -off)) + ex2f(fmaf(v.y, scale_log2, -off))) +
         (ex2f(fmaf(v.z, scale_log2, -off)) + ex2f(fmaf(v.w, scale_log2, -off)));
  }
  s = block_sum(s, scratch);
  const float inv = 1.0f / s;
  for (int64_t c = threadIdx.x * 4; c < cols; c += 256 * 4) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    float o[4] = {ex2f(fmaf(v.x, scale_log2, -off)) * inv, ex2f(fmaf(v.y, scale_log2, -off)) * inv,
                  ex2f(fmaf(v.z, scale_log2, -off)) * inv, ex2f(fmaf(v.w, scale_log2, -off)) * inv};
    if (round_tf32) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(o[k]));
        o[k] = __uint_as_float(u);
      }
    }
    *reinterpret_cast<float4*>(yr + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------------ PickScore head
// One warp per image.  txt row = b % n_txt.  bf16 mode (the reference's bf16 model, pickscore_scorer.py:44-51): features are
// normalised with bf16 rounding of the norm and the quotient, the dot is an fp32 accumulation rounded to bf16, then
// logit_scale.exp() * dot and / 26 each rounded to bf16.  f32 mode: plain fp32.
__global__ void __launch_bounds__(kWarps * 32)
pickscore_head_kernel(const __nv_bfloat16* __restrict__ img, const __nv_bfloat16* __restrict__ txt,
                      const int64_t* __restrict__ txt_index, const void* __restrict__ logit_scale, int logit_scale_is_bf16,
                      float* __restrict__ out, int64_t B, int64_t n_txt, int D, int bf16_mode) {
  const int64_t b = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  const __nv_bfloat16* ir = img + b * D;
  int64_t ti = txt_index ? txt_index[b] : b % n_txt;
  ti = ti < 0 ? 0 : (ti >= n_txt ? n_txt - 1 : ti);
  const __nv_bfloat16* tr = txt + ti * D;
  const float ls = logit_scale_is_bf16 ? __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(logit_scale))
                                       : *reinterpret_cast<const float*>(logit_scale);
  const float logit_scale_exp = expf(ls);
  float si = 0.f, st = 0.f;
  for (int v = lane; v < (D >> 3); v += 32) {
    float fi[8], ft[8];
    unpack8(*reinterpret_cast<const bf16x8*>(ir + v * 8), fi);
    unpack8(*reinterpret_cast<const bf16x8*>(tr + v * 8), ft);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      si = fmaf(fi[k], fi[k], si);
      st = fmaf(ft[k], ft[k], st);
    }
  }
  float ni = sqrtf(warp_sum(si)), nt = sqrtf(warp_sum(st));
  if (bf16_mode) {
    ni = bf16_round(ni);
    nt = bf16_round(nt);
  }
  float dot = 0.f;
  for (int v = lane; v < (D >> 3); v += 32) {
    float fi[8], ft[8];
    unpack8(*reinterpret_cast<const bf16x8*>(ir + v * 8), fi);
    unpack8(*reinterpret_cast<const bf16x8*>(tr + v * 8), ft);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float a = __fdiv_rn(fi[k], ni), c = __fdiv_rn(ft[k], nt);
      if (bf16_mode) {
        a = bf16_round(a);
        c = bf16_round(c);
      }
      dot = fmaf(a, c, dot);
    }
  }
  dot = warp_sum(dot);
  if (lane == 0) {
    if (bf16_mode) out[b] = bf16_round(__fdiv_rn(bf16_round(bf16_round(logit_scale_exp) * bf16_round(dot)), 26.0f));
    else out[b] = logit_scale_exp * dot / 26.0f;
  }
}

}  // namespace
}  // namespace advgrpo

using namespace advgrpo;

extern "C" {

int advgrpo_gather_rows_l2norm(const void* feats, const int64_t* idx, void* out, int64_t B, int64_t T, int64_t n, int64_t D,
                               int l2norm, float eps, advgrpo_stream_t stream) {
  if (B == 0) return ADVGRPO_OK;
  ADVGRPO_CHECK_ARG(feats && out && (n == 0 || idx), "gather_rows_l2norm: null pointer");
  ADVGRPO_CHECK_ARG(B > 0 && T >= 1 && n >= 0 && (n == 0 || T >= 2) && D >= 8 && D % 8 == 0,
                    "gather_rows_l2norm: bad sizes B=%lld T=%lld n=%lld D=%lld (D must be a multiple of 8)", (long long)B,
                    (long long)T, (long long)n, (long long)D);
  ADVGRPO_CHECK_ARG(aligned16(feats) && aligned16(out), "gather_rows_l2norm: tensors must be 16-byte aligned");
  const int64_t rows = B * (1 + n);
  gather_rows_l2norm_kernel<<<(unsigned)((rows + kWarps - 1) / kWarps), kWarps * 32, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)feats, idx, (__nv_bfloat16*)out, B, T, n, (int)D, l2norm, eps);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_head_logits(const void* a, const void* w2, const void* b2, float* logits, int64_t R, int64_t Hd, int round_bf16,
                        advgrpo_stream_t stream) {
  if (R == 0) return ADVGRPO_OK;
  ADVGRPO_CHECK_ARG(a && w2 && logits && R > 0 && Hd >= 8 && Hd % 8 == 0, "head_logits: bad arguments (Hd must be a multiple of 8)");
  ADVGRPO_CHECK_ARG(aligned16(a) && aligned16(w2), "head_logits: tensors must be 16-byte aligned");
  head_logits_kernel<<<(unsigned)((R + kWarps - 1) / kWarps), kWarps * 32, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)a, (const __nv_bfloat16*)w2, (const __nv_bfloat16*)b2, logits, R, (int)Hd, round_bf16);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_dino_hybrid_score(const float* logits, float* hybrid, int64_t B, int64_t n, float cls_weight, int round_bf16,
                              advgrpo_stream_t stream) {
  if (B == 0) return ADVGRPO_OK;
  ADVGRPO_CHECK_ARG(logits && hybrid && B > 0 && n >= 0, "dino_hybrid_score: bad arguments");
  dino_hybrid_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(logits, hybrid, B, n, cls_weight, round_bf16);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_dino_hinge_loss(const float* logits, float* dlogits, float* out3, int64_t B_real, int64_t B_fake, int64_t n,
                            float patch_loss_weight, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(logits && dlogits && out3 && B_real > 0 && B_fake > 0 && n >= 0, "dino_hinge_loss: bad arguments");
  dino_hinge_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, dlogits, out3, B_real, B_fake, n, patch_loss_weight);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_head_dz(const float* dlogits, const void* w2, const void* z, void* dz, int64_t R, int64_t Hd,
                    advgrpo_stream_t stream) {
  if (R == 0) return ADVGRPO_OK;
  ADVGRPO_CHECK_ARG(dlogits && w2 && z && dz && R > 0 && Hd >= 8 && Hd % 8 == 0, "head_dz: bad arguments");
  ADVGRPO_CHECK_ARG(aligned16(w2) && aligned16(z) && aligned16(dz), "head_dz: tensors must be 16-byte aligned");
  head_dz_kernel<<<(unsigned)((R + kWarps - 1) / kWarps), kWarps * 32, 0, (cudaStream_t)stream>>>(
      dlogits, (const __nv_bfloat16*)w2, (const __nv_bfloat16*)z, (__nv_bfloat16*)dz, R, (int)Hd);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

size_t advgrpo_col_sum_workspace_bytes(int64_t rows, int64_t C) {
  return (size_t)(2 * col_sum_ny(rows)) * (size_t)(C > 0 ? C : 1) * sizeof(float) + 256;
}

int advgrpo_col_sum(const void* a, int64_t lda, const void* b, int64_t ldb, const float* row_scale, float* out, int64_t rows,
                    int64_t C, void* workspace, size_t workspace_bytes, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(out && C >= 0 && rows >= 0, "col_sum: bad arguments");
  if (C == 0) return ADVGRPO_OK;
  if (rows == 0) {
    ADVGRPO_CUDA_CALL(cudaMemsetAsync(out, 0, (size_t)C * sizeof(float), (cudaStream_t)stream));
    return ADVGRPO_OK;
  }
  ADVGRPO_CHECK_ARG(a && C % 8 == 0 && lda % 8 == 0 && (!b || ldb % 8 == 0) && aligned16(a) && (!b || aligned16(b)),
                    "col_sum: C and the leading dimensions must be multiples of 8, tensors 16-byte aligned");
  ADVGRPO_CHECK_ARG(!(b && row_scale), "col_sum: either a second matrix or a row scale, not both");
  ADVGRPO_CHECK_ARG(workspace && workspace_bytes >= advgrpo_col_sum_workspace_bytes(rows, C), "col_sum: workspace too small");
  const int ny = col_sum_ny(rows);
  const int64_t rpb = (rows + ny - 1) / ny;
  dim3 grid((unsigned)((C + 255) / 256), (unsigned)ny);
  float* part = (float*)workspace;
  if (b)
    col_sum_partial_kernel<1><<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, lda, (const __nv_bfloat16*)b,
                                                                             ldb, nullptr, nullptr, part, rows, C, rpb);
  else
    col_sum_partial_kernel<0><<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, lda, nullptr, 0, row_scale,
                                                                             nullptr, part, rows, C, rpb);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  col_sum_finish_kernel<<<(unsigned)((C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(part, out, nullptr, C, ny);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

size_t advgrpo_layer_norm_affine_bwd_workspace_bytes(int64_t rows, int64_t D) {
  return (size_t)(rows > 0 ? rows : 1) * sizeof(float2) + advgrpo_col_sum_workspace_bytes(rows, D) + 256;
}

int advgrpo_layer_norm_affine_bwd(const void* x, const void* weight, const void* dy, void* dx, float* dweight, float* dbias,
                                  int64_t rows, int64_t D, float eps, void* workspace, size_t workspace_bytes,
                                  advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(x && weight && dy && dx, "layer_norm_affine_bwd: null pointer");
  ADVGRPO_CHECK_ARG(D >= 8 && D % 8 == 0 && D <= 2048, "layer_norm_affine_bwd: D=%lld must be a multiple of 8 (<= 2048)", (long long)D);
  ADVGRPO_CHECK_ARG((dweight == nullptr) == (dbias == nullptr), "layer_norm_affine_bwd: dweight and dbias come together");
  ADVGRPO_CHECK_ARG(aligned16(x) && aligned16(weight) && aligned16(dy) && aligned16(dx), "layer_norm_affine_bwd: 16-byte alignment");
  if (rows <= 0) {
    if (dweight) {
      ADVGRPO_CUDA_CALL(cudaMemsetAsync(dweight, 0, (size_t)D * sizeof(float), (cudaStream_t)stream));
      ADVGRPO_CUDA_CALL(cudaMemsetAsync(dbias, 0, (size_t)D * sizeof(float), (cudaStream_t)stream));
    }
    return ADVGRPO_OK;
  }
  float2* stats = nullptr;
  float* part = nullptr;
  if (dweight) {
    ADVGRPO_CHECK_ARG(workspace && workspace_bytes >= advgrpo_layer_norm_affine_bwd_workspace_bytes(rows, D),
                      "layer_norm_affine_bwd: workspace too small");
    stats = (float2*)workspace;
    part = (float*)((uint8_t*)workspace + (((size_t)rows * sizeof(float2) + 255) & ~(size_t)255));
  }
  ln_affine_bwd_kernel<<<(unsigned)((rows + kWarps - 1) / kWarps), kWarps * 32, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)weight, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, stats, rows, (int)D, eps);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  if (dweight) {
    const int ny = col_sum_ny(rows);
    const int64_t rpb = (rows + ny - 1) / ny;
    dim3 grid((unsigned)((D + 255) / 256), (unsigned)ny);
    col_sum_partial_kernel<2><<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy, D, (const __nv_bfloat16*)x, D,
                                                                             nullptr, stats, part, rows, D, rpb);
    ADVGRPO_CUDA_LAUNCH_CHECK();
    col_sum_finish_kernel<<<(unsigned)((D + 255) / 256), 256, 0, (cudaStream_t)stream>>>(part, dweight, dbias, D, ny);
    ADVGRPO_CUDA_LAUNCH_CHECK();
  }
  return ADVGRPO_OK;
}

size_t advgrpo_ln_modulation_grads_workspace_bytes(int64_t nseg, int64_t rows_per_seg, int64_t D) {
  const int ny = col_sum_ny(rows_per_seg);
  return (((size_t)(nseg * rows_per_seg) * sizeof(float2) + 255) & ~(size_t)255) + (size_t)nseg * 2 * ny * (size_t)D * sizeof(float) + 256;
}

int advgrpo_ln_modulation_grads(const void* x, const void* dy, float* dshift, float* dscale, int64_t nseg, int64_t rows_per_seg,
                                int64_t D, float eps, void* workspace, size_t workspace_bytes, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(x && dy && dshift && dscale, "ln_modulation_grads: null pointer");
  ADVGRPO_CHECK_ARG(nseg >= 1 && nseg < 65536 && rows_per_seg >= 1 && D >= 8 && D % 8 == 0,
                    "ln_modulation_grads: bad sizes nseg=%lld rows_per_seg=%lld D=%lld", (long long)nseg, (long long)rows_per_seg,
                    (long long)D);
  ADVGRPO_CHECK_ARG(aligned16(x) && aligned16(dy), "ln_modulation_grads: tensors must be 16-byte aligned");
  ADVGRPO_CHECK_ARG(workspace && workspace_bytes >= advgrpo_ln_modulation_grads_workspace_bytes(nseg, rows_per_seg, D),
                    "ln_modulation_grads: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = nseg * rows_per_seg;
  float2* stats = (float2*)workspace;
  float* part = (float*)((uint8_t*)workspace + (((size_t)rows * sizeof(float2) + 255) & ~(size_t)255));
  row_stats_kernel<<<(unsigned)((rows + kWarps - 1) / kWarps), kWarps * 32, 0, st>>>((const __nv_bfloat16*)x, stats, rows, (int)D, eps);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  const int ny = col_sum_ny(rows_per_seg);
  const int64_t rpb = (rows_per_seg + ny - 1) / ny;
  dim3 grid((unsigned)((D + 255) / 256), (unsigned)ny, (unsigned)nseg);
  col_sum_partial_kernel<2><<<grid, kWarps * 32, 0, st>>>((const __nv_bfloat16*)dy, D, (const __nv_bfloat16*)x, D, nullptr, stats, part,
                                                         rows_per_seg, D, rpb);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  col_sum_finish_kernel<<<dim3((unsigned)((D + 255) / 256), (unsigned)nseg), 256, 0, st>>>(part, dscale, dshift, D, ny);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_adam_torch_order(void* param, void* grad, void* exp_avg, void* exp_avg_sq, int64_t n, int param_is_bf16,
                             int grad_is_bf16, double lr, double beta1, double beta2, double eps, int64_t step, int zero_grad,
                             advgrpo_stream_t stream) {
  if (n == 0) return ADVGRPO_OK;
  ADVGRPO_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adam_torch_order: bad arguments");
  AdamT a;
  // the Python floats of torch.optim.Adam are doubles that ATen narrows to fp32 per op: narrow the same quantities
  a.w = (float)(1.0 - beta1);
  a.beta2 = (float)beta2;
  a.beta2_c = (float)(1.0 - beta2);
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  a.bc2_sqrt = (float)sqrt(bc2);
  a.eps = (float)eps;
  a.neg_step_size = (float)(-(lr / bc1));
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (param_is_bf16 && grad_is_bf16)
    adam_torch_order_kernel<__nv_bfloat16, __nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(
        (__nv_bfloat16*)param, (const __nv_bfloat16*)grad, (__nv_bfloat16*)exp_avg, (__nv_bfloat16*)exp_avg_sq, n, a, zero_grad,
        (__nv_bfloat16*)grad);
  else if (param_is_bf16)
    adam_torch_order_kernel<__nv_bfloat16, float><<<(unsigned)blocks, 256, 0, st>>>(
        (__nv_bfloat16*)param, (const float*)grad, (__nv_bfloat16*)exp_avg, (__nv_bfloat16*)exp_avg_sq, n, a, zero_grad, (float*)grad);
  else if (!grad_is_bf16)
    adam_torch_order_kernel<float, float><<<(unsigned)blocks, 256, 0, st>>>((float*)param, (const float*)grad, (float*)exp_avg,
                                                                          (float*)exp_avg_sq, n, a, zero_grad, (float*)grad);
  else
    adam_torch_order_kernel<float, __nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(
        (float*)param, (const __nv_bfloat16*)grad, (float*)exp_avg, (float*)exp_avg_sq, n, a, zero_grad, (__nv_bfloat16*)grad);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_row_softmax_f32(const float* x, float* y, int64_t rows, int64_t cols, float scale, int round_tf32,
                            advgrpo_stream_t stream) {
  if (rows == 0) return ADVGRPO_OK;
  ADVGRPO_CHECK_ARG(x && y && rows > 0 && cols >= 4 && cols % 4 == 0 && rows < ((int64_t)1 << 31),
                    "row_softmax_f32: cols must be a multiple of 4 (rows=%lld cols=%lld)", (long long)rows, (long long)cols);
  ADVGRPO_CHECK_ARG(aligned16(x) && aligned16(y), "row_softmax_f32: tensors must be 16-byte aligned");
  row_softmax_f32_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(x, y, cols, scale * 1.4426950408889634f, round_tf32);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_pickscore_head(const void* img_feat, const void* txt_feat, const int64_t* txt_index, const void* logit_scale,
                           int logit_scale_is_bf16, float* scores, int64_t B, int64_t n_txt, int64_t D, int bf16_arithmetic,
                           advgrpo_stream_t stream) {
  if (B == 0) return ADVGRPO_OK;
  ADVGRPO_CHECK_ARG(img_feat && txt_feat && logit_scale && scores && B > 0 && n_txt > 0 && D >= 8 && D % 8 == 0,
                    "pickscore_head: bad arguments");
  ADVGRPO_CHECK_ARG(aligned16(img_feat) && aligned16(txt_feat), "pickscore_head: tensors must be 16-byte aligned");
  pickscore_head_kernel<<<(unsigned)((B + kWarps - 1) / kWarps), kWarps * 32, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)img_feat, (const __nv_bfloat16*)txt_feat, txt_index, logit_scale, logit_scale_is_bf16, scores, B, n_txt,
      (int)D, bf16_arithmetic);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

}  // extern "C"

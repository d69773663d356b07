// GroupNorm(32 groups, affine) + optional SiLU for NHWC fp32 activations: the normalisation between the
// cuDNN convolutions of the SD3 VAE decoder (fast.py:669 -> diffusers AutoencoderKL.decode, ResnetBlock2D /
// mid-block attention / conv_norm_out).  PyTorch's group_norm on channels_last tensors round-trips through
// NCHW copies and launches SiLU separately; this is two streaming passes (statistics, apply) at HBM speed.
//   stats : per-thread channel-quad partial sums -> shared-memory double atomics per group -> global
//   apply : y = (x - mean) * rstd * gamma[c] + beta[c]   (then x * sigmoid(x) if silu)
#include "common.cuh"

namespace advgrpo {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxGroups = 64;     // checked in the entry point

__global__ void __launch_bounds__(kThreads)
gn_stats_kernel(const float* __restrict__ x, const float* __restrict__ in_bias, double* __restrict__ sums, int64_t HW,
                int C, int groups, int pixels_per_block) {
  __shared__ double s_sum[64], s_sq[64];
  const int n = blockIdx.y;
  const int quads = C / 4;                       // float4 per pixel
  const int ppi = kThreads / quads;              // pixels per iteration
  const int q = threadIdx.x % quads, pl = threadIdx.x / quads;
  const int cpg = C / groups;
  const int g = (q * 4) / cpg;
  if (threadIdx.x < 64) { s_sum[threadIdx.x] = 0.0; s_sq[threadIdx.x] = 0.0; }
  __syncthreads();
  const int64_t p0 = (int64_t)blockIdx.x * pixels_per_block;
  int64_t p1 = p0 + pixels_per_block;
  if (p1 > HW) p1 = HW;
  const float* base = x + (int64_t)n * HW * C;
  float s = 0.f, ss = 0.f;
  const float4 ib = in_bias ? *reinterpret_cast<const float4*>(in_bias + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (pl < ppi) {
    for (int64_t p = p0 + pl; p < p1; p += ppi) {
      float4 v = *reinterpret_cast<const float4*>(base + p * C + q * 4);
      v.x += ib.x; v.y += ib.y; v.z += ib.z; v.w += ib.w;
      s += (v.x + v.y) + (v.z + v.w);
      ss = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, ss))));
    }
    atomicAdd(&s_sum[g], (double)s);
    atomicAdd(&s_sq[g], (double)ss);
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    atomicAdd(&sums[((int64_t)n * groups + threadIdx.x) * 2], s_sum[threadIdx.x]);
    atomicAdd(&sums[((int64_t)n * groups + threadIdx.x) * 2 + 1], s_sq[threadIdx.x]);
  }
}

// Round to the nearest TF32 value (10-bit mantissa, ties away): what cuDNN / cuBLAS do to fp32 operands of a TF32
// tensor-core product.  The tcgen05 TF32 convolution reads its operands straight from TMA-filled shared memory, where
// the tensor core would TRUNCATE the low 13 mantissa bits, so the producer of every convolution input rounds here.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__global__ void __launch_bounds__(kThreads)
gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ in_bias, const double* __restrict__ sums, const float* __restrict__ gamma,
                const float* __restrict__ beta, float* __restrict__ y, int64_t HW, int C, int groups, float eps,
                int silu) {
  const int n = blockIdx.y;
  const int quads = C / 4;
  const int64_t total = HW * quads;
  const int cpg = C / groups;
  const double cnt = (double)HW * cpg;
  // per-group mean / rstd once per CTA (double arithmetic on the f64 sums), then float lookups in the streaming loop:
  // recomputing them per float4 put an FP64 divide + sqrt on every element quad (FP64 runs at 1/64 rate on this part)
  __shared__ float s_mean[kMaxGroups], s_rstd[kMaxGroups];
  for (int g = threadIdx.x; g < groups; g += kThreads) {
    const double m = sums[((int64_t)n * groups + g) * 2] / cnt;
    double var = sums[((int64_t)n * groups + g) * 2 + 1] / cnt - m * m;
    if (var < 0) var = 0;
    s_mean[g] = (float)m;
    s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const float* xb = x + (int64_t)n * HW * C;
  float* yb = y + (int64_t)n * HW * C;
  // channel-quad index carried incrementally (no 64-bit modulo per element); group index by shift when cpg is 2^k
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int dq = (int)(stride % quads);
  const int cpg_shift = (cpg & (cpg - 1)) == 0 ? __ffs(cpg) - 1 : -1;
  int q = (int)(((int64_t)blockIdx.x * kThreads + threadIdx.x) % quads);
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride, q = (q + dq >= quads ? q + dq - quads : q + dq)) {
    const int c = q * 4;
    const int g = cpg_shift >= 0 ? (c >> cpg_shift) : c / cpg;
    const float mean = s_mean[g], rstd = s_rstd[g];
    float4 v = *reinterpret_cast<const float4*>(xb + i * 4);
    if (in_bias) {
      const float4 ib = *reinterpret_cast<const float4*>(in_bias + c);
      v.x += ib.x; v.y += ib.y; v.z += ib.z; v.w += ib.w;
    }
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c);
    const float4 be = *reinterpret_cast<const float4*>(beta + c);
    float o[4] = {(v.x - mean) * rstd * ga.x + be.x, (v.y - mean) * rstd * ga.y + be.y,
                  (v.z - mean) * rstd * ga.z + be.z, (v.w - mean) * rstd * ga.w + be.w};
    if (silu & 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = o[j] / (1.0f + __expf(-o[j]));
    }
    if (silu & 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = round_tf32(o[j]);
    }
    *reinterpret_cast<float4*>(yb + i * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// out = a + b + bias[c]   (residual add of a ResnetBlock2D with the conv bias folded in)
__global__ void __launch_bounds__(kThreads)
add_bias_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ bias,
                float* __restrict__ out, int64_t total4, int C) {
  const int quads = C / 4;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  const int dq = (int)(stride % quads);
  int q = (int)(((int64_t)blockIdx.x * kThreads + threadIdx.x) % quads);
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total4; i += stride, q = (q + dq >= quads ? q + dq - quads : q + dq)) {
    const float4 x = reinterpret_cast<const float4*>(a)[i];
    const float4 y = reinterpret_cast<const float4*>(b)[i];
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bb = *reinterpret_cast<const float4*>(bias + q * 4);
    reinterpret_cast<float4*>(out)[i] = make_float4(x.x + y.x + bb.x, x.y + y.y + bb.y, x.z + y.z + bb.z, x.w + y.w + bb.w);
  }
}

// nearest-neighbour 2x upsampling, NHWC fp32: one float4 read, four float4 writes per thread
__global__ void __launch_bounds__(kThreads)
upsample2x_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t B, int H, int W, int C) {
  const int quads = C / 4;
  const int64_t total = B * H * W * quads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int q = (int)(i % quads);
    int64_t pix = i / quads;
    const int xw = (int)(pix % W);
    pix /= W;
    const int yh = (int)(pix % H);
    const int64_t n = pix / H;
    float4 v = reinterpret_cast<const float4*>(x)[i];
    // the upsampled map is only ever the input of the upsampler's convolution: hand it over as TF32 values
    v = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
    float* base = y + (((n * 2 * H + 2 * yh) * 2 * W + 2 * xw) * (int64_t)C) + q * 4;
    const int64_t row = (int64_t)2 * W * C;
    *reinterpret_cast<float4*>(base) = v;
    *reinterpret_cast<float4*>(base + C) = v;
    *reinterpret_cast<float4*>(base + row) = v;
    *reinterpret_cast<float4*>(base + row + C) = v;
  }
}

}  // namespace
}  // namespace advgrpo

using namespace advgrpo;

extern "C" {

size_t advgrpo_group_norm_workspace_bytes(int64_t B, int64_t groups) { return (size_t)B * groups * 2 * sizeof(double); }

int advgrpo_group_norm_silu_nhwc(const float* x, const float* in_bias, const float* gamma, const float* beta, float* y, int64_t B,
                                 int64_t HW, int64_t C, int64_t groups, float eps, int silu, void* workspace,
                                 size_t workspace_bytes, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(x && gamma && beta && y, "group_norm_silu_nhwc: null pointer");
  ADVGRPO_CHECK_ARG(B >= 1 && HW >= 1 && groups >= 1 && groups <= 64 && C % groups == 0 && (C / groups) % 4 == 0 &&
                        C / 4 <= kThreads && kThreads % (C / 4) == 0,
                    "group_norm_silu_nhwc: unsupported C=%lld groups=%lld", (long long)C, (long long)groups);
  ADVGRPO_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), "group_norm_silu_nhwc: alignment");
  if (!workspace || workspace_bytes < advgrpo_group_norm_workspace_bytes(B, groups))
    return set_error(ADVGRPO_ERR_WORKSPACE, "group_norm_silu_nhwc: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  double* sums = (double*)workspace;
  ADVGRPO_CUDA_CALL(cudaMemsetAsync(sums, 0, advgrpo_group_norm_workspace_bytes(B, groups), st));
  const int ppi = kThreads / (int)(C / 4);
  int64_t blocks = (int64_t)sm_count() * 8 / B;
  if (blocks < 1) blocks = 1;
  int64_t ppb = (HW + blocks - 1) / blocks;
  ppb = (ppb + ppi - 1) / ppi * ppi;
  blocks = (HW + ppb - 1) / ppb;
  gn_stats_kernel<<<dim3((unsigned)blocks, (unsigned)B), kThreads, 0, st>>>(x, in_bias, sums, HW, (int)C, (int)groups, (int)ppb);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  int64_t ablocks = (HW * (C / 4) + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)sm_count() * 16 / B + 1;
  if (ablocks > cap) ablocks = cap;
  gn_apply_kernel<<<dim3((unsigned)ablocks, (unsigned)B), kThreads, 0, st>>>(x, in_bias, sums, gamma, beta, y, HW, (int)C,
                                                                            (int)groups, eps, silu);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_add_bias_nhwc(const float* a, const float* b, const float* bias, float* out, int64_t rows, int64_t C,
                          advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(a && b && out, "add_bias_nhwc: null pointer");
  ADVGRPO_CHECK_ARG(rows >= 1 && C >= 4 && C % 4 == 0, "add_bias_nhwc: C must be a multiple of 4");
  ADVGRPO_CHECK_ARG(aligned16(a) && aligned16(b) && aligned16(out) && (!bias || aligned16(bias)), "add_bias_nhwc: alignment");
  const int64_t total4 = rows * (C / 4);
  int64_t blocks = (total4 + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  add_bias_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(a, b, bias, out, total4, (int)C);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_upsample_nearest2x_nhwc(const float* x, float* y, int64_t B, int64_t H, int64_t W, int64_t C,
                                    advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(x && y, "upsample_nearest2x_nhwc: null pointer");
  ADVGRPO_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && C >= 4 && C % 4 == 0, "upsample_nearest2x_nhwc: bad shape");
  ADVGRPO_CHECK_ARG(aligned16(x) && aligned16(y), "upsample_nearest2x_nhwc: alignment");
  const int64_t total = B * H * W * (C / 4);
  int64_t blocks = (total + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  upsample2x_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(x, y, B, (int)H, (int)W, (int)C);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

}  // extern "C"

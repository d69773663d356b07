// Global-norm gradient clipping + AdamW on the flat fp32 LoRA master parameter (18.78 M values at
// SD3.5-medium r = 32), two HBM-bound passes:
//   1. grad_sumsq_kernel   : per-block partial sums of g^2 (float4 loads, f64 partials)
//   2. clip_adamw_kernel   : every block re-reduces the partials in one fixed order (deterministic, no
//                            atomics, no host read of the norm), then p, m, v are updated in one pass and
//                            the gradient is cleared (or left clipped, as torch does in place).
// Algorithmic traffic: pass 1 reads 4 B/param; pass 2 reads 16 B/param (p, g, m, v) and writes 16 B/param
// (p, m, v, g) => 36 B/param, 0.68 GB per optimizer step at 18.78 M parameters.
//
// Reference: scripts/train_sd3_fast_pickscore.py:1165-1171 (accelerator.clip_grad_norm_(params,
// max_grad_norm) -> optimizer.step() -> optimizer.zero_grad(), optimizer = torch.optim.AdamW :515-521).
// Arithmetic order follows torch: g <- g * min(max_norm / (norm + 1e-6), 1);  p <- p - lr*wd*p;
// m <- m + (1 - b1)(g - m);  v <- b2 v + (1 - b2) g^2;  p <- p - (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps).
#include "common.cuh"

namespace advgrpo {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ double block_sum_f64(double v, double* scratch) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  double t = (l < nw) ? scratch[l] : 0.0;
  return warp_sum(t);
}

__global__ void __launch_bounds__(kThreads) grad_sumsq_kernel(const float* __restrict__ g, int64_t n,
                                                              double* __restrict__ partials) {
  __shared__ double scratch[32];
  const int64_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
    const float4 v = g4[i];
    a0 = fmaf(v.x, v.x, a0);
    a1 = fmaf(v.y, v.y, a1);
    a2 = fmaf(v.z, v.z, a2);
    a3 = fmaf(v.w, v.w, a3);
  }
  double acc = ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {           // ragged tail (n not a multiple of 4)
    const float v = g[(n4 << 2) + threadIdx.x];
    acc += (double)v * (double)v;
  }
  acc = block_sum_f64(acc, scratch);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

struct AdamArgs {
  float* p;
  float* g;
  float* m;
  float* v;
  int64_t n;
  float lr_wd;        // lr * weight_decay
  float beta1_c;      // 1 - beta1
  float beta2;
  float beta2_c;      // 1 - beta2
  float step_size;    // lr / (1 - beta1^step)
  float bc2_sqrt;     // sqrt(1 - beta2^step)
  float eps;
  float max_norm;     // <= 0: no clipping
  int zero_grad;
  const double* partials;
  int n_partials;
  float* norm_out;    // may be null
};

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, float coef, const AdamArgs& a) {
  g = g * coef;
  p = p - a.lr_wd * p;
  m = m + a.beta1_c * (g - m);
  v = a.beta2 * v + a.beta2_c * g * g;
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p = p - a.step_size * m / denom;
}

__global__ void __launch_bounds__(kThreads) clip_adamw_kernel(const AdamArgs a) {
  __shared__ double scratch[32];
  float coef = 1.0f;
  if (a.max_norm > 0.f) {
    double s = 0.0;
    for (int i = threadIdx.x; i < a.n_partials; i += kThreads) s += a.partials[i];
    s = block_sum_f64(s, scratch);
    const float norm = (float)sqrt(s);
    coef = fminf(a.max_norm / (norm + 1e-6f), 1.0f);     // torch.nn.utils.clip_grad_norm_
    if (a.norm_out && blockIdx.x == 0 && threadIdx.x == 0) a.norm_out[0] = norm;
  }
  const int64_t n4 = a.n >> 2;
  float4* p4 = reinterpret_cast<float4*>(a.p);
  float4* g4 = reinterpret_cast<float4*>(a.g);
  float4* m4 = reinterpret_cast<float4*>(a.m);
  float4* v4 = reinterpret_cast<float4*>(a.v);
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
    float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
    adam_one(p.x, g.x, m.x, v.x, coef, a);
    adam_one(p.y, g.y, m.y, v.y, coef, a);
    adam_one(p.z, g.z, m.z, v.z, coef, a);
    adam_one(p.w, g.w, m.w, v.w, coef, a);
    p4[i] = p;
    m4[i] = m;
    v4[i] = v;
    g4[i] = a.zero_grad ? make_float4(0.f, 0.f, 0.f, 0.f) : g;
  }
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    float p = a.p[i], g = a.g[i], m = a.m[i], v = a.v[i];
    adam_one(p, g, m, v, coef, a);
    a.p[i] = p;
    a.m[i] = m;
    a.v[i] = v;
    a.g[i] = a.zero_grad ? 0.f : g;
  }
}

int grid_for(int64_t n) {
  const int64_t want = (n / 4 + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)sm_count() * 8;              // 8 resident 256-thread blocks per SM
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace
}  // namespace advgrpo

using namespace advgrpo;

extern "C" {

size_t advgrpo_clip_adamw_workspace_bytes(int64_t n) { return (size_t)grid_for(n) * sizeof(double); }

int advgrpo_clip_adamw(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                       double beta1, double beta2, double eps, double weight_decay, int64_t step,
                       double max_grad_norm, int zero_grad, float* grad_norm_out, void* workspace,
                       size_t workspace_bytes, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "clip_adamw: null pointer");
  ADVGRPO_CHECK_ARG(n >= 0 && step >= 1, "clip_adamw: n must be >= 0 and step >= 1 (got n=%lld step=%lld)",
                    (long long)n, (long long)step);
  ADVGRPO_CHECK_ARG(aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq),
                    "clip_adamw: buffers must be 16-byte aligned");
  ADVGRPO_CHECK_ARG(beta1 >= 0 && beta1 < 1 && beta2 >= 0 && beta2 < 1 && eps >= 0 && lr >= 0,
                    "clip_adamw: bad hyper-parameters");
  if (n == 0) return ADVGRPO_OK;
  const int grid = grid_for(n);
  cudaStream_t st = (cudaStream_t)stream;
  AdamArgs a;
  a.p = param; a.g = grad; a.m = exp_avg; a.v = exp_avg_sq; a.n = n;
  a.lr_wd = (float)(lr * weight_decay);
  a.beta1_c = (float)(1.0 - beta1);
  a.beta2 = (float)beta2;
  a.beta2_c = (float)(1.0 - beta2);
  a.step_size = (float)(lr / (1.0 - pow(beta1, (double)step)));
  a.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
  a.eps = (float)eps;
  a.max_norm = (float)max_grad_norm;
  a.zero_grad = zero_grad;
  a.partials = (const double*)workspace;
  a.n_partials = grid;
  a.norm_out = grad_norm_out;
  if (max_grad_norm > 0) {
    if (!workspace || workspace_bytes < (size_t)grid * sizeof(double))
      return set_error(ADVGRPO_ERR_WORKSPACE, "clip_adamw: workspace too small");
    grad_sumsq_kernel<<<grid, kThreads, 0, st>>>(grad, n, (double*)workspace);
    ADVGRPO_CUDA_LAUNCH_CHECK();
  }
  clip_adamw_kernel<<<grid, kThreads, 0, st>>>(a);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

}  // extern "C"

// Reward-model image preprocessing on the device (no device->host->PIL->device trip).
//
// clip_preprocess: reproduces, bit for bit on the 8-bit image, what the reference does at
//   adv_grpo/rewards.py:581-584  ((images*255).round().clamp(0,255).to(uint8) on a bf16 tensor)
//   + CLIPProcessor (adv_grpo/pickscore_scorer.py:21-28): Pillow antialiased BICUBIC resize
//   (ImagingResample: a = -0.5 kernel, support 2*scale, horizontal pass then vertical pass with an
//   8-bit intermediate, coefficients quantised to 22-bit fixed point), rescale 1/255, normalise.
// dino_preprocess: adv_grpo/rewards.py:379-391 (torch bicubic A = -0.75, align_corners = False,
//   ImageNet mean/std, bf16).
#include "common.cuh"

namespace advgrpo {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Pillow: PRECISION_BITS

__device__ __forceinline__ double pil_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(a + 2.0, x), (a + 3.0)), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), a);
  return 0.0;
}

__device__ __forceinline__ double pil_bilinear(double x) {      // Pillow bilinear_filter (support 1)
  if (x < 0.0) x = -x;
  return x < 1.0 ? __dsub_rn(1.0, x) : 0.0;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for one axis.  One thread per output index.
// filter: 0 = BICUBIC (support 2), 1 = BILINEAR (support 1).
__global__ void pil_coeffs_kernel(int in_size, int out_size, int ksize, int* __restrict__ bounds,
                                  int* __restrict__ kk, int filter = 0) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= out_size) return;
  const double scale = (double)in_size / (double)out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = (filter == 1 ? 1.0 : 2.0) * filterscale;
  const double center = __dmul_rn((double)xx + 0.5, scale);
  const double ss = 1.0 / filterscale;
  int xmin = (int)(center - support + 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)(center + support + 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double w[64];
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    const double arg = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
    w[x] = filter == 1 ? pil_bilinear(arg) : pil_bicubic(arg);
    ww = __dadd_rn(ww, w[x]);
  }
  for (int x = 0; x < ksize; ++x) {
    double v = 0.0;
    if (x < xmax) v = (ww != 0.0) ? w[x] / ww : w[x];
    const double s = __dmul_rn(v, (double)(1 << kPrecisionBits));
    kk[xx * ksize + x] = v < 0 ? (int)(-0.5 + s) : (int)(0.5 + s);
  }
  bounds[xx * 2] = xmin;
  bounds[xx * 2 + 1] = xmax;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__device__ __forceinline__ int quantise_bf16(__nv_bfloat16 v) {
  // (images * 255).round().clamp(0, 255).to(uint8) evaluated on a bf16 tensor (quirk Q6)
  float p = bf16_round(__bfloat162float(v) * 255.0f);
  p = rintf(p);
  p = fminf(fmaxf(p, 0.f), 255.f);
  return (int)p;
}

__device__ __forceinline__ int quantise_bf16(uint8_t v) { return (int)v; }   // already 8-bit (PIL input)

// horizontal pass: bf16 or u8 [P, H, W] -> u8 [P, H, out]      (P = B*3 planes)
template <typename InT>
__global__ void pil_horizontal_kernel(const InT* __restrict__ img, int H, int W, int out,
                                      int ksize, const int* __restrict__ bounds,
                                      const int* __restrict__ kk, uint8_t* __restrict__ tmp) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)gridDim.y * H * out;
  (void)total;
  const int plane = blockIdx.y;
  if (idx >= (int64_t)H * out) return;
  const int y = (int)(idx / out), xx = (int)(idx % out);
  const int xmin = bounds[xx * 2], xmax = bounds[xx * 2 + 1];
  const InT* row = img + ((int64_t)plane * H + y) * W;
  int ss0 = 1 << (kPrecisionBits - 1);
  for (int x = 0; x < xmax; ++x) ss0 += quantise_bf16(row[xmin + x]) * kk[xx * ksize + x];
  tmp[((int64_t)plane * H + y) * out + xx] = clip8(ss0);
}

// vertical pass + rescale + normalise: u8 [P, H, out] -> [P, out, out]
template <typename OutT>
__global__ void pil_vertical_kernel(const uint8_t* __restrict__ tmp, int H, int out, int ksize,
                                    const int* __restrict__ bounds, const int* __restrict__ kk,
                                    const float* __restrict__ mean3, const float* __restrict__ std3,
                                    OutT* __restrict__ pixels, uint8_t* __restrict__ u8_out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int plane = blockIdx.y;
  if (idx >= (int64_t)out * out) return;
  const int yy = (int)(idx / out), xx = (int)(idx % out);
  const int ymin = bounds[yy * 2], ymax = bounds[yy * 2 + 1];
  int ss0 = 1 << (kPrecisionBits - 1);
  for (int y = 0; y < ymax; ++y)
    ss0 += (int)tmp[((int64_t)plane * H + ymin + y) * out + xx] * kk[yy * ksize + y];
  const uint8_t v = clip8(ss0);
  const int64_t o = ((int64_t)plane * out + yy) * out + xx;
  if (u8_out) u8_out[o] = v;
  const int ch = plane % 3;
  const float f = (float)((double)v * (1.0 / 255.0));          // rescale (float64 product -> float32)
  const float n = (f - mean3[ch]) / std3[ch];                  // normalize in float32
  if constexpr (sizeof(OutT) == 2) pixels[o] = __float2bfloat16_rn(n);
  else pixels[o] = n;
}

// ---- reference-image resize (train_sd3_fast_pickscore.py:791-797: transforms.Resize((S, S)) on a PIL image + ToTensor) ----
// horizontal pass on the interleaved RGB bytes PIL decodes: u8 [H, W, 3] -> u8 [3, H, out_w]
__global__ void pil_horizontal_hwc_kernel(const uint8_t* __restrict__ img, int H, int W, int out_w, int ksize,
                                          const int* __restrict__ bounds, const int* __restrict__ kk,
                                          uint8_t* __restrict__ tmp) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)H * out_w) return;
  const int y = (int)(idx / out_w), xx = (int)(idx % out_w);
  const int xmin = bounds[xx * 2], xmax = bounds[xx * 2 + 1];
  const uint8_t* row = img + ((int64_t)y * W + xmin) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < xmax; ++x) {
    const int k = kk[xx * ksize + x];
    s0 += (int)row[3 * x] * k;
    s1 += (int)row[3 * x + 1] * k;
    s2 += (int)row[3 * x + 2] * k;
  }
  const int64_t o = (int64_t)y * out_w + xx, plane = (int64_t)H * out_w;
  tmp[o] = clip8(s0);
  tmp[plane + o] = clip8(s1);
  tmp[2 * plane + o] = clip8(s2);
}

// vertical pass + ToTensor: u8 [3, H, out_w] -> f32 [3, out_h, out_w] = v / 255 (and optionally the bytes themselves)
__global__ void pil_vertical_totensor_kernel(const uint8_t* __restrict__ tmp, int H, int out_h, int out_w, int ksize,
                                             const int* __restrict__ bounds, const int* __restrict__ kk,
                                             float* __restrict__ out, uint8_t* __restrict__ u8_out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int plane = blockIdx.y;
  if (idx >= (int64_t)out_h * out_w) return;
  const int yy = (int)(idx / out_w), xx = (int)(idx % out_w);
  const int ymin = bounds[yy * 2], ymax = bounds[yy * 2 + 1];
  int ss0 = 1 << (kPrecisionBits - 1);
  for (int y = 0; y < ymax; ++y) ss0 += (int)tmp[((int64_t)plane * H + ymin + y) * out_w + xx] * kk[yy * ksize + y];
  const uint8_t v = clip8(ss0);
  const int64_t o = ((int64_t)plane * out_h + yy) * out_w + xx;
  if (u8_out) u8_out[o] = v;
  out[o] = __fdiv_rn((float)v, 255.0f);                        // ToTensor: byte_tensor.float().div(255)
}

__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  float x = t + 1.0f;
  w[0] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
  x = t;
  w[1] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
  x = 1.0f - t;
  w[2] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
  x = 2.0f - t;
  w[3] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
}

template <typename InT>
__global__ void dino_preprocess_kernel(const InT* __restrict__ img, int H, int W, int out,
                                       const float* __restrict__ mean3, const float* __restrict__ std3,
                                       __nv_bfloat16* __restrict__ pixels) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int plane = blockIdx.y;
  if (idx >= (int64_t)out * out) return;
  const int oy = (int)(idx / out), ox = (int)(idx % out);
  const float sy = (float)H / (float)out, sx = (float)W / (float)out;
  const float fy = sy * ((float)oy + 0.5f) - 0.5f;
  const float fx = sx * ((float)ox + 0.5f) - 0.5f;
  const int iy = (int)floorf(fy), ix = (int)floorf(fx);
  float wy[4], wx[4];
  cubic_coeffs(fy - (float)iy, wy);
  cubic_coeffs(fx - (float)ix, wx);
  const InT* base = img + (int64_t)plane * H * W;
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int y = iy - 1 + a;
    y = y < 0 ? 0 : (y > H - 1 ? H - 1 : y);
    float r = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int x = ix - 1 + b;
      x = x < 0 ? 0 : (x > W - 1 ? W - 1 : x);
      float v;
      if constexpr (sizeof(InT) == 2) v = __bfloat162float(base[(int64_t)y * W + x]);
      else v = base[(int64_t)y * W + x];
      r += v * wx[b];
    }
    acc += r * wy[a];
  }
  if constexpr (sizeof(InT) == 2) acc = bf16_round(acc);   // F.interpolate returns the input dtype
  const int ch = plane % 3;
  pixels[((int64_t)plane * out + oy) * out + ox] = __float2bfloat16_rn((acc - mean3[ch]) / std3[ch]);
}

int pil_ksize(int in_size, int out_size, double filter_support = 2.0) {
  double scale = (double)in_size / (double)out_size;
  double fs = scale < 1.0 ? 1.0 : scale;
  return (int)ceil(filter_support * fs) * 2 + 1;
}

}  // namespace
}  // namespace advgrpo

using namespace advgrpo;

extern "C" {

size_t advgrpo_clip_preprocess_workspace_bytes(int64_t B, int64_t H, int64_t W, int64_t out) {
  (void)W;
  const int ks = pil_ksize((int)H, (int)out);
  size_t coeff = (size_t)out * (ks + 2) * sizeof(int);
  coeff = (coeff + 255) & ~(size_t)255;
  return coeff + (size_t)B * 3 * H * out + 256;
}

int advgrpo_clip_preprocess(const void* images, int images_u8, int64_t B, int64_t H, int64_t W, int64_t out,
                            const float* mean3, const float* std3, void* pixels, int pixels_f32,
                            uint8_t* u8_out, void* workspace, size_t workspace_bytes,
                            advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(images && mean3 && std3 && pixels, "clip_preprocess: null pointer");
  ADVGRPO_CHECK_ARG(B >= 1 && H == W && H >= 1 && out >= 1, "clip_preprocess: square images required (H=%lld W=%lld out=%lld)",
                    (long long)H, (long long)W, (long long)out);
  const int ks = pil_ksize((int)H, (int)out);
  ADVGRPO_CHECK_ARG(ks <= 64, "clip_preprocess: downscale factor too large");
  if (!workspace || workspace_bytes < advgrpo_clip_preprocess_workspace_bytes(B, H, W, out))
    return set_error(ADVGRPO_ERR_WORKSPACE, "clip_preprocess: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int* bounds = (int*)workspace;
  int* kk = bounds + 2 * out;
  size_t coeff = (size_t)out * (ks + 2) * sizeof(int);
  coeff = (coeff + 255) & ~(size_t)255;
  uint8_t* tmp = (uint8_t*)workspace + coeff;
  pil_coeffs_kernel<<<(unsigned)((out + 127) / 128), 128, 0, st>>>((int)H, (int)out, ks, bounds, kk);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  const unsigned planes = (unsigned)(B * 3);
  if (images_u8)
    pil_horizontal_kernel<uint8_t><<<dim3((unsigned)((H * out + 255) / 256), planes), 256, 0, st>>>(
        (const uint8_t*)images, (int)H, (int)W, (int)out, ks, bounds, kk, tmp);
  else
    pil_horizontal_kernel<__nv_bfloat16><<<dim3((unsigned)((H * out + 255) / 256), planes), 256, 0, st>>>(
        (const __nv_bfloat16*)images, (int)H, (int)W, (int)out, ks, bounds, kk, tmp);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  dim3 g2((unsigned)((out * out + 255) / 256), planes);
  if (pixels_f32)
    pil_vertical_kernel<float><<<g2, 256, 0, st>>>(tmp, (int)H, (int)out, ks, bounds, kk, mean3, std3, (float*)pixels, u8_out);
  else
    pil_vertical_kernel<__nv_bfloat16><<<g2, 256, 0, st>>>(tmp, (int)H, (int)out, ks, bounds, kk, mean3, std3, (__nv_bfloat16*)pixels, u8_out);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

static size_t pil_coeff_bytes(int out, int ks) { return (((size_t)out * (ks + 2) * sizeof(int)) + 255) & ~(size_t)255; }

size_t advgrpo_pil_resize_bilinear_workspace_bytes(int64_t H, int64_t W, int64_t out_h, int64_t out_w) {
  const int ksx = pil_ksize((int)W, (int)out_w, 1.0), ksy = pil_ksize((int)H, (int)out_h, 1.0);
  return pil_coeff_bytes((int)out_w, ksx) + pil_coeff_bytes((int)out_h, ksy) + (size_t)3 * H * out_w + 256;
}

int advgrpo_pil_resize_bilinear_u8(const uint8_t* img_hwc, int64_t H, int64_t W, int64_t out_h, int64_t out_w, float* out_chw,
                                   uint8_t* u8_out_chw, void* workspace, size_t workspace_bytes, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(img_hwc && out_chw, "pil_resize_bilinear_u8: null pointer");
  ADVGRPO_CHECK_ARG(H >= 1 && W >= 1 && out_h >= 1 && out_w >= 1 && H < (1 << 15) && W < (1 << 15) && out_h < (1 << 15) &&
                        out_w < (1 << 15),
                    "pil_resize_bilinear_u8: bad sizes H=%lld W=%lld -> %lld x %lld", (long long)H, (long long)W, (long long)out_h,
                    (long long)out_w);
  const int ksx = pil_ksize((int)W, (int)out_w, 1.0), ksy = pil_ksize((int)H, (int)out_h, 1.0);
  ADVGRPO_CHECK_ARG(ksx <= 64 && ksy <= 64, "pil_resize_bilinear_u8: downscale factor above 31 is not supported");
  if (!workspace || workspace_bytes < advgrpo_pil_resize_bilinear_workspace_bytes(H, W, out_h, out_w))
    return set_error(ADVGRPO_ERR_WORKSPACE, "pil_resize_bilinear_u8: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int* bx = (int*)workspace;
  int* kx = bx + 2 * out_w;
  int* by = (int*)((uint8_t*)workspace + pil_coeff_bytes((int)out_w, ksx));
  int* ky = by + 2 * out_h;
  uint8_t* tmp = (uint8_t*)workspace + pil_coeff_bytes((int)out_w, ksx) + pil_coeff_bytes((int)out_h, ksy);
  pil_coeffs_kernel<<<(unsigned)((out_w + 127) / 128), 128, 0, st>>>((int)W, (int)out_w, ksx, bx, kx, 1);
  pil_coeffs_kernel<<<(unsigned)((out_h + 127) / 128), 128, 0, st>>>((int)H, (int)out_h, ksy, by, ky, 1);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  pil_horizontal_hwc_kernel<<<(unsigned)((H * out_w + 255) / 256), 256, 0, st>>>(img_hwc, (int)H, (int)W, (int)out_w, ksx, bx, kx, tmp);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  pil_vertical_totensor_kernel<<<dim3((unsigned)((out_h * out_w + 255) / 256), 3), 256, 0, st>>>(tmp, (int)H, (int)out_h, (int)out_w, ksy,
                                                                                                by, ky, out_chw, u8_out_chw);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_dino_preprocess(const void* images, int images_f32, int64_t B, int64_t H, int64_t W,
                            int64_t out, const float* mean3, const float* std3, void* pixels,
                            advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(images && mean3 && std3 && pixels, "dino_preprocess: null pointer");
  ADVGRPO_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && out >= 1, "dino_preprocess: bad sizes");
  dim3 g((unsigned)((out * out + 255) / 256), (unsigned)(B * 3));
  if (images_f32)
    dino_preprocess_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)images, (int)H, (int)W, (int)out, mean3, std3, (__nv_bfloat16*)pixels);
  else
    dino_preprocess_kernel<__nv_bfloat16><<<g, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)images, (int)H, (int)W, (int)out, mean3, std3, (__nv_bfloat16*)pixels);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

}  // extern "C"

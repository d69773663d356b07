// Differentiable attention for SHORT sequences with head sizes the tcgen05 kernels do not take in the backward direction
// (CLIP-ViT-H/14: 257 tokens, head_dim 80): forward + backward in fp32 on the CUDA cores, K / V (or Q / dO) of one
// (sample, head) resident in shared memory, warp-shuffle reductions, bf16 in / out.
//
// Used ONLY by the one or two trainable vision blocks of the PickScore discriminator step
// (scripts/train_sd3_fast_pickscore.py:1016-1029, tune_layer = -1 / -2; adv_grpo/pick_score_training.py:94-106): 27 GFLOP
// per step against the ~12 TFLOP of the 31 frozen blocks in front of them, which run on the tcgen05 attention kernel.
// The D = 64 tcgen05 backward (attn_bwd.cu) keeps S^T, dP^T, P^T, dV, dK and dQ in the 512 TMEM columns; a 128-wide
// (80 padded) head does not fit that layout, and at this size the op is latency-, not throughput-bound.
//
// Layout: q, k, v, o, do, dq, dk, dv are bf16 [B, S, H, Dh] (the [B, S, H * Dh] outputs of the q / k / v projections
// viewed per head: no transposes around the op).  lse, delta: f32 [B, H, S].
//   forward : one warp per query row: lanes over keys for s = scale q.k, softmax, then lanes over d for o = p V.
//   backward: pass A, one warp per query row (K, V in smem):  p = exp(scale q.k - lse), dp = do.v,
//                      ds = p (dp - delta) scale, dq = ds K;   delta = do.o is written here.
//             pass B, one warp per key row (Q, dO, lse, delta in smem): the same p / ds down a column,
//                      dv = p^T dO, dk = ds^T Q.
#include <math.h>

#include "common.cuh"

namespace advgrpo {
namespace {

constexpr int kWarpsS = 8;
constexpr int kMaxD = 128;

__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

// Stage rows [0, S) of one (b, h) slice of a [B, S, H, Dh] tensor into shared memory as bf16 with a padded row stride of
// Dh + 2 elements (an odd number of 32-bit words: rows read by consecutive lanes fall into different banks).
__device__ __forceinline__ void stage_rows(const __nv_bfloat16* __restrict__ src, uint32_t* __restrict__ dst, int S, int H,
                                           int Dh, int b, int h) {
  const int w = Dh >> 1, ws = w + 1;
  for (int idx = threadIdx.x; idx < S * w; idx += blockDim.x) {
    const int r = idx / w, c = idx - r * w;
    dst[r * ws + c] = reinterpret_cast<const uint32_t*>(src + (((int64_t)b * S + r) * H + h) * Dh)[c];
  }
}

// dot of a staged bf16 row with an fp32 vector in shared memory (both Dh long)
__device__ __forceinline__ float dot_row(const uint32_t* __restrict__ row, const float* __restrict__ vec, int w) {
  float acc = 0.f;
  for (int c = 0; c < w; ++c) {
    const float2 kk = bf2_to_f2(row[c]);
    acc = fmaf(kk.x, vec[2 * c], acc);
    acc = fmaf(kk.y, vec[2 * c + 1], acc);
  }
  return acc;
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(kWarpsS * 32)
attn_small_fwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                      const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ o, float* __restrict__ lse, int S, int H,
                      int Dh, float scale, int rows_per_block, int causal) {
  extern __shared__ uint32_t sm_u[];
  const int w = Dh >> 1, ws = w + 1;
  uint32_t* sK = sm_u;
  uint32_t* sV = sK + S * ws;
  float* sQ = reinterpret_cast<float*>(sV + S * ws);          // kWarpsS x kMaxD
  float* sP = sQ + kWarpsS * kMaxD;                            // kWarpsS x S_pad
  const int S_pad = (S + 31) & ~31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows(k, sK, S, H, Dh, b, h);
  stage_rows(v, sV, S, H, Dh, b, h);
  __syncthreads();
  float* myQ = sQ + warp * kMaxD;
  float* myP = sP + warp * S_pad;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(S, r0 + rows_per_block);
  for (int i = r0 + warp; i < r1; i += kWarpsS) {
    const int64_t row_off = (((int64_t)b * S + i) * H + h) * Dh;
    for (int d = lane; d < Dh; d += 32) myQ[d] = __bfloat162float(q[row_off + d]);
    __syncwarp();
    const int jn = causal ? i + 1 : S;                            // keys this query sees
    float mx = -INFINITY;
    for (int j = lane; j < jn; j += 32) {
      const float s = dot_row(sK + j * ws, myQ, w) * scale;
      myP[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float l = 0.f;
    for (int j = lane; j < jn; j += 32) {
      const float p = __expf(myP[j] - mx);
      myP[j] = p;
      l += p;
    }
    l = warp_sum(l);
    __syncwarp();
    const float inv = 1.0f / l;
    for (int d2 = lane; d2 < w; d2 += 32) {                      // two output columns per lane step
      float a0 = 0.f, a1 = 0.f;
      for (int j = 0; j < jn; ++j) {
        const float2 vv = bf2_to_f2(sV[j * ws + d2]);
        const float p = myP[j];
        a0 = fmaf(p, vv.x, a0);
        a1 = fmaf(p, vv.y, a1);
      }
      *reinterpret_cast<__nv_bfloat162*>(o + row_off + 2 * d2) = __floats2bfloat162_rn(a0 * inv, a1 * inv);
    }
    if (lane == 0 && lse) lse[((int64_t)b * H + h) * S + i] = mx + __logf(l);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ backward, pass A
__global__ void __launch_bounds__(kWarpsS * 32)
attn_small_bwd_dq_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                         const __nv_bfloat16* __restrict__ v, const __nv_bfloat16* __restrict__ o,
                         const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse, float* __restrict__ delta,
                         __nv_bfloat16* __restrict__ dq, int S, int H, int Dh, float scale, int rows_per_block, int causal) {
  extern __shared__ uint32_t sm_u[];
  const int w = Dh >> 1, ws = w + 1;
  uint32_t* sK = sm_u;
  uint32_t* sV = sK + S * ws;
  float* sQ = reinterpret_cast<float*>(sV + S * ws);          // kWarpsS x kMaxD   (q_i)
  float* sD = sQ + kWarpsS * kMaxD;                            // kWarpsS x kMaxD   (do_i)
  float* sP = sD + kWarpsS * kMaxD;                            // kWarpsS x S_pad   (ds_i.)
  const int S_pad = (S + 31) & ~31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows(k, sK, S, H, Dh, b, h);
  stage_rows(v, sV, S, H, Dh, b, h);
  __syncthreads();
  float* myQ = sQ + warp * kMaxD;
  float* myD = sD + warp * kMaxD;
  float* myP = sP + warp * S_pad;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(S, r0 + rows_per_block);
  for (int i = r0 + warp; i < r1; i += kWarpsS) {
    const int64_t row_off = (((int64_t)b * S + i) * H + h) * Dh;
    float dl = 0.f;
    for (int d = lane; d < Dh; d += 32) {
      const float dd = __bfloat162float(dout[row_off + d]);
      myQ[d] = __bfloat162float(q[row_off + d]);
      myD[d] = dd;
      dl = fmaf(dd, __bfloat162float(o[row_off + d]), dl);
    }
    dl = warp_sum(dl);
    const int64_t stat = ((int64_t)b * H + h) * S + i;
    if (lane == 0) delta[stat] = dl;
    const float li = lse[stat];
    const int jn = causal ? i + 1 : S;
    __syncwarp();
    for (int j = lane; j < jn; j += 32) {
      const float p = __expf(dot_row(sK + j * ws, myQ, w) * scale - li);
      const float dp = dot_row(sV + j * ws, myD, w);
      myP[j] = p * (dp - dl) * scale;
    }
    __syncwarp();
    for (int d2 = lane; d2 < w; d2 += 32) {
      float a0 = 0.f, a1 = 0.f;
      for (int j = 0; j < jn; ++j) {
        const float2 kk = bf2_to_f2(sK[j * ws + d2]);
        const float ds = myP[j];
        a0 = fmaf(ds, kk.x, a0);
        a1 = fmaf(ds, kk.y, a1);
      }
      *reinterpret_cast<__nv_bfloat162*>(dq + row_off + 2 * d2) = __floats2bfloat162_rn(a0, a1);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ backward, pass B
__global__ void __launch_bounds__(kWarpsS * 32)
attn_small_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                          const __nv_bfloat16* __restrict__ v, const __nv_bfloat16* __restrict__ dout,
                          const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dk,
                          __nv_bfloat16* __restrict__ dv, int S, int H, int Dh, float scale, int rows_per_block, int causal) {
  extern __shared__ uint32_t sm_u[];
  const int w = Dh >> 1, ws = w + 1;
  const int S_pad = (S + 31) & ~31;
  uint32_t* sQ = sm_u;
  uint32_t* sDO = sQ + S * ws;
  float* sL = reinterpret_cast<float*>(sDO + S * ws);         // S_pad  lse
  float* sDl = sL + S_pad;                                     // S_pad  delta
  float* sKj = sDl + S_pad;                                    // kWarpsS x kMaxD   (k_j)
  float* sVj = sKj + kWarpsS * kMaxD;                          // kWarpsS x kMaxD   (v_j)
  float* sP = sVj + kWarpsS * kMaxD;                           // kWarpsS x S_pad   (p_.j)
  float* sS = sP + kWarpsS * S_pad;                            // kWarpsS x S_pad   (ds_.j)
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_rows(q, sQ, S, H, Dh, b, h);
  stage_rows(dout, sDO, S, H, Dh, b, h);
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    sL[i] = lse[((int64_t)b * H + h) * S + i];
    sDl[i] = delta[((int64_t)b * H + h) * S + i];
  }
  __syncthreads();
  float* myK = sKj + warp * kMaxD;
  float* myV = sVj + warp * kMaxD;
  float* myP = sP + warp * S_pad;
  float* myS = sS + warp * S_pad;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(S, r0 + rows_per_block);
  for (int j = r0 + warp; j < r1; j += kWarpsS) {
    const int64_t row_off = (((int64_t)b * S + j) * H + h) * Dh;
    for (int d = lane; d < Dh; d += 32) {
      myK[d] = __bfloat162float(k[row_off + d]);
      myV[d] = __bfloat162float(v[row_off + d]);
    }
    __syncwarp();
    const int i0 = causal ? j : 0;                                // queries that see this key
    for (int i = i0 + lane; i < S; i += 32) {
      const float p = __expf(dot_row(sQ + i * ws, myK, w) * scale - sL[i]);
      const float dp = dot_row(sDO + i * ws, myV, w);
      myP[i] = p;
      myS[i] = p * (dp - sDl[i]) * scale;
    }
    __syncwarp();
    for (int d2 = lane; d2 < w; d2 += 32) {
      float v0 = 0.f, v1 = 0.f, k0 = 0.f, k1 = 0.f;
      for (int i = i0; i < S; ++i) {
        const float2 dd = bf2_to_f2(sDO[i * ws + d2]);
        const float2 qq = bf2_to_f2(sQ[i * ws + d2]);
        const float p = myP[i], ds = myS[i];
        v0 = fmaf(p, dd.x, v0);
        v1 = fmaf(p, dd.y, v1);
        k0 = fmaf(ds, qq.x, k0);
        k1 = fmaf(ds, qq.y, k1);
      }
      *reinterpret_cast<__nv_bfloat162*>(dv + row_off + 2 * d2) = __floats2bfloat162_rn(v0, v1);
      *reinterpret_cast<__nv_bfloat162*>(dk + row_off + 2 * d2) = __floats2bfloat162_rn(k0, k1);
    }
    __syncwarp();
  }
}

size_t smem_fwd(int S, int Dh) {
  const int S_pad = (S + 31) & ~31;
  return (size_t)2 * S * (Dh / 2 + 1) * 4 + (size_t)kWarpsS * kMaxD * 4 + (size_t)kWarpsS * S_pad * 4;
}
size_t smem_dq(int S, int Dh) { return smem_fwd(S, Dh) + (size_t)kWarpsS * kMaxD * 4; }
size_t smem_dkv(int S, int Dh) {
  const int S_pad = (S + 31) & ~31;
  return (size_t)2 * S * (Dh / 2 + 1) * 4 + (size_t)2 * S_pad * 4 + (size_t)2 * kWarpsS * kMaxD * 4 + (size_t)2 * kWarpsS * S_pad * 4;
}
constexpr size_t kSmemCap = 227 * 1024;

int check_common(const char* who, int64_t B, int64_t S, int64_t H, int64_t Dh, size_t smem) {
  ADVGRPO_CHECK_ARG(B >= 1 && S >= 1 && H >= 1 && H < 65536 && B < 65536, "%s: bad sizes B=%lld S=%lld H=%lld", who, (long long)B,
                    (long long)S, (long long)H);
  ADVGRPO_CHECK_ARG(Dh >= 2 && Dh % 2 == 0 && Dh <= kMaxD, "%s: head_dim %lld must be even and <= %d", who, (long long)Dh, kMaxD);
  if (smem > kSmemCap)
    return set_error(ADVGRPO_ERR_UNSUPPORTED, "%s: S=%lld x head_dim=%lld does not fit the shared-memory-resident short-sequence "
                     "kernel (%zu > %zu bytes); use advgrpo_attn_fwd / advgrpo_attn_bwd (head_dim 64)", who, (long long)S,
                     (long long)Dh, smem, kSmemCap);
  return ADVGRPO_OK;
}

// rows per block so that the grid has ~2 waves of CTAs without re-staging K / V more often than needed
int rows_per_block(int64_t S, int64_t BH) {
  const int64_t want_blocks = (int64_t)sm_count() * 2;
  int64_t per_bh = (want_blocks + BH - 1) / BH;
  if (per_bh < 1) per_bh = 1;
  int64_t rpb = (S + per_bh - 1) / per_bh;
  rpb = (rpb + kWarpsS - 1) / kWarpsS * kWarpsS;
  if (rpb < kWarpsS) rpb = kWarpsS;
  return (int)rpb;
}

}  // namespace
}  // namespace advgrpo

using namespace advgrpo;

extern "C" {

int advgrpo_attn_small_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int64_t B, int64_t S, int64_t H,
                           int64_t Dh, float scale, int causal, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(q && k && v && o, "attn_small_fwd: null pointer");
  const size_t smem = smem_fwd((int)S, (int)Dh);
  int rc = check_common("attn_small_fwd", B, S, H, Dh, smem);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    ADVGRPO_CUDA_CALL(cudaFuncSetAttribute(attn_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap));
    attr_set = true;
  }
  const int rpb = rows_per_block(S, B * H);
  dim3 grid((unsigned)((S + rpb - 1) / rpb), (unsigned)H, (unsigned)B);
  attn_small_fwd_kernel<<<grid, kWarpsS * 32, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (__nv_bfloat16*)o, lse, (int)S, (int)H, (int)Dh,
      scale, rpb, causal);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int advgrpo_attn_small_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                           float* delta, void* dq, void* dk, void* dv, int64_t B, int64_t S, int64_t H, int64_t Dh, float scale,
                           int causal, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(q && k && v && o && dout && lse && delta && dq && dk && dv, "attn_small_bwd: null pointer");
  const size_t s_a = smem_dq((int)S, (int)Dh), s_b = smem_dkv((int)S, (int)Dh);
  int rc = check_common("attn_small_bwd", B, S, H, Dh, s_a > s_b ? s_a : s_b);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    ADVGRPO_CUDA_CALL(cudaFuncSetAttribute(attn_small_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap));
    ADVGRPO_CUDA_CALL(cudaFuncSetAttribute(attn_small_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCap));
    attr_set = true;
  }
  const int rpb = rows_per_block(S, B * H);
  dim3 grid((unsigned)((S + rpb - 1) / rpb), (unsigned)H, (unsigned)B);
  attn_small_bwd_dq_kernel<<<grid, kWarpsS * 32, s_a, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (const __nv_bfloat16*)o, (const __nv_bfloat16*)dout,
      lse, delta, (__nv_bfloat16*)dq, (int)S, (int)H, (int)Dh, scale, rpb, causal);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  attn_small_bwd_dkv_kernel<<<grid, kWarpsS * 32, s_b, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (const __nv_bfloat16*)dout, lse, delta,
      (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, (int)S, (int)H, (int)Dh, scale, rpb, causal);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

}  // extern "C"

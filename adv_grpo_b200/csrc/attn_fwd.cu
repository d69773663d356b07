// Flash-attention forward on tcgen05 + TMA for sm_100a.
//
// One CTA owns NQ query tiles of 128 rows of one (batch, head) and streams the K/V
// sequence in 128-row tiles:
//   warp 4*NQ     : TMA producer   (Q once, then K_j / V_j into a STAGES-deep ring)
//   warp 4*NQ + 1 : MMA issuer     (S = Q K_j^T into TMEM;  O += P_j V_j with P read from TMEM)
//   warps 0..4NQ-1: softmax        (thread == query row: tcgen05.ld S, online softmax with
//                                   lazy rescale, bf16 P written back to TMEM, final O / l)
// TMEM per query tile: S 128 cols (fp32) | P 64 cols (packed bf16) | O D cols (fp32).
// Q/K/V are read straight out of the token-major joint buffer [B, S, 3, H, D] through one
// 4-D tensor map (128B swizzle), O is written token-major [B, S, H, D]: no transposes.
// S(j+1) = Q K_{j+1}^T is issued before P_j V_j so the softmax of tile j+1 overlaps the
// PV MMA of tile j; with NQ = 1 two CTAs share an SM (256 TMEM columns each), with
// NQ = 2 the two query tiles ping-pong inside one CTA.
//
// Replaces F.scaled_dot_product_attention in diffusers JointAttnProcessor2_0 / CLIPAttention /
// timm Attention (see include/advgrpo_b200.h for the reference call sites).
#include <math.h>

#include "common.cuh"
#include "sm100.cuh"

namespace advgrpo {
namespace {

using namespace sm100;

constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int STAGES = 2;
constexpr float kRescaleThreshold = 8.0f;  // log2 domain: skip O rescale while max grows < 2^8

template <int D, int NQ>
struct Cfg {
  static constexpr int kHalves = D / 64;                 // 64-column (128 B) swizzle atoms per row
  static constexpr int kTileBytes = BQ * D * 2;          // one Q/K/V tile
  static constexpr int kSmemTiles = NQ * kTileBytes + STAGES * 2 * kTileBytes;
  static constexpr int kSmemBytes = kSmemTiles + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kColsPerQ = 128 + 64 + D;
  static constexpr int kTmemCols = (NQ * kColsPerQ <= 256) ? 256 : 512;
  static constexpr int kSoftmaxWarps = 4 * NQ;
  static constexpr int kThreads = (kSoftmaxWarps + 2) * 32;
  static_assert(NQ * kColsPerQ <= 512, "TMEM budget");
};

struct Params {
  __nv_bfloat16* out;   // [B, S, H, D]   (rows [0, S_split) when out2 is set: [B, S_split, H, D])
  __nv_bfloat16* out2;  // rows [S_split, S): [B, S - S_split, H, D], or null
  float* lse;           // [B, H, S] or null
  int S, H, S_split;
  float scale_log2;     // softmax scale * log2(e)
  int causal;
};

template <int D, int NQ>
__global__ void __launch_bounds__(Cfg<D, NQ>::kThreads, (NQ == 1 && D == 64) ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  using C = Cfg<D, NQ>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;                                   // NQ tiles
  uint8_t* k_smem = smem + NQ * C::kTileBytes;              // STAGES tiles
  uint8_t* v_smem = k_smem + STAGES * C::kTileBytes;        // STAGES tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kSmemTiles);
  uint64_t* bar_q_full = bars;                  // 1
  uint64_t* bar_k_full = bars + 1;              // STAGES
  uint64_t* bar_v_full = bar_k_full + STAGES;   // STAGES
  uint64_t* bar_kv_empty = bar_v_full + STAGES; // STAGES
  uint64_t* bar_s_full = bar_kv_empty + STAGES; // NQ
  uint64_t* bar_p_full = bar_s_full + NQ;       // NQ
  uint64_t* bar_pv_done = bar_p_full + NQ;      // NQ
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bar_pv_done + NQ);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (BQ * NQ);     // first query row of this CTA
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int S = p.S;

  // number of K/V tiles this CTA visits
  int kv_len = S;
  if (p.causal) {
    int qend = q0 + BQ * NQ;
    kv_len = qend < S ? qend : S;
  }
  const int nkv = (kv_len + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    mbar_init(bar_q_full, 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&bar_k_full[i], 1);
      mbar_init(&bar_v_full[i], 1);
      mbar_init(&bar_kv_empty[i], 1);
    }
    for (int i = 0; i < NQ; ++i) {
      mbar_init(&bar_s_full[i], 1);
      mbar_init(&bar_p_full[i], 128);
      mbar_init(&bar_pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == C::kSoftmaxWarps + 1) {
    tmem_alloc(tmem_base_smem, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == C::kSoftmaxWarps) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      prefetch_tmap(&tmap);
      mbar_expect_tx(bar_q_full, NQ * C::kTileBytes);
      for (int q = 0; q < NQ; ++q)
        for (int hf = 0; hf < C::kHalves; ++hf)
          tma_load_4d(q_smem + q * C::kTileBytes + hf * (BQ * 128), &tmap, bar_q_full, hf * 64,
                      0 * p.H + h, q0 + q * BQ, b);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % STAGES;
        mbar_wait(&bar_kv_empty[st], ((j / STAGES) & 1) ^ 1);
        mbar_expect_tx(&bar_k_full[st], C::kTileBytes);
        for (int hf = 0; hf < C::kHalves; ++hf)
          tma_load_4d(k_smem + st * C::kTileBytes + hf * (BKV * 128), &tmap, &bar_k_full[st],
                      hf * 64, 1 * p.H + h, j * BKV, b);
        mbar_expect_tx(&bar_v_full[st], C::kTileBytes);
        for (int hf = 0; hf < C::kHalves; ++hf)
          tma_load_4d(v_smem + st * C::kTileBytes + hf * (BKV * 128), &tmap, &bar_v_full[st],
                      hf * 64, 2 * p.H + h, j * BKV, b);
      }
    }
  } else if (warp == C::kSoftmaxWarps + 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, D, 0, 1);
      auto issue_qk = [&](int q, int st) {
        const uint32_t qa = smem_u32(q_smem + q * C::kTileBytes);
        const uint32_t ka = smem_u32(k_smem + st * C::kTileBytes);
        const uint32_t s_tmem = tmem_base + q * C::kColsPerQ;
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k / 4) * (BQ * 128) + (k % 4) * 32;
          mma_ss(s_tmem, make_smem_desc_sw128(qa + off, 16, 1024),
                 make_smem_desc_sw128(ka + off, 16, 1024), idesc_qk, k > 0);
        }
      };
      auto issue_pv = [&](int q, int st, bool acc) {
        const uint32_t va = smem_u32(v_smem + st * C::kTileBytes);
        const uint32_t p_tmem = tmem_base + q * C::kColsPerQ + 128;
        const uint32_t o_tmem = tmem_base + q * C::kColsPerQ + 192;
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          mma_ts(o_tmem, p_tmem + k * 8, make_smem_desc_sw128(va + k * 2048, BKV * 128, 1024),
                 idesc_pv, (acc || k > 0) ? 1u : 0u);
        }
      };
      mbar_wait(bar_q_full, 0);
      mbar_wait(&bar_k_full[0], 0);
      tc_fence_after();
      for (int q = 0; q < NQ; ++q) {
        issue_qk(q, 0);
        mma_commit(&bar_s_full[q]);
      }
      for (int j = 0; j < nkv; ++j) {
        const int st = j % STAGES;
        for (int q = 0; q < NQ; ++q) {
          mbar_wait(&bar_p_full[q], j & 1);
          tc_fence_after();
          if (j + 1 < nkv) {
            const int st1 = (j + 1) % STAGES;
            if (q == 0) {
              mbar_wait(&bar_k_full[st1], ((j + 1) / STAGES) & 1);
              tc_fence_after();
            }
            issue_qk(q, st1);
            mma_commit(&bar_s_full[q]);
          }
          if (q == 0) {
            mbar_wait(&bar_v_full[st], (j / STAGES) & 1);
            tc_fence_after();
          }
          issue_pv(q, st, j > 0);
          mma_commit(&bar_pv_done[q]);
        }
        mma_commit(&bar_kv_empty[st]);
      }
    }
  } else {
    // ============================== softmax / epilogue ==============================
    const int q = warp / 4;                         // query tile handled by this warp
    const int row = (warp % 4) * 32 + lane;         // row inside the tile == TMEM lane
    const int q_idx = q0 + q * BQ + row;            // global query index
    const uint32_t lane_addr = static_cast<uint32_t>((warp % 4) * 32) << 16;
    const uint32_t s_tmem = tmem_base + q * C::kColsPerQ + lane_addr;
    const uint32_t p_tmem = s_tmem + 128;
    const uint32_t o_tmem = s_tmem + 192;
    const float sl2 = p.scale_log2;
    float m = -INFINITY;   // running max (scaled, log2 domain)
    float l = 0.f;         // running denominator
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&bar_s_full[q], j & 1);
      tc_fence_after();
      const int kv0 = j * BKV;
      // columns >= lim are masked out
      int lim = S - kv0;
      if (p.causal) {
        int c = q_idx - kv0 + 1;
        lim = c < lim ? c : lim;
      }
      const bool need_mask = lim < BKV;
      // ---- pass 1: row max ----
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(s_tmem + c * 32, r);
        tmem_wait_ld();
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < lim) mx = fmaxf(mx, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      float m_new = fmaxf(m, mx * sl2);
      // lazy rescale: keep the stale max while it is within 2^8 of the true one
      if (m != -INFINITY && m_new - m <= kRescaleThreshold) m_new = m;
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;   // fully masked row so far
      const float alpha = (m == -INFINITY) ? 1.f : exp2f(m - m_use);
      if (j > 0) {
        mbar_wait(&bar_pv_done[q], (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
          for (int c = 0; c < D / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(o_tmem + c * 32, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st32(o_tmem + c * 32, r);
          }
        }
      }
      // ---- pass 2: p = exp2(s * scale - m), row sum, bf16 P -> TMEM ----
      float rowsum = 0.f;
      const float neg_m = -m_use;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(s_tmem + c * 32, r);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float e0 = exp2f(fmaf(__uint_as_float(r[i]), sl2, neg_m));
          float e1 = exp2f(fmaf(__uint_as_float(r[i + 1]), sl2, neg_m));
          if (need_mask) {
            if (c * 32 + i >= lim) e0 = 0.f;
            if (c * 32 + i + 1 >= lim) e1 = 0.f;
          }
          rowsum += e0 + e1;
          pk[i / 2] = pack_bf16(e0, e1);
        }
        tmem_st16(p_tmem + c * 16, pk);
      }
      l = l * alpha + rowsum;
      m = m_new;
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bar_p_full[q]);
    }
    // ---- epilogue: O / l -> bf16, token-major store ----
    mbar_wait(&bar_pv_done[q], (nkv - 1) & 1);
    tc_fence_after();
    const float inv_l = l > 0.f ? 1.0f / l : 0.f;
    const bool valid = q_idx < S;
    __nv_bfloat16* orow;
    if (p.out2 == nullptr) orow = p.out + (((int64_t)b * S + q_idx) * p.H + h) * D;
    else if (q_idx < p.S_split) orow = p.out + (((int64_t)b * p.S_split + q_idx) * p.H + h) * D;
    else orow = p.out2 + (((int64_t)b * (S - p.S_split) + (q_idx - p.S_split)) * p.H + h) * D;
#pragma unroll
    for (int c = 0; c < D / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(o_tmem + c * 32, r);
      tmem_wait_ld();
      if (valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(r[i + 0]) * inv_l, __uint_as_float(r[i + 1]) * inv_l);
          v.y = pack_bf16(__uint_as_float(r[i + 2]) * inv_l, __uint_as_float(r[i + 3]) * inv_l);
          v.z = pack_bf16(__uint_as_float(r[i + 4]) * inv_l, __uint_as_float(r[i + 5]) * inv_l);
          v.w = pack_bf16(__uint_as_float(r[i + 6]) * inv_l, __uint_as_float(r[i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + i) = v;
        }
      }
    }
    if (valid && p.lse) {
      const float mm = (m == -INFINITY) ? 0.f : m;
      p.lse[((int64_t)b * p.H + h) * S + q_idx] = (mm + log2f(l)) * 0.6931471805599453f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == C::kSoftmaxWarps + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

template <int D, int NQ>
int launch(const void* qkv, void* out, void* out2, int64_t S_split, float* lse, int64_t B, int64_t S, int64_t H, float scale,
           int causal, cudaStream_t st) {
  using C = Cfg<D, NQ>;
  CUtensorMap tmap;
  const uint64_t dims[4] = {(uint64_t)D, (uint64_t)(3 * H), (uint64_t)S, (uint64_t)B};
  const uint64_t strides[4] = {0, (uint64_t)D * 2, (uint64_t)(3 * H * D) * 2, (uint64_t)(S * 3 * H * D) * 2};
  const uint32_t box[4] = {64, 1, BQ, 1};
  int rc = make_tmap_bf16(&tmap, qkv, 4, dims, strides, box, true);
  if (rc != ADVGRPO_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    ADVGRPO_CUDA_CALL(cudaFuncSetAttribute(attn_fwd_kernel<D, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  Params p;
  p.out = (__nv_bfloat16*)out;
  p.out2 = (__nv_bfloat16*)out2;
  p.S_split = (int)S_split;
  p.lse = lse;
  p.S = (int)S;
  p.H = (int)H;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  dim3 grid((unsigned)((S + BQ * NQ - 1) / (BQ * NQ)), (unsigned)H, (unsigned)B);
  attn_fwd_kernel<D, NQ><<<grid, C::kThreads, C::kSmemBytes, st>>>(tmap, p);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

}  // namespace

// variant: 0 = auto, 1 = one query tile per CTA (2 CTAs/SM), 2 = two query tiles per CTA
int attn_fwd_dispatch(const void* qkv, void* out, void* out2, int64_t S_split, float* lse, int64_t B, int64_t S, int64_t H,
                      int64_t D, float scale, int causal, int variant, cudaStream_t st) {
  if (D == 64) {
    if (variant == 2) return launch<64, 2>(qkv, out, out2, S_split, lse, B, S, H, scale, causal, st);
    return launch<64, 1>(qkv, out, out2, S_split, lse, B, S, H, scale, causal, st);
  }
  if (D == 128) return launch<128, 1>(qkv, out, out2, S_split, lse, B, S, H, scale, causal, st);
  return set_error(ADVGRPO_ERR_UNSUPPORTED, "attn_fwd: head_dim %lld not in {64, 128}", (long long)D);
}

}  // namespace advgrpo

using namespace advgrpo;

extern "C" {

int advgrpo_attn_fwd(const void* qkv, void* out, void* out2, int64_t S_split, float* lse, int64_t B,
                     int64_t S, int64_t H, int64_t D, float scale, int causal,
                     advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(qkv && out, "attn_fwd: null pointer");
  ADVGRPO_CHECK_ARG(!out2 || (S_split > 0 && S_split < S && aligned16(out2)), "attn_fwd: out2 needs 0 < S_split < S");
  ADVGRPO_CHECK_ARG(B >= 1 && S >= 1 && H >= 1 && B <= 65535 && H <= 65535, "attn_fwd: bad sizes B=%lld S=%lld H=%lld",
                    (long long)B, (long long)S, (long long)H);
  ADVGRPO_CHECK_ARG(aligned16(qkv) && aligned16(out), "attn_fwd: tensors must be 16-byte aligned");
  return attn_fwd_dispatch(qkv, out, out2, S_split, lse, B, S, H, D, scale, causal, 0, (cudaStream_t)stream);
}

// Test/bench hook (not part of the reference-facing surface): pick the CTA shape explicitly.
int advgrpo_attn_fwd_variant(const void* qkv, void* out, float* lse, int64_t B, int64_t S, int64_t H,
                             int64_t D, float scale, int causal, int variant,
                             advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(qkv && out, "attn_fwd: null pointer");
  return attn_fwd_dispatch(qkv, out, nullptr, 0, lse, B, S, H, D, scale, causal, variant, (cudaStream_t)stream);
}

}  // extern "C"

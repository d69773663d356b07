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
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

namespace advgrpo {
namespace {

using namespace sm100;

constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int STAGES = 2;
constexpr float kRescaleThreshold = 8.0f;  // log2 domain: skip O rescale while max grows < 2^8

template <int D, int NQ, int RS = 1>
struct Cfg {
  static constexpr int kHalves = D / 64;                 // 64-column (128 B) swizzle atoms per row
  static constexpr int kTileBytes = BQ * D * 2;          // one Q/K/V tile
  static constexpr int kSmemTiles = NQ * kTileBytes + STAGES * 2 * kTileBytes;
  static constexpr int kXchBytes = 3 * NQ * RS * 128 * 4;   // row max (2 parities) + row sum exchange between row-split threads
  static constexpr int kSmemBytes = kSmemTiles + 1024 /*align*/ + 256 /*barriers*/ + kXchBytes;
  static constexpr int kColsPerQ = 128 + 64 + D;
  static constexpr int kTmemCols = (NQ * kColsPerQ <= 256) ? 256 : 512;
  static constexpr int kSoftmaxWarps = 4 * NQ * RS;   // RS threads share one query row (column slices)
  static constexpr int kThreads = (kSoftmaxWarps + 4) * 32;   // + one utility warpgroup (TMA, MMA, 2 idle)
  // register re-balancing (setmaxnreg works per 4-warp group): utility warps shrink, softmax warps grow
  static constexpr bool kRebalance = (D == 64);
  static constexpr int kRegsUtility = (RS == 2) ? 40 : 56;
  // budgets: (softmax threads) * kRegsSoftmax + 128 * kRegsUtility <= threads * (registers at launch)
  //   RS=1: NQ=2 256*216+128*56 <= 384*168 ; NQ=1 128*200+128*56 <= 256*128
  //   RS=2: NQ=2 512*104+128*40 <= 640*96  ; NQ=1 256*96+128*40 <= 384*80
  static constexpr int kRegsSoftmax = (RS == 2) ? ((NQ == 2) ? 104 : 96) : ((NQ == 2) ? 216 : 200);
  static_assert(NQ * kColsPerQ <= 512, "TMEM budget");
};

struct Params {
  __nv_bfloat16* out;   // [B, S, H, D]   (rows [0, S_split) when out2 is set: [B, S_split, H, D])
  __nv_bfloat16* out2;  // rows [S_split, S): [B, S - S_split, H, D], or null
  float* lse;           // [B, H, S] or null
  int S, H, S_split;
  float scale_log2;     // softmax scale * log2(e)
  int causal;
  int park;             // bit 0: MMA warp waits parked, bit 1: softmax S waits parked (experiment switches)
  const float* bias;    // optional additive score bias [H, S, S] (T5 relative positions); persistent kernel only
  long long* trace;     // debug timeline (quad kernel, TRACE instantiation only): [cta < 4][role 2][tile < 128][8] clock64 stamps
};

// 2^x for x <= ~8 on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax, rel. err 7.5e-5): used for
// EMU of every 4 element pairs so the 16/clk/SM MUFU unit is not the only exp engine (it would cap the
// tensor pipe at 50% for head_dim 64).  Two elements at a time with the packed f32x2 FMA / ADD.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);      // 1.5 * 2^23: round to nearest integer
  const float2 xr = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(xr, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));         // f in [-0.5, 0.5]
  float2 pl = __ffma2_rn(f, make_float2(0.055171650f, 0.055171650f), make_float2(0.24261113f, 0.24261113f));
  pl = __ffma2_rn(pl, f, make_float2(0.69326097f, 0.69326097f));
  pl = __ffma2_rn(pl, f, make_float2(0.99992806f, 0.99992806f));
  float2 r;
  r.x = __int_as_float(__float_as_int(pl.x) + (__float_as_int(xr.x) << 23));
  r.y = __int_as_float(__float_as_int(pl.y) + (__float_as_int(xr.y) << 23));
  return r;
}

template <int D, int NQ, int EMU, int RS>
__global__ void __launch_bounds__(Cfg<D, NQ, RS>::kThreads, (NQ == 1 && D == 64) ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  using C = Cfg<D, NQ, RS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;                                   // NQ tiles
  uint8_t* k_smem = smem + NQ * C::kTileBytes;              // STAGES tiles
  uint8_t* v_smem = k_smem + STAGES * C::kTileBytes;        // STAGES tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kSmemTiles);
  uint64_t* bar_q_full = bars;                  // 1
  uint64_t* bar_k_full = bars + 1;              // STAGES
  uint64_t* bar_v_full = bar_k_full + STAGES;   // STAGES
  uint64_t* bar_k_empty = bar_v_full + STAGES;  // STAGES   K(j) is free once QK(j) completed (long before PV(j))
  uint64_t* bar_v_empty = bar_k_empty + STAGES; // STAGES   V(j) is free once PV(j) completed
  uint64_t* bar_s_full = bar_v_empty + STAGES;  // NQ
  uint64_t* bar_p_full = bar_s_full + NQ;       // NQ
  uint64_t* bar_pv_done = bar_p_full + NQ;      // NQ
  uint64_t* bar_s_free = bar_pv_done + NQ;      // NQ   softmax has S in registers: QK(j+1) may overwrite it
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bar_s_free + NQ);
  float* xch = reinterpret_cast<float*>(smem + C::kSmemTiles + 256);   // [3][NQ][RS][128]

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
      mbar_init(&bar_k_empty[i], 1);
      mbar_init(&bar_v_empty[i], 1);
    }
    for (int i = 0; i < NQ; ++i) {
      mbar_init(&bar_s_full[i], 1);
      mbar_init(&bar_p_full[i], 128 * RS);
      mbar_init(&bar_pv_done[i], 1);
      mbar_init(&bar_s_free[i], 128 * RS);
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

  if (warp >= C::kSoftmaxWarps) {
  if constexpr (C::kRebalance) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::kRegsUtility));
  if (warp == C::kSoftmaxWarps) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      prefetch_tmap(&tmap);
      mbar_expect_tx(bar_q_full, NQ * C::kTileBytes);
      for (int q = 0; q < NQ; ++q)
        for (int hf = 0; hf < C::kHalves; ++hf)
          tma_load_4d(q_smem + q * C::kTileBytes + hf * (BQ * 128), &tmap, bar_q_full, hf * 64,
                      0 * p.H + h, q0 + q * BQ, b);
      // K runs STAGES tiles ahead of the QK MMAs: its stage is released by QK(j) itself, so K(j + STAGES) is
      // already in flight while the softmax of tile j is still running (a K stage released only after PV(j)
      // left one softmax period for the whole L2 / HBM round trip and stalled S(j+1)); V follows one tile behind.
      auto load_k = [&](int j) {
        const int st = j % STAGES;
        mbar_expect_tx(&bar_k_full[st], C::kTileBytes);
        for (int hf = 0; hf < C::kHalves; ++hf)
          tma_load_4d(k_smem + st * C::kTileBytes + hf * (BKV * 128), &tmap, &bar_k_full[st],
                      hf * 64, 1 * p.H + h, j * BKV, b);
      };
      for (int j = 0; j < STAGES && j < nkv; ++j) load_k(j);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % STAGES;
        mbar_wait_parked(&bar_v_empty[st], ((j / STAGES) & 1) ^ 1);
        mbar_expect_tx(&bar_v_full[st], C::kTileBytes);
        for (int hf = 0; hf < C::kHalves; ++hf)
          tma_load_4d(v_smem + st * C::kTileBytes + hf * (BKV * 128), &tmap, &bar_v_full[st],
                      hf * 64, 2 * p.H + h, j * BKV, b);
        if (j + STAGES < nkv) {
          mbar_wait_parked(&bar_k_empty[st], (j / STAGES) & 1);        // QK(j) done
          load_k(j + STAGES);
        }
      }
    }
  } else if (warp == C::kSoftmaxWarps + 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, D, 0, 1);
      // descriptors are built once; per MMA only the 16-byte-unit start address is advanced
      const uint64_t q_d0 = make_smem_desc_sw128(smem_u32(q_smem), 16, 1024);
      const uint64_t k_d0 = make_smem_desc_sw128(smem_u32(k_smem), 16, 1024);
      const uint64_t v_d0 = make_smem_desc_sw128(smem_u32(v_smem), BKV * 128, 1024);
      auto issue_qk = [&](int q, int st) {
        const uint32_t s_tmem = tmem_base + q * C::kColsPerQ;
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = ((k / 4) * (BQ * 128) + (k % 4) * 32) >> 4;
          mma_ss(s_tmem, q_d0 + (uint64_t)(q * (C::kTileBytes >> 4) + off),
                 k_d0 + (uint64_t)(st * (C::kTileBytes >> 4) + off), idesc_qk, k > 0);
        }
      };
      auto issue_pv = [&](int q, int st, bool acc) {
        const uint32_t p_tmem = tmem_base + q * C::kColsPerQ + 128;
        const uint32_t o_tmem = tmem_base + q * C::kColsPerQ + 192;
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          mma_ts(o_tmem, p_tmem + k * 8, v_d0 + (uint64_t)(st * (C::kTileBytes >> 4) + k * 128), idesc_pv,
                 (acc || k > 0) ? 1u : 0u);
        }
      };
      const bool pk = p.park & 1;
      auto wait = [&](uint64_t* bar, uint32_t par) { if (pk) mbar_wait_parked(bar, par); else mbar_wait(bar, par); };
      wait(bar_q_full, 0);
      wait(&bar_k_full[0], 0);
      tc_fence_after();
      for (int q = 0; q < NQ; ++q) {
        issue_qk(q, 0);
        mma_commit(&bar_s_full[q]);
      }
      mma_commit(&bar_k_empty[0]);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % STAGES;
        if (j + 1 < nkv) {
          const int st1 = (j + 1) % STAGES;
          for (int q = 0; q < NQ; ++q) {
            wait(&bar_s_free[q], j & 1);           // S(j) is in the softmax warps' registers
            if (q == 0) wait(&bar_k_full[st1], ((j + 1) / STAGES) & 1);
            tc_fence_after();
            issue_qk(q, st1);
            mma_commit(&bar_s_full[q]);
          }
          mma_commit(&bar_k_empty[st1]);
        }
        for (int q = 0; q < NQ; ++q) {
          wait(&bar_p_full[q], j & 1);
          if (q == 0) wait(&bar_v_full[st], (j / STAGES) & 1);
          tc_fence_after();
          issue_pv(q, st, j > 0);
          mma_commit(&bar_pv_done[q]);
        }
        mma_commit(&bar_v_empty[st]);
      }
    }
  }
  } else {
    // ============================== softmax / epilogue ==============================
    if constexpr (C::kRebalance) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::kRegsSoftmax));
    constexpr int NCH = 4 / RS;                     // 32-column chunks of the S row owned by this thread
    constexpr int CPT = 128 / RS;                   // S columns per thread
    constexpr int OCH = (D / 32) / RS;              // 32-column chunks of the O row owned by this thread
    const int wg = warp / 4;
    const int q = wg / RS;                          // query tile handled by this warp
    const int half = wg % RS;                       // column slice of the row (RS threads share a row)
    const int row = (warp % 4) * 32 + lane;         // row inside the tile == TMEM lane
    const int q_idx = q0 + q * BQ + row;            // global query index
    const uint32_t lane_addr = static_cast<uint32_t>((warp % 4) * 32) << 16;
    const uint32_t s_tmem = tmem_base + q * C::kColsPerQ + lane_addr + half * CPT;
    const uint32_t p_tmem = tmem_base + q * C::kColsPerQ + 128 + lane_addr + half * (CPT / 2);
    const uint32_t o_tmem = tmem_base + q * C::kColsPerQ + 192 + lane_addr + half * (OCH * 32);
    float* xmax = xch + (q * RS) * 128 + row;       // + parity * NQ*RS*128 + slice * 128
    float* xsum = xch + 2 * NQ * RS * 128 + (q * RS) * 128 + row;
    const float sl2 = p.scale_log2;
    float m = -INFINITY;   // running max (scaled, log2 domain)
    float l = 0.f;         // running denominator (this thread's column slice)
    for (int j = 0; j < nkv; ++j) {
      if (p.park & 2) mbar_wait_parked(&bar_s_full[q], j & 1); else mbar_wait(&bar_s_full[q], j & 1);
      tc_fence_after();
      const int kv0 = j * BKV + half * CPT;         // first kv index of my column slice
      // columns >= lim (relative to my slice) are masked out
      int lim = S - kv0;
      if (p.causal) {
        int c = q_idx - kv0 + 1;
        lim = c < lim ? c : lim;
      }
      const bool need_mask = lim < CPT;
      // ---- my slice of the S row -> registers (single TMEM read), then release S for QK(j+1) ----
      uint32_t sv[NCH][32];
#pragma unroll
      for (int c = 0; c < NCH; ++c) tmem_ld32(s_tmem + c * 32, sv[c]);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&bar_s_free[q]);
      if (need_mask) {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= lim) sv[c][i] = 0xff800000u;   // -inf
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // 4 independent FMNMX3 chains
#pragma unroll
      for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          mx4[(i / 2) & 3] = fmaxf(mx4[(i / 2) & 3], fmaxf(__uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1])));
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      if constexpr (RS == 2) {
        // the two threads of a row agree on the row max through shared memory (double-buffered by tile parity)
        float* slot = xmax + (j & 1) * (NQ * RS * 128);
        slot[half * 128] = mx;
        named_bar_sync(1 + q, 128 * RS);
        mx = fmaxf(mx, slot[(half ^ 1) * 128]);
      }
      float m_new = fmaxf(m, mx * sl2);
      // lazy rescale: keep the stale max while it is within 2^8 of the true one
      if (m != -INFINITY && m_new - m <= kRescaleThreshold) m_new = m;
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;   // fully masked row so far
      const float alpha = (m == -INFINITY) ? 1.f : ex2(m - m_use);
      // ---- p = exp2(s * scale - m) and the row sum, bf16 P kept in registers: the exponentials do not depend
      //      on PV(j-1), so they run while the tensor pipe is still busy with it ----
      float2 sum2 = make_float2(0.f, 0.f);
      const float2 sc2 = make_float2(sl2, sl2), nm2 = make_float2(-m_use, -m_use);
      uint32_t pk[NCH][16];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1])), sc2, nm2);
          float2 e;
          if ((i / 2) % 4 < EMU) {
            e = ex2_poly2(x);
          } else {
            e.x = ex2(x.x);
            e.y = ex2(x.y);
          }
          sum2 = __fadd2_rn(sum2, e);
          pk[c][i / 2] = pack_bf16(e.x, e.y);
        }
      }
      // ---- only now wait for PV(j-1): it reads P(j-1) from the TMEM columns P(j) is about to overwrite, and O
      //      may only be rescaled between PV(j-1) and PV(j) ----
      if (j > 0) {
        mbar_wait(&bar_pv_done[q], (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
          for (int c = 0; c < OCH; ++c) {
            uint32_t r[32];
            tmem_ld32(o_tmem + c * 32, r);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st32(o_tmem + c * 32, r);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < NCH; ++c) tmem_st16(p_tmem + c * 16, pk[c]);
      const float rowsum = sum2.x + sum2.y;
      l = l * alpha + rowsum;
      m = m_new;
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bar_p_full[q]);
    }
    // ---- epilogue: O / l -> bf16, token-major store ----
    mbar_wait(&bar_pv_done[q], (nkv - 1) & 1);
    tc_fence_after();
    if constexpr (RS == 2) {
      xsum[half * 128] = l;
      named_bar_sync(1 + q, 128 * RS);
      l += xsum[(half ^ 1) * 128];
    }
    const float inv_l = l > 0.f ? 1.0f / l : 0.f;
    const bool valid = q_idx < S;
    __nv_bfloat16* orow;
    if (p.out2 == nullptr) orow = p.out + (((int64_t)b * S + q_idx) * p.H + h) * D;
    else if (q_idx < p.S_split) orow = p.out + (((int64_t)b * p.S_split + q_idx) * p.H + h) * D;
    else orow = p.out2 + (((int64_t)b * (S - p.S_split) + (q_idx - p.S_split)) * p.H + h) * D;
    orow += half * (OCH * 32);
#pragma unroll
    for (int c = 0; c < OCH; ++c) {
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
    if (valid && p.lse && half == 0) {
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

// ---------------------------------------------------------------------------------------------------------
// Persistent variant (default): one CTA per SM slot walks a static round-robin list of (query tile, head, sample)
// work items.  All pipelines (Q double buffer, K / V rings, S / P / O in TMEM) run on global tile counters, so the
// Q + first K loads and the first QK MMA of item n+1 are issued while the softmax warps are still finishing and
// writing out item n: the per-CTA prologue (tensor-map fetch, TMEM allocation, Q/K round trip) and the epilogue
// no longer idle the SM between 10-tile work items.
template <int D>
struct PCfg {
  static constexpr int kHalves = D / 64;
  static constexpr int kTileBytes = BQ * D * 2;
  static constexpr int kSmemTiles = 2 * kTileBytes + STAGES * 2 * kTileBytes;   // 2 Q buffers + K ring + V ring
  static constexpr int kSmemBytes = kSmemTiles + 1024 + 256;
  static constexpr int kColsPerQ = 128 + 64 + D;
  static constexpr int kTmemCols = (kColsPerQ <= 256) ? 256 : 512;
  static constexpr int kThreads = 256;
  static constexpr bool kRebalance = (D == 64);
};

template <int D, int EMU>
__global__ void __launch_bounds__(256, (D == 64) ? 2 : 1)
attn_fwd_persist_kernel(const __grid_constant__ CUtensorMap tmap, const Params p, const int n_items, const int nqt) {
  using C = PCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;                                   // 2 tiles
  uint8_t* k_smem = smem + 2 * C::kTileBytes;               // STAGES tiles
  uint8_t* v_smem = k_smem + STAGES * C::kTileBytes;        // STAGES tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kSmemTiles);
  uint64_t* bar_q_full = bars;                   // 2
  uint64_t* bar_q_empty = bars + 2;              // 2
  uint64_t* bar_k_full = bars + 4;               // STAGES
  uint64_t* bar_v_full = bar_k_full + STAGES;    // STAGES
  uint64_t* bar_k_empty = bar_v_full + STAGES;   // STAGES
  uint64_t* bar_v_empty = bar_k_empty + STAGES;  // STAGES
  uint64_t* bar_s_full = bar_v_empty + STAGES;   // 1
  uint64_t* bar_p_full = bar_s_full + 1;         // 1
  uint64_t* bar_pv_done = bar_p_full + 1;        // 1
  uint64_t* bar_s_free = bar_pv_done + 1;        // 1
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bar_s_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.S;
  const int H = p.H;

  // item -> (query tile, head, sample) and its number of K/V tiles
  auto item_q0 = [&](int it) { return (it % nqt) * BQ; };
  auto item_h = [&](int it) { return (it / nqt) % H; };
  auto item_b = [&](int it) { return it / (nqt * H); };
  auto item_nkv = [&](int it) {
    int kv_len = S;
    if (p.causal) {
      const int qend = item_q0(it) + BQ;
      kv_len = qend < S ? qend : S;
    }
    return (kv_len + BKV - 1) / BKV;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_q_full[i], 1);
      mbar_init(&bar_q_empty[i], 1);
    }
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&bar_k_full[i], 1);
      mbar_init(&bar_v_full[i], 1);
      mbar_init(&bar_k_empty[i], 1);
      mbar_init(&bar_v_empty[i], 1);
    }
    mbar_init(bar_s_full, 1);
    mbar_init(bar_p_full, 128);
    mbar_init(bar_pv_done, 1);
    mbar_init(bar_s_free, 128);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_base_smem, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_trigger();   // programmatic dependent launch (common.cuh): no-ops unless the launch carries the attribute
  pdl_wait();

  if (warp >= 4) {
    if constexpr (C::kRebalance) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(56));
    // Utility warps run their loops warp-uniformly and only the TMA / tcgen05 instructions sit under elect_one() (a
    // divergent `lane == 0` branch makes ptxas wrap each of them in an ELECT + BRA.U.ANY loop); QK and PV are issued
    // by two different warps.  See attn_fwd_quad_kernel.
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t a_q_full = bar0, a_q_empty = bar0 + 16, a_k_full = bar0 + 32, a_v_full = a_k_full + 8 * STAGES,
                   a_k_empty = a_v_full + 8 * STAGES, a_v_empty = a_k_empty + 8 * STAGES, a_s_full = a_v_empty + 8 * STAGES,
                   a_p_full = a_s_full + 8, a_pv_done = a_s_full + 16, a_s_free = a_s_full + 24;
    const uint32_t q_sm = smem_u32(q_smem), k_sm = q_sm + 2 * C::kTileBytes, v_sm = k_sm + STAGES * C::kTileBytes;
    if (warp == 4) {
      // ============================== TMA producer ==============================
      if (elect_one()) prefetch_tmap(&tmap);
      // K cursor (runs STAGES tiles ahead, loads the Q tile when it enters a new item) and V cursor
      int k_item = blockIdx.x, k_j = 0, k_n = 0, k_g = 0;          // item, tile in item, local item index, global tile
      int k_nkv = k_item < n_items ? item_nkv(k_item) : 0;
      auto advance_k = [&]() {
        if (k_item >= n_items) return;
        if (k_j == 0) {
          const int qb = k_n & 1;
          mbar_wait_a(a_q_empty + 8 * qb, ((k_n >> 1) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx_a(a_q_full + 8 * qb, C::kTileBytes);
            for (int hf = 0; hf < C::kHalves; ++hf)
              tma_load_4d_a(q_sm + qb * C::kTileBytes + hf * (BQ * 128), &tmap, a_q_full + 8 * qb, hf * 64,
                            0 * H + item_h(k_item), item_q0(k_item), item_b(k_item));
          }
        }
        const int st = k_g % STAGES;
        mbar_wait_a(a_k_empty + 8 * st, ((k_g / STAGES) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx_a(a_k_full + 8 * st, C::kTileBytes);
          for (int hf = 0; hf < C::kHalves; ++hf)
            tma_load_4d_a(k_sm + st * C::kTileBytes + hf * (BKV * 128), &tmap, a_k_full + 8 * st, hf * 64,
                          1 * H + item_h(k_item), k_j * BKV, item_b(k_item));
        }
        ++k_g;
        if (++k_j == k_nkv) {
          k_j = 0;
          ++k_n;
          k_item += gridDim.x;
          k_nkv = k_item < n_items ? item_nkv(k_item) : 0;
        }
      };
      for (int i = 0; i < STAGES; ++i) advance_k();
      int g = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int nkv = item_nkv(item), h = item_h(item), b = item_b(item);
        for (int j = 0; j < nkv; ++j, ++g) {
          const int st = g % STAGES;
          mbar_wait_a(a_v_empty + 8 * st, ((g / STAGES) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx_a(a_v_full + 8 * st, C::kTileBytes);
            for (int hf = 0; hf < C::kHalves; ++hf)
              tma_load_4d_a(v_sm + st * C::kTileBytes + hf * (BKV * 128), &tmap, a_v_full + 8 * st, hf * 64, 2 * H + h,
                            j * BKV, b);
          }
          advance_k();
        }
      }
    } else if (warp == 5) {
      // ============================== QK issuer: S(g) = Q K_g^T ==============================
      constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);
      const uint64_t q_d0 = make_smem_desc_sw128(q_sm, 16, 1024);
      const uint64_t k_d0 = make_smem_desc_sw128(k_sm, 16, 1024);
      const uint32_t s_tmem = tmem_base;
      int g = 0, n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const int nkv = item_nkv(item);
        const int qb = n & 1;
        for (int j = 0; j < nkv; ++j, ++g) {
          const int st = g % STAGES;
          if (g > 0) mbar_wait_a(a_s_free, (g - 1) & 1);          // S(g-1) is in the softmax warps' registers
          if (j == 0) mbar_wait_a(a_q_full + 8 * qb, (n >> 1) & 1);
          mbar_wait_a(a_k_full + 8 * st, (g / STAGES) & 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < D / 16; ++k) {
              const uint32_t off = ((k / 4) * (BQ * 128) + (k % 4) * 32) >> 4;
              mma_ss(s_tmem, q_d0 + (uint64_t)(qb * (C::kTileBytes >> 4) + off),
                     k_d0 + (uint64_t)(st * (C::kTileBytes >> 4) + off), idesc_qk, k > 0);
            }
            mma_commit_a(a_s_full);
            mma_commit_a(a_k_empty + 8 * st);
            if (j == nkv - 1) mma_commit_a(a_q_empty + 8 * qb);   // last QK of the item read the Q tile
          }
          __syncwarp();
        }
      }
    } else if (warp == 6) {
      // ============================== PV issuer: O += P_g V_g (P read from TMEM) ==============================
      constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, D, 0, 1);
      const uint64_t v_d0 = make_smem_desc_sw128(v_sm, BKV * 128, 1024);
      const uint32_t p_tmem = tmem_base + 128, o_tmem = tmem_base + 192;
      int g = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int nkv = item_nkv(item);
        for (int j = 0; j < nkv; ++j, ++g) {
          const int st = g % STAGES;
          mbar_wait_a(a_p_full, g & 1);
          mbar_wait_a(a_v_full + 8 * st, (g / STAGES) & 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BKV / 16; ++k)
              mma_ts(o_tmem, p_tmem + k * 8, v_d0 + (uint64_t)(st * (C::kTileBytes >> 4) + k * 128), idesc_pv,
                     (j > 0 || k > 0) ? 1u : 0u);
            mma_commit_a(a_pv_done);
            mma_commit_a(a_v_empty + 8 * st);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ============================== softmax / epilogue ==============================
    if constexpr (C::kRebalance) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(200));
    const int row = warp * 32 + lane;               // row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_addr;
    const uint32_t p_tmem = tmem_base + 128 + lane_addr;
    const uint32_t o_tmem = tmem_base + 192 + lane_addr;
    const float sl2 = p.bias ? 1.0f : p.scale_log2;   // with a bias the scores are moved to the log2 domain first
    int g = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int nkv = item_nkv(item), h = item_h(item), b = item_b(item);
      const int q_idx = item_q0(item) + row;
      float m = -INFINITY;   // running max (scaled, log2 domain)
      float l = 0.f;         // running denominator
      for (int j = 0; j < nkv; ++j, ++g) {
        mbar_wait(bar_s_full, g & 1);
        tc_fence_after();
        const int kv0 = j * BKV;
        int lim = S - kv0;
        if (p.causal) {
          const int c = q_idx - kv0 + 1;
          lim = c < lim ? c : lim;
        }
        const bool need_mask = lim < 128;
        uint32_t sv[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(s_tmem + c * 32, sv[c]);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(bar_s_free);
        if (p.bias) {
          // t = s * scale * log2(e) + bias * log2(e): this thread's row of the [S, S] bias plane of head h
          const float* brow = p.bias + ((int64_t)h * S + (q_idx < S ? q_idx : S - 1)) * S + kv0;
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < lim)
                sv[c][i] = __float_as_uint(fmaf(__uint_as_float(sv[c][i]), p.scale_log2, brow[c * 32 + i] * 1.4426950408889634f));
        }
        if (need_mask) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= lim) sv[c][i] = 0xff800000u;   // -inf
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            mx4[(i / 2) & 3] = fmaxf(mx4[(i / 2) & 3], fmaxf(__uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1])));
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        float m_new = fmaxf(m, mx * sl2);
        if (m != -INFINITY && m_new - m <= kRescaleThreshold) m_new = m;   // lazy rescale
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = (m == -INFINITY) ? 1.f : ex2(m - m_use);
        float2 sum2 = make_float2(0.f, 0.f);
        const float2 sc2 = make_float2(sl2, sl2), nm2 = make_float2(-m_use, -m_use);
        uint32_t pk[4][16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1])), sc2, nm2);
            float2 e;
            if ((i / 2) % 4 < EMU) {
              e = ex2_poly2(x);
            } else {
              e.x = ex2(x.x);
              e.y = ex2(x.y);
            }
            sum2 = __fadd2_rn(sum2, e);
            pk[c][i / 2] = pack_bf16(e.x, e.y);
          }
        }
        // PV(g-1) reads P(g-1) from the columns P(g) overwrites; O may only be rescaled between PV(g-1) and PV(g).
        // For the first tile of an item the epilogue of the previous item already consumed that phase.
        if (j > 0) {
          mbar_wait(bar_pv_done, (g - 1) & 1);
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
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st16(p_tmem + c * 16, pk[c]);
        l = l * alpha + (sum2.x + sum2.y);
        m = m_new;
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bar_p_full);
      }
      // ---- epilogue: O / l -> bf16, token-major store (the MMA warp is already computing S of the next item) ----
      mbar_wait(bar_pv_done, (g - 1) & 1);
      tc_fence_after();
      const float inv_l = l > 0.f ? 1.0f / l : 0.f;
      const bool valid = q_idx < S;
      __nv_bfloat16* orow;
      if (p.out2 == nullptr) orow = p.out + (((int64_t)b * S + q_idx) * H + h) * D;
      else if (q_idx < p.S_split) orow = p.out + (((int64_t)b * p.S_split + q_idx) * H + h) * D;
      else orow = p.out2 + (((int64_t)b * (S - p.S_split) + (q_idx - p.S_split)) * H + h) * D;
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
      tc_fence_before();      // O is in registers: orders the TMEM reads before the next item's PV(0) (after p_full)
      if (valid && p.lse) {
        const float mm = (m == -INFINITY) ? 0.f : m;
        p.lse[((int64_t)b * H + h) * S + q_idx] = (mm + log2f(l)) * 0.6931471805599453f;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Quad-layout persistent variant (head_dim 64, the MMDiT / DINOv2 shape): the TMA and MMA pipelines are those of
// attn_fwd_persist_kernel, but the softmax of a 128-row query tile runs on EIGHT warps instead of four.  S is read
// with the 16-lane tcgen05.ld shape (16x256b: a row is spread over the 4 threads of a quad, like an mma.sync
// accumulator), so a warp owns 16 rows x 128 columns and a thread 2 rows x 32 columns = 64 scores per tile: ~100
// registers instead of ~200, which lets 16 softmax warps (4 per SM sub-partition, two CTAs per SM) hide the MUFU /
// FMA latencies the 8-warp version exposed (ncu: 0.21 IPC per softmax warp, XU pipe 49 %, tensor pipe 30 %).  The row
// max needs two quad shuffles per tile, the row sum is reduced over the quad once per work item, and P goes back to
// TMEM with the matching 16x128b store (bf16 pairs of adjacent columns are already in one thread).
struct QCfg {
  static constexpr int kTileBytes = BQ * 64 * 2;
  static constexpr int kSmemTiles = 2 * kTileBytes + STAGES * 2 * kTileBytes;
  static constexpr int kSmemBytes = kSmemTiles + 1024 + 256;
  static constexpr int kTmemCols = 256;                  // S 128 | P 64 | O 64
  static constexpr int kSoftmaxWarps = 8;
  static constexpr int kThreads = (kSoftmaxWarps + 4) * 32;
  // 2 CTAs / SM: 384 threads x 80 registers at launch = 256 x kRegsSoftmax + 128 x kRegsUtility
  static constexpr int kRegsSoftmax = 104;
  static constexpr int kRegsUtility = 32;
};

template <int EMU, bool TRACE = false>
__global__ void __launch_bounds__(QCfg::kThreads, 2)
attn_fwd_quad_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_o,
                     const __grid_constant__ CUtensorMap tmap_o2, const Params p, const int n_items, const int nqt,
                     const int tma_out) {
  using C = QCfg;
  constexpr int D = 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;                                   // 2 tiles
  uint8_t* k_smem = smem + 2 * C::kTileBytes;               // STAGES tiles
  uint8_t* v_smem = k_smem + STAGES * C::kTileBytes;        // STAGES tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kSmemTiles);
  uint64_t* bar_q_full = bars;                   // 2
  uint64_t* bar_q_empty = bars + 2;              // 2
  uint64_t* bar_k_full = bars + 4;               // STAGES
  uint64_t* bar_v_full = bar_k_full + STAGES;    // STAGES
  uint64_t* bar_k_empty = bar_v_full + STAGES;   // STAGES
  uint64_t* bar_v_empty = bar_k_empty + STAGES;  // STAGES
  uint64_t* bar_s_full = bar_v_empty + STAGES;   // 1
  uint64_t* bar_p_full = bar_s_full + 1;         // 1
  uint64_t* bar_pv_done = bar_p_full + 1;        // 1
  uint64_t* bar_s_free = bar_pv_done + 1;        // 1
  uint64_t* bar_o_ready = bar_s_free + 1;        // 1   normalised O tile of the finished item is staged in its Q buffer
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bar_o_ready + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.S;
  const int H = p.H;

  auto item_q0 = [&](int it) { return (it % nqt) * BQ; };
  auto item_h = [&](int it) { return (it / nqt) % H; };
  auto item_b = [&](int it) { return it / (nqt * H); };
  auto item_nkv = [&](int it) {
    int kv_len = S;
    if (p.causal) {
      const int qend = item_q0(it) + BQ;
      kv_len = qend < S ? qend : S;
    }
    return (kv_len + BKV - 1) / BKV;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_q_full[i], 1);
      mbar_init(&bar_q_empty[i], 1);
    }
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&bar_k_full[i], 1);
      mbar_init(&bar_v_full[i], 1);
      mbar_init(&bar_k_empty[i], 1);
      mbar_init(&bar_v_empty[i], 1);
    }
    mbar_init(bar_s_full, 1);
    mbar_init(bar_p_full, C::kSoftmaxWarps * 32);
    mbar_init(bar_pv_done, 1);
    mbar_init(bar_s_free, C::kSoftmaxWarps * 32);
    mbar_init(bar_o_ready, C::kSoftmaxWarps * 32);
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
  pdl_trigger();
  pdl_wait();

  if (warp >= C::kSoftmaxWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::kRegsUtility));
    // Utility warps run their loops WARP-UNIFORMLY (all 32 lanes wait on the barriers) and only the TMA / tcgen05
    // instructions themselves sit under elect_one(): under a divergent `lane == 0` branch ptxas wraps every UTCHMMA /
    // UTCBAR / UTMALDG in an ELECT + BRA.U.ANY loop (~7 extra instructions each), which made the single issuing
    // thread -- not the softmax -- the critical path of the tile loop (clock64 timeline: ~2200 clocks per tile).
    // The QK and the PV MMAs are issued by two different warps, so each pays only two barrier waits per tile.
    auto pin = [](uint32_t v) { uint32_t r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; };
    const uint32_t bar0 = pin(smem_u32(bars));
    const uint32_t a_q_full = bar0, a_q_empty = bar0 + 16, a_k_full = bar0 + 32, a_v_full = a_k_full + 8 * STAGES,
                   a_k_empty = a_v_full + 8 * STAGES, a_v_empty = a_k_empty + 8 * STAGES, a_s_full = a_v_empty + 8 * STAGES,
                   a_p_full = a_s_full + 8, a_pv_done = a_s_full + 16, a_s_free = a_s_full + 24;
    const uint32_t q_sm = pin(smem_u32(q_smem)), k_sm = q_sm + 2 * C::kTileBytes, v_sm = k_sm + STAGES * C::kTileBytes;
    if (warp == C::kSoftmaxWarps) {
      // ============================== TMA producer ==============================
      if (elect_one()) prefetch_tmap(&tmap);
      int k_item = blockIdx.x, k_j = 0, k_n = 0, k_g = 0;          // item, tile in item, local item index, global tile
      int k_nkv = k_item < n_items ? item_nkv(k_item) : 0;
      auto advance_k = [&]() {
        if (k_item >= n_items) return;
        if (k_j == 0) {
          const int qb = k_n & 1;
          mbar_wait_a(a_q_empty + 8 * qb, ((k_n >> 1) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx_a(a_q_full + 8 * qb, C::kTileBytes);
            tma_load_4d_a(q_sm + qb * C::kTileBytes, &tmap, a_q_full + 8 * qb, 0, 0 * H + item_h(k_item), item_q0(k_item),
                          item_b(k_item));
          }
        }
        const int st = k_g % STAGES;
        mbar_wait_a(a_k_empty + 8 * st, ((k_g / STAGES) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx_a(a_k_full + 8 * st, C::kTileBytes);
          tma_load_4d_a(k_sm + st * C::kTileBytes, &tmap, a_k_full + 8 * st, 0, 1 * H + item_h(k_item), k_j * BKV,
                        item_b(k_item));
        }
        ++k_g;
        if (++k_j == k_nkv) {
          k_j = 0;
          ++k_n;
          k_item += gridDim.x;
          k_nkv = k_item < n_items ? item_nkv(k_item) : 0;
        }
      };
      for (int i = 0; i < STAGES; ++i) advance_k();
      int g = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int nkv = item_nkv(item), h = item_h(item), b = item_b(item);
        for (int j = 0; j < nkv; ++j, ++g) {
          const int st = g % STAGES;
          mbar_wait_a(a_v_empty + 8 * st, ((g / STAGES) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx_a(a_v_full + 8 * st, C::kTileBytes);
            tma_load_4d_a(v_sm + st * C::kTileBytes, &tmap, a_v_full + 8 * st, 0, 2 * H + h, j * BKV, b);
          }
          advance_k();
        }
      }
    } else if (warp == C::kSoftmaxWarps + 1) {
      // ============================== QK issuer: S(g) = Q K_g^T ==============================
      constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);
      const uint64_t q_d0 = make_smem_desc_sw128(q_sm, 16, 1024);
      const uint64_t k_d0 = make_smem_desc_sw128(k_sm, 16, 1024);
      const uint32_t s_tmem = tmem_base;
      const bool tracer = TRACE && blockIdx.x < 4 && lane == 0;
      long long* tr = TRACE ? p.trace + ((int64_t)blockIdx.x * 2 + 1) * 128 * 8 : nullptr;
      int g = 0, n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const int nkv = item_nkv(item);
        const int qb = n & 1;
        const uint64_t q_d = q_d0 + (uint64_t)(qb * (C::kTileBytes >> 4));
        for (int j = 0; j < nkv; ++j, ++g) {
          const int st = g % STAGES;
          if (tracer && g < 128) tr[g * 8 + 0] = clock64();
          if (g > 0) mbar_wait_a(a_s_free, (g - 1) & 1);          // S(g-1) is in the softmax warps' registers
          if (j == 0) mbar_wait_a(a_q_full + 8 * qb, (n >> 1) & 1);
          mbar_wait_a(a_k_full + 8 * st, (g / STAGES) & 1);
          tc_fence_after();
          if (tracer && g < 128) tr[g * 8 + 1] = clock64();
          if (elect_one()) {
            const uint64_t k_d = k_d0 + (uint64_t)(st * (C::kTileBytes >> 4));
            mma_ss_c<false>(s_tmem, q_d, k_d, idesc_qk);
            mma_ss_c<true>(s_tmem, q_d + 2, k_d + 2, idesc_qk);
            mma_ss_c<true>(s_tmem, q_d + 4, k_d + 4, idesc_qk);
            mma_ss_c<true>(s_tmem, q_d + 6, k_d + 6, idesc_qk);
            mma_commit_a(a_s_full);
            mma_commit_a(a_k_empty + 8 * st);
            // last QK of the item read the Q tile; with the TMA-store epilogue the Q buffer doubles as the O staging
            // tile and is released by the store warp instead
            if (j == nkv - 1 && !tma_out) mma_commit_a(a_q_empty + 8 * qb);
          }
          __syncwarp();
          if (tracer && g < 128) tr[g * 8 + 2] = clock64();
        }
      }
    } else if (warp == C::kSoftmaxWarps + 2) {
      // ============================== PV issuer: O += P_g V_g (P read from TMEM) ==============================
      constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, D, 0, 1);
      const uint64_t v_d0 = make_smem_desc_sw128(v_sm, BKV * 128, 1024);
      const uint32_t p_tmem = tmem_base + 128, o_tmem = tmem_base + 192;
      const bool tracer = TRACE && blockIdx.x < 4 && lane == 0;
      long long* tr = TRACE ? p.trace + ((int64_t)blockIdx.x * 2 + 1) * 128 * 8 : nullptr;
      int g = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int nkv = item_nkv(item);
        for (int j = 0; j < nkv; ++j, ++g) {
          const int st = g % STAGES;
          if (tracer && g < 128) tr[g * 8 + 4] = clock64();
          mbar_wait_a(a_p_full, g & 1);
          if (tracer && g < 128) tr[g * 8 + 5] = clock64();
          mbar_wait_a(a_v_full + 8 * st, (g / STAGES) & 1);
          tc_fence_after();
          if (tracer && g < 128) tr[g * 8 + 6] = clock64();
          if (elect_one()) {
            const uint64_t v_d = v_d0 + (uint64_t)(st * (C::kTileBytes >> 4));
            if (j > 0) mma_ts_c<true>(o_tmem, p_tmem, v_d, idesc_pv);
            else mma_ts_c<false>(o_tmem, p_tmem, v_d, idesc_pv);
#pragma unroll
            for (int k = 1; k < BKV / 16; ++k) mma_ts_c<true>(o_tmem, p_tmem + k * 8, v_d + (uint64_t)(k * 128), idesc_pv);
            mma_commit_a(a_pv_done);
            mma_commit_a(a_v_empty + 8 * st);
          }
          __syncwarp();
          if (tracer && g < 128) tr[g * 8 + 7] = clock64();
        }
      }
    } else if (tma_out) {
      // ============================== O store: staged tile (Q buffer of the finished item) -> global by TMA ==============
      const uint32_t a_o_ready = a_s_free + 8;
      int n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const int qb = n & 1;
        mbar_wait_a(a_o_ready, n & 1);
        if (elect_one()) {
          const int q0 = item_q0(item), h = item_h(item), b = item_b(item);
          if (p.out2 == nullptr || q0 < p.S_split) tma_store_4d_a(&tmap_o, q_sm + qb * C::kTileBytes, 0, h, q0, b);
          else tma_store_4d_a(&tmap_o2, q_sm + qb * C::kTileBytes, 0, h, q0 - p.S_split, b);
          tma_store_commit();
          tma_store_wait_read();                      // the staged tile has been read: the Q buffer may be reloaded
          mbar_arrive_a(a_q_empty + 8 * qb);
        }
        __syncwarp();
      }
    }
  } else {
    // ============================== softmax / epilogue (8 warps, quad layout) ==============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::kRegsSoftmax));
    const int lane0 = (warp & 3) * 32 + (warp >> 2) * 16;     // first TMEM lane (= tile row) of this warp's 16 rows
    const int row0 = lane0 + (lane >> 2);                      // this thread: rows row0 and row0 + 8
    const int cq = lane & 3;                                   // columns 8k + 2 cq + {0, 1}
    // pin(): an opaque move, so that ptxas keeps the value in a register instead of re-deriving it (shared-window
    // conversion, thread-id arithmetic: ~10 instructions each) at every use inside the tile loop
    auto pin = [](uint32_t v) { uint32_t r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; };
    const uint32_t s_tmem = pin(tmem_base + (static_cast<uint32_t>(lane0) << 16));
    const uint32_t p_tmem = s_tmem + 128;
    const uint32_t o_tmem = s_tmem + 192;
    const uint32_t bar0 = pin(smem_u32(bars));                 // 32-bit shared addresses of the barriers, computed once
    const uint32_t a_s_full = bar0 + 8 * (4 + 4 * STAGES), a_p_full = a_s_full + 8, a_pv_done = a_s_full + 16,
                   a_s_free = a_s_full + 24;
    const float sl2 = p.scale_log2;
    const float2 sc2 = make_float2(sl2, sl2);
    constexpr float kNegBig = -1.0e30f;                        // finite "-inf" for the running max: no special cases
    const bool tracer = TRACE && blockIdx.x < 4 && threadIdx.x == 0;
    long long* tr = TRACE ? p.trace + (int64_t)blockIdx.x * 2 * 128 * 8 : nullptr;
    const uint32_t a_o_ready = a_s_free + 8;
    const uint32_t q_sm = pin(smem_u32(q_smem));
    // Epilogue of a finished item: O / l -> bf16.  TMA mode: the tile is staged (128B-swizzled, conflict-free 4-byte
    // stores) in the item's Q buffer and written out by the store warp; otherwise token-major 4-byte global stores.
    auto finish_item = [&](int item, int n, const float2 (&l2)[2], const float (&m)[2]) {
      const int h = item_h(item), b = item_b(item);
      const int q_idx0 = item_q0(item) + row0;
      float inv_l[2], lse2[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float l = l2[r].x + l2[r].y;
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        inv_l[r] = l > 0.f ? __frcp_rn(l) : 0.f;
        lse2[r] = (m[r] + __log2f(l)) * 0.6931471805599453f;
      }
      const uint32_t stage = q_sm + (n & 1) * C::kTileBytes;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[16];                                        // o[4k + 2r + c] = O[row0 + 8r][32 hh + 8k + 2cq + c]
        tmem_ld_16x256b_x4(o_tmem + 32 * hh, o);
        tmem_wait_ld();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int row = row0 + 8 * r;
          if (tma_out) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              sts_u32(stage + row * 128 + (((4 * hh + k) ^ (row & 7)) << 4) + 4 * cq,
                      pack_bf16(__uint_as_float(o[4 * k + 2 * r]) * inv_l[r], __uint_as_float(o[4 * k + 2 * r + 1]) * inv_l[r]));
          } else {
            const int q_idx = q_idx0 + 8 * r;
            if (q_idx < S) {
              __nv_bfloat16* orow;
              if (p.out2 == nullptr) orow = p.out + (((int64_t)b * S + q_idx) * H + h) * D;
              else if (q_idx < p.S_split) orow = p.out + (((int64_t)b * p.S_split + q_idx) * H + h) * D;
              else orow = p.out2 + (((int64_t)b * (S - p.S_split) + (q_idx - p.S_split)) * H + h) * D;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                *reinterpret_cast<uint32_t*>(orow + 32 * hh + 8 * k + 2 * cq) =
                    pack_bf16(__uint_as_float(o[4 * k + 2 * r]) * inv_l[r], __uint_as_float(o[4 * k + 2 * r + 1]) * inv_l[r]);
            }
          }
        }
      }
      tc_fence_before();      // O is in registers: orders the TMEM reads before the next PV(0) (issued after p_full)
      if (tma_out) {
        fence_proxy_async_smem();                              // generic-proxy smem writes -> visible to the TMA store
        mbar_arrive_a(a_o_ready);
      }
      if (p.lse && cq == 0) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
          if (q_idx0 + 8 * r < S) p.lse[((int64_t)b * H + h) * S + q_idx0 + 8 * r] = lse2[r];
      }
    };
    int g = 0, n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const int nkv = item_nkv(item);
      const int q_idx0 = item_q0(item) + row0;                 // second row: q_idx0 + 8
      float m[2] = {kNegBig, kNegBig};                         // running max (scaled, log2 domain)
      float2 l2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};   // running denominators (this thread's columns)
      for (int j = 0; j < nkv; ++j, ++g) {
        if (tracer && g < 128) tr[g * 8 + 0] = clock64();
        mbar_wait_a(a_s_full, g & 1);
        tc_fence_after();
        if (tracer && g < 128) tr[g * 8 + 1] = clock64();
        uint32_t sa[32], sb[32];                               // s?[4k + 2r + c] = S[row0 + 8r][(64 if B) + 8k + 2cq + c]
        tmem_ld_16x256b_x8(s_tmem, sa);
        tmem_ld_16x256b_x8(s_tmem + 64, sb);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive_a(a_s_free);
        if (tracer && g < 128) tr[g * 8 + 2] = clock64();
        const int kv0 = j * BKV;
        int lim0 = S - kv0, lim1 = S - kv0;
        if (p.causal) {
          const int c0 = q_idx0 - kv0 + 1, c1 = c0 + 8;
          lim0 = c0 < lim0 ? c0 : lim0;
          lim1 = c1 < lim1 ? c1 : lim1;
        }
        if (lim0 < BKV || lim1 < BKV) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int col = 8 * k + 2 * cq + c;
              if (col >= lim0) sa[4 * k + c] = 0xff800000u;          // -inf
              if (col >= lim1) sa[4 * k + 2 + c] = 0xff800000u;
              if (col + 64 >= lim0) sb[4 * k + c] = 0xff800000u;
              if (col + 64 >= lim1) sb[4 * k + 2 + c] = 0xff800000u;
            }
        }
        float alpha[2];
        float2 nm2[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          float mx0 = kNegBig, mx1 = kNegBig;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sa[4 * k + 2 * r]), __uint_as_float(sa[4 * k + 2 * r + 1])));
            mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sb[4 * k + 2 * r]), __uint_as_float(sb[4 * k + 2 * r + 1])));
          }
          float mx = fmaxf(mx0, mx1);
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          float mn = fmaxf(m[r], mx * sl2);
          if (mn - m[r] <= kRescaleThreshold) mn = m[r];       // lazy rescale (never taken on the first tile: m = -1e30)
          alpha[r] = ex2(m[r] - mn);                           // 1 when the max is kept, 0 on the first tile
          nm2[r] = make_float2(-mn, -mn);
          m[r] = mn;
        }
        if (tracer && g < 128) tr[g * 8 + 3] = clock64();
        // ---- p = exp2(s * scale - m), bf16 pairs; does not depend on PV(g-1) ----
        uint32_t pk[32];                                       // pk[2k + r] = P[row0 + 8r][cols 8k + 2cq, +1], k < 16
        float2 sum2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          const uint32_t(&sv)[32] = hb ? sb : sa;
#pragma unroll
          for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[4 * k + 2 * r]), __uint_as_float(sv[4 * k + 2 * r + 1])),
                                          sc2, nm2[r]);
              float2 e;
              if ((2 * k + r) % 4 < EMU) {
                e = ex2_poly2(x);
              } else {
                e.x = ex2(x.x);
                e.y = ex2(x.y);
              }
              sum2[r] = __fadd2_rn(sum2[r], e);
              pk[16 * hb + 2 * k + r] = pack_bf16(e.x, e.y);
            }
        }
        if (tracer && g < 128) tr[g * 8 + 4] = clock64();
        // PV(g-1) reads P(g-1) from the columns P(g) overwrites; O may only be rescaled between PV(g-1) and PV(g).
        // For the first tile of an item the epilogue of the previous item already consumed that phase.
        if (j > 0) {
          mbar_wait_a(a_pv_done, (g - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, alpha[0] != 1.f || alpha[1] != 1.f)) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t o[16];                                  // o[4k + 2r + c] = O[row0 + 8r][32 hh + 8k + 2cq + c]
              tmem_ld_16x256b_x4(o_tmem + 32 * hh, o);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha[(i >> 1) & 1]);
              tmem_st_16x256b_x4(o_tmem + 32 * hh, o);
            }
          }
        }
        if (tracer && g < 128) tr[g * 8 + 5] = clock64();
        tmem_st_16x128b_x16(p_tmem, pk);
#pragma unroll
        for (int r = 0; r < 2; ++r) l2[r] = __ffma2_rn(l2[r], make_float2(alpha[r], alpha[r]), sum2[r]);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive_a(a_p_full);
        if (tracer && g < 128) tr[g * 8 + 6] = clock64();
      }
      // ---- epilogue (the QK warp is already computing S of the next item).  Deferring it into the first tile of the
      //      next item was measured slower: the extra live state costs more in the tile loop than the hidden wait ----
      mbar_wait_a(a_pv_done, (g - 1) & 1);
      tc_fence_after();
      finish_item(item, n, l2, m);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == C::kSoftmaxWarps + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Pair kernel (head_dim 64, non-causal; the MMDiT joint-attention shape and the default there): ONE persistent CTA per
// SM owns TWO 128-row query tiles of one (sample, head) and runs their softmaxes on two groups of eight quad-layout
// warps (G0: warps 0-7, G1: warps 8-15; S/P/O of tile t in TMEM columns [256 t, 256 t + 256)).
//
// Why (measured, profiles/r2_attn_pair_*): K / V tiles are shared by the two query tiles (half the shared-memory and L2
// traffic per FLOP and half the utility warps polling barriers per SM).  Two experiments on top of it were measured
// SLOWER and removed: an explicit "exp token" (named barriers) that forced the two groups into antiphase -- one in its
// MUFU-bound exp phase while the other loads S / publishes P -- 682 vs 757 TFLOP/s at B=16 H=24 S=1229; and a
// stale-max fast path (exponentials of tile j > 0 from the previous max, the row max only checked afterwards: no quad
// shuffles on the common path, but S must stay live for a possible redo -> spills) 680 TFLOP/s.  3-stage K and V rings,
// Q double-buffered across work items, O staged in the finished item's Q buffers and written by TMA.
struct PairCfg {
  static constexpr int kTileBytes = BQ * 64 * 2;                       // 16 KB
  static constexpr int kKStages = 3, kVStages = 3;
  static constexpr int kSmemTiles = (4 + kKStages + kVStages) * kTileBytes;   // 4 Q buffers (2 items x 2 tiles)
  static constexpr int kNumBars = 4 + 2 * kKStages + 2 * kVStages + 10;
  static constexpr int kSmemBytes = kSmemTiles + 1024 + 8 * kNumBars + 16;
  static constexpr int kTmemCols = 512;
  static constexpr int kGroupWarps = 8;
  static constexpr int kSoftmaxWarps = 16;
  static constexpr int kThreads = (kSoftmaxWarps + 4) * 32;            // 640
  // 640 threads x 96 registers at launch = 512 x kRegsSoftmax + 128 x kRegsUtility
  static constexpr int kRegsSoftmax = 112;
  static constexpr int kRegsUtility = 32;
};

template <int EMU>
__global__ void __launch_bounds__(PairCfg::kThreads, 1)
attn_fwd_pair_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_o,
                     const __grid_constant__ CUtensorMap tmap_o2, const Params p, const int n_items, const int npair,
                     const int tma_out, const int flags) {
  using C = PairCfg;
  constexpr int D = 64;
  constexpr int KS = C::kKStages, VS = C::kVStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kSmemTiles);
  // barrier indices (8 bytes each)
  constexpr int I_Q_FULL = 0, I_Q_EMPTY = 2, I_K_FULL = 4, I_K_EMPTY = I_K_FULL + KS, I_V_FULL = I_K_EMPTY + KS,
                I_V_EMPTY = I_V_FULL + VS, I_S_FULL = I_V_EMPTY + VS, I_P_FULL = I_S_FULL + 2, I_PV_DONE = I_P_FULL + 2,
                I_S_FREE = I_PV_DONE + 2, I_O_READY = I_S_FREE + 2;
  static_assert(I_O_READY + 2 == C::kNumBars, "barrier count");
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + C::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int S = p.S;
  const int H = p.H;

  auto item_q0 = [&](int it) { return (it % npair) * (2 * BQ); };
  auto item_h = [&](int it) { return (it / npair) % H; };
  auto item_b = [&](int it) { return it / (npair * H); };
  const int nkv = (S + BKV - 1) / BKV;                         // non-causal: every item walks all K/V tiles

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[I_Q_FULL + i], 1);
      mbar_init(&bars[I_Q_EMPTY + i], tma_out ? 1 : 2);        // store warp, or the last QK of both groups' issuers
      mbar_init(&bars[I_S_FULL + i], 1);
      mbar_init(&bars[I_P_FULL + i], C::kGroupWarps * 32);
      mbar_init(&bars[I_PV_DONE + i], 1);
      mbar_init(&bars[I_S_FREE + i], C::kGroupWarps * 32);
      mbar_init(&bars[I_O_READY + i], C::kGroupWarps * 32);
    }
    for (int i = 0; i < KS; ++i) {
      mbar_init(&bars[I_K_FULL + i], 1);
      mbar_init(&bars[I_K_EMPTY + i], 2);                      // QK of both groups read the stage
    }
    for (int i = 0; i < VS; ++i) {
      mbar_init(&bars[I_V_FULL + i], 1);
      mbar_init(&bars[I_V_EMPTY + i], 2);                      // PV of both groups read the stage
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
  pdl_trigger();
  pdl_wait();

  auto pin = [](uint32_t v) { uint32_t r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; };
  const uint32_t bar0 = pin(smem_u32(bars));
  const uint32_t q_sm = pin(smem_u32(smem));                   // 4 Q tiles: [item parity][tile]
  const uint32_t k_sm = q_sm + 4 * C::kTileBytes, v_sm = k_sm + KS * C::kTileBytes;
  auto B = [&](int idx) { return bar0 + 8u * (uint32_t)idx; };

  if (warp >= C::kSoftmaxWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::kRegsUtility));
    // warp-uniform loops, only the TMA / tcgen05 instructions under elect_one() (see attn_fwd_quad_kernel)
    if (warp == C::kSoftmaxWarps) {
      // ============================== TMA producer ==============================
      if (elect_one()) prefetch_tmap(&tmap);
      int k_item = blockIdx.x, k_j = 0, k_n = 0, k_g = 0;      // K cursor: item, tile in item, local item index, global tile
      auto advance_k = [&]() {
        if (k_item >= n_items) return;
        if (k_j == 0) {
          const int qb = k_n & 1;
          mbar_wait_a(B(I_Q_EMPTY + qb), ((k_n >> 1) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx_a(B(I_Q_FULL + qb), 2 * C::kTileBytes);
            // rows past S (a dead second tile when the tile count is odd) are zero-filled by the TMA unit
            for (int t = 0; t < 2; ++t)
              tma_load_4d_a(q_sm + (2 * qb + t) * C::kTileBytes, &tmap, B(I_Q_FULL + qb), 0, 0 * H + item_h(k_item),
                            item_q0(k_item) + t * BQ, item_b(k_item));
          }
        }
        const int st = k_g % KS;
        mbar_wait_a(B(I_K_EMPTY + st), ((k_g / KS) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx_a(B(I_K_FULL + st), C::kTileBytes);
          tma_load_4d_a(k_sm + st * C::kTileBytes, &tmap, B(I_K_FULL + st), 0, 1 * H + item_h(k_item), k_j * BKV,
                        item_b(k_item));
        }
        ++k_g;
        if (++k_j == nkv) {
          k_j = 0;
          ++k_n;
          k_item += gridDim.x;
        }
      };
      for (int i = 0; i < KS - 1; ++i) advance_k();
      int g = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int h = item_h(item), b = item_b(item);
        for (int j = 0; j < nkv; ++j, ++g) {
          advance_k();                                         // K runs KS - 1 tiles ahead of V
          const int st = g % VS;
          mbar_wait_a(B(I_V_EMPTY + st), ((g / VS) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx_a(B(I_V_FULL + st), C::kTileBytes);
            tma_load_4d_a(v_sm + st * C::kTileBytes, &tmap, B(I_V_FULL + st), 0, 2 * H + h, j * BKV, b);
          }
        }
      }
    } else if (warp <= C::kSoftmaxWarps + 2) {
      // ============================== MMA issuer of group t: S_t(g) = Q_t K_g^T and O_t += P_t(g) V_g ==================
      // One issuer warp per softmax group, so the two groups are coupled only through the K / V stages: with a single
      // QK warp and a single PV warp walking (tile 0, tile 1) in fixed order, the group that ran ahead waited for the
      // other's s_free / p_full (ncu: 5.8 % of the softmax warps' time at the s_full wait).  Issue order per group
      // follows its softmax: s_free(g) -> QK(g+1), then p_full(g) -> PV(g).
      const int t = warp - (C::kSoftmaxWarps + 1);
      constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, D, 0, 1);
      const uint64_t q_d0 = make_smem_desc_sw128(q_sm + t * C::kTileBytes, 16, 1024);
      const uint64_t k_d0 = make_smem_desc_sw128(k_sm, 16, 1024);
      const uint64_t v_d0 = make_smem_desc_sw128(v_sm, BKV * 128, 1024);
      const uint32_t s_tmem = tmem_base + 256 * t, p_tmem = s_tmem + 128, o_tmem = s_tmem + 192;
      const uint32_t a_s_full = B(I_S_FULL + t), a_s_free = B(I_S_FREE + t), a_p_full = B(I_P_FULL + t),
                     a_pv_done = B(I_PV_DONE + t);
      const int my_items = blockIdx.x < n_items ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      const int total = my_items * nkv;                        // K/V tiles this CTA walks
      // QK(gq): gq = global tile index, (nq, jq) = (local item, tile in item)
      auto issue_qk = [&](int gq, int nq, int jq) {
        const int qb = nq & 1, st = gq % KS;
        if (jq == 0) mbar_wait_a(B(I_Q_FULL + qb), (nq >> 1) & 1);
        mbar_wait_a(B(I_K_FULL + st), (gq / KS) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t q_d = q_d0 + (uint64_t)(2 * qb * (C::kTileBytes >> 4));
          const uint64_t k_d = k_d0 + (uint64_t)(st * (C::kTileBytes >> 4));
          mma_ss_c<false>(s_tmem, q_d, k_d, idesc_qk);
          mma_ss_c<true>(s_tmem, q_d + 2, k_d + 2, idesc_qk);
          mma_ss_c<true>(s_tmem, q_d + 4, k_d + 4, idesc_qk);
          mma_ss_c<true>(s_tmem, q_d + 6, k_d + 6, idesc_qk);
          mma_commit_a(a_s_full);
          mma_commit_a(B(I_K_EMPTY + st));
          // last QK of the item read the Q tile; with the TMA-store epilogue the Q buffers double as the O staging tiles
          // and are released by the store warp instead
          if (jq == nkv - 1 && !tma_out) mma_commit_a(B(I_Q_EMPTY + qb));
        }
        __syncwarp();
      };
      if (total > 0) issue_qk(0, 0, 0);
      int n = 0, j = 0;
      for (int g = 0; g < total; ++g) {
        int n1 = n, j1 = j + 1;
        if (j1 == nkv) { j1 = 0; ++n1; }
        if (g + 1 < total) {
          mbar_wait_a(a_s_free, g & 1);                        // S_t(g) is in the softmax warps' registers
          issue_qk(g + 1, n1, j1);
        }
        const int st = g % VS;
        mbar_wait_a(a_p_full, g & 1);
        mbar_wait_a(B(I_V_FULL + st), (g / VS) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t v_d = v_d0 + (uint64_t)(st * (C::kTileBytes >> 4));
          if (j > 0) mma_ts_c<true>(o_tmem, p_tmem, v_d, idesc_pv);
          else mma_ts_c<false>(o_tmem, p_tmem, v_d, idesc_pv);
#pragma unroll
          for (int k = 1; k < BKV / 16; ++k) mma_ts_c<true>(o_tmem, p_tmem + k * 8, v_d + (uint64_t)(k * 128), idesc_pv);
          mma_commit_a(a_pv_done);
          mma_commit_a(B(I_V_EMPTY + st));
        }
        __syncwarp();
        n = n1;
        j = j1;
      }
    } else if (tma_out) {
      // ============================== O store: staged tiles (Q buffers of the finished item) -> global by TMA ==========
      int n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const int qb = n & 1;
        const int h = item_h(item), b = item_b(item);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait_a(B(I_O_READY + t), n & 1);
          if (elect_one()) {
            const int q0 = item_q0(item) + t * BQ;             // a tile fully past S is clipped away by the TMA unit
            if (p.out2 == nullptr || q0 < p.S_split) tma_store_4d_a(&tmap_o, q_sm + (2 * qb + t) * C::kTileBytes, 0, h, q0, b);
            else tma_store_4d_a(&tmap_o2, q_sm + (2 * qb + t) * C::kTileBytes, 0, h, q0 - p.S_split, b);
            tma_store_commit();
          }
          __syncwarp();
        }
        if (elect_one()) {
          tma_store_wait_read();                               // both staged tiles have been read: reload the Q buffers
          mbar_arrive_a(B(I_Q_EMPTY + qb));
        }
        __syncwarp();
      }
    }
  } else {
    // ============================== softmax / epilogue: group t = warp / 8 owns query tile t ==============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::kRegsSoftmax));
    const int t = warp >> 3;
    const int wl = warp & 7;
    const int lane0 = (wl & 3) * 32 + (wl >> 2) * 16;         // first TMEM lane (= tile row) of this warp's 16 rows
    const int row0 = lane0 + (lane >> 2);                      // this thread: rows row0 and row0 + 8
    const int cq = lane & 3;                                   // columns 8k + 2 cq + {0, 1}
    const uint32_t s_tmem = pin(tmem_base + 256 * t + (static_cast<uint32_t>(lane0) << 16));
    const uint32_t p_tmem = s_tmem + 128;
    const uint32_t o_tmem = s_tmem + 192;
    const uint32_t a_s_full = B(I_S_FULL + t), a_p_full = B(I_P_FULL + t), a_pv_done = B(I_PV_DONE + t),
                   a_s_free = B(I_S_FREE + t), a_o_ready = B(I_O_READY + t);
    const float sl2 = p.scale_log2;
    const float2 sc2 = make_float2(sl2, sl2);
    constexpr float kNegBig = -1.0e30f;                        // finite "-inf" for the running max: no special cases
    auto finish_item = [&](int item, int n, const float2 (&l2)[2], const float (&m)[2]) {
      const int h = item_h(item), b = item_b(item);
      const int q_idx0 = item_q0(item) + t * BQ + row0;
      float inv_l[2], lse2[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float l = l2[r].x + l2[r].y;
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        inv_l[r] = l > 0.f ? __frcp_rn(l) : 0.f;
        lse2[r] = (m[r] + __log2f(l)) * 0.6931471805599453f;
      }
      const uint32_t stage = q_sm + (2 * (n & 1) + t) * C::kTileBytes;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[16];                                        // o[4k + 2r + c] = O[row0 + 8r][32 hh + 8k + 2cq + c]
        tmem_ld_16x256b_x4(o_tmem + 32 * hh, o);
        tmem_wait_ld();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int row = row0 + 8 * r;
          if (tma_out) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              sts_u32(stage + row * 128 + (((4 * hh + k) ^ (row & 7)) << 4) + 4 * cq,
                      pack_bf16(__uint_as_float(o[4 * k + 2 * r]) * inv_l[r], __uint_as_float(o[4 * k + 2 * r + 1]) * inv_l[r]));
          } else {
            const int q_idx = q_idx0 + 8 * r;
            if (q_idx < S) {
              __nv_bfloat16* orow;
              if (p.out2 == nullptr) orow = p.out + (((int64_t)b * S + q_idx) * H + h) * D;
              else if (q_idx < p.S_split) orow = p.out + (((int64_t)b * p.S_split + q_idx) * H + h) * D;
              else orow = p.out2 + (((int64_t)b * (S - p.S_split) + (q_idx - p.S_split)) * H + h) * D;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                *reinterpret_cast<uint32_t*>(orow + 32 * hh + 8 * k + 2 * cq) =
                    pack_bf16(__uint_as_float(o[4 * k + 2 * r]) * inv_l[r], __uint_as_float(o[4 * k + 2 * r + 1]) * inv_l[r]);
            }
          }
        }
      }
      tc_fence_before();      // O is in registers: orders the TMEM reads before the next PV(0) (issued after p_full)
      if (tma_out) {
        fence_proxy_async_smem();                              // generic-proxy smem writes -> visible to the TMA store
        mbar_arrive_a(a_o_ready);
      }
      if (p.lse && cq == 0) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
          if (q_idx0 + 8 * r < S) p.lse[((int64_t)b * H + h) * S + q_idx0 + 8 * r] = lse2[r];
      }
    };

    int g = 0, n = 0;
    // The epilogue of a finished item is DEFERRED into the first tile of the next one (between its exponentials and its
    // P store): waiting for the last PV right after publishing the last P exposed the whole p_full -> issue -> MMA ->
    // commit chain once per item (ncu: 5.7 % of the softmax warps' time).
    float m_prev[2] = {0.f, 0.f};
    float2 l2_prev[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    int item_prev = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      float m[2] = {kNegBig, kNegBig};                         // running max (scaled, log2 domain)
      float2 l2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};   // running denominators (this thread's columns)
      for (int j = 0; j < nkv; ++j, ++g) {
        mbar_wait_a(a_s_full, g & 1);
        tc_fence_after();
        uint32_t sa[32], sb[32];                               // s?[4k + 2r + c] = S[row0 + 8r][(64 if B) + 8k + 2cq + c]
        tmem_ld_16x256b_x8(s_tmem, sa);
        tmem_ld_16x256b_x8(s_tmem + 64, sb);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive_a(a_s_free);
        const int lim = S - j * BKV;                           // columns >= lim are past the sequence end
        if (lim < BKV) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int col = 8 * k + 2 * cq + c;
              if (col >= lim) { sa[4 * k + c] = 0xff800000u; sa[4 * k + 2 + c] = 0xff800000u; }          // -inf
              if (col + 64 >= lim) { sb[4 * k + c] = 0xff800000u; sb[4 * k + 2 + c] = 0xff800000u; }
            }
        }
        float alpha[2];
        float2 nm2[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          float mx0 = kNegBig, mx1 = kNegBig;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sa[4 * k + 2 * r]), __uint_as_float(sa[4 * k + 2 * r + 1])));
            mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sb[4 * k + 2 * r]), __uint_as_float(sb[4 * k + 2 * r + 1])));
          }
          float mx = fmaxf(mx0, mx1);
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          float mn = fmaxf(m[r], mx * sl2);
          if (mn - m[r] <= kRescaleThreshold) mn = m[r];       // lazy rescale (never taken on the first tile: m = -1e30)
          alpha[r] = ex2(m[r] - mn);                           // 1 when the max is kept, 0 on the first tile
          nm2[r] = make_float2(-mn, -mn);
          m[r] = mn;
        }
        // the PV(g-1) barrier is polled BEFORE the exponentials (a try_wait costs ~100 clocks even when the phase is
        // complete; here that latency hides under the exp stream) and only re-polled afterwards if it was not done yet
        const bool pv_ok = g > 0 ? mbar_try_wait_a(a_pv_done, (g - 1) & 1) : true;
        // ---- p = exp2(s * scale - m), bf16 pairs; does not depend on PV(g-1) ----
        uint32_t pk[32];                                       // pk[2k + r] = P[row0 + 8r][cols 8k + 2cq, +1], k < 16
        float2 sum2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          const uint32_t(&sv)[32] = hb ? sb : sa;
#pragma unroll
          for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[4 * k + 2 * r]), __uint_as_float(sv[4 * k + 2 * r + 1])),
                                          sc2, nm2[r]);
              float2 e;
              if ((2 * k + r) % 4 < EMU) {
                e = ex2_poly2(x);
              } else {
                e.x = ex2(x.x);
                e.y = ex2(x.y);
              }
              sum2[r] = __fadd2_rn(sum2[r], e);
              pk[16 * hb + 2 * k + r] = pack_bf16(e.x, e.y);
            }
        }
        // PV(g-1) reads P(g-1) from the columns P(g) overwrites; O may only be rescaled (or, for the first tile of an
        // item, read out by the previous item's epilogue) between PV(g-1) and PV(g).
        if (g > 0) {
          if (!pv_ok) mbar_wait_a(a_pv_done, (g - 1) & 1);
          tc_fence_after();
          if (j == 0) {
            finish_item(item_prev, n - 1, l2_prev, m_prev);
          } else if (__any_sync(0xffffffffu, alpha[0] != 1.f || alpha[1] != 1.f)) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t o[16];                                  // o[4k + 2r + c] = O[row0 + 8r][32 hh + 8k + 2cq + c]
              tmem_ld_16x256b_x4(o_tmem + 32 * hh, o);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha[(i >> 1) & 1]);
              tmem_st_16x256b_x4(o_tmem + 32 * hh, o);
            }
          }
        }
        tmem_st_16x128b_x16(p_tmem, pk);
#pragma unroll
        for (int r = 0; r < 2; ++r) l2[r] = __ffma2_rn(l2[r], make_float2(alpha[r], alpha[r]), sum2[r]);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive_a(a_p_full);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        m_prev[r] = m[r];
        l2_prev[r] = l2[r];
      }
      item_prev = item;
    }
    if (n > 0) {                                               // epilogue of this CTA's last item
      mbar_wait_a(a_pv_done, (g - 1) & 1);
      tc_fence_after();
      finish_item(item_prev, n - 1, l2_prev, m_prev);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == C::kSoftmaxWarps + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

long long* g_attn_trace = nullptr;   // debug: set by advgrpo_debug_set_attn_trace; consumed by dispatch variant 19

template <int EMU, bool TRACE = false>
int launch_quad(const void* qkv, void* out, void* out2, int64_t S_split, float* lse, int64_t B, int64_t S, int64_t H,
                float scale, int causal, cudaStream_t st) {
  using C = QCfg;
  CUtensorMap tmap;
  const uint64_t dims[4] = {64, (uint64_t)(3 * H), (uint64_t)S, (uint64_t)B};
  const uint64_t strides[4] = {0, 64 * 2, (uint64_t)(3 * H * 64) * 2, (uint64_t)(S * 3 * H * 64) * 2};
  const uint32_t box[4] = {64, 1, BQ, 1};
  int rc = make_tmap_bf16(&tmap, qkv, 4, dims, strides, box, true);
  if (rc != ADVGRPO_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    ADVGRPO_CUDA_CALL(cudaFuncSetAttribute(attn_fwd_quad_kernel<EMU, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
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
  p.park = 0;
  p.bias = nullptr;
  p.trace = TRACE ? g_attn_trace : nullptr;
  ADVGRPO_CHECK_ARG(!TRACE || p.trace, "attn_fwd: trace variant needs advgrpo_debug_set_attn_trace");
  const int nqt = (int)((S + BQ - 1) / BQ);
  const int64_t items = (int64_t)nqt * H * B;
  ADVGRPO_CHECK_ARG(items < (int64_t)1 << 30, "attn_fwd: too many work items");
  int grid = sm_count() * 2;
  if (grid > items) grid = (int)items;
  // O goes out by TMA (tile staged in the finished item's Q buffer) unless the image / text split cuts a 128-row tile
  static const int tma_env = getenv("ADVGRPO_ATTN_TMA_OUT") ? atoi(getenv("ADVGRPO_ATTN_TMA_OUT")) : 1;
  const int tma_out = (tma_env && (out2 == nullptr || S_split % BQ == 0)) ? 1 : 0;
  CUtensorMap tmap_o = tmap, tmap_o2 = tmap;
  if (tma_out) {
    const int64_t S1 = out2 ? S_split : S;
    const uint64_t od[4] = {64, (uint64_t)H, (uint64_t)S1, (uint64_t)B};
    const uint64_t os[4] = {0, 64 * 2, (uint64_t)(H * 64) * 2, (uint64_t)(S1 * H * 64) * 2};
    rc = make_tmap_bf16(&tmap_o, out, 4, od, os, box, true);
    if (rc != ADVGRPO_OK) return rc;
    if (out2) {
      const uint64_t od2[4] = {64, (uint64_t)H, (uint64_t)(S - S_split), (uint64_t)B};
      const uint64_t os2[4] = {0, 64 * 2, (uint64_t)(H * 64) * 2, (uint64_t)((S - S_split) * H * 64) * 2};
      rc = make_tmap_bf16(&tmap_o2, out2, 4, od2, os2, box, true);
      if (rc != ADVGRPO_OK) return rc;
    }
  }
  ADVGRPO_CUDA_CALL(launch_chain(attn_fwd_quad_kernel<EMU, TRACE>, dim3(grid), dim3(C::kThreads), C::kSmemBytes, st, 1, tmap,
                                 tmap_o, tmap_o2, p, (int)items, nqt, tma_out));
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

template <int EMU>
int launch_pair(const void* qkv, void* out, void* out2, int64_t S_split, float* lse, int64_t B, int64_t S, int64_t H,
                float scale, int flags, cudaStream_t st) {
  using C = PairCfg;
  CUtensorMap tmap;
  const uint64_t dims[4] = {64, (uint64_t)(3 * H), (uint64_t)S, (uint64_t)B};
  const uint64_t strides[4] = {0, 64 * 2, (uint64_t)(3 * H * 64) * 2, (uint64_t)(S * 3 * H * 64) * 2};
  const uint32_t box[4] = {64, 1, BQ, 1};
  int rc = make_tmap_bf16(&tmap, qkv, 4, dims, strides, box, true);
  if (rc != ADVGRPO_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    ADVGRPO_CUDA_CALL(cudaFuncSetAttribute(attn_fwd_pair_kernel<EMU>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
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
  p.causal = 0;
  p.park = 0;
  p.bias = nullptr;
  p.trace = nullptr;
  const int nqt = (int)((S + BQ - 1) / BQ);
  const int npair = (nqt + 1) / 2;
  const int64_t items = (int64_t)npair * H * B;
  ADVGRPO_CHECK_ARG(items < (int64_t)1 << 30, "attn_fwd: too many work items");
  int grid = sm_count();
  if (grid > items) grid = (int)items;
  static const int tma_env = getenv("ADVGRPO_ATTN_TMA_OUT") ? atoi(getenv("ADVGRPO_ATTN_TMA_OUT")) : 1;
  const int tma_out = (tma_env && (out2 == nullptr || S_split % BQ == 0)) ? 1 : 0;
  CUtensorMap tmap_o = tmap, tmap_o2 = tmap;
  if (tma_out) {
    const int64_t S1 = out2 ? S_split : S;
    const uint64_t od[4] = {64, (uint64_t)H, (uint64_t)S1, (uint64_t)B};
    const uint64_t os[4] = {0, 64 * 2, (uint64_t)(H * 64) * 2, (uint64_t)(S1 * H * 64) * 2};
    rc = make_tmap_bf16(&tmap_o, out, 4, od, os, box, true);
    if (rc != ADVGRPO_OK) return rc;
    if (out2) {
      const uint64_t od2[4] = {64, (uint64_t)H, (uint64_t)(S - S_split), (uint64_t)B};
      const uint64_t os2[4] = {0, 64 * 2, (uint64_t)(H * 64) * 2, (uint64_t)((S - S_split) * H * 64) * 2};
      rc = make_tmap_bf16(&tmap_o2, out2, 4, od2, os2, box, true);
      if (rc != ADVGRPO_OK) return rc;
    }
  }
  ADVGRPO_CUDA_CALL(launch_chain(attn_fwd_pair_kernel<EMU>, dim3(grid), dim3(C::kThreads), C::kSmemBytes, st, 1, tmap, tmap_o,
                                 tmap_o2, p, (int)items, npair, tma_out, flags));
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

template <int D, int EMU>
int launch_persist(const void* qkv, void* out, void* out2, int64_t S_split, float* lse, int64_t B, int64_t S, int64_t H,
                   float scale, int causal, cudaStream_t st, const float* bias = nullptr) {
  using C = PCfg<D>;
  CUtensorMap tmap;
  const uint64_t dims[4] = {(uint64_t)D, (uint64_t)(3 * H), (uint64_t)S, (uint64_t)B};
  const uint64_t strides[4] = {0, (uint64_t)D * 2, (uint64_t)(3 * H * D) * 2, (uint64_t)(S * 3 * H * D) * 2};
  const uint32_t box[4] = {64, 1, BQ, 1};
  int rc = make_tmap_bf16(&tmap, qkv, 4, dims, strides, box, true);
  if (rc != ADVGRPO_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    ADVGRPO_CUDA_CALL(cudaFuncSetAttribute(attn_fwd_persist_kernel<D, EMU>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
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
  p.park = 0;
  p.bias = bias;
  p.trace = nullptr;
  const int nqt = (int)((S + BQ - 1) / BQ);
  const int64_t items = (int64_t)nqt * H * B;
  ADVGRPO_CHECK_ARG(items < (int64_t)1 << 30, "attn_fwd: too many work items");
  const int per_sm = (D == 64) ? 2 : 1;
  int grid = sm_count() * per_sm;
  if (grid > items) grid = (int)items;
  ADVGRPO_CUDA_CALL(launch_chain(attn_fwd_persist_kernel<D, EMU>, dim3(grid), dim3(C::kThreads), C::kSmemBytes, st, 1, tmap, p,
                                 (int)items, nqt));
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

template <int D, int NQ, int EMU, int RS = 1>
int launch(const void* qkv, void* out, void* out2, int64_t S_split, float* lse, int64_t B, int64_t S, int64_t H, float scale,
           int causal, cudaStream_t st, int park = 0) {
  using C = Cfg<D, NQ, RS>;
  CUtensorMap tmap;
  const uint64_t dims[4] = {(uint64_t)D, (uint64_t)(3 * H), (uint64_t)S, (uint64_t)B};
  const uint64_t strides[4] = {0, (uint64_t)D * 2, (uint64_t)(3 * H * D) * 2, (uint64_t)(S * 3 * H * D) * 2};
  const uint32_t box[4] = {64, 1, BQ, 1};
  int rc = make_tmap_bf16(&tmap, qkv, 4, dims, strides, box, true);
  if (rc != ADVGRPO_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    ADVGRPO_CUDA_CALL(cudaFuncSetAttribute(attn_fwd_kernel<D, NQ, EMU, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
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
  p.park = park;
  p.bias = nullptr;
  p.trace = nullptr;
  dim3 grid((unsigned)((S + BQ * NQ - 1) / (BQ * NQ)), (unsigned)H, (unsigned)B);
  attn_fwd_kernel<D, NQ, EMU, RS><<<grid, C::kThreads, C::kSmemBytes, st>>>(tmap, p);
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

}  // namespace

// variant: 0 = auto; 1..6 = {1,2} query tiles per CTA x {0,1,2} of every 4 exp pairs emulated on the FMA pipe;
// 7..10 = the same with two threads per query row (column-split softmax, 16 softmax warps per SM)
int attn_fwd_dispatch(const void* qkv, void* out, void* out2, int64_t S_split, float* lse, int64_t B, int64_t S, int64_t H,
                      int64_t D, float scale, int causal, int variant, cudaStream_t st) {
#define ADVGRPO_ATTN_ARGS qkv, out, out2, S_split, lse, B, S, H, scale, causal, st
  if (D == 64) {
    switch (variant) {
      case 1: return launch<64, 1, 0>(ADVGRPO_ATTN_ARGS);
      case 2: return launch<64, 2, 0>(ADVGRPO_ATTN_ARGS);
      case 3: return launch<64, 1, 1>(ADVGRPO_ATTN_ARGS);
      case 4: return launch<64, 2, 1>(ADVGRPO_ATTN_ARGS);
      case 5: return launch<64, 1, 2>(ADVGRPO_ATTN_ARGS);
      case 6: return launch<64, 2, 2>(ADVGRPO_ATTN_ARGS);
      case 7: return launch<64, 1, 0, 2>(ADVGRPO_ATTN_ARGS);   // two threads per query row
      case 8: return launch<64, 1, 1, 2>(ADVGRPO_ATTN_ARGS);
      case 9: return launch<64, 2, 0, 2>(ADVGRPO_ATTN_ARGS);
      case 10: return launch<64, 2, 1, 2>(ADVGRPO_ATTN_ARGS);
      case 11: return launch<64, 1, 1>(ADVGRPO_ATTN_ARGS, 1);
      case 12: return launch<64, 1, 1>(ADVGRPO_ATTN_ARGS, 2);
      case 13: return launch<64, 1, 1>(ADVGRPO_ATTN_ARGS, 3);
      case 14: return launch_persist<64, 0>(ADVGRPO_ATTN_ARGS);
      case 15: return launch_persist<64, 2>(ADVGRPO_ATTN_ARGS);
      case 16: return launch_quad<0>(ADVGRPO_ATTN_ARGS);      // 8 softmax warps per query tile (quad layout)
      case 17: return launch_quad<1>(ADVGRPO_ATTN_ARGS);
      case 18: return launch_quad<2>(ADVGRPO_ATTN_ARGS);
      case 19: return launch_quad<1, true>(ADVGRPO_ATTN_ARGS);   // timeline trace (debug)
      case 20: return launch_persist<64, 1>(ADVGRPO_ATTN_ARGS);   // round-1 default (thread-per-row persistent kernel)
      // pair kernel (one CTA per SM, two query tiles in antiphase); non-causal only
      case 21: if (!causal) return launch_pair<1>(qkv, out, out2, S_split, lse, B, S, H, scale, 0, st); break;
      case 22: if (!causal) return launch_pair<2>(qkv, out, out2, S_split, lse, B, S, H, scale, 0, st); break;
      case 23: if (!causal) return launch_pair<0>(qkv, out, out2, S_split, lse, B, S, H, scale, 0, st); break;
      default:
        // fastest measured (profiles/r2_*): the pair kernel whenever there are two query tiles to pair up
        if (!causal && S > BQ) return launch_pair<1>(qkv, out, out2, S_split, lse, B, S, H, scale, 0, st);
        return launch_quad<1>(ADVGRPO_ATTN_ARGS);
    }
    return launch_quad<1>(ADVGRPO_ATTN_ARGS);   // pair variants asked for a causal problem
  }
  if (D == 128) return variant == 1 ? launch<128, 1, 0>(ADVGRPO_ATTN_ARGS) : launch_persist<128, 0>(ADVGRPO_ATTN_ARGS);
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

int advgrpo_attn_fwd_bias(const void* qkv, const float* bias, void* out, float* lse, int64_t B, int64_t S, int64_t H,
                          int64_t D, float scale, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(qkv && out && bias, "attn_fwd_bias: null pointer");
  ADVGRPO_CHECK_ARG(B >= 1 && S >= 1 && H >= 1, "attn_fwd_bias: bad sizes B=%lld S=%lld H=%lld", (long long)B, (long long)S,
                    (long long)H);
  ADVGRPO_CHECK_ARG(D == 64, "attn_fwd_bias: head_dim must be 64 (got %lld)", (long long)D);
  ADVGRPO_CHECK_ARG(aligned16(qkv) && aligned16(out), "attn_fwd_bias: tensors must be 16-byte aligned");
  return launch_persist<64, 1>(qkv, out, nullptr, 0, lse, B, S, H, scale, 0, (cudaStream_t)stream, bias);
}

// Debug: device buffer of 4 * 2 * 128 * 8 int64 clock stamps filled by dispatch variant 19.
void advgrpo_debug_set_attn_trace(void* buf) { g_attn_trace = (long long*)buf; }

// Test/bench hook (not part of the reference-facing surface): pick the CTA shape explicitly.
int advgrpo_attn_fwd_variant(const void* qkv, void* out, float* lse, int64_t B, int64_t S, int64_t H,
                             int64_t D, float scale, int causal, int variant,
                             advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(qkv && out, "attn_fwd: null pointer");
  return attn_fwd_dispatch(qkv, out, nullptr, 0, lse, B, S, H, D, scale, causal, variant, (cudaStream_t)stream);
}

}  // extern "C"

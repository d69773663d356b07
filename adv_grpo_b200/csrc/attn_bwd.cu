// Flash-attention backward on tcgen05 + TMA for sm_100a (head_dim 64).
//
// One CTA owns one 128-row K/V tile of one (batch, head) and streams the query tiles.  The
// scores are computed TRANSPOSED (S^T = K Q^T, lanes = kv rows) so that P^T and dS^T are directly
// the A operands of the dV / dK products:
//   S^T  = K_j Q_i^T            (SS)            dP^T = V_j dO_i^T            (SS)
//   P^T  = exp2(S^T c - lse_i)  (registers -> TMEM, bf16)
//   dS^T = P^T (dP^T - delta_i) (registers -> shared, bf16, 128B-swizzled by hand)
//   dV_j += P^T dO_i            (A = P^T from TMEM,   B = dO_i  MN-major)
//   dK_j += dS^T Q_i            (A = dS^T K-major,    B = Q_i   MN-major)
//   dQ_i  = dS K_j              (A = dS^T read MN-major, B = K_j MN-major) -> fp32 TMA reduce-add
// Warp roles (512 threads, the last two warps idle): warps 0-7 compute (thread = kv row; two warpgroups split the 128
// query columns), warps 8-11 drain dQ (TMEM -> swizzled smem -> cp.reduce.async.bulk.tensor add into
// the fp32 dQ accumulator), warp 12 TMA producer, warp 13 MMA issuer.
// TMEM columns: S^T 0-127 | dP^T 128-255 | P^T 256-319 | dV 320-383 | dK 384-447 | dQ 448-511.
// A prologue kernel computes delta = rowsum(dO . O), converts lse to the log2 domain into a
// 128-padded layout (+inf on padding so padded query columns give P = 0) and clears the dQ
// accumulator; an epilogue kernel scales and casts dQ to bf16 into dqkv.
//
// Replaces autograd through F.scaled_dot_product_attention at
// scripts/train_sd3_fast_pickscore.py:1165 (MMDiT joint attention and attn2 in the replay step).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

namespace advgrpo {
namespace {

using namespace sm100;

constexpr int T = 128;          // tile rows (both q and kv)
constexpr int D = 64;
constexpr int QSTAGES = 3;   // Q / dO / stats ring: a stage is only released by the back-half MMAs of its tile, so two
                             // stages left S^T / dP^T of tile it+1 waiting for a load issued one tile too late (ncu: 27 %)
constexpr int kTile = T * D * 2;            // 16 KB  bf16 [128 x 64]
constexpr int kDsBytes = T * T * 2;         // 32 KB  bf16 [128 x 128]
constexpr int kStatBytes = 2 * T * 4;       // lse2 + delta
constexpr int kDqBytes = T * 32 * 4;        // 16 KB  fp32 [128 x 32]: the dQ tile is drained as two half-width boxes
constexpr int OFF_K = 0;
constexpr int OFF_V = OFF_K + kTile;
constexpr int OFF_Q = OFF_V + kTile;                      // QSTAGES
constexpr int OFF_DO = OFF_Q + QSTAGES * kTile;           // QSTAGES
constexpr int OFF_DS = OFF_DO + QSTAGES * kTile;          // 2 (dK / dQ MMAs of tile it-1 still read one while tile it writes the other)
constexpr int OFF_DQ = OFF_DS + 2 * kDsBytes;             // 1
constexpr int OFF_STAT = OFF_DQ + kDqBytes;               // QSTAGES
constexpr int OFF_BAR = OFF_STAT + QSTAGES * kStatBytes;
constexpr int kSmemBytes = OFF_BAR + 256 + 1024;
// NWG compute warpgroups split the 128 query columns of a tile (thread = kv row): 2 x 64 columns (176 registers per compute
// thread, 8 compute warps) or 4 x 32 columns (96 registers, 16 compute warps: four per SM sub-partition hide the TMEM /
// LDS / MUFU latencies the 8-warp version exposed -- ncu r2: issue slots 31 % busy, compute warps 25 % long-scoreboard).
// Warps: [0, 4 NWG) compute, then 4 dQ-drain warps, then TMA, MMA and 2 idle warps (every setmaxnreg group is complete).
template <int NWG> struct BCfg {
  static constexpr int kComputeWarps = 4 * NWG;
  static constexpr int kThreads = (kComputeWarps + 8) * 32;
  static constexpr int kRegsCompute = NWG == 2 ? 176 : 96;   // 512 x 128 = 256 x 176 + 128 x 88 + 128 x 72
  static constexpr int kRegsDrain = NWG == 2 ? 88 : 48;      // 768 x 80  = 512 x 96 + 128 x 48 + 128 x 48
  static constexpr int kRegsUtility = NWG == 2 ? 72 : 48;
};
constexpr int COL_S = 0, COL_DP = 128, COL_P = 256, COL_DV = 320, COL_DK = 384, COL_DQ = 448;

struct BParams {
  __nv_bfloat16* dqkv;     // [B, S, 3, H, D]
  const float* lse2;       // [B, H, S_pad]  (log2 domain, +inf on padding)
  const float* delta;      // [B, H, S_pad]
  int S, S_pad, H;
  float scale, scale_log2;
  int causal;
};

template <int NWG>
__global__ void __launch_bounds__(BCfg<NWG>::kThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dq, const BParams p) {
  using C = BCfg<NWG>;
  constexpr int CW = C::kComputeWarps;                 // compute warps; drain = CW..CW+3, TMA = CW+4, MMA = CW+5
  constexpr int NCT = CW * 32;                         // compute threads
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* bar_kv_full = bars;                 // 1
  uint64_t* bar_q_full = bars + 1;              // QSTAGES
  uint64_t* bar_q_empty = bar_q_full + QSTAGES; // QSTAGES
  uint64_t* bar_s_full = bar_q_empty + QSTAGES; // 1   S^T and dP^T in TMEM
  uint64_t* bar_s_free = bar_s_full + 1;        // 1   compute finished reading them (256)
  uint64_t* bar_p_full = bar_s_free + 1;        // 1   P^T in TMEM (256)
  uint64_t* bar_pv_done = bar_p_full + 1;       // 1   dV MMA finished reading P^T
  uint64_t* bar_ds_full = bar_pv_done + 1;      // 2   dS^T in smem (256)
  uint64_t* bar_ds_empty = bar_ds_full + 2;     // 2   dK / dQ MMAs finished reading it
  uint64_t* bar_dq_full = bar_ds_empty + 2;     // 1
  uint64_t* bar_dq_free = bar_dq_full + 1;      // 1   drain finished reading dQ TMEM (128)
  uint64_t* bar_dkv_full = bar_dq_free + 1;     // 1
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bar_dkv_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * T;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int S = p.S;
  const int nq_total = (S + T - 1) / T;
  const int i_begin = (p.causal & 1) ? (kv0 / T) : 0;      // query tiles strictly above the diagonal see nothing
  const int nq = nq_total - i_begin;

  if (threadIdx.x == 0) {
    mbar_init(bar_kv_full, 1);
    for (int i = 0; i < QSTAGES; ++i) {
      mbar_init(&bar_q_full[i], 1);
      mbar_init(&bar_q_empty[i], 1);
    }
    mbar_init(bar_s_full, 1);
    mbar_init(bar_s_free, NCT);
    mbar_init(bar_p_full, NCT);
    mbar_init(bar_pv_done, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_ds_full[i], NCT);
      mbar_init(&bar_ds_empty[i], 1);
    }
    mbar_init(bar_dq_full, 1);
    mbar_init(bar_dq_free, 128);
    mbar_init(bar_dkv_full, 1);
    fence_barrier_init();
  }
  if (warp == CW + 5) {
    tmem_alloc(tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_trigger();   // programmatic dependent launch (common.cuh): no-ops unless the launch carries the attribute
  pdl_wait();
  const int64_t stat_row = ((int64_t)b * p.H + h) * p.S_pad;

  // register re-balancing (setmaxnreg is per 4-warp group): 256 compute threads x 176 + 128 drain x 88 + 128 utility x 72
  // = 512 x 128, the registers at launch
  if (warp >= CW + 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::kRegsUtility));
  if (warp == CW + 4) {
    // ============================== TMA producer ==============================
    // (utility warps run their loops warp-uniformly; only the TMA / tcgen05 instructions sit under elect_one(): under
    //  a divergent `lane == 0` branch ptxas wraps every UTMALDG / UTCHMMA / UTCBAR in an ELECT + BRA.U.ANY loop)
    if (elect_one()) {
      prefetch_tmap(&tm_qkv);
      prefetch_tmap(&tm_do);
      mbar_expect_tx(bar_kv_full, 2 * kTile);
      tma_load_4d(smem + OFF_K, &tm_qkv, bar_kv_full, 0, 1 * p.H + h, kv0, b);
      tma_load_4d(smem + OFF_V, &tm_qkv, bar_kv_full, 0, 2 * p.H + h, kv0, b);
    }
    __syncwarp();
    for (int it = 0; it < nq; ++it) {
      const int i = i_begin + it;
      const int st = it % QSTAGES;
      mbar_wait(&bar_q_empty[st], ((it / QSTAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&bar_q_full[st], 2 * kTile + kStatBytes);
        tma_load_4d(smem + OFF_Q + st * kTile, &tm_qkv, &bar_q_full[st], 0, 0 * p.H + h, i * T, b);
        tma_load_4d(smem + OFF_DO + st * kTile, &tm_do, &bar_q_full[st], 0, h, i * T, b);
        bulk_load_1d(smem + OFF_STAT + st * kStatBytes, p.lse2 + stat_row + i * T, T * 4, &bar_q_full[st]);
        bulk_load_1d(smem + OFF_STAT + st * kStatBytes + T * 4, p.delta + stat_row + i * T, T * 4, &bar_q_full[st]);
      }
      __syncwarp();
    }
  } else {
    // ============================== MMA issuers ==============================
    // Two issuer warps: CW+5 issues the FRONT half of a tile (S^T = K Q^T, dP^T = V dO^T), CW+6 the BACK half
    // (dV += P^T dO, dK += dS^T Q, dQ = dS K).  With a single issuer (round 1) S^T / dP^T of tile it+1 could only be issued
    // after the back half of tile it-1, whose ~250 instructions of descriptor rebuilding, modulo arithmetic and barrier polls
    // ran in ONE serial stream: ncu r2 showed that warp executing ~440 instructions per tile with no dominant stall, i.e. it
    // was the critical path (compute warps 12 % at the s_full wait, tensor pipe 32 %).  Descriptors are built once; per tile
    // only the 16-byte-unit start address moves (stage * 1024, dS buffer * 2048, K step offsets as immediates).
    constexpr uint32_t idesc_s = make_idesc_bf16(T, T, 0, 0);     // A K-major, B K-major, N = 128
    constexpr uint32_t idesc_kn = make_idesc_bf16(T, D, 0, 1);    // A K-major (smem or TMEM), B MN-major
    constexpr uint32_t idesc_mn = (make_idesc_bf16(T, D, 1, 1));  // A MN-major, B MN-major
    const uint32_t k_s = smem_u32(smem + OFF_K), v_s = smem_u32(smem + OFF_V);
    const uint32_t q_s0 = smem_u32(smem + OFF_Q), do_s0 = smem_u32(smem + OFF_DO), ds_s0 = smem_u32(smem + OFF_DS);
    if (warp == CW + 5) {
      const uint64_t k_kmaj = make_smem_desc_sw128(k_s, 16, 1024), v_kmaj = make_smem_desc_sw128(v_s, 16, 1024);
      const uint64_t q_k0 = make_smem_desc_sw128(q_s0, 16, 1024), do_k0 = make_smem_desc_sw128(do_s0, 16, 1024);
      mbar_wait(bar_kv_full, 0);
      int st = 0, ph = 0;                                         // Q / dO stage of tile `it` and its phase bit
      for (int it = 0; it < nq; ++it) {
        mbar_wait(&bar_q_full[st], ph);
        if (it > 0) mbar_wait(bar_s_free, (it - 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t q_k = q_k0 + (uint64_t)(st * (kTile >> 4)), do_k = do_k0 + (uint64_t)(st * (kTile >> 4));
          mma_ss_c<false>(tmem_base + COL_S, k_kmaj, q_k, idesc_s);
          mma_ss_c<true>(tmem_base + COL_S, k_kmaj + 2, q_k + 2, idesc_s);
          mma_ss_c<true>(tmem_base + COL_S, k_kmaj + 4, q_k + 4, idesc_s);
          mma_ss_c<true>(tmem_base + COL_S, k_kmaj + 6, q_k + 6, idesc_s);
          mma_ss_c<false>(tmem_base + COL_DP, v_kmaj, do_k, idesc_s);
          mma_ss_c<true>(tmem_base + COL_DP, v_kmaj + 2, do_k + 2, idesc_s);
          mma_ss_c<true>(tmem_base + COL_DP, v_kmaj + 4, do_k + 4, idesc_s);
          mma_ss_c<true>(tmem_base + COL_DP, v_kmaj + 6, do_k + 6, idesc_s);
          mma_commit(bar_s_full);
        }
        __syncwarp();
        if (++st == QSTAGES) { st = 0; ph ^= 1; }
      }
    } else if (warp == CW + 6) {
      const uint64_t k_mn = make_smem_desc_sw128(k_s, 16384, 1024);
      const uint64_t q_mn0 = make_smem_desc_sw128(q_s0, 16384, 1024), do_mn0 = make_smem_desc_sw128(do_s0, 16384, 1024);
      const uint64_t ds_k0 = make_smem_desc_sw128(ds_s0, 16, 1024), ds_mn0 = make_smem_desc_sw128(ds_s0, 16384, 1024);
      mbar_wait(bar_kv_full, 0);                                  // K (dQ's B operand) landed: observed by this thread too
      int st = 0, ph = 0;
      for (int it = 0; it < nq; ++it) {
        const int bb = it & 1;
        mbar_wait(&bar_q_full[st], ph);                           // Q / dO of this tile (long complete; acquire for this thread)
        const uint64_t q_mn = q_mn0 + (uint64_t)(st * (kTile >> 4)), do_mn = do_mn0 + (uint64_t)(st * (kTile >> 4));
        const uint64_t ds_k = ds_k0 + (uint64_t)(bb * (kDsBytes >> 4)), ds_mn = ds_mn0 + (uint64_t)(bb * (kDsBytes >> 4));
        // dV += P^T dO
        mbar_wait(bar_p_full, it & 1);
        tc_fence_after();
        if (elect_one()) {
          if (it > 0) mma_ts_c<true>(tmem_base + COL_DV, tmem_base + COL_P, do_mn, idesc_kn);
          else mma_ts_c<false>(tmem_base + COL_DV, tmem_base + COL_P, do_mn, idesc_kn);
#pragma unroll
          for (int k = 1; k < T / 16; ++k)
            mma_ts_c<true>(tmem_base + COL_DV, tmem_base + COL_P + k * 8, do_mn + (uint64_t)(k * 128), idesc_kn);
          mma_commit(bar_pv_done);
        }
        __syncwarp();
        // dK += dS^T Q
        mbar_wait(&bar_ds_full[bb], (it >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          if (it > 0) mma_ss_c<true>(tmem_base + COL_DK, ds_k, q_mn, idesc_kn);
          else mma_ss_c<false>(tmem_base + COL_DK, ds_k, q_mn, idesc_kn);
#pragma unroll
          for (int k = 1; k < T / 16; ++k)
            mma_ss_c<true>(tmem_base + COL_DK, ds_k + (uint64_t)((k / 4) * 1024 + (k % 4) * 2), q_mn + (uint64_t)(k * 128), idesc_kn);
        }
        __syncwarp();
        // dQ = dS K   (fresh accumulator every query tile)
        if (it > 0) {
          mbar_wait(bar_dq_free, (it - 1) & 1);
          tc_fence_after();
        }
        if (elect_one()) {
          mma_ss_c<false>(tmem_base + COL_DQ, ds_mn, k_mn, idesc_mn);
#pragma unroll
          for (int k = 1; k < T / 16; ++k)
            mma_ss_c<true>(tmem_base + COL_DQ, ds_mn + (uint64_t)(k * 128), k_mn + (uint64_t)(k * 128), idesc_mn);
          mma_commit(&bar_ds_empty[bb]);
          mma_commit(bar_dq_full);
          mma_commit(&bar_q_empty[st]);
        }
        __syncwarp();
        if (++st == QSTAGES) { st = 0; ph ^= 1; }
      }
      if (elect_one()) mma_commit(bar_dkv_full);
      __syncwarp();
    }
  }
  } else if (warp < CW) {
    // ============================== compute: P^T and dS^T ==============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::kRegsCompute));
    constexpr int NCH = 4 / NWG;                      // 32-column chunks of the query axis per thread
    const int wg = warp / 4;                          // which slice of the query columns
    const int row = (warp % 4) * 32 + lane;           // kv row inside the tile == TMEM lane
    const int kv_idx = kv0 + row;
    const bool kv_ok = kv_idx < S;
    const bool kv_ok_tile = kv0 + T <= S;             // every kv row of this CTA's tile is a real token
    const uint32_t lane_addr = static_cast<uint32_t>((warp % 4) * 32) << 16;
    const uint32_t t_s = tmem_base + COL_S + lane_addr;
    const uint32_t t_dp = tmem_base + COL_DP + lane_addr;
    const uint32_t t_p = tmem_base + COL_P + lane_addr;
    for (int it = 0; it < nq; ++it) {
      const int i = i_begin + it;
      const int st = it % QSTAGES, bb = it & 1;
      const float* lse2 = reinterpret_cast<const float*>(smem + OFF_STAT + st * kStatBytes);
      const uint32_t ds_base = smem_u32(smem + OFF_DS + bb * kDsBytes + row * 128);   // + 16384 per 64-column block
      const uint32_t stat_s = smem_u32(lse2);          // lse2[128] then delta[128] (explicit shared-space loads)
      mbar_wait(&bar_q_full[st], (it / QSTAGES) & 1);     // lse2 / delta visible
      mbar_wait(bar_s_full, it & 1);
      tc_fence_after();
      // my query columns of S^T and dP^T -> registers in one go, then release both TMEM regions at once: the MMA
      // warp issues S^T / dP^T of the NEXT query tile while this tile's exponentials are still running (they were
      // serialised behind the whole tile before: ncu showed 12 % of the compute warps' time waiting on s_full)
      uint32_t rs[NCH][32], rp[NCH][32];
#pragma unroll
      for (int cc = 0; cc < NCH; ++cc) {
        tmem_ld32(t_s + (wg * NCH + cc) * 32, rs[cc]);
        tmem_ld32(t_dp + (wg * NCH + cc) * 32, rp[cc]);
      }
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(bar_s_free);
      // P^T and dS^T of all chunks into registers first: the exponentials do not depend on the dV MMA of the previous
      // tile, which is still reading P^T(it-1) from the TMEM columns the stores below overwrite
      uint32_t pk[NCH][16], dk[NCH][16];
      // element-wise masking (causal diagonal tile, kv rows beyond S) only where a tile needs it: the common tile
      // runs 5 instructions per element (FFMA, MUFU.EX2, FADD, FMUL, half a pack pair) instead of ~19
      const bool masked = !kv_ok_tile || ((p.causal & 1) && i * T < kv0 + T);
      if (masked) {
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) {
          const int c = wg * NCH + cc;                  // 32-column chunk of the query axis
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 l4 = lds_f4(stat_s + (c * 32 + e) * 4);
            const float4 d4 = lds_f4(stat_s + T * 4 + (c * 32 + e) * 4);
            const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
              const int ee = e + u;
              const int q_a = i * T + c * 32 + ee;
              float p0 = ex2(fmaf(__uint_as_float(rs[cc][ee]), p.scale_log2, -ls[u]));
              float p1 = ex2(fmaf(__uint_as_float(rs[cc][ee + 1]), p.scale_log2, -ls[u + 1]));
              if (!kv_ok) { p0 = 0.f; p1 = 0.f; }
              if (p.causal & 1) {
                if (q_a < kv_idx) p0 = 0.f;
                if (q_a + 1 < kv_idx) p1 = 0.f;
              }
              const float d0 = p0 * (__uint_as_float(rp[cc][ee]) - dl[u]);
              const float d1 = p1 * (__uint_as_float(rp[cc][ee + 1]) - dl[u + 1]);
              pk[cc][ee / 2] = pack_bf16(p0, p1);
              dk[cc][ee / 2] = pack_bf16(d0, d1);
            }
          }
        }
      } else {
#pragma unroll
        for (int cc = 0; cc < NCH; ++cc) {
          const int c = wg * NCH + cc;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 l4 = lds_f4(stat_s + (c * 32 + e) * 4);
            const float4 d4 = lds_f4(stat_s + T * 4 + (c * 32 + e) * 4);
            const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int u = 0; u < 4; u += 2) {
              const int ee = e + u;
              const float p0 = ex2(fmaf(__uint_as_float(rs[cc][ee]), p.scale_log2, -ls[u]));
              const float p1 = ex2(fmaf(__uint_as_float(rs[cc][ee + 1]), p.scale_log2, -ls[u + 1]));
              const float d0 = p0 * (__uint_as_float(rp[cc][ee]) - dl[u]);
              const float d1 = p1 * (__uint_as_float(rp[cc][ee + 1]) - dl[u + 1]);
              pk[cc][ee / 2] = pack_bf16(p0, p1);
              dk[cc][ee / 2] = pack_bf16(d0, d1);
            }
          }
        }
      }
      if (it > 0) mbar_wait(bar_pv_done, (it - 1) & 1);   // P^T region free
      if (it >= 2) mbar_wait(&bar_ds_empty[bb], ((it - 2) >> 1) & 1);   // dS buffer free (dK / dQ MMAs of tile it-2 done)
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < NCH; ++cc) {
        const int c = wg * NCH + cc;
        tmem_st16(t_p + c * 16, pk[cc]);
        // dS^T row: 64 bytes of this chunk = four 16-byte pieces inside 64-column block c / 2, hand swizzled (128B pattern)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int piece = (c & 1) * 4 + k;
          sts_u4(ds_base + (c >> 1) * 16384 + ((piece ^ (row & 7)) * 16),
                 make_uint4(dk[cc][4 * k], dk[cc][4 * k + 1], dk[cc][4 * k + 2], dk[cc][4 * k + 3]));
        }
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(bar_p_full);
      fence_proxy_async_smem();
      mbar_arrive(&bar_ds_full[bb]);
    }
    // ---- epilogue: dV (first half of the warpgroups) and dK (second half) -> bf16, 64 / (NWG / 2) columns per thread ----
    {
      mbar_wait(bar_dkv_full, 0);
      tc_fence_after();
      constexpr int ECH = 4 / NWG;                      // 32-column chunks of the 64-wide dV / dK row per thread
      __nv_bfloat16* base = p.dqkv + ((int64_t)b * S + kv_idx) * 3 * p.H * D;
      const int which = wg / (NWG / 2);                 // 0: dV, 1: dK
      const int col0 = (wg % (NWG / 2)) * (ECH * 32);
      const uint32_t t_acc = tmem_base + (which == 0 ? COL_DV : COL_DK) + lane_addr + col0;
      const float sc = which == 0 ? 1.0f : p.scale;
      __nv_bfloat16* dst = base + ((which == 0 ? 2 : 1) * p.H + h) * D + col0;
      uint32_t r[ECH][32];
#pragma unroll
      for (int c = 0; c < ECH; ++c) tmem_ld32(t_acc + c * 32, r[c]);
      tmem_wait_ld();
      if (kv_ok) {
#pragma unroll
        for (int c = 0; c < ECH; ++c)
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(r[c][e + 0]) * sc, __uint_as_float(r[c][e + 1]) * sc);
            v.y = pack_bf16(__uint_as_float(r[c][e + 2]) * sc, __uint_as_float(r[c][e + 3]) * sc);
            v.z = pack_bf16(__uint_as_float(r[c][e + 4]) * sc, __uint_as_float(r[c][e + 5]) * sc);
            v.w = pack_bf16(__uint_as_float(r[c][e + 6]) * sc, __uint_as_float(r[c][e + 7]) * sc);
            *reinterpret_cast<uint4*>(dst + c * 32 + e) = v;
          }
      }
    }
  } else {
    // ============================== dQ drain (4 warps after the compute warps) ==============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::kRegsDrain));
    const int row = (warp % 4) * 32 + lane;           // query row inside the tile == TMEM lane
    const uint32_t t_dq = tmem_base + COL_DQ + (static_cast<uint32_t>((warp % 4) * 32) << 16);
    uint8_t* stage = smem + OFF_DQ;
    const int tid = threadIdx.x - NCT;
    for (int it = 0; it < nq; ++it) {
      const int i = i_begin + it;
      mbar_wait(bar_dq_full, it & 1);
      tc_fence_after();
      // two [128 rows x 32 fp32] boxes through ONE 16 KB staging tile (128B-swizzled rows), one after the other; the
      // TMEM columns are released as soon as the second half is in registers
      const int grow = (int)(stat_row + (int64_t)i * T);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t r[32];
        tmem_ld32(t_dq + hf * 32, r);
        tmem_wait_ld();
        if (hf == 1) {
          tc_fence_before();
          mbar_arrive(bar_dq_free);
        }
        if (tid == 0) tma_store_wait_read();          // the previous reduce finished reading the staging tile
        named_bar_sync(1, 128);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          sts_u4(smem_u32(stage) + row * 128 + ((k ^ (row & 7)) * 16), make_uint4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]));
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (tid == 0 && !(p.causal & 2)) {   // bit 1: timing experiment only (skip the dQ reduce)
          tma_reduce_add_2d(&tm_dq, stage, hf * 32, grow);
          tma_store_commit();
        }
      }
    }
    if (tid == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == CW + 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// delta = rowsum(dO . O); lse -> log2 domain, padded; clear the dQ accumulator.
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                     const float* __restrict__ lse, float* __restrict__ lse2,
                                     float* __restrict__ delta, float* __restrict__ dq_acc, int B, int S,
                                     int S_pad, int H) {
  pdl_trigger();   // programmatic dependent launch (common.cuh): no-ops unless the launch carries the attribute
  pdl_wait();
  // one 8-lane group per (b, h, s_pad) row
  const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 8;
  const int sub = threadIdx.x & 7;
  const int64_t total = (int64_t)B * H * S_pad;
  if (gid >= total) return;
  const int s = (int)(gid % S_pad);
  const int h = (int)((gid / S_pad) % H);
  const int b = (int)(gid / ((int64_t)S_pad * H));
  float acc = 0.f;
  if (s < S) {
    const int64_t off = (((int64_t)b * S + s) * H + h) * D + sub * 8;
    float fo[8], fd[8];
    unpack8(*reinterpret_cast<const bf16x8*>(out + off), fo);
    unpack8(*reinterpret_cast<const bf16x8*>(dout + off), fd);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(fo[j], fd[j], acc);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (sub == 0) {
    delta[gid] = acc;
    lse2[gid] = s < S ? lse[((int64_t)b * H + h) * S + s] * 1.4426950408889634f : INFINITY;
  }
  float4* dq = reinterpret_cast<float4*>(dq_acc + gid * D + sub * 8);
  dq[0] = make_float4(0.f, 0.f, 0.f, 0.f);
  dq[1] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void attn_bwd_dq_finish_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv,
                                          int B, int S, int S_pad, int H, float scale) {
  pdl_trigger();   // programmatic dependent launch (common.cuh): no-ops unless the launch carries the attribute
  pdl_wait();
  const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 8;
  const int sub = threadIdx.x & 7;
  const int64_t total = (int64_t)B * H * S;
  if (gid >= total) return;
  const int s = (int)(gid % S);
  const int h = (int)((gid / S) % H);
  const int b = (int)(gid / ((int64_t)S * H));
  const float4* src = reinterpret_cast<const float4*>(dq_acc + ((((int64_t)b * H + h) * S_pad + s) * D + sub * 8));
  float4 a = src[0], c = src[1];
  float f[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, c.x * scale, c.y * scale, c.z * scale, c.w * scale};
  *reinterpret_cast<bf16x8*>(dqkv + ((((int64_t)b * S + s) * 3 + 0) * H + h) * D + sub * 8) = pack8(f);
}

inline int64_t pad128(int64_t s) { return (s + 127) / 128 * 128; }

}  // namespace
}  // namespace advgrpo

using namespace advgrpo;

extern "C" {

size_t advgrpo_attn_bwd_workspace_bytes(int64_t B, int64_t S, int64_t H, int64_t D) {
  const int64_t sp = pad128(S);
  return (size_t)(B * H * sp * D * 4 + 2 * B * H * sp * 4 + 1024);
}

int advgrpo_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                     void* dqkv, int64_t B, int64_t S, int64_t H, int64_t D, float scale,
                     int causal, void* workspace, size_t workspace_bytes,
                     advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(qkv && out && dout && lse && dqkv, "attn_bwd: null pointer");
  if (D != 64) return set_error(ADVGRPO_ERR_UNSUPPORTED, "attn_bwd: head_dim %lld not supported (64 only)", (long long)D);
  ADVGRPO_CHECK_ARG(B >= 1 && S >= 1 && H >= 1 && B <= 65535 && H <= 65535, "attn_bwd: bad sizes");
  ADVGRPO_CHECK_ARG(aligned16(qkv) && aligned16(out) && aligned16(dout) && aligned16(dqkv), "attn_bwd: 16-byte alignment");
  if (!workspace || workspace_bytes < advgrpo_attn_bwd_workspace_bytes(B, S, H, D))
    return set_error(ADVGRPO_ERR_WORKSPACE, "attn_bwd: workspace too small");
  ADVGRPO_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "attn_bwd: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t sp = pad128(S);
  float* dq_acc = (float*)workspace;
  float* lse2 = dq_acc + B * H * sp * D;
  float* delta = lse2 + B * H * sp;
  {
    const int64_t groups = B * H * sp;
    const int threads = 256;
    const int64_t blocks = (groups * 8 + threads - 1) / threads;
    ADVGRPO_CUDA_CALL(launch_chain(attn_bwd_prep_kernel, dim3((unsigned)blocks), dim3(threads), 0, st, 1,
                                   (const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, (const float*)lse, lse2, delta,
                                   dq_acc, (int)B, (int)S, (int)sp, (int)H));
    ADVGRPO_CUDA_LAUNCH_CHECK();
  }
  CUtensorMap tm_qkv, tm_do, tm_dq;
  {
    const uint64_t dims[4] = {(uint64_t)D, (uint64_t)(3 * H), (uint64_t)S, (uint64_t)B};
    const uint64_t str[4] = {0, (uint64_t)D * 2, (uint64_t)(3 * H * D) * 2, (uint64_t)(S * 3 * H * D) * 2};
    const uint32_t box[4] = {64, 1, 128, 1};
    int rc = make_tmap_bf16(&tm_qkv, qkv, 4, dims, str, box, true);
    if (rc) return rc;
    const uint64_t dims2[4] = {(uint64_t)D, (uint64_t)H, (uint64_t)S, (uint64_t)B};
    const uint64_t str2[4] = {0, (uint64_t)D * 2, (uint64_t)(H * D) * 2, (uint64_t)(S * H * D) * 2};
    rc = make_tmap_bf16(&tm_do, dout, 4, dims2, str2, box, true);
    if (rc) return rc;
    const uint64_t dims3[2] = {(uint64_t)D, (uint64_t)(B * H * sp)};
    const uint64_t str3[2] = {0, (uint64_t)D * 4};
    const uint32_t box3[2] = {32, 128};
    rc = make_tmap(&tm_dq, dq_acc, 2, dims3, str3, box3, true, true);
    if (rc) return rc;
  }
  // 16 compute warps (4 x 32 query columns) by default; ADVGRPO_ATTN_BWD_WG=2 selects the 8-warp layout (measurement)
  static const int nwg = (getenv("ADVGRPO_ATTN_BWD_WG") && atoi(getenv("ADVGRPO_ATTN_BWD_WG")) == 2) ? 2 : 4;
  static bool attr_set = false;
  if (!attr_set) {
    ADVGRPO_CUDA_CALL(cudaFuncSetAttribute(attn_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    ADVGRPO_CUDA_CALL(cudaFuncSetAttribute(attn_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  BParams p;
  p.dqkv = (__nv_bfloat16*)dqkv;
  p.lse2 = lse2;
  p.delta = delta;
  p.S = (int)S; p.S_pad = (int)sp; p.H = (int)H;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  dim3 grid((unsigned)(sp / T), (unsigned)H, (unsigned)B);
  if (nwg == 2) {
    ADVGRPO_CUDA_CALL(launch_chain(attn_bwd_kernel<2>, grid, dim3(BCfg<2>::kThreads), kSmemBytes, st, 1, tm_qkv, tm_do, tm_dq, p));
  } else {
    ADVGRPO_CUDA_CALL(launch_chain(attn_bwd_kernel<4>, grid, dim3(BCfg<4>::kThreads), kSmemBytes, st, 1, tm_qkv, tm_do, tm_dq, p));
  }
  ADVGRPO_CUDA_LAUNCH_CHECK();
  {
    const int64_t groups = B * H * S;
    const int threads = 256;
    const int64_t blocks = (groups * 8 + threads - 1) / threads;
    ADVGRPO_CUDA_CALL(launch_chain(attn_bwd_dq_finish_kernel, dim3((unsigned)blocks), dim3(threads), 0, st, 1,
                                   (const float*)dq_acc, (__nv_bfloat16*)dqkv, (int)B, (int)S, (int)sp, (int)H, scale));
    ADVGRPO_CUDA_LAUNCH_CHECK();
  }
  return ADVGRPO_OK;
}

}  // extern "C"

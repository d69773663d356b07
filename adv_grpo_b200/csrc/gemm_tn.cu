// Skinny "TN" GEMM on tcgen05 + TMA for sm_100a: the LoRA weight-gradient products of the replay step.
//
//   C[m, n] = sum_k A[k, m] * B[k, n]      A: [Kt, Ms] bf16 row-major, Ms <= 256 (the LoRA side, rank-padded)
//                                          B: [Kt, Nb] bf16 row-major (an activation / output-gradient matrix)
//   dA_lora  = dt^T x   : A = dt [tokens, r_pad], B = x  [tokens, K]   -> [r_pad, K]
//   d(sB)    = dy^T t   : A = t  [tokens, r_pad], B = dy [tokens, N]   -> [r_pad, N], written TRANSPOSED as [N, r_pad]
//
// The contraction runs over the TOKEN axis (16 384 rows per CFG batch), the output is tiny (<= 256 x 6144), so the op
// is an HBM streaming reduction of B (50-200 MB), not a FLOP problem.  Both operands are consumed exactly as they lie
// in memory: a TMA box of {64 columns, 64 tokens} lands as an MN-major UMMA operand tile (128-byte swizzle), so there
// is no transpose pass (torch's `dt.t() @ x` -> cuBLAS TN did the same through a library kernel + split-K reduce).
// Split-K over token chunks fills the machine (grid = column tiles x splits x row tiles, >= 2 CTAs per SM when the
// problem allows); each CTA writes its fp32 partial tile to a workspace slab and a finish kernel sums the slabs in a
// FIXED order, casts to bf16 and (optionally) transposes: deterministic, no atomics.
//
// Warps: 0 = TMA producer, 1 = MMA issuer (elect_one), 2-5 = epilogue (TMEM lane quarters 2, 3, 0, 1).
// Replaces `da = dt.t() @ x2`, `dw2 = dy2.t() @ t` of the LoRA autograd nodes (scripts/train_sd3_fast_pickscore.py:1165,
// peft LoRA Linear backward) that round 1 left on cuBLAS (nvjet_* in the launch list).
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

namespace advgrpo {
namespace {

using namespace sm100;

constexpr int BM = 128;                 // rows of C per CTA (columns of A)
constexpr int BN = 128;                 // columns of C per CTA (columns of B)
constexpr int BK = 64;                  // tokens per pipeline stage
constexpr int STG = 4;
constexpr int kAtom = BK * 128;         // bytes of one 64-column x BK-token atom
constexpr int kStageBytes = 4 * kAtom;  // A: 2 atoms, B: 2 atoms
constexpr int kSmem = STG * kStageBytes + 1024 + 128;
constexpr int kThreads = 192;

__global__ void __launch_bounds__(kThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, float* __restrict__ part,
               int Kt, int Ms, int Nb, int chunk) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STG * kStageBytes);
  uint64_t* bar_full = bars;             // STG
  uint64_t* bar_empty = bars + STG;      // STG
  uint64_t* bar_done = bars + 2 * STG;   // 1
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 2 * STG + 1);
  const int warp = threadIdx.x >> 5;
  const int n0 = blockIdx.x * BN, split = blockIdx.y, m0 = blockIdx.z * BM;
  const int k_begin = split * chunk;
  const int k_end = min(Kt, k_begin + chunk);
  const int nk = (k_end - k_begin + BK - 1) / BK;        // >= 1 by construction of the grid

  if (threadIdx.x == 0) {
    for (int i = 0; i < STG; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_smem, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      prefetch_tmap(&tm_a);
      prefetch_tmap(&tm_b);
    }
    for (int it = 0; it < nk; ++it) {
      const int st = it % STG;
      mbar_wait(&bar_empty[st], ((it / STG) & 1) ^ 1);
      if (elect_one()) {
        uint8_t* s = smem + st * kStageBytes;
        const int k = k_begin + it * BK;                   // tokens past Kt (and past this split's end, see below) are
        mbar_expect_tx(&bar_full[st], kStageBytes);        // zero-filled by the TMA unit / masked by the split bounds
        tma_load_2d(s, &tm_a, &bar_full[st], m0, k);
        tma_load_2d(s + kAtom, &tm_a, &bar_full[st], m0 + 64, k);
        tma_load_2d(s + 2 * kAtom, &tm_b, &bar_full[st], n0, k);
        tma_load_2d(s + 3 * kAtom, &tm_b, &bar_full[st], n0 + 64, k);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 1, 1);          // A and B MN-major
    const uint32_t s0 = smem_u32(smem);
    // MN-major operand: 64-element (128 B) atoms along M / N are kAtom bytes apart (LBO), 8 token rows 1024 B apart (SBO)
    const uint64_t a_d0 = make_smem_desc_sw128(s0, kAtom, 1024);
    const uint64_t b_d0 = make_smem_desc_sw128(s0 + 2 * kAtom, kAtom, 1024);
    for (int it = 0; it < nk; ++it) {
      const int st = it % STG;
      mbar_wait(&bar_full[st], (it / STG) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a_d = a_d0 + (uint64_t)(st * (kStageBytes >> 4));
        const uint64_t b_d = b_d0 + (uint64_t)(st * (kStageBytes >> 4));
        if (it > 0) mma_ss_c<true>(tmem_base, a_d, b_d, idesc);
        else mma_ss_c<false>(tmem_base, a_d, b_d, idesc);
#pragma unroll
        for (int k = 1; k < BK / 16; ++k) mma_ss_c<true>(tmem_base, a_d + (uint64_t)(k * 128), b_d + (uint64_t)(k * 128), idesc);
        mma_commit(&bar_empty[st]);
        if (it == nk - 1) mma_commit(bar_done);
      }
      __syncwarp();
    }
  } else {
    // ============================== epilogue: fp32 partial tile -> workspace slab of this split ==============================
    const int q = warp & 3;                                  // TMEM lane quarter this warp may read
    const int row = q * 32 + (threadIdx.x & 31);             // row of C inside the tile == TMEM lane
    const int m = m0 + row;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    float* dst = part + ((int64_t)split * Ms + m) * Nb + n0;
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
      tmem_wait_ld();
      if (m < Ms) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const int n = n0 + c * 32 + i;
          if (n + 3 < Nb) {
            *reinterpret_cast<float4*>(dst + c * 32 + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                     __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
          } else {
            for (int j = 0; j < 4; ++j)
              if (n + j < Nb) dst[c * 32 + i + j] = __uint_as_float(r[i + j]);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// out[m, n] (or out[n, m]) = bf16(sum over splits, in split order)
__global__ void gemm_tn_finish_kernel(const float* __restrict__ part, __nv_bfloat16* __restrict__ out, int splits, int Ms,
                                      int Nb, int transpose) {
  pdl_trigger();
  pdl_wait();
  const int64_t total = (int64_t)Ms * Nb;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[(int64_t)k * total + i];
    if (transpose) {
      const int64_t m = i / Nb, n = i - m * Nb;
      out[n * Ms + m] = __float2bfloat16_rn(s);
    } else {
      out[i] = __float2bfloat16_rn(s);
    }
  }
}

int plan_splits(int64_t Kt, int64_t Ms, int64_t Nb, int* chunk_out) {
  const int64_t tiles = ((Nb + BN - 1) / BN) * ((Ms + BM - 1) / BM);
  int64_t want = (2 * (int64_t)sm_count() + tiles - 1) / tiles;             // ~2 CTAs per SM
  const int64_t max_splits = (Kt + 4 * BK - 1) / (4 * BK);                  // at least 4 pipeline stages of work per CTA
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  int64_t chunk = ((Kt + want - 1) / want + BK - 1) / BK * BK;              // whole stages per split: no token is seen twice
  *chunk_out = (int)chunk;
  return (int)((Kt + chunk - 1) / chunk);
}

}  // namespace
}  // namespace advgrpo

using namespace advgrpo;

extern "C" {

size_t advgrpo_gemm_tn_skinny_workspace_bytes(int64_t Kt, int64_t Ms, int64_t Nb) {
  if (Kt <= 0 || Ms <= 0 || Nb <= 0) return 16;
  int chunk = 0;
  const int splits = plan_splits(Kt, Ms, Nb, &chunk);
  return (size_t)splits * (size_t)Ms * (size_t)Nb * sizeof(float) + 256;
}

int advgrpo_gemm_tn_skinny(const void* a, const void* b, void* out, int64_t Kt, int64_t Ms, int64_t Nb, int transpose_out,
                           void* workspace, size_t workspace_bytes, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(a && b && out, "gemm_tn_skinny: null pointer");
  ADVGRPO_CHECK_ARG(Kt >= 1 && Ms >= 1 && Ms <= 256 && Nb >= 1 && Ms % 8 == 0 && Nb % 8 == 0 && Kt < ((int64_t)1 << 30) &&
                        Nb < ((int64_t)1 << 30),
                    "gemm_tn_skinny: need 1 <= Ms <= 256, Ms %% 8 == 0, Nb %% 8 == 0 (got Kt=%lld Ms=%lld Nb=%lld)", (long long)Kt,
                    (long long)Ms, (long long)Nb);
  ADVGRPO_CHECK_ARG(aligned16(a) && aligned16(b) && aligned16(out), "gemm_tn_skinny: tensors must be 16-byte aligned");
  if (!workspace || workspace_bytes < advgrpo_gemm_tn_skinny_workspace_bytes(Kt, Ms, Nb))
    return set_error(ADVGRPO_ERR_WORKSPACE, "gemm_tn_skinny: workspace too small");
  ADVGRPO_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "gemm_tn_skinny: workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int chunk = 0;
  const int splits = plan_splits(Kt, Ms, Nb, &chunk);
  CUtensorMap tm_a, tm_b;
  {
    const uint32_t box[2] = {64, BK};
    const uint64_t da[2] = {(uint64_t)Ms, (uint64_t)Kt}, sa[2] = {0, (uint64_t)Ms * 2};
    int rc = make_tmap_bf16(&tm_a, a, 2, da, sa, box, true);
    if (rc) return rc;
    const uint64_t db[2] = {(uint64_t)Nb, (uint64_t)Kt}, sb[2] = {0, (uint64_t)Nb * 2};
    rc = make_tmap_bf16(&tm_b, b, 2, db, sb, box, true);
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    ADVGRPO_CUDA_CALL(cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set = true;
  }
  dim3 grid((unsigned)((Nb + BN - 1) / BN), (unsigned)splits, (unsigned)((Ms + BM - 1) / BM));
  ADVGRPO_CUDA_CALL(launch_chain(gemm_tn_kernel, grid, dim3(kThreads), kSmem, st, 1, tm_a, tm_b, (float*)workspace, (int)Kt,
                                 (int)Ms, (int)Nb, chunk));
  ADVGRPO_CUDA_LAUNCH_CHECK();
  const int64_t total = Ms * Nb;
  const int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  if (blocks > 4 * (int64_t)sm_count()) blocks = 4 * (int64_t)sm_count();
  ADVGRPO_CUDA_CALL(launch_chain(gemm_tn_finish_kernel, dim3((unsigned)blocks), dim3(threads), 0, st, 1,
                                 (const float*)workspace, (__nv_bfloat16*)out, splits, (int)Ms, (int)Nb, transpose_out));
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

}  // extern "C"

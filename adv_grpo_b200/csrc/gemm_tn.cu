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
// Split-K WITHOUT a global workspace: one thread-block CLUSTER of 8 CTAs owns a 128 x 128 output tile, CTA r streams the
// r-th eighth of the tokens into its own TMEM accumulator, parks the fp32 tile in its shared memory, and after a cluster
// barrier every CTA sums 16 rows of the tile over the eight CTAs' shared memories (distributed shared memory, fixed rank
// order: bit-reproducible, no atomics), casts to bf16 and writes the output -- optionally transposed.  One launch.  (A
// first version wrote fp32 partial tiles to a workspace and summed them in a second kernel: the partials cost almost as
// much traffic as the operand stream, 30.8 us vs cuBLAS 25.5 us at 16384 x 128 x 1536.)
//
// Warps: 0 = TMA producer, 1 = MMA issuer (elect_one), 2-5 = epilogue / reduction (TMEM lane quarters 2, 3, 0, 1).
// Replaces `da = dt.t() @ x2`, `dw2 = dy2.t() @ t` of the LoRA autograd nodes (scripts/train_sd3_fast_pickscore.py:1165,
// peft LoRA Linear backward) that round 1 left on cuBLAS (nvjet_* in the launch list).
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

namespace advgrpo {
namespace {

using namespace sm100;

constexpr int BM = 128;                 // rows of C per cluster (columns of A)
// BN = columns of C per cluster (columns of B): 128, or 256 for wide B (the skinny operand is re-read once per column
// tile, through L2: at Nb = 4608 the 36 re-reads of a 128-wide tile moved as many bytes as B itself)
constexpr int BK = 64;                  // tokens per pipeline stage
constexpr int STG = 4;
constexpr int CL = 8;                   // CTAs per cluster = split-K factor
constexpr int kAtom = BK * 128;         // bytes of one 64-column x BK-token atom
constexpr int kThreads = 192;
template <int BN> struct TnCfg {
  static constexpr int kStageBytes = (2 + BN / 64) * kAtom;   // A: 2 atoms, B: BN / 64 atoms
  static constexpr int kSmem = STG * kStageBytes + 1024 + 128;
  static_assert(BM * BN * 4 <= STG * kStageBytes, "the fp32 tile is parked in the (drained) pipeline stages");
};

__device__ __forceinline__ float4 ld_cluster_f4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr));
  return v;
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, __nv_bfloat16* __restrict__ out,
               int Kt, int Ms, int Nb, int chunk, int transpose) {
  constexpr int kStageBytes = TnCfg<BN>::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STG * kStageBytes);
  uint64_t* bar_full = bars;             // STG
  uint64_t* bar_empty = bars + STG;      // STG
  uint64_t* bar_done = bars + 2 * STG;   // 1
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 2 * STG + 1);
  const int warp = threadIdx.x >> 5;
  const int rank = (int)cluster_ctarank();                   // which eighth of the tokens
  const int n0 = (blockIdx.x / CL) * BN, m0 = blockIdx.z * BM;
  const int k_begin = rank * chunk;
  const int k_end = min(Kt, k_begin + chunk);
  const int nk = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;   // 0 for the tail ranks of a short token axis

  if (threadIdx.x == 0) {
    for (int i = 0; i < STG; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_base_smem, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_trigger();
  pdl_wait();
  const uint32_t tile_sm = smem_u32(smem);                   // fp32 [128][128] tile, 16-byte chunks XOR-swizzled by row

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
        const int k = k_begin + it * BK;                   // tokens past Kt are zero-filled by the TMA unit; a chunk is
        mbar_expect_tx(&bar_full[st], kStageBytes);        // a whole number of stages, so no token is read by two ranks
        tma_load_2d(s, &tm_a, &bar_full[st], m0, k);
        tma_load_2d(s + kAtom, &tm_a, &bar_full[st], m0 + 64, k);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) tma_load_2d(s + (2 + j) * kAtom, &tm_b, &bar_full[st], n0 + 64 * j, k);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 1, 1);          // A and B MN-major
    // MN-major operand: 64-element (128 B) atoms along M / N are kAtom bytes apart (LBO), 8 token rows 1024 B apart (SBO)
    const uint64_t a_d0 = make_smem_desc_sw128(tile_sm, kAtom, 1024);
    const uint64_t b_d0 = make_smem_desc_sw128(tile_sm + 2 * kAtom, kAtom, 1024);
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
    // ============================== epilogue, part 1: my fp32 partial tile -> my shared memory ==============================
    const int q = warp & 3;                                  // TMEM lane quarter this warp may read
    const int row = q * 32 + (threadIdx.x & 31);             // row of C inside the tile == TMEM lane
    if (nk > 0) {
      mbar_wait(bar_done, 0);                                // every MMA finished: the pipeline stages are free
      tc_fence_after();
    }
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      if (nk > 0) {
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0u;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)                            // 16-byte chunk (8 c + i) of the row, swizzled: conflict-free
        sts_u4(tile_sm + row * (BN * 4) + (((8 * c + i) ^ (row & 31)) << 4), make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]));
    }
    tc_fence_before();
  }
  cluster_sync_all();                                        // all eight partial tiles are parked (release / acquire)
  if (warp >= 2) {
    // ============================== epilogue, part 2: rows [16 rank, 16 rank + 16) summed over the cluster ==============================
    const int t = threadIdx.x - 64;                          // 0..127
    uint32_t peer[CL];
#pragma unroll
    for (int r = 0; r < CL; ++r) peer[r] = mapa_u32(tile_sm, (uint32_t)r);
#pragma unroll
    for (int i = 0; i < (BM / CL) * (BN / 4) / 128; ++i) {
      const int idx = t + 128 * i;
      const int row = rank * (BM / CL) + idx / (BN / 4), c4 = idx % (BN / 4);   // 16-byte chunk c4 of the row
      const uint32_t off = (uint32_t)(row * (BN * 4) + ((c4 ^ (row & 31)) << 4));
      float4 acc = ld_cluster_f4(peer[0] + off);
#pragma unroll
      for (int r = 1; r < CL; ++r) {                         // fixed order: bit-reproducible
        const float4 v = ld_cluster_f4(peer[r] + off);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      const int m = m0 + row, n = n0 + 4 * c4;
      if (m < Ms && n < Nb) {                                // Nb is a multiple of 8: a 4-column chunk is in or out as a whole
        if (!transpose) {
          uint2 pk;
          pk.x = pack_bf16(acc.x, acc.y);
          pk.y = pack_bf16(acc.z, acc.w);
          *reinterpret_cast<uint2*>(out + (int64_t)m * Nb + n) = pk;
        } else {
          out[(int64_t)(n + 0) * Ms + m] = __float2bfloat16_rn(acc.x);
          out[(int64_t)(n + 1) * Ms + m] = __float2bfloat16_rn(acc.y);
          out[(int64_t)(n + 2) * Ms + m] = __float2bfloat16_rn(acc.z);
          out[(int64_t)(n + 3) * Ms + m] = __float2bfloat16_rn(acc.w);
        }
      }
    }
  }
  cluster_sync_all();                                        // nobody exits while a peer still reads its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

}  // namespace
}  // namespace advgrpo

using namespace advgrpo;

extern "C" {

size_t advgrpo_gemm_tn_skinny_workspace_bytes(int64_t Kt, int64_t Ms, int64_t Nb) {
  (void)Kt; (void)Ms; (void)Nb;
  return 0;   // the split-K partials live in the cluster's shared memory; kept in the ABI for callers that size buffers
}

int advgrpo_gemm_tn_skinny(const void* a, const void* b, void* out, int64_t Kt, int64_t Ms, int64_t Nb, int transpose_out,
                           void* workspace, size_t workspace_bytes, advgrpo_stream_t stream) {
  (void)workspace; (void)workspace_bytes;
  ADVGRPO_CHECK_ARG(a && b && out, "gemm_tn_skinny: null pointer");
  ADVGRPO_CHECK_ARG(Kt >= 1 && Ms >= 1 && Ms <= 16384 && Nb >= 1 && Ms % 8 == 0 && Nb % 8 == 0 && Kt < ((int64_t)1 << 30) &&
                        Nb < ((int64_t)1 << 24),
                    "gemm_tn_skinny: need 1 <= Ms <= 16384, Ms %% 8 == 0, Nb %% 8 == 0 (got Kt=%lld Ms=%lld Nb=%lld)", (long long)Kt,
                    (long long)Ms, (long long)Nb);
  ADVGRPO_CHECK_ARG(aligned16(a) && aligned16(b) && aligned16(out), "gemm_tn_skinny: tensors must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunk = (int)((((Kt + CL - 1) / CL) + BK - 1) / BK * BK);     // whole pipeline stages per cluster rank
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
    ADVGRPO_CUDA_CALL(cudaFuncSetAttribute(gemm_tn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TnCfg<128>::kSmem));
    ADVGRPO_CUDA_CALL(cudaFuncSetAttribute(gemm_tn_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TnCfg<256>::kSmem));
    attr_set = true;
  }
  static const int bn_env = getenv("ADVGRPO_GEMM_TN_BN") ? atoi(getenv("ADVGRPO_GEMM_TN_BN")) : 0;
  // 256-column tiles when B is wide (its 128-column tiles would re-read the A tile as often as B itself) or when A is a
  // full weight-gradient operand (Ms >= 512: the grid already has hundreds of clusters, halving A's re-reads wins)
  const int bn = bn_env == 128 || bn_env == 256 ? bn_env : ((Nb >= 3072 || (Ms >= 512 && Nb >= 256)) ? 256 : 128);
  dim3 grid((unsigned)(((Nb + bn - 1) / bn) * CL), 1, (unsigned)((Ms + BM - 1) / BM));
  if (bn == 256) {
    ADVGRPO_CUDA_CALL(launch_chain(gemm_tn_kernel<256>, grid, dim3(kThreads), TnCfg<256>::kSmem, st, CL, tm_a, tm_b,
                                   (__nv_bfloat16*)out, (int)Kt, (int)Ms, (int)Nb, chunk, transpose_out));
  } else {
    ADVGRPO_CUDA_CALL(launch_chain(gemm_tn_kernel<128>, grid, dim3(kThreads), TnCfg<128>::kSmem, st, CL, tm_a, tm_b,
                                   (__nv_bfloat16*)out, (int)Kt, (int)Ms, (int)Nb, chunk, transpose_out));
  }
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

}  // extern "C"

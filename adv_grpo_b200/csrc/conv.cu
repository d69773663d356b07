// 3x3 (stride 1, zero padding 1) and 1x1 convolutions on channels-last fp32 activations as an implicit GEMM on
// tcgen05 tensor cores in TF32 (kind::tf32, fp32 accumulation in TMEM) -- the convolutions of the SD3 VAE decoder,
// which the reference runs as fp32 `AutoencoderKL.decode` (TF32 tensor-core convolutions under PyTorch's default
// cudnn.allow_tf32; reference call site adv_grpo/diffusers_patch/sd3_pipeline_with_logprob_fast.py:667-670).
//
//   y[b, y, x, n] = sum_{tap, c} x[b, y + dy(tap), x + dx(tap), c] * w[n, tap, c]  (+ bias[n])
//
// GEMM view: M = pixels (128 per tile: a bw x bh patch of one image), N = output channels, K = taps * Cin walked as
// (tap, 32-channel block).  The A tile of one K block is ONE 4-D TMA box {32 channels, bw, bh, 1} of the NHWC input at
// the tap's spatial offset: rows that fall outside the image are zero-filled by the TMA unit, which IS the
// convolution's zero padding -- no im2col buffer, no halo copies, no padded activations.  The B tile is a 2-D box of
// the [Cout, taps * Cin] weight matrix.  Both land 128B-swizzled (32 fp32 = 128 B rows) and feed tcgen05.mma directly.
// Structure as in gemm.cu: persistent CTAs, warp 8 TMA producer, warp 9 MMA issuer, two 4-warp epilogue teams that
// drain the two TMEM accumulators (TMEM -> registers -> (+bias) -> swizzled smem tile -> 4-D TMA store into NHWC y).
#include "common.cuh"
#include "sm100.cuh"

namespace advgrpo {
namespace {

using namespace sm100;

constexpr int CBM = 128;   // pixels per tile
constexpr int CBK = 32;    // fp32 channels per K block (128 bytes)

// MT = 128-pixel sub-tiles per CTA tile (two vertically adjacent patches share one B tile): with 128 output
// channels a 128 x 128 tile needs 32 KB of operands per 256 tensor clocks, which starves the pipe; 256 x 128 halves it.
// TWO = CTA pair (cta_group::2): the pair owns 256 pixels (two stacked patches, one per CTA) x 256 channels; each CTA
// stages its own patch and HALF of the weight tile, which halves the per-CTA operand traffic and allows 5 stages.
template <int BN, int MT, bool TWO = false>
struct CCfg {
  // stage = MT x 16 KB of pixels + BN x 128 B of weights; + 64 KB of staging tiles: 208-224 KB of shared memory
  static constexpr int kStages = TWO ? 5 : ((BN * MT >= 256) ? 3 : ((BN == 32 && MT == 2) ? 4 : 5));
  static constexpr int kABytes = MT * CBM * 128;
  static constexpr int kBBytes = (TWO ? BN / 2 : BN) * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kNBuf = 4;                            // 2 staging tiles [128 x 32 fp32] per epilogue team
  static constexpr int kStagingBytes = kNBuf * CBM * 128;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 + 256;
  static constexpr int kTmemCols = 2 * BN * MT <= 256 ? 256 : 512;
  static constexpr int kThreads = 320;
};

struct CParams {
  const float* bias;      // [Cout] or null
  int Cout, Cin, taps, kb_per_tap;
  int tiles_n, tiles_x, tiles_y, num_tiles;
  int bw, bh;
};

// kind::tf32 instruction descriptor: c_format f32 (1) at [4,6), a/b_format TF32 (2) at [7,10) / [10,13), K-major A and B
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void mma_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mma_ss_tf32_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

template <int BN, int MT, bool TWO>
__global__ void __launch_bounds__(320, 1)
conv_tf32_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                 const __grid_constant__ CUtensorMap tm_y, const CParams p) {
  using G = CCfg<BN, MT, TWO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage = smem + G::kStages * G::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + G::kStagingBytes);
  uint64_t* bar_full = bars;
  uint64_t* bar_empty = bars + G::kStages;
  uint64_t* bar_acc_full = bar_empty + G::kStages;   // 2
  uint64_t* bar_acc_empty = bar_acc_full + 2;        // 2
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_tiles;
  const int kb_total = p.taps * p.kb_per_tap;
  const uint32_t rank = TWO ? cluster_ctarank() : 0u;          // 0 = leader CTA of the pair
  const int worker = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_workers = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int kRowMul = TWO ? 2 : MT;                       // patches stacked per tile

  if (threadIdx.x == 0) {
    for (int i = 0; i < G::kStages; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_acc_full[i], 1);
      mbar_init(&bar_acc_empty[i], TWO ? 256 : 128);
    }
    fence_barrier_init();
  }
  if constexpr (TWO) cluster_sync_all();
  if (warp == 9) {
    if constexpr (TWO) {
      tmem_alloc_2cta(tmem_base_smem, G::kTmemCols);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(tmem_base_smem, G::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (TWO) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  // tile -> (output-channel block, image patch)
  auto decode = [&](int tile, int& n0, int& x0, int& y0, int& b) {
    n0 = (tile % p.tiles_n) * BN;
    const int m = tile / p.tiles_n;
    x0 = (m % p.tiles_x) * p.bw;
    y0 = ((m / p.tiles_x) % p.tiles_y) * (p.bh * kRowMul) + (TWO ? (int)rank * p.bh : 0);
    b = m / (p.tiles_x * p.tiles_y);
  };

  if (warp == 8) {
    // ============================== TMA producer (warp-uniform loop; elect_one() around the TMA issue) ==========
    {
      if (elect_one()) {
        prefetch_tmap(&tm_x);
        prefetch_tmap(&tm_w);
      }
      int it = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        int n0, x0, y0, b;
        decode(tile, n0, x0, y0, b);
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
          for (int cb = 0; cb < p.kb_per_tap; ++cb, ++it) {
            const int st = it % G::kStages;
            mbar_wait(&bar_empty[st], ((it / G::kStages) & 1) ^ 1);
            if (elect_one()) {
              uint8_t* sa = smem + st * G::kStageBytes;
              // rows outside the image (negative or >= W / H coordinates) arrive as zeros: the conv's zero padding
              if constexpr (TWO) {
                const uint32_t full_leader = mapa_u32(smem_u32(&bar_full[st]), 0);
                if (rank == 0) mbar_expect_tx(&bar_full[st], 2 * G::kStageBytes);
                tma_load_4d_2sm(sa, &tm_x, full_leader, cb * CBK, x0 + dx, y0 + dy, b);
                tma_load_2d_2sm(sa + G::kABytes, &tm_w, full_leader, tap * p.Cin + cb * CBK, n0 + (int)rank * (BN / 2));
              } else {
                mbar_expect_tx(&bar_full[st], G::kStageBytes);
                tma_load_4d(sa, &tm_x, &bar_full[st], cb * CBK, x0 + dx, y0 + dy, b);
                tma_load_2d(sa + G::kABytes, &tm_w, &bar_full[st], tap * p.Cin + cb * CBK, n0);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 9) {
    // ============================== MMA issuer (warp-uniform loop; elect_one() around the tcgen05 issue) =========
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(TWO ? 2 * CBM : CBM, BN);
      const uint32_t smem_base = smem_u32(smem);
      int it = 0, local = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers, ++local) {
        const int acc = local & 1;
        mbar_wait(&bar_acc_empty[acc], ((local >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (BN * MT);
        for (int kb = 0; kb < kb_total; ++kb, ++it) {
          const int st = it % G::kStages;
          mbar_wait(&bar_full[st], (it / G::kStages) & 1);
          tc_fence_after();
          if (elect_one()) {
            // descriptors built once per stage; + k * 32 bytes = + 2, + mt * CBM * 128 bytes in 16-byte address units
            const uint64_t da = make_smem_desc_sw128(smem_base + st * G::kStageBytes, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(smem_base + st * G::kStageBytes + G::kABytes, 16, 1024);
#pragma unroll
            for (int k = 0; k < CBK / 8; ++k)   // K = 8 tf32 (32 bytes) per instruction
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                if constexpr (TWO)
                  mma_ss_tf32_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                else
                  mma_ss_tf32(d_tmem + mt * BN, da + (uint64_t)(mt * (CBM * 128 / 16) + 2 * k), db + 2 * k, idesc,
                              (kb > 0 || k > 0) ? 1u : 0u);
              }
            if constexpr (TWO) mma_commit_2cta_mc(&bar_empty[st], 3); else mma_commit(&bar_empty[st]);
            if (kb == kb_total - 1) {
              if constexpr (TWO) mma_commit_2cta_mc(&bar_acc_full[acc], 3); else mma_commit(&bar_acc_full[acc]);
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ============================== epilogue ==============================
    const int team = warp >> 2;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int tid = threadIdx.x & 127;
    const int lrow = (warp & 3) * 32 + lane;
    const uint32_t bar_a = 1 + 2 * team, bar_b = 2 + 2 * team;
    uint8_t* tstage = stage + team * (2 * CBM * 128);
    constexpr int NG = BN / 32;
    uint32_t gc = 0;
    int round = 0;
    for (int tile = worker + team * num_workers; tile < num_tiles; tile += 2 * num_workers, ++round) {
      int n0, x0, y0, b;
      decode(tile, n0, x0, y0, b);
      int ng = (p.Cout - n0 + 31) / 32;
      ng = ng < NG ? ng : NG;
      mbar_wait(&bar_acc_full[team], round & 1);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + team * (BN * MT) + lane_addr;
#pragma unroll 1
      for (int gg = 0; gg < ng * MT; ++gg, ++gc) {
        const int mt = gg / ng, g = gg - mt * ng;     // sub-tile (patch rows y0 + mt * bh ..) and 32-column group
        uint32_t r[32];
        tmem_ld32(t_acc + mt * BN + g * 32, r);
        const uint32_t buf = gc & 1;
        if (tid == 0) tma_store_wait_read<1>();        // this staging tile's previous store (two groups ago) drained
        named_bar_sync(bar_a, 128);
        uint8_t* sbuf = tstage + buf * (CBM * 128);
        const int colg = n0 + g * 32;
        float4 bv[8];
        if (p.bias) {
#pragma unroll
          for (int q = 0; q < 8; ++q) bv[q] = *reinterpret_cast<const float4*>(p.bias + colg + q * 4);
        }
        tmem_wait_ld();
        if (gg == ng * MT - 1) {
          tc_fence_before();
          if constexpr (TWO) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_acc_empty[team]), 0));
          else mbar_arrive(&bar_acc_empty[team]);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 v = make_float4(__uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]), __uint_as_float(r[q * 4 + 2]),
                                 __uint_as_float(r[q * 4 + 3]));
          if (p.bias) {
            v.x += bv[q].x; v.y += bv[q].y; v.z += bv[q].z; v.w += bv[q].w;
          }
          sts_u4(smem_u32(sbuf) + lrow * 128 + ((q ^ (lrow & 7)) * 16),
                 make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)));
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_b, 128);
        if (tid == 0) {
          tma_store_4d(&tm_y, sbuf, colg, x0, y0 + mt * p.bh, b);   // clipped at the image edges for ragged patches
          tma_store_commit();
        }
      }
    }
    if (tid == 0) tma_store_wait_all();
  }

  tc_fence_before();
  if constexpr (TWO) cluster_sync_all(); else __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    if constexpr (TWO) tmem_dealloc_2cta(tmem_base, G::kTmemCols); else tmem_dealloc(tmem_base, G::kTmemCols);
  }
}

template <int BN, int MT, bool TWO>
int launch_conv(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, const CParams& p, cudaStream_t st) {
  using G = CCfg<BN, MT, TWO>;
  static bool attr_set = false;
  if (!attr_set) {
    ADVGRPO_CUDA_CALL(cudaFuncSetAttribute(conv_tf32_kernel<BN, MT, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::kSmemBytes));
    attr_set = true;
  }
  int workers = TWO ? sm_count() / 2 : sm_count();
  if (workers > p.num_tiles) workers = p.num_tiles;
  if constexpr (TWO) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * workers);
    cfg.blockDim = dim3(G::kThreads);
    cfg.dynamicSmemBytes = G::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ADVGRPO_CUDA_CALL(cudaLaunchKernelEx(&cfg, conv_tf32_kernel<BN, MT, TWO>, mx, mw, my, p));
  } else {
    conv_tf32_kernel<BN, MT, TWO><<<workers, G::kThreads, G::kSmemBytes, st>>>(mx, mw, my, p);
  }
  ADVGRPO_CUDA_LAUNCH_CHECK();
  return ADVGRPO_OK;
}

int g_conv_variant = 0;   // 0 = auto, 1 = single-CTA tiles only (test hook)

}  // namespace
}  // namespace advgrpo

using namespace advgrpo;

extern "C" {

// Test/bench hook (not part of the reference-facing surface): force the convolution CTA shape.
void advgrpo_debug_set_conv_variant(int v) { g_conv_variant = v; }

int advgrpo_conv2d_nhwc_tf32(const float* x, const float* w, const float* bias, float* y, int64_t B, int64_t H, int64_t W,
                             int64_t Cin, int64_t Cout, int ksize, advgrpo_stream_t stream) {
  ADVGRPO_CHECK_ARG(x && w && y, "conv2d_nhwc_tf32: null pointer");
  ADVGRPO_CHECK_ARG(ksize == 1 || ksize == 3, "conv2d_nhwc_tf32: kernel size must be 1 or 3 (got %d)", ksize);
  ADVGRPO_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && Cin >= 32 && Cin % 32 == 0 && Cout >= 32 && Cout % 32 == 0,
                    "conv2d_nhwc_tf32: Cin and Cout must be multiples of 32 (B=%lld H=%lld W=%lld Cin=%lld Cout=%lld)",
                    (long long)B, (long long)H, (long long)W, (long long)Cin, (long long)Cout);
  ADVGRPO_CHECK_ARG(aligned16(x) && aligned16(w) && aligned16(y) && (!bias || aligned16(bias)), "conv2d_nhwc_tf32: 16-byte alignment");
  const int taps = ksize * ksize;
  // patch = bw x bh pixels with bw * bh = 128: the widest power of two <= min(W, 128)
  int bw = 128;
  while (bw > W) bw >>= 1;
  const int bh = CBM / bw;
  CParams p;
  p.bias = bias;
  p.Cout = (int)Cout; p.Cin = (int)Cin; p.taps = taps; p.kb_per_tap = (int)(Cin / CBK);
  p.bw = bw; p.bh = bh;
  p.tiles_x = (int)((W + bw - 1) / bw);
  p.tiles_y = (int)((H + bh - 1) / bh);
  // 32-column tiles for the decoder's conv_out (3 output channels zero-padded to 32): a 128-column tile would spend 4x the
  // tensor-core time on zero filters (g_conv_variant 2 = test hook: keep the 128-column tiles)
  const int BN = Cout >= 256 ? 256 : ((Cout <= 32 && g_conv_variant != 2) ? 32 : 128);
  const bool pair = BN == 256 && H >= 2 * bh && g_conv_variant != 1;
  const int MT = (BN <= 128 && H >= 2 * bh) ? 2 : 1;
  const int rows = pair ? 2 : MT;                                          // patches stacked per tile
  p.tiles_y = (int)((H + bh * rows - 1) / (bh * rows));
  p.tiles_n = (int)((Cout + BN - 1) / BN);
  const int64_t tiles = (int64_t)p.tiles_n * p.tiles_x * p.tiles_y * B;
  ADVGRPO_CHECK_ARG(tiles < ((int64_t)1 << 30), "conv2d_nhwc_tf32: too many tiles");
  p.num_tiles = (int)tiles;
  CUtensorMap mx, mw, my;
  const uint64_t dx[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t sx[4] = {0, (uint64_t)Cin * 4, (uint64_t)(W * Cin) * 4, (uint64_t)(H * W * Cin) * 4};
  const uint32_t bx[4] = {CBK, (uint32_t)bw, (uint32_t)(bh * MT), 1};     // load box: MT stacked patches
  const uint32_t by[4] = {CBK, (uint32_t)bw, (uint32_t)bh, 1};            // store box: one patch per 128-row sub-tile
  int rc = make_tmap(&mx, x, 4, dx, sx, bx, true, true);
  if (rc) return rc;
  const uint64_t dw[2] = {(uint64_t)(taps * Cin), (uint64_t)Cout};
  const uint64_t sw[2] = {0, (uint64_t)(taps * Cin) * 4};
  const uint32_t bwt[2] = {CBK, (uint32_t)(pair ? BN / 2 : BN)};
  rc = make_tmap(&mw, w, 2, dw, sw, bwt, true, true);
  if (rc) return rc;
  const uint64_t dy[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t sy[4] = {0, (uint64_t)Cout * 4, (uint64_t)(W * Cout) * 4, (uint64_t)(H * W * Cout) * 4};
  rc = make_tmap(&my, y, 4, dy, sy, by, true, true);
  if (rc) return rc;
  if (pair) return launch_conv<256, 1, true>(mx, mw, my, p, (cudaStream_t)stream);
  if (BN == 256) return launch_conv<256, 1, false>(mx, mw, my, p, (cudaStream_t)stream);
  if (BN == 32) {
    if (MT == 2) return launch_conv<32, 2, false>(mx, mw, my, p, (cudaStream_t)stream);
    return launch_conv<32, 1, false>(mx, mw, my, p, (cudaStream_t)stream);
  }
  if (MT == 2) return launch_conv<128, 2, false>(mx, mw, my, p, (cudaStream_t)stream);
  return launch_conv<128, 1, false>(mx, mw, my, p, (cudaStream_t)stream);
}

}  // extern "C"

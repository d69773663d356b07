// Shared host/device helpers for libadvgrpo_b200: error reporting across the C ABI,
// warp/block reductions, bf16 vector access, TMA tensor-map creation.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/advgrpo_b200.h"

namespace advgrpo {

// ---- error state (thread-local: the ABI is re-entrant, reward thread + main thread) ----
extern thread_local char g_last_error[512];
int set_error(int code, const char* fmt, ...);

#define ADVGRPO_CHECK_ARG(cond, ...)                                       \
  do {                                                                     \
    if (!(cond)) return ::advgrpo::set_error(ADVGRPO_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)

#define ADVGRPO_CUDA_LAUNCH_CHECK()                                                        \
  do {                                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess)                                                                \
      return ::advgrpo::set_error(ADVGRPO_ERR_CUDA, "%s:%d CUDA error %d: %s", __FILE__, __LINE__, \
                                  (int)e__, cudaGetErrorString(e__));                      \
  } while (0)

#define ADVGRPO_CUDA_CALL(x)                                                               \
  do {                                                                                     \
    cudaError_t e__ = (x);                                                                 \
    if (e__ != cudaSuccess)                                                                \
      return ::advgrpo::set_error(ADVGRPO_ERR_CUDA, "%s:%d CUDA error %d: %s", __FILE__, __LINE__, \
                                  (int)e__, cudaGetErrorString(e__));                      \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int sm_count();

// ---- programmatic dependent launch (experimental, off by default: env ADVGRPO_PDL=1 or advgrpo_debug_set_pdl) ----
// Kernels of the MMDiT forward / backward chain call pdl_trigger() + pdl_wait() once their prologue (barrier init,
// TMEM allocation, tensor-map prefetch) is done and before their first global-memory access.  When the launch carries
// cudaLaunchAttributeProgrammaticStreamSerialization, the next kernel's CTAs are scheduled as soon as every CTA of
// this one has started, run their own prologue, and block in griddepcontrol.wait until this grid has completed and
// flushed: the launch latency and the prologue move off the critical path.  Without the attribute both instructions
// are no-ops, so the default (plain <<<>>>) behaviour is unchanged.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- TMA tensor maps (driver entry point resolved at run time: no link-time libcuda) ----
// dims/strides innermost first; strides in BYTES for dims 1..rank-1. bf16 elements.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128);
int make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128, bool f32);

// ---- device helpers ----
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum; every thread gets the result. `scratch` holds >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) scratch[w] = v;
  __syncthreads();
  float t = (l < nw) ? scratch[l] : 0.f;
  return warp_sum(t);
}

// Eight bf16 values as ONE 16-byte vector.  The payload is a built-in uint4 so that `*reinterpret_cast<const bf16x8*>(p)`
// compiles to a single LDG.128 / STG.128 (an array of four __nv_bfloat162 members was split into four 32-bit accesses,
// which quadrupled the L1 wavefronts of every streaming kernel).
struct alignas(16) bf16x8 {
  uint4 u;
};
__device__ __forceinline__ void unpack8(const bf16x8& p, float (&f)[8]) {
  const uint32_t w[4] = {p.u.x, p.u.y, p.u.z, p.u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);               // low half = element 2i
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);   // high half = element 2i + 1
  }
}
__device__ __forceinline__ bf16x8 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  bf16x8 p;
  p.u = make_uint4(w[0], w[1], w[2], w[3]);
  return p;
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
#endif

}  // namespace advgrpo

// Microbenchmark: how fast can the SM sub-partitions run the softmax instruction stream of the attention forward
// (row max, exp2(s*c - m) on MUFU and/or the FMA-pipe polynomial, row sum, bf16 pack) when nothing else (TMEM,
// mbarriers, MMA) is in the way?  Reports clocks per 128x128 score tile per SM for several (elements per thread,
// warps per SM, emulated-exp share) points.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_pipe softmax_pipe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  const float2 xr = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(xr, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
  float2 pl = __ffma2_rn(f, make_float2(0.055171650f, 0.055171650f), make_float2(0.24261113f, 0.24261113f));
  pl = __ffma2_rn(pl, f, make_float2(0.69326097f, 0.69326097f));
  pl = __ffma2_rn(pl, f, make_float2(0.99992806f, 0.99992806f));
  float2 r;
  r.x = __int_as_float(__float_as_int(pl.x) + (__float_as_int(xr.x) << 23));
  r.y = __int_as_float(__float_as_int(pl.y) + (__float_as_int(xr.y) << 23));
  return r;
}

// EPT scores per thread per tile; EMU of every 4 pairs on the FMA pipe; flags drop parts of the stream.
template <int EPT, int EMU, bool MAX, bool SUM, bool PACK, int THREADS, int MINB, int LAT = 0, bool SYNC = false>
__global__ void __launch_bounds__(THREADS, MINB) softmax_stream(float* out, int iters, float seed, long long* clk) {
  float s[EPT];
#pragma unroll
  for (int i = 0; i < EPT; ++i) s[i] = seed * (float)(i + 1) + (float)threadIdx.x * 1e-3f;
  float m = -1e30f, l = 0.f;
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < EPT; ++i) asm volatile("" : "+f"(s[i]));      // a fresh tile as far as the compiler knows
    if (LAT > 0) {                       // a per-tile latency bubble (mbarrier poll + TMEM load/store round trips)
      const long long t = clock64();
      while (clock64() - t < LAT) {}
    }
    if (SYNC) asm volatile("bar.sync 1, %0;" ::"r"(THREADS) : "memory");   // the warps of a query tile move in lockstep
    float mx = m;
    if (MAX) {
      float m4[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
#pragma unroll
      for (int i = 0; i < EPT; i += 2) m4[(i / 2) & 3] = fmaxf(m4[(i / 2) & 3], fmaxf(s[i], s[i + 1]));
      mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      mx = fmaxf(m, mx * 0.18f);
    }
    const float2 sc2 = make_float2(0.18f, 0.18f), nm2 = make_float2(-mx, -mx);
    float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < EPT; i += 2) {
      const float2 x = __ffma2_rn(make_float2(s[i], s[i + 1]), sc2, nm2);
      float2 e;
      if ((i / 2) % 4 < EMU) e = ex2_poly2(x);
      else { e.x = ex2(x.x); e.y = ex2(x.y); }
      if (SUM) sum2 = __fadd2_rn(sum2, e);
      if (PACK) { uint32_t p = pack_bf16(e.x, e.y); asm volatile("" :: "r"(p)); }
      else { asm volatile("" :: "f"(e.x), "f"(e.y)); }
    }
    l = l + sum2.x + sum2.y;
    m = mx;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  if (l == 12345.f) out[0] = l + m + acc;
}

template <int EPT, int EMU, bool MAX, bool SUM, bool PACK, int THREADS, int MINB, int LAT = 0, bool SYNC = false>
void run(const char* name, int ctas_per_sm, float* out, long long* clk, int sms) {
  auto k = softmax_stream<EPT, EMU, MAX, SUM, PACK, THREADS, MINB, LAT, SYNC>;
  const int iters = 2000;
  const int grid = sms * ctas_per_sm;
  k<<<grid, THREADS>>>(out, 10, 1.0f, clk);
  cudaDeviceSynchronize();
  k<<<grid, THREADS>>>(out, iters, 1.0f, clk);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  static long long h[4096];
  cudaMemcpy(h, clk, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
  // elements per SM per iteration / 16384 = tiles per SM per iteration
  const double tiles_per_iter = (double)ctas_per_sm * THREADS * EPT / 16384.0;
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
  printf("%-44s warps/SM %2d regs %3d : %7.1f clk per 128x128 tile per SM\n", name, ctas_per_sm * THREADS / 32, fa.numRegs,
         avg / iters / tiles_per_iter);
}

int main() {
  float* out; long long* clk; int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaMalloc(&out, 16); cudaMalloc(&clk, 4096 * 8);
  //   EPT EMU MAX SUM PACK THREADS MINB LAT SYNC
  run<128, 1, true, true, true, 128, 2>("row/thread 128 elts, 1/4 poly", 2, out, clk, sms);
  run<128, 1, true, true, true, 128, 2, 400>("row/thread 128 elts, 1/4 poly, lat 400", 2, out, clk, sms);
  run<128, 1, true, true, true, 128, 2, 400, true>("row/thread 128 elts, 1/4 poly, lat 400 sync", 2, out, clk, sms);
  run<64, 1, true, true, true, 256, 2>("quad 64 elts, 1/4 poly", 2, out, clk, sms);
  run<64, 1, true, true, true, 320, 2>("quad 64 elts, 1/4 poly, 96-reg cap (20 warps)", 2, out, clk, sms);
  run<64, 1, true, true, true, 256, 2, 200>("quad 64 elts, 1/4 poly, lat 200", 2, out, clk, sms);
  run<64, 1, true, true, true, 256, 2, 400>("quad 64 elts, 1/4 poly, lat 400", 2, out, clk, sms);
  run<64, 1, true, true, true, 256, 2, 800>("quad 64 elts, 1/4 poly, lat 800", 2, out, clk, sms);
  run<64, 1, true, true, true, 256, 2, 400, true>("quad 64 elts, 1/4 poly, lat 400 sync", 2, out, clk, sms);
  run<64, 1, true, true, true, 256, 2, 800, true>("quad 64 elts, 1/4 poly, lat 800 sync", 2, out, clk, sms);
  run<64, 0, true, true, true, 256, 2, 400, true>("quad 64 elts, MUFU only, lat 400 sync", 2, out, clk, sms);
  run<32, 1, true, true, true, 256, 2, 200, true>("32 elts (BKV 64), 1/4 poly, lat 200 sync", 2, out, clk, sms);
  run<32, 1, true, true, true, 256, 2, 0, true>("32 elts (BKV 64), 1/4 poly, lat 0 sync", 2, out, clk, sms);
  run<32, 1, true, true, true, 256, 2, 0, false>("32 elts (BKV 64), 1/4 poly, 16 warps", 2, out, clk, sms);
  run<32, 0, true, true, true, 256, 2, 0, false>("32 elts (BKV 64), MUFU only, 16 warps", 2, out, clk, sms);
  run<64, 0, true, true, true, 256, 2>("quad 64 elts, MUFU only", 2, out, clk, sms);
  return 0;
}

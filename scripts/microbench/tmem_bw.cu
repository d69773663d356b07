// Microbenchmark: TMEM -> register (tcgen05.ld) and register -> TMEM (tcgen05.st) throughput per SM for the shapes the
// attention softmax uses, as a function of the number of warps issuing them.  Is the 64 KB fp32 S tile read per
// 128x128 attention tile a bandwidth limit?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../adv_grpo_b200/csrc -o tmem_bw tmem_bw.cu
#include <cstdio>
#include "sm100.cuh"
using namespace sm100;

template <int MODE>
__global__ void __launch_bounds__(512, 1) tmem_rw(long long* clk, float* sink, int iters, int nwarps) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_base_s;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    // warp w may only touch TMEM lanes [32 (w % 4), +32); spread the warps of one quarter over different columns
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col = (uint32_t)((warp >> 2) * 128) & 511;
    const uint32_t addr = base + lane_base + col;
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0) {          // 2 x 16x256b.x8 (16 lanes x 64 columns each) + wait: one warp's share of an S tile (8 KB)
        uint32_t a[32], b[32];
        tmem_ld_16x256b_x8(addr, a);
        tmem_ld_16x256b_x8(addr + 64, b);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(a[i]) + __uint_as_float(b[i]);
      } else if (MODE == 1) {   // 2 x 32x32b.x32 (32 lanes x 32 columns each) + wait: 8 KB
        uint32_t a[32], b[32];
        tmem_ld32(addr, a);
        tmem_ld32(addr + 32, b);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(a[i]) + __uint_as_float(b[i]);
      } else if (MODE == 2) {   // store 16x128b.x16 (the packed bf16 P tile share of one warp: 16 lanes x 64 columns = 4 KB)
        uint32_t a[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = it + i;
        tmem_st_16x128b_x16(addr, a);
        tmem_wait_st();
      } else if (MODE == 3) {   // 4 x 16x256b.x8 back to back, one wait: 16 KB in flight per warp
        uint32_t a[32], b[32], c[32], d[32];
        tmem_ld_16x256b_x8(addr, a);
        tmem_ld_16x256b_x8(addr + 64, b);
        tmem_ld_16x256b_x8(addr, c);
        tmem_ld_16x256b_x8(addr + 64, d);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(a[i]) + __uint_as_float(b[i]) + __uint_as_float(c[i]) + __uint_as_float(d[i]);
      }
    }
    t1 = clock64();
  }
  if ((threadIdx.x & 31) == 0 && warp < nwarps) clk[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 12345.678f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(base, 512); }
}

template <int MODE>
void run(const char* name, int bytes_per_warp_iter, long long* clk, float* sink) {
  for (int nw : {1, 4, 8, 16}) {
    const int iters = 2000;
    tmem_rw<MODE><<<148, 512>>>(clk, sink, 10, nw);
    cudaDeviceSynchronize();
    tmem_rw<MODE><<<148, 512>>>(clk, sink, iters, nw);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    static long long h[148 * 16];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int b = 0; b < 148; ++b) for (int w = 0; w < nw; ++w) worst = h[b * 16 + w] > worst ? (double)h[b * 16 + w] : worst;
    const double per_iter = worst / iters;
    printf("%-44s warps/SM %2d : %7.1f clk per iteration, %7.1f B/clk/SM\n", name, nw, per_iter, nw * (double)bytes_per_warp_iter / per_iter);
  }
}

int main() {
  long long* clk; float* sink;
  cudaMalloc(&clk, 148 * 16 * 8); cudaMalloc(&sink, 16);
  run<0>("ld 2 x 16x256b.x8 + wait (8 KB/warp)", 8192, clk, sink);
  run<3>("ld 4 x 16x256b.x8 + wait (16 KB/warp)", 16384, clk, sink);
  run<1>("ld 2 x 32x32b.x32 + wait (8 KB/warp)", 8192, clk, sink);
  run<2>("st 16x128b.x16 + wait (4 KB/warp)", 4096, clk, sink);
  return 0;
}

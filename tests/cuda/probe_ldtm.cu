// Microbenchmark: cycles per tcgen05.ld (32x32b.xN) + wait as a function of N and of the number of warps reading.
#include <cstdio>
#include "common.cuh"
using namespace adaface;

template <int N>
__device__ __forceinline__ uint32_t ld_n(uint32_t taddr) {
  uint32_t acc = 0;
  if constexpr (N == 64) {
    uint32_t v[64];
    tmem_ld_32x32b_x64_wait(taddr, v);
#pragma unroll
    for (int i = 0; i < 64; ++i) acc ^= v[i];
  } else {
    uint32_t v[16];
    tmem_ld_32x32b_x16(taddr, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= v[i];
  }
  return acc;
}

template <int N>
__global__ void __launch_bounds__(256) probe(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 256);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  for (int i = 0; i < 4; ++i) acc ^= ld_n<N>(tb);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) acc ^= ld_n<N>(tb + (i & 1) * 64);
  const long long t1 = clock64();
  if (acc == 0x12345678) sink[0] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 256); }
}

int main() {
  long long* d; uint32_t* sink; cudaMalloc(&d, 8); cudaMalloc(&sink, 4);
  const int iters = 4096;
  for (int ctas : {148, 296})
    for (int warps : {1, 4, 8}) {
      for (int n : {16, 64}) {
        if (n == 16) probe<16><<<ctas, warps * 32>>>(iters, d, sink); else probe<64><<<ctas, warps * 32>>>(iters, d, sink);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
        printf("ctas=%3d warps/CTA=%d  ld.x%-2d : %8.1f cycles per ld+wait (+%d XORs)  => %.1f B/clk per warp\n", ctas, warps, n, (double)c / iters, n,
               32.0 * n * 4 / ((double)c / iters));
      }
    }
  return 0;
}

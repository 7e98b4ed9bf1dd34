// Microbenchmark: cycles per tcgen05.mma for the shapes the projection GEMM / implicit-GEMM convolution issue,
//   (a) cta_group::1, M = 128, N in {128, 160, 192, 256}, operands from shared memory (SS), one CTA per SM;
//   (b) cta_group::2, M = 256 (128 rows per CTA of a 2-CTA cluster), same N: one thread of the leader CTA issues for both SMs,
//       every CTA holds its 128 rows of A and HALF of the B rows (N / 2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../adaface-dev_b200/csrc probe_mma_pair.cu -o probe_mma_pair
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace adaface;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

template <int PAIR>
__global__ void __launch_bounds__(128) probe(int N, int iters, long long* out, int nacc) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc(&slot, 512);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 1 && rank == 0 && elect_one()) {
    const int M = PAIR ? 256 : 128;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t aA = smem_u32(smem), aB = smem_u32(smem + 32768);
    for (int rep = 0; rep < 2; ++rep) {     // rep 0 = warm-up
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint64_t da = make_smem_desc_sw128(aA + (i & 3) * 32 + ((i >> 2) & 1) * 16384);
        const uint64_t db = make_smem_desc_sw128(aB + (i & 3) * 32 + ((i >> 2) & 1) * 32768);
        const uint32_t td = tb + (uint32_t)((i / 20) % nacc) * 256;      // 20 MMAs per accumulator, then the other one
        if (PAIR) umma_bf16_2cta(td, da, db, idesc, 1);
        else umma_bf16(td, da, db, idesc, 1);
      }
      if (PAIR) umma_commit_2cta(&bar, 1);
      else umma_commit(&bar);
      mbar_wait(&bar, rep);
      const long long t1 = clock64();
      if (blockIdx.x == 0 && rep == 1) out[0] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
    else tmem_dealloc(tb, 512);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 4000;
  printf("%-28s %4s %5s %10s %12s\n", "mode", "N", "nacc", "cyc/mma", "FLOP/clk/SM");
  for (int pair = 0; pair < 2; ++pair)
    for (int nacc : {1, 2})
      for (int N : {16, 32, 48, 64, 96, 128, 160, 192, 256}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = 100 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = pair ? 2 : 1;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaError_t e = pair ? cudaLaunchKernelEx(&cfg, probe<1>, N, iters, d, nacc) : cudaLaunchKernelEx(&cfg, probe<0>, N, iters, d, nacc);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        long long cy = 0;
        cudaMemcpy(&cy, d, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
          printf("ERROR %s\n", cudaGetErrorString(e));
          return 1;
        }
        const double c = (double)cy / iters;
        printf("%-28s %4d %5d %10.1f %12.0f\n", pair ? "cta_group::2 M=256 (pair)" : "cta_group::1 M=128", N, nacc, c, 2.0 * 128 * N * 16 / c);
      }
  return 0;
}

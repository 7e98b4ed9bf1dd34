// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16 -> fp32) as a function of M, N and operand source
// (A from shared memory "SS" vs A from tensor memory "TS"), one CTA per SM, back-to-back issue from one thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../adaface-dev_b200/csrc probe_mma_rate.cu -o probe_mma_rate
#include <cstdio>
#include "common.cuh"
using namespace adaface;

__global__ void __launch_bounds__(128) probe(int M, int N, int ts, int b_mn, int iters, long long* out, int tmem_cols, int issuers) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar2[4];
  uint64_t& bar = bar2[threadIdx.x >> 5];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i % 7;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar2[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, tmem_cols);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (warp >= 1 && warp <= issuers && elect_one()) {
    const uint32_t tb = slot + (warp - 1) * 64;   // separate accumulators per issuing warp
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t aA = smem_u32(smem), aB = smem_u32(smem + 32768);
    const uint32_t ta = slot + 256 - 64;
    // warm-up
    for (int i = 0; i < 8; ++i) {
      const uint64_t da = make_smem_desc_sw128(aA + (i & 3) * 32);
      const uint64_t db = b_mn ? make_smem_desc_sw128_mn(aB + (i & 3) * 2048, 8192) : make_smem_desc_sw128(aB + (i & 3) * 32);
      if (ts) umma_bf16_ts(tb, ta + (i & 3) * 8, db, idesc, 1); else umma_bf16(tb, da, db, idesc, 1);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint64_t da = make_smem_desc_sw128(aA + (i & 3) * 32);
      const uint64_t db = b_mn ? make_smem_desc_sw128_mn(aB + (i & 3) * 2048, 8192) : make_smem_desc_sw128(aB + (i & 3) * 32);
      if (ts) umma_bf16_ts(tb, ta + (i & 3) * 8, db, idesc, 1); else umma_bf16(tb, da, db, idesc, 1);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 1);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && warp == 1) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, tmem_cols); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2048;
  printf("%4s %4s %3s %4s %12s\n", "M", "N", "src", "Bmaj", "cyc/mma");
  struct Cfg { int ctas, cols, issuers; const char* name; };
  for (Cfg c : {Cfg{148, 256, 1, "1 CTA/SM, 1 issuer"}, Cfg{296, 256, 1, "2 CTA/SM, 1 issuer each"}, Cfg{148, 256, 2, "1 CTA/SM, 2 issuing warps"},
                Cfg{148, 256, 3, "1 CTA/SM, 3 issuing warps"}}) {
    printf("-- %s\n", c.name);
    for (int M : {128}) for (int ts : {0, 1}) for (int bmn : {0}) for (int N : {48, 64, 128}) {
      probe<<<c.ctas, 128, 100 * 1024>>>(M, N, ts, bmn, iters, d, c.cols, c.issuers);
      cudaError_t e = cudaDeviceSynchronize();
      long long cy = 0; cudaMemcpy(&cy, d, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
      printf("%4d %4d %3s %4s %12.1f cyc per mma per issuer\n", M, N, ts ? "TS" : "SS", bmn ? "MN" : "K", (double)cy / iters);
    }
  }
  return 0;
}

// Microbenchmark: TMA smem-fill bandwidth of ONE CTA per SM as a function of the bytes in flight (ring depth x box size), for the
// GEMM's operand boxes: [rows, 64] bf16 boxes (128-byte swizzle) of a row-major [M, K] matrix, K = 320 / 1280, either L2-resident
// (every CTA loops over the same 2 MB) or streamed from HBM (each CTA walks its own part of a 1 GB matrix).
#include <cstdio>
#include <cuda.h>
#include "common.cuh"
using namespace adaface;
__device__ __forceinline__ void wait_mode(uint64_t* bar, uint32_t parity, int mode) {
  if (mode == 0) { mbar_wait(bar, parity); return; }
  uint32_t ok = 0;
  while (!ok) {
    if (mode == 1)
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    else
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tm, int iters, int depth, int box_rows, int kblocks, int m_tiles_per_cta,
                                            int stream, long long* out, int nwarps, int mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_all[64];
  if (threadIdx.x == 0) { for (int i = 0; i < 64; ++i) mbar_init(&bar_all[i], 1); fence_barrier_init(); }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if (warp < nwarps && elect_one()) {
    uint64_t* bar = bar_all + warp * 16;
    smem += warp * depth * box_rows * 128;
    const int bytes = box_rows * 128;
    const long long t0 = clock64();
    for (int i = 0; i < iters + depth; ++i) {
      const int s = i % depth;
      if (i >= depth) wait_mode(&bar[s], ((i - depth) / depth) & 1, mode);
      if (i < iters) {
        const int kb = i % kblocks, mt = (i / kblocks + warp * 3) % m_tiles_per_cta;
        const int row = (stream ? blockIdx.x * m_tiles_per_cta + mt : mt) * box_rows;
        mbar_arrive_expect_tx(&bar[s], bytes);
        tma_load_2d(smem + s * bytes, &tm, &bar[s], kb * 64, row);
      }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && warp == 0) out[0] = t1 - t0;
  }
}

int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  const size_t bytes_total = 1ull << 30;
  void* buf; cudaMalloc(&buf, bytes_total); cudaMemset(buf, 1, bytes_total);
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int K : {320}) for (int box_rows : {128, 256}) for (int promo : {1}) {
    const uint64_t M = bytes_total / (K * 2);
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)K, M}; cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    for (int stream : {0, 1}) for (int mode : {0, 1, 2}) for (int nwarps : {1, 2}) for (int depth : {4, 8}) {
      if (nwarps * depth * box_rows * 128 > 200 * 1024) continue;
      const int kblocks = K / 64, iters = 2000;
      const int mt = stream ? (int)(M / box_rows / 148) : 8;
      probe<<<148, 128, 220 * 1024>>>(tm, iters, depth, box_rows, kblocks, mt, stream, d, nwarps, mode);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
      const double per_box = (double)c / iters;
      printf("K=%4d box=%3dx64 %-12s mode=%d issuers=%d depth=%2d (%3d KB in flight): %7.1f cyc/box/issuer  %6.1f B/clk/SM\n", K, box_rows,
             stream ? "HBM stream" : "L2 resident", mode, nwarps, depth, nwarps * depth * box_rows * 128 / 1024, per_box, nwarps * box_rows * 128.0 / per_box);
    }
  }
  return 0;
}

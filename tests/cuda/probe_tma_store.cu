// Microbenchmark: TMA bulk tensor STORE cost per box as a function of the box shape (rows x 64- or 128-byte rows), one issuing
// thread per CTA, one CTA per SM, 64 boxes issued back to back from ONE smem slab (reads may overlap), each its own bulk group.
#include <cstdio>
#include <cuda.h>
#include "common.cuh"
using namespace adaface;
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void __launch_bounds__(64) probe(const __grid_constant__ CUtensorMap tm, int box_rows, int box_cols, int nbox, int N, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x; i < 32 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x < 32 && elect_one()) {
    const int per_row = N / box_cols;
    for (int rep = 0; rep < 2; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < nbox; ++i) {
        const int b = blockIdx.x * nbox + i;
        tma_store_2d(&tm, smem, (b % per_row) * box_cols, (b / per_row) * box_rows);
        tma_store_commit();
      }
      const long long t1 = clock64();
      tma_store_wait_read<0>();
      const long long t2 = clock64();
      tma_store_wait_all();
      const long long t3 = clock64();
      if (blockIdx.x == 0 && rep == 1) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = t3 - t0; }
    }
  }
}
int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  const int N = 960; const uint64_t M = 1 << 18;
  void* buf; cudaMalloc(&buf, M * N * 2);
  long long* d; cudaMalloc(&d, 24);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  struct S { int rows, cols; CUtensorMapSwizzle sw; const char* name; };
  for (S s : {S{32, 32, CU_TENSOR_MAP_SWIZZLE_64B, "32x32 sw64 (2 KB)"}, S{32, 64, CU_TENSOR_MAP_SWIZZLE_128B, "32x64 sw128 (4 KB)"},
              S{64, 64, CU_TENSOR_MAP_SWIZZLE_128B, "64x64 sw128 (8 KB)"}, S{128, 32, CU_TENSOR_MAP_SWIZZLE_64B, "128x32 sw64 (8 KB)"},
              S{128, 64, CU_TENSOR_MAP_SWIZZLE_128B, "128x64 sw128 (16 KB)"}, S{32, 192, CU_TENSOR_MAP_SWIZZLE_NONE, "32x192 none (12 KB)"},
              S{128, 64, CU_TENSOR_MAP_SWIZZLE_NONE, "128x64 none (16 KB)"}}) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)N, M}; cuuint64_t gstr[1] = {(cuuint64_t)N * 2};
    cuuint32_t box[2] = {(cuuint32_t)s.cols, (cuuint32_t)s.rows}, es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, s.sw,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d for %s\n", (int)r, s.name); continue; }
    for (int ctas : {1, 148}) {
      const int nbox = 64;
      probe<<<ctas, 64, 40 * 1024>>>(tm, s.rows, s.cols, nbox, N, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long c[3]; cudaMemcpy(c, d, 24, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
      const double bytes = 2.0 * s.rows * s.cols;
      printf("%-24s ctas=%3d: issue %6.1f clk/box | smem read done %7.1f clk/box (%5.1f B/clk/SM) | all written %7.1f clk/box (%5.1f B/clk/SM)\n", s.name, ctas,
             (double)c[0] / nbox, (double)c[1] / nbox, bytes * nbox / c[1], (double)c[2] / nbox, bytes * nbox / c[2]);
    }
  }
  return 0;
}

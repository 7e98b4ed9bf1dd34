// Microbenchmark: (1) cycles the issuing thread spends per cp.async.bulk.tensor instruction (12 boxes issued back to back, no
// waits), (2) cycles until the first / the last of them has landed, (3) steady-state loop cost with and without the wait.
#include <cstdio>
#include <cuda.h>
#include "common.cuh"
using namespace adaface;
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(64) probe(const __grid_constant__ CUtensorMap tm, int box_rows, int nbox, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[16];
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x < 32 && elect_one()) {
    const int bytes = box_rows * 128;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < nbox; ++i) {
        mbar_arrive_expect_tx(&bar[i], bytes);
        tma_load_2d(smem + i * bytes, &tm, &bar[i], (i % 5) * 64, (blockIdx.x * 4 + rep) * box_rows);
      }
      const long long t1 = clock64();
      mbar_wait(&bar[0], rep & 1);
      const long long t2 = clock64();
      for (int i = 1; i < nbox; ++i) mbar_wait(&bar[i], rep & 1);
      const long long t3 = clock64();
      if (blockIdx.x == 0) { out[rep * 3 + 0] = t1 - t0; out[rep * 3 + 1] = t2 - t0; out[rep * 3 + 2] = t3 - t0; }
    }
  }
}

int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  const size_t bytes_total = 1ull << 28; const int K = 320;
  void* buf; cudaMalloc(&buf, bytes_total); cudaMemset(buf, 1, bytes_total);
  long long* d; cudaMalloc(&d, 9 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int box_rows : {64, 128, 256}) {
    const uint64_t M = bytes_total / (K * 2);
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)K, M}; cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
    enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int ctas : {1, 148}) for (int nbox : {1, 2, 4, 6}) {
      if (nbox * box_rows * 128 > 200 * 1024) continue;
      probe<<<ctas, 64, 220 * 1024>>>(tm, box_rows, nbox, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long c[9]; cudaMemcpy(c, d, 72, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
      printf("box=%3dx64 ctas=%3d nbox=%d : issue %5lld clk total, first landed %5lld, all landed %5lld  (rep 1; rep 2: %lld %lld %lld)\n", box_rows, ctas, nbox,
             c[3], c[4], c[5], c[6], c[7], c[8]);
    }
  }
  return 0;
}

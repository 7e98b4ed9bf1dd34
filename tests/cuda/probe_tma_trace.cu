// Trace of a single-thread TMA producer ring: per iteration, cycles spent in the mbarrier wait and in the issue.
#include <cstdio>
#include <cuda.h>
#include "common.cuh"
using namespace adaface;
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int NT = 40;
__global__ void __launch_bounds__(64) probe(const __grid_constant__ CUtensorMap tm, int depth, int box_rows, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[16];
  __shared__ long long tr[NT * 3];
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x < 32 && elect_one()) {
    const int bytes = box_rows * 128;
    const long long t00 = clock64();
    for (int i = 0; i < NT + depth; ++i) {
      const int s = i % depth;
      const long long ta = clock64();
      if (i >= depth) mbar_wait(&bar[s], ((i - depth) / depth) & 1);
      const long long tb = clock64();
      if (i < NT) {
        mbar_arrive_expect_tx(&bar[s], bytes);
        tma_load_2d(smem + s * bytes, &tm, &bar[s], (i % 5) * 64, (blockIdx.x * 16 + i / 5) * box_rows);
        const long long tc = clock64();
        tr[i * 3] = ta - t00; tr[i * 3 + 1] = tb - t00; tr[i * 3 + 2] = tc - t00;
      }
    }
    if (blockIdx.x == 0) for (int i = 0; i < NT * 3; ++i) out[i] = tr[i];
  }
}
int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  const size_t bytes_total = 1ull << 28; const int K = 320;
  void* buf; cudaMalloc(&buf, bytes_total); cudaMemset(buf, 1, bytes_total);
  long long* d; cudaMalloc(&d, NT * 3 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int box_rows = 128;
  const uint64_t M = bytes_total / (K * 2);
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)K, M}; cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, es[2] = {1, 1};
  enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  for (int ctas : {1, 148}) for (int depth : {4, 8}) {
    for (int rep = 0; rep < 2; ++rep) probe<<<ctas, 64, 220 * 1024>>>(tm, depth, box_rows, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[NT * 3]; cudaMemcpy(c, d, sizeof(c), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
    printf("== ctas=%d depth=%d: i: t_wait_begin t_wait_end t_issued\n", ctas, depth);
    for (int i = 0; i < NT; ++i) printf("%2d: %6lld %6lld %6lld\n", i, c[i * 3], c[i * 3 + 1], c[i * 3 + 2]);
  }
  return 0;
}

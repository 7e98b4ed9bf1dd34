// Microbenchmark: TMA tile-load rate per SM for the K/V tile shapes of the attention kernel
// (64 rows x 64-element box, bf16, 128B swizzle) from an L2-resident tensor, for three global layouts.
#include <cstdio>
#include <vector>
#include <cuda.h>
#include "common.cuh"
using namespace adaface;

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(64) probe(const __grid_constant__ CUtensorMap tm, int iters, int rows_total, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4];
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    const int head = blockIdx.x % 8, batch = (blockIdx.x / 8) % 8;
    for (int i = 0; i < iters + 4; ++i) {
      const int s = i & 3;
      if (i >= 4) mbar_wait(&bar[s], ((i - 4) >> 2) & 1);          // stage s landed: reuse it
      if (i < iters) {
        mbar_arrive_expect_tx(&bar[s], 64 * 128);
        tma_load_4d(smem + s * 8192, &tm, &bar[s], 0, head, (i * 64) % rows_total, batch);
      }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
}

int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  const uint64_t B = 8, H = 8, N = 4096;
  void* buf; cudaMalloc(&buf, B * N * 3 * 8 * 64 * 2 + (1 << 20)); cudaMemset(buf, 1, B * N * 3 * 8 * 64 * 2);
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  struct L { const char* name; uint64_t dd, sh, sn, sb; };
  // element strides: head, token, batch
  L layouts[] = {{"interleaved [B,N,3*H*40] (d=40, OOB fill)", 40, 40, 960, N * 960},
                 {"head-major dense [B,H,N,40] (OOB fill)", 40, N * 40, 40, H * N * 40},
                 {"head-major padded [B,H,N,64] (full rows)", 64, N * 64, 64, H * N * 64},
                 {"interleaved [B,N,3*H*80] (d=80 first atom)", 80, 80, 1920, N * 1920}};
  for (auto& l : layouts) {
    CUtensorMap tm;
    cuuint64_t gdim[4] = {l.dd, H, N, B};
    cuuint64_t gstr[3] = {l.sh * 2, l.sn * 2, l.sb * 2};
    cuuint32_t box[4] = {64, 1, 64, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d for %s\n", (int)r, l.name); continue; }
    for (int ctas : {148, 296, 592}) {
      const int iters = 2048;
      probe<<<ctas, 64, 40 * 1024>>>(tm, iters, (int)N, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
      const double per_box = (double)c / iters;
      const double useful = 64.0 * (l.dd < 64 ? l.dd : 64) * 2;
      printf("%-46s ctas/SM=%d : %7.1f cyc per 64-row box per CTA -> %6.1f useful B/clk/SM\n", l.name, ctas / 148, per_box,
             useful * (ctas / 148) / per_box);
    }
  }
  return 0;
}

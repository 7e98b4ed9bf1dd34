// Microbenchmark: per-SM throughput of the instructions the softmax inner loop is made of.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) { uint32_t y; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }

template <int OP>
__global__ void k(int iters, float* out, long long* cyc) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i * 0.01f - 1.f;
  float2 s2 = make_float2(0.f, 0.f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) a[i] = ex2(a[i]);                                           // MUFU.EX2
      if (OP == 1) a[i] = fmaf(a[i], 1.0001f, -0.5f);                          // FFMA
      if (OP == 2) a[i] = __uint_as_float(__float_as_uint(a[i]) & 0xffff0ff0u) ; // LOP3
      if (OP == 3) a[i] = fmaxf(a[i], a[(i + 1) & 15]);                        // FMNMX
      if (OP == 4) { float2 t = __ffma2_rn(make_float2(a[i], a[(i + 1) & 15]), make_float2(1.0001f, 1.0001f), make_float2(-0.5f, -0.5f)); a[i] = t.x; a[(i + 1) & 15] = t.y; ++i; }  // FFMA2
      if (OP == 5) { a[i] = ex2(a[i]); a[i] = fmaf(a[i], 0.5f, -1.f); }        // MUFU + FFMA mix
      if (OP == 6) a[i] = __uint_as_float(__byte_perm(__float_as_uint(a[i]), __float_as_uint(a[(i + 1) & 15]), 0x7632)); // PRMT
      if (OP == 7) a[i] = __uint_as_float(ex2_bf16x2(__float_as_uint(a[i])));   // MUFU.EX2 bf16x2 (2 results per lane-op)
      if (OP == 8) a[i] = __uint_as_float(ex2_f16x2(__float_as_uint(a[i])));    // MUFU.EX2 f16x2
      if (OP == 9) a[i] = __uint_as_float(cvt_bf16x2(a[i], a[(i + 1) & 15]));   // F2FP pack
      if (OP == 10) a[i] = fmaxf(fmaxf(a[i], a[(i + 1) & 15]), a[(i + 2) & 15]); // FMNMX3 ?
      if (OP == 11) a[i] = __uint_as_float(__float_as_uint(a[i]) + 0x8000u);    // IADD
      if (OP == 12) {   // candidate softmax inner step for one PAIR: FFMA2, 2 x IADD, PRMT, EX2.bf16x2
        float2 t = __ffma2_rn(make_float2(a[i], a[(i + 1) & 15]), make_float2(1.0001f, 1.0001f), make_float2(-0.5f, -0.5f));
        const uint32_t u = __byte_perm(__float_as_uint(t.x) + 0x8000u, __float_as_uint(t.y) + 0x8000u, 0x7632);
        a[i] = __uint_as_float(ex2_bf16x2(u)); a[(i + 1) & 15] = t.y * 0.f + a[(i + 1) & 15]; ++i; }
    }
  }
  const long long t1 = clock64();
  float s = s2.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const char* names[] = {"MUFU.EX2", "FFMA", "LOP3", "FMNMX", "FFMA2(pairs)", "EX2+FFMA", "PRMT", "EX2.bf16x2", "EX2.f16x2", "F2FP.bf16x2", "FMNMX x2", "IADD", "pair-step"};
  const int iters = 2000;
  for (int warps : {8, 16, 32}) {
    for (int op = 0; op < 13; ++op) {
      switch (op) {
        case 0: k<0><<<148, warps * 32>>>(iters, out, cyc); break;
        case 1: k<1><<<148, warps * 32>>>(iters, out, cyc); break;
        case 2: k<2><<<148, warps * 32>>>(iters, out, cyc); break;
        case 3: k<3><<<148, warps * 32>>>(iters, out, cyc); break;
        case 4: k<4><<<148, warps * 32>>>(iters, out, cyc); break;
        case 5: k<5><<<148, warps * 32>>>(iters, out, cyc); break;
        case 6: k<6><<<148, warps * 32>>>(iters, out, cyc); break;
        case 7: k<7><<<148, warps * 32>>>(iters, out, cyc); break;
        case 8: k<8><<<148, warps * 32>>>(iters, out, cyc); break;
        case 9: k<9><<<148, warps * 32>>>(iters, out, cyc); break;
        case 10: k<10><<<148, warps * 32>>>(iters, out, cyc); break;
        case 11: k<11><<<148, warps * 32>>>(iters, out, cyc); break;
        case 12: k<12><<<148, warps * 32>>>(iters, out, cyc); break;
      }
      cudaDeviceSynchronize();
      long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double ops = (double)iters * 16 * warps * 32;   // lane-ops per SM (FFMA2 counts 2 lanes-ops per instr-lane)
      printf("warps/SM=%2d %-13s %7.1f lane-ops/clk/SM\n", warps, names[op], ops / c);
    }
  }
  return 0;
}

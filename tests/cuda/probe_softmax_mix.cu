// Microbenchmark: the softmax warp's step of the level-A attention kernel WITHOUT any hand-off -- TMEM read of 64 scores, (row maximum),
// (scale / shift FFMA2), exp2 (MUFU, part on the FMA-pipe polynomial), bf16 pack, TMEM store -- on 1..4 warps per scheduler.
// Answers: how many clocks does a scheduler need per warp-step for a given instruction mix, i.e. what is the compute floor of the
// kernel's period (4 tiles -> 4 warp-steps per scheduler per 64-key step)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --use_fast_math -I adaface-dev_b200/csrc -o build/probe_softmax_mix tests/cuda/probe_softmax_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include "common.cuh"
using namespace adaface;

template <int DEG>
__device__ __forceinline__ float2 emu2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.f, 12582912.f));
  const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);
  float2 q = __ffma2_rn(f, make_float2(0.23842894f, 0.23842894f), make_float2(0.7034480f, 0.7034480f));
  q = __ffma2_rn(q, f, make_float2(1.0004431f, 1.0004431f));
  float2 r;
  r.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
  return r;
}

// MAXP: row-maximum pass; SCALE: FFMA2 scale / shift in front of the exp2; EMU: pairs of every 8 on the polynomial; ORC: OR-reduce the packed words
template <bool MAXP, bool SCALE, int EMU, bool ORC, bool TM>
__global__ void __launch_bounds__(512, 1) k(int iters, float scale, float* out, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
  float m_ref = 0.5f, acc = 0.f;
  uint32_t v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = __float_as_uint(-0.01f * i - 0.001f * threadIdx.x);
  if (TM) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = v[c * 16 + i];
      tmem_st_32x32b_x16(tb + c * 16, w);
    }
    tmem_st_wait();
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (TM) tmem_ld_32x32b_x64_wait(tb, v);
    if (MAXP) {
      float m4[4] = {__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])};
#pragma unroll
      for (int i = 4; i < 64; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) m4[u] = fmaxf(m4[u], __uint_as_float(v[i + u]));
      }
      const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * scale;
      if (mx > m_ref + 8.f) m_ref = mx;
    }
    const float2 sc2 = make_float2(scale, scale), nm2 = make_float2(-m_ref, -m_ref);
    uint32_t any = 0;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float2 t = make_float2(__uint_as_float(v[hf * 32 + 2 * i]), __uint_as_float(v[hf * 32 + 2 * i + 1]));
        if (SCALE) t = __ffma2_rn(t, sc2, nm2);
        const float2 e = ((i & 7) < EMU) ? emu2<2>(t) : make_float2(fast_exp2(t.x), fast_exp2(t.y));
        pk[i] = __byte_perm(__float_as_uint(e.x), __float_as_uint(e.y), 0x7632);
      }
      if (ORC) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) any |= pk[i] | pk[i + 1];
      }
      if (TM) {
        tmem_st_32x32b_x16(tb + 64 + hf * 16, pk);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[hf * 32 + i] = (pk[i] & 0x3fff0000u) | 0x80000000u | (v[hf * 32 + i] & 0xffffu);      // keep the loop live: feed the results back as (negative) scores
      }
    }
    if (ORC && (any & 0x40004000u)) m_ref += 1.f;
    if (TM) tmem_st_wait();
    acc += m_ref;
  }
  const long long t1 = clock64();
  float s = acc;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += __uint_as_float(v[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

template <bool MAXP, bool SCALE, int EMU, bool ORC, bool TM>
void run(const char* name, float* out, long long* cyc) {
  const int iters = 2000;
  printf("%-58s", name);
  for (int warps : {4, 8, 12, 16}) {
    k<MAXP, SCALE, EMU, ORC, TM><<<148, warps * 32>>>(iters, 0.2281f, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf(" ERR %s", cudaGetErrorString(e)); break; }
    // clocks per warp-step per scheduler: (total clocks / iterations) / (warps per scheduler)
    printf("  %2d warps: %6.1f clk/step (%5.1f per warp-step)", warps, (double)c / iters, (double)c / iters / (warps / 4));
  }
  printf("\n");
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  run<true, true, 2, false, true>("kernel's step: max + FFMA2 + exp2 (2/8 emulated) + TMEM", out, cyc);
  run<true, true, 0, false, true>("  same, no emulation", out, cyc);
  run<true, true, 3, false, true>("  same, 3/8 emulated", out, cyc);
  run<false, true, 2, true, true>("no max pass, OR check: FFMA2 + exp2 (2/8) + TMEM", out, cyc);
  run<false, false, 2, true, true>("no max, no FFMA2: exp2 (2/8) on raw scores + TMEM", out, cyc);
  run<false, false, 0, true, true>("  same, no emulation", out, cyc);
  run<false, false, 3, true, true>("  same, 3/8 emulated", out, cyc);
  run<true, true, 2, false, false>("kernel's step without the TMEM traffic (registers only)", out, cyc);
  run<false, false, 0, false, false>("MUFU + PRMT only (registers only)", out, cyc);
  return 0;
}

// K1: projection GEMM on tcgen05 / TMEM, operands fed by TMA, LoRA/DoRA folded into the same tile.
//
//   Y[M,N] = act( colscale[N] * ( X[M,K] W[N,K]^T + T[M,R] Bs[N,R]^T ) + bias[N] ) + residual[M,N]
//
// One CTA computes a 128 x BN output tile.  Warp roles (192 threads):
//   warp 0    TMA producer: streams 64-wide K slabs of X/W (then of T/Bs -- the rank-R LoRA tail simply
//             continues the same accumulation) into a STAGES-deep 128B-swizzled smem ring.
//   warp 1    allocates TMEM, then one elected lane issues tcgen05.mma (M=128, N=BN, K=16) per 32-byte K step;
//             tcgen05.commit releases smem stages and finally signals the epilogue.
//   warps 2-5 epilogue: tcgen05.ld the fp32 accumulator (one row per thread, 32 TMEM lanes per warp),
//             apply DoRA column scale, bias, activation, residual, convert, store.
// Reference arithmetic: nn.Linear / peft lora.Linear(+DoRA) call sites dalc:235-249, 280-288, 328-331
// (SURVEY.md 8a rows A1, A4); ldm/modules/attention.py:31-58, 156-164; HF CLIP q/k/v/out_proj, fc1, fc2.
#include "common.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;          // 64 bf16 = 128 B = one swizzle span
constexpr int GEMM_THREADS = 192;
constexpr int A_STAGE_BYTES = GEMM_BM * GEMM_BK * 2;

struct GemmEpilogue {
  const float* colscale;
  const float* bias;
  const void* residual;
  void* y;
  long long ldr, ldy;
  int M, N;
  int num_kb1, num_kb2;
  int act, y_f32, res_f32;
  // head-scatter store (hs_d > 0): output column c = which*hs_C + h*hs_d + dd of row b*hs_rows + n goes to
  // y[which][b][h][n][dd] with rows padded to hs_dpad elements -- the head-major, 128-byte-row layout the
  // attention kernel's TMA loads want (TMA boxes that run out of bounds inside a row are ~3x slower).
  int hs_d, hs_dpad, hs_C, hs_H, hs_rows, hs_B;
};

template <int BN>
struct GemmCfg {
  static constexpr int B_STAGE_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN <= 160) ? 3 : 4;   // <=110 KB: two CTAs co-reside per SM (epilogue/mainloop overlap)
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_tn_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                        const __grid_constant__ CUtensorMap tmB,
                                                                        const __grid_constant__ CUtensorMap tmA2,
                                                                        const __grid_constant__ CUtensorMap tmB2,
                                                                        const GemmEpilogue ep) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * GEMM_BM;
  const int num_kb = ep.num_kb1 + ep.num_kb2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (ep.num_kb2 > 0) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  } else if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {   // elect.sync: ptxas then knows one thread is active -> plain R2UR, no waterfall loops
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + A_STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        if (kb < ep.num_kb1) {
          tma_load_2d(sa, &tmA, &full_bar[s], kb * GEMM_BK, m0);
          tma_load_2d(sb, &tmB, &full_bar[s], kb * GEMM_BK, n0);
        } else {
          tma_load_2d(sa, &tmA2, &full_bar[s], (kb - ep.num_kb1) * GEMM_BK, m0);
          tma_load_2d(sb, &tmB2, &full_bar[s], (kb - ep.num_kb1) * GEMM_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {   // elect.sync: ptxas then knows one thread is active -> plain R2UR, no waterfall loops
      constexpr uint32_t idesc = make_idesc_bf16_f32(GEMM_BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          // advancing 16 K-elements inside the swizzle span = +32 bytes on the start address
          const uint64_t da = make_smem_desc_sw128(sa + k * 32);
          const uint64_t db = make_smem_desc_sw128(sb + k * 32);
          umma_bf16(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);   // smem stage reusable once these MMAs have read it
      }
      umma_commit(tmem_full_bar);     // accumulator complete
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const bool row_ok = row < ep.M;
    const bool geglu = ep.act == ADAFACE_ACT_GEGLU;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      if (n0 + c0 >= ep.N && !geglu) break;           // warp-uniform
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      float f[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = n0 + c0 + j;
        float a = __uint_as_float(v[j]);
        if (col < ep.N) {
          if (ep.colscale) a *= __ldg(ep.colscale + col);
          if (ep.bias) a += __ldg(ep.bias + col);
        }
        f[j] = a;
      }
      int out_col0 = n0 + c0;          // first output column of this 16-wide group
      int out_n = ep.N;
      if (geglu) {
        // tile columns [0, BN/2) are the "a" half, [BN/2, BN) the gates of the same output columns.
        if (c0 >= BN / 2) break;
        uint32_t g[16];
        tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c0 + BN / 2), g);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = n0 + c0 + BN / 2 + j;
          float gt = __uint_as_float(g[j]);
          if (col < ep.N) {
            if (ep.colscale) gt *= __ldg(ep.colscale + col);
            if (ep.bias) gt += __ldg(ep.bias + col);
          }
          f[j] = f[j] * gelu_erf(gt);
        }
        out_col0 = blockIdx.x * (BN / 2) + c0;
        out_n = ep.N / 2;
      } else if (ep.act == ADAFACE_ACT_QUICK_GELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = f[j] / (1.f + __expf(-1.702f * f[j]));
      }
      if (!row_ok) continue;
      if (ep.residual) {
        if (ep.res_f32) {
          const float* r = reinterpret_cast<const float*>(ep.residual) + (long long)row * ep.ldr + out_col0;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (out_col0 + j < out_n) f[j] += r[j];
        } else {
          const bf16* r = reinterpret_cast<const bf16*>(ep.residual) + (long long)row * ep.ldr + out_col0;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (out_col0 + j < out_n) f[j] += __bfloat162float(r[j]);
        }
      }
      if (ep.y_f32) {
        float* y = reinterpret_cast<float*>(ep.y) + (long long)row * ep.ldy + out_col0;
        if (out_col0 + 16 <= out_n && ((reinterpret_cast<uintptr_t>(y) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(y + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (out_col0 + j < out_n) y[j] = f[j];
        }
      } else if (ep.hs_d > 0) {
        const int bb = row / ep.hs_rows, nn = row - bb * ep.hs_rows;
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          const int col = out_col0 + g8 * 8;             // 8-column groups never straddle a head (hs_d % 8 == 0)
          if (col < out_n) {
            const int which = col / ep.hs_C, rem = col - which * ep.hs_C;
            const int hh = rem / ep.hs_d, dd = rem - hh * ep.hs_d;
            bf16* dst = reinterpret_cast<bf16*>(ep.y) +
                        ((((long long)which * ep.hs_B + bb) * ep.hs_H + hh) * ep.hs_rows + nn) * ep.hs_dpad + dd;
            *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(f[g8 * 8 + 0], f[g8 * 8 + 1]), pack_bf16(f[g8 * 8 + 2], f[g8 * 8 + 3]),
                                                        pack_bf16(f[g8 * 8 + 4], f[g8 * 8 + 5]), pack_bf16(f[g8 * 8 + 6], f[g8 * 8 + 7]));
          }
        }
      } else {
        bf16* y = reinterpret_cast<bf16*>(ep.y) + (long long)row * ep.ldy + out_col0;
        if (out_col0 + 16 <= out_n && ((reinterpret_cast<uintptr_t>(y) & 15) == 0)) {
          uint4 p0 = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
          uint4 p1 = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]),
                                pack_bf16(f[14], f[15]));
          *reinterpret_cast<uint4*>(y) = p0;
          *reinterpret_cast<uint4*>(y + 8) = p1;
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (out_col0 + j < out_n) y[j] = __float2bfloat16(f[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
extern long long g_launch_count;

template <int BN>
static int launch_gemm(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tA2, const CUtensorMap& tB2,
                       const GemmEpilogue& ep, int n_tiles, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    AF_CUDA(cudaFuncSetAttribute(gemm_tn_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg::SMEM_BYTES));
    configured = true;
  }
  dim3 grid(n_tiles, (ep.M + GEMM_BM - 1) / GEMM_BM);
  gemm_tn_tcgen05_kernel<BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tA, tB, tA2, tB2, ep);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

int proj_lora_fwd(const void* x, int64_t ldx, const void* w, const void* t, int64_t ldt, const void* bs,
                  const float* colscale, const float* bias, const void* residual, int64_t ldr, int residual_dtype,
                  void* y, int64_t ldy, int y_dtype, int64_t M, int64_t N, int64_t K, int64_t R, int act,
                  int64_t hs_heads, int64_t hs_d, int64_t hs_dpad, int64_t hs_rows, cudaStream_t stream) {
  AF_CHECK(x && w && y, "proj_lora_fwd: null x/w/y");
  AF_CHECK(M > 0 && N > 0 && K > 0, "proj_lora_fwd: empty problem M=%lld N=%lld K=%lld", (long long)M, (long long)N,
           (long long)K);
  AF_CHECK(K % 8 == 0 && ldx % 8 == 0, "proj_lora_fwd: K (%lld) and ldx (%lld) must be multiples of 8", (long long)K,
           (long long)ldx);
  AF_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
           "proj_lora_fwd: x / w must be 16-byte aligned");
  const bool lora = (t != nullptr) || (bs != nullptr) || R > 0;
  if (lora) {
    AF_CHECK(t && bs && R > 0, "proj_lora_fwd: t, bs and R must be given together");
    AF_CHECK(R % 8 == 0 && ldt % 8 == 0, "proj_lora_fwd: R (%lld) and ldt (%lld) must be multiples of 8", (long long)R,
             (long long)ldt);
    AF_CHECK((reinterpret_cast<uintptr_t>(t) & 15) == 0 && (reinterpret_cast<uintptr_t>(bs) & 15) == 0,
             "proj_lora_fwd: t / bs must be 16-byte aligned");
  }
  AF_CHECK(act >= 0 && act <= 2, "proj_lora_fwd: bad act %d", act);
  if (hs_d > 0) {
    AF_CHECK(y_dtype == ADAFACE_BF16 && act != ADAFACE_ACT_GEGLU && !residual, "proj_lora_fwd: head-scatter output is plain bf16");
    AF_CHECK(hs_heads > 0 && hs_d % 8 == 0 && hs_dpad % 8 == 0 && hs_dpad >= hs_d && hs_rows > 0 && M % hs_rows == 0 &&
                 N % (hs_heads * hs_d) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
             "proj_lora_fwd: bad head-scatter geometry (heads=%lld d=%lld dpad=%lld rows=%lld M=%lld N=%lld)", (long long)hs_heads,
             (long long)hs_d, (long long)hs_dpad, (long long)hs_rows, (long long)M, (long long)N);
  }
  AF_CHECK(M < (1ll << 31) && N < (1ll << 31), "proj_lora_fwd: M/N too large");

  int BN;
  if (act == ADAFACE_ACT_GEGLU) {
    AF_CHECK(N % 128 == 0, "proj_lora_fwd: GEGLU needs N %% 128 == 0 (packed [a|g] tiles), got %lld", (long long)N);
    BN = 128;
  } else if (N % 160 == 0 && M >= 2048) {
    BN = 160;
  } else if (N % 128 == 0 || N > 512) {
    BN = 128;
  } else if (N <= 64 || N % 128 <= 64) {
    BN = 64;
  } else {
    BN = 128;
  }
  // Small problems: prefer more, narrower tiles so that more SMs get work.
  if (BN == 128 && act != ADAFACE_ACT_GEGLU && ((M + 127) / 128) * ((N + 127) / 128) < 148 && N % 64 == 0) BN = 64;

  CUtensorMap tA, tB, tA2, tB2;
  if (make_tmap_bf16_2d(&tA, x, (uint64_t)M, (uint64_t)K, (uint64_t)ldx, GEMM_BM)) return 3;
  if (make_tmap_bf16_2d(&tB, w, (uint64_t)N, (uint64_t)K, (uint64_t)K, (uint32_t)BN)) return 3;
  if (lora) {
    if (make_tmap_bf16_2d(&tA2, t, (uint64_t)M, (uint64_t)R, (uint64_t)ldt, GEMM_BM)) return 3;
    if (make_tmap_bf16_2d(&tB2, bs, (uint64_t)N, (uint64_t)R, (uint64_t)R, (uint32_t)BN)) return 3;
  } else {
    tA2 = tA;
    tB2 = tB;
  }
  GemmEpilogue ep;
  ep.colscale = colscale;
  ep.bias = bias;
  ep.residual = residual;
  ep.y = y;
  ep.ldr = ldr;
  ep.ldy = ldy;
  ep.M = (int)M;
  ep.N = (int)N;
  ep.num_kb1 = (int)((K + GEMM_BK - 1) / GEMM_BK);
  ep.num_kb2 = lora ? (int)((R + GEMM_BK - 1) / GEMM_BK) : 0;
  ep.act = act;
  ep.y_f32 = y_dtype == ADAFACE_F32;
  ep.res_f32 = residual_dtype == ADAFACE_F32;
  ep.hs_d = (int)hs_d;
  ep.hs_dpad = (int)hs_dpad;
  ep.hs_H = (int)hs_heads;
  ep.hs_C = (int)(hs_heads * hs_d);
  ep.hs_rows = (int)hs_rows;
  ep.hs_B = hs_rows > 0 ? (int)(M / hs_rows) : 0;
  const int n_tiles = (int)((N + BN - 1) / BN);
  switch (BN) {
    case 64: return launch_gemm<64>(tA, tB, tA2, tB2, ep, n_tiles, stream);
    case 128: return launch_gemm<128>(tA, tB, tA2, tB2, ep, n_tiles, stream);
    case 160: return launch_gemm<160>(tA, tB, tA2, tB2, ep, n_tiles, stream);
  }
  set_error("proj_lora_fwd: unreachable tile width %d", BN);
  return 1;
}

}  // namespace adaface

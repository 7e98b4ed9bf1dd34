// K2, level A (d = 40), generation 4: THREE query tiles per CTA, every hand-off of the four-tile kernel taken off the softmax chain.
//
// What the trace of the four-tile kernel showed (profiles/r02_attn_trace_quad.txt, CTA 0, 64-key steps of 2150 clk):
//   * a tile's step is a CHAIN: scores ready -> TMEM read + row max (250 clk) -> exp2 + pack (900-1200 clk, two to three warps share the
//     16-op/clk MUFU) -> P ready -> the issuer notices, issues P.V (4 MMAs) and the next Q.K^T (3 MMAs), the tensor pipe runs them behind
//     the other tiles' MMAs -> next scores ready: 700-970 clk in which the tile's four softmax warps have nothing to do;
//   * with four tiles there are on average TWO warps per scheduler in their exp2 phase: the MUFU pipe is 70 % busy, the tensor pipe
//     (SS Q.K^T at N = 64: 75 clk per MMA, TS P.V: 34 clk) 67 % -- neither is the bound, the chain is.
// S and P shared one TMEM region there (Q.K^T(j+1) overwrites P(j)), so the round trip could not overlap the tile's own softmax.
// Here a tile owns FOUR regions -- S 64 | P 32 | O 48 | Q 24 columns = 168, three tiles = 504 of the 512 columns:
//   * the softmax warps hand S back as soon as the 64 scores are in registers (s_free); the Q.K^T issuer queues Q.K^T(g, j+1) at once, so the
//     next scores are ready long before the tile's exp2 work on step j ends: the chain is softmax-only;
//   * P(j) goes to its own region; P.V(g, j) is issued by a SECOND issuer warp when it is complete (p_full) and P / O are only touched
//     again after that MMA has retired (pv_done) -- normally a whole step later;
//   * Q lives in TMEM (written once per CTA by the softmax warps, 24 columns per tile): Q.K^T is a TS MMA like P.V and costs 42 instead
//     of 75 clk of the tensor pipe per k16 step (profiles/r01_microbench.md) -- three tiles need 3 x (3 x 42 + 4 x 34) = 790 clk per step;
//   * two single-thread issuers with blocking in-order loops (Q.K^T: s_free -> MMA; P.V: p_full -> MMA); neither waits for the other.
// Everything else is the four-tile kernel's: K/V tiles arrive by TMA once per CTA (ring of three), row sums come from a ones column in V
// (column 40 of the accumulator), lazy rescale at 2^8, part of the exp2 on the FMA pipe.
// Units: a (batch, head) has ceil(Lq / 128) query tiles, cut into n3 three-tile and n2 two-tile units (a two-tile unit leaves the third
// slot idle); the launcher picks (n3, n2) by simulating the greedy block scheduler on 148 SMs, three-tile units first.
// Reference arithmetic: F.scaled_dot_product_attention at dalc:321 / ldm attention.py:181-204 (no mask).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "attn_tc_params.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

extern long long g_launch_count;

#ifdef AF_ATTN_TRACE
#define T3_TR(...) __VA_ARGS__
#else
#define T3_TR(...)
#endif

// Issuer-side wait: non-blocking test_wait in a tight loop (one thread of a control warp; it has nothing else to do and must react within a
// few clocks -- the suspending try_wait of mbar_wait was seen to notice a completed phase several hundred clocks late).
__device__ __forceinline__ void t3_wait_spin(uint64_t* bar, uint32_t parity) {
#ifdef T3_NO_SPIN
  mbar_wait(bar, parity);
#else
  uint32_t ok = 0, polls = 0;
  do {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (++polls > (1u << 28)) __trap();
  } while (!ok);
#endif
}

constexpr int T3_BM = 128, T3_BN = 64, T3_D = 40, T3_DO = 48, T3_KT = 3, T3_G = 3, T3_ST = 3;
constexpr int T3_Q_BYTES = T3_BM * 128, T3_KV_BYTES = T3_BN * 128;      // 128-byte (64-column) swizzled rows
// TMEM columns, grouped by kind so that every region keeps its natural alignment
constexpr int T3_S = 0, T3_P = T3_G * 64, T3_O = T3_P + T3_G * 32, T3_Q = T3_O + T3_G * T3_DO;
static_assert(T3_Q + T3_G * 24 <= 512, "TMEM budget");
constexpr int T3_WARPS = 4 * T3_G + 1 + T3_G;                          // softmax warps, TMA producer, one MMA issuer per tile
constexpr int kT3Tma = 4 * T3_G, kT3Iss = 4 * T3_G + 1;

template <int DEG>
__device__ __forceinline__ float2 t3_exp2_emu2(float2 x) {
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.f, 12582912.f));
  const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);
  float2 q;
  if constexpr (DEG == 3) {
    q = __ffma2_rn(f, make_float2(0.05517164f, 0.05517164f), make_float2(0.24261113f, 0.24261113f));
    q = __ffma2_rn(q, f, make_float2(0.69326097f, 0.69326097f));
    q = __ffma2_rn(q, f, make_float2(0.99992806f, 0.99992806f));
  } else {
    q = __ffma2_rn(f, make_float2(0.23842894f, 0.23842894f), make_float2(0.7034480f, 0.7034480f));
    q = __ffma2_rn(q, f, make_float2(1.0004431f, 1.0004431f));
  }
  float2 r;
  r.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
  return r;
}

struct T3Units {
  int n3, n2;      // three-tile / two-tile units per (batch, head)
  int bh;          // batch * heads
  int stagger;     // clocks between the first Q.K^T of consecutive tiles (0 = all at once)
};

// EMU = how many of every 8 exp2 pairs go to the degree-2 FMA-pipe polynomial instead of MUFU.EX2.
template <int EMU>
__global__ void __launch_bounds__(T3_WARPS * 32, 1)
attn_fwd_tcgen05_tri_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                            const __grid_constant__ CUtensorMap tmV, const TaParams p, const T3Units un, const int H) {
  constexpr int D = T3_D, DO = T3_DO, BN = T3_BN, G = T3_G, ST = T3_ST, KT = T3_KT;
  extern __shared__ uint8_t smem_raw_t3[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_t3) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                             // [G][128][128 B]
  uint8_t* sK = sQ + G * T3_Q_BYTES;                              // [ST][64][128 B]
  uint8_t* sV = sK + ST * T3_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ST * T3_KV_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                                   // [ST]
  uint64_t* kv_empty = kv_full + ST;                              // [ST]: every tile's issuer is done with the stage
  uint64_t* v_ready = kv_empty + ST;                              // [ST]: the ones column has been written into V stage s
  uint64_t* q_ready = v_ready + ST;                               // [G]: Q tile g sits in TMEM
  uint64_t* s_full = q_ready + G;                                 // [G]
  uint64_t* s_free = s_full + G;                                  // [G]: the scores are in registers
  uint64_t* p_full = s_free + G;                                  // [G]
  uint64_t* pv_done = p_full + G;                                 // [G]: P.V(g, j) has retired (P and O may be touched)
  uint64_t* o_full = pv_done + G;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // unit -> (batch * head, first query tile, active tiles): all three-tile units first, the two-tile units close the grid
  int bh, tile0, n_act;
  {
    const int unit = blockIdx.x, u3 = un.bh * un.n3;
    if (unit < u3) {
      bh = unit / un.n3;
      tile0 = (unit - bh * un.n3) * 3;
      n_act = 3;
    } else {
      const int u = unit - u3;
      bh = u / un.n2;
      tile0 = un.n3 * 3 + (u - bh * un.n2) * 2;
      n_act = 2;
    }
  }
  {
    const int n_qt = (p.Lq + T3_BM - 1) / T3_BM;
    if (tile0 + n_act > n_qt) n_act = n_qt - tile0;               // ragged end of a (batch, head): 1 .. 3 tiles
  }
  const int h = bh % H, b = bh / H;
  const int m0 = tile0 * T3_BM;
  const int n_tiles = (p.Lk + BN - 1) / BN;

  if (warp == kT3Tma && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], n_act);
      mbar_init(&v_ready[s], 1);
    }
    for (int g = 0; g < G; ++g) {
      mbar_init(&q_ready[g], 128);
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 128);
      mbar_init(&p_full[g], 128);
      mbar_init(&pv_done[g], 1);
    }
    mbar_init(o_full, n_act);
    fence_barrier_init();
  } else if (warp == kT3Iss) {
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // global memory is touched only after the predecessor kernel has completed
  AF_PDL_TRIGGER_EARLY();

  if (warp == kT3Tma) {
    // whole warp: one elected lane issues the TMA loads (two tiles ahead), then all 32 lanes write the ONES COLUMN into the V tile
    // that has just landed (column D of every V row := 1.0: column D of the P.V accumulator is the row sum of P)
    const bool leader = elect_one();
    auto issue_kv = [&](int j) {
      const int s = j % ST;
      mbar_wait(&kv_empty[s], ((j / ST) & 1) ^ 1);
      mbar_arrive_expect_tx(&kv_full[s], 2 * T3_KV_BYTES);
      tma_load_4d(sK + s * T3_KV_BYTES, &tmK, &kv_full[s], 0, h, j * BN, b);
      tma_load_4d(sV + s * T3_KV_BYTES, &tmV, &kv_full[s], 0, h, j * BN, b);
    };
    if (leader) {
      mbar_arrive_expect_tx(q_full, n_act * T3_Q_BYTES);
      for (int g = 0; g < n_act; ++g) tma_load_4d(sQ + g * T3_Q_BYTES, &tmQ, q_full, 0, h, m0 + g * T3_BM, b);
      issue_kv(0);
      if (n_tiles > 1) issue_kv(1);
    }
    for (int j = 0; j < n_tiles; ++j) {
      const int s = j % ST;
      mbar_wait(&kv_full[s], (j / ST) & 1);
#pragma unroll
      for (int r = lane; r < BN; r += 32)      // 128B-swizzled tile: element D of row r sits in chunk (D/8) ^ (r & 7)
        *reinterpret_cast<uint16_t*>(sV + s * T3_KV_BYTES + r * 128 + ((((D >> 3) ^ (r & 7)) << 4) | ((D & 7) << 1))) = 0x3F80;
      fence_proxy_async_smem();
      __syncwarp();
      if (leader) {
        mbar_arrive(&v_ready[s]);
        if (j + 2 < n_tiles) issue_kv(j + 2);
      }
      __syncwarp();
    }
  } else if (warp >= kT3Iss) {
    // one issuer per tile, one tile per scheduler (warps 13 / 14 / 15 sit on sub-partitions 1 / 2 / 3): a tcgen05.mma that waits for room in
    // the tensor queue holds its sub-partition's dispatch port, so a single issuer made ITS sub-partition's softmax warps 800 clk per step
    // slower than the others (traced: the warp that shares a scheduler with the Q.K^T issuer arrived last at every p_full)
    const int g = warp - kT3Iss;
    if (g < n_act && elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc_bf16_f32(T3_BM, BN, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16_f32(T3_BM, DO, true);
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV);
      auto issue_qk = [&](int j) {
        const int s = j % ST;
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
          const uint64_t db = make_smem_desc_sw128(aK + s * T3_KV_BYTES + kk * 32);
          umma_bf16_ts(tmem_base + (uint32_t)(T3_S + g * 64), tmem_base + (uint32_t)(T3_Q + g * 24 + kk * 8), db, idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[g]);
      };
      mbar_wait(&kv_full[0], 0);
      mbar_wait(&q_ready[g], 0);
      tc_fence_after();
      if (g > 0 && un.stagger > 0) {            // spread the tiles over the step: their TMEM reads / row maxima then fall into each other's exp2 phases
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)un.stagger * g) {}
      }
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % ST;
        T3_TR(const bool tr = p.trace && blockIdx.x == 0 && j >= 8 && j < 16; if (tr) p.trace[(j - 8) * 16 + g * 3] = clock64();)
        if (j + 1 < n_tiles) {
          mbar_wait(&kv_full[(j + 1) % ST], ((j + 1) / ST) & 1);
          mbar_wait(&s_free[g], j & 1);          // the scores of step j are in registers
          T3_TR(if (tr) p.trace[(j - 8) * 16 + g * 3 + 1] = clock64();)
          tc_fence_after();
          issue_qk(j + 1);
          T3_TR(if (tr) p.trace[(j - 8) * 16 + g * 3 + 2] = clock64();)
        }
        T3_TR(if (tr) p.trace[128 + (j - 8) * 16 + g * 3] = clock64();)
        mbar_wait(&v_ready[s], (j / ST) & 1);      // V_j carries its ones column
        mbar_wait(&p_full[g], j & 1);              // P_j is in TMEM
        T3_TR(if (tr) p.trace[128 + (j - 8) * 16 + g * 3 + 1] = clock64();)
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          const uint64_t db = make_smem_desc_sw128_mn(aV + s * T3_KV_BYTES + k * 2048, T3_KV_BYTES);
          umma_bf16_ts(tmem_base + (uint32_t)(T3_O + g * DO), tmem_base + (uint32_t)(T3_P + g * 32 + k * 8), db, idesc_pv, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&pv_done[g]);
        umma_commit(&kv_empty[s]);
        T3_TR(if (tr) p.trace[128 + (j - 8) * 16 + g * 3 + 2] = clock64();)
      }
      umma_commit(o_full);
    }
  } else if (warp < 4 * G && (warp >> 2) < n_act) {
    const int g = warp >> 2, qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qd * 32) << 16);
    const uint32_t tS = t_lane + (uint32_t)(T3_S + g * 64), tP = t_lane + (uint32_t)(T3_P + g * 32), tO = t_lane + (uint32_t)(T3_O + g * DO);
    {
      // Q row -> TMEM (two consecutive-K bf16 per 32-bit column): 48 elements = six 16-byte chunks of the swizzled row
      mbar_wait(q_full, 0);
      const uint8_t* qrow = sQ + g * T3_Q_BYTES + row * 128;
      uint32_t qv[24];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const uint4 w = *reinterpret_cast<const uint4*>(qrow + ((c ^ (row & 7)) << 4));
        qv[c * 4 + 0] = w.x; qv[c * 4 + 1] = w.y; qv[c * 4 + 2] = w.z; qv[c * 4 + 3] = w.w;
      }
      uint32_t q16[16], q8[8];
#pragma unroll
      for (int i = 0; i < 16; ++i) q16[i] = qv[i];
#pragma unroll
      for (int i = 0; i < 8; ++i) q8[i] = qv[16 + i];
      tmem_st_32x32b_x16(t_lane + (uint32_t)(T3_Q + g * 24), q16);
      tmem_st_32x32b_x8(t_lane + (uint32_t)(T3_Q + g * 24 + 16), q8);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&q_ready[g]);
    }
    float m_ref = -INFINITY;
    // Waits on the softmax chain: a non-blocking test first (a completed phase answers in a few tens of clocks; the suspending try_wait
    // of mbar_wait was traced at ~130 clk even then).
    auto wait_fast = [&](uint64_t* bar, uint32_t parity) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
      if (!ok) mbar_wait(bar, parity);
    };
    for (int j = 0; j < n_tiles; ++j) {
      T3_TR(const bool str_ = p.trace && blockIdx.x == 0 && lane == 0 && qd == 0 && j >= 8 && j < 16; long long* tp = p.trace + 256 + g * 64 + (j - 8) * 8;
            if (str_) tp[0] = clock64();)
      wait_fast(&s_full[g], j & 1);
      T3_TR(if (str_) tp[1] = clock64();)
      tc_fence_after();
      const int valid = p.Lk - j * BN;
      // the scores arrive in two halves: the exponentials of the first half run while the second is still on its way
      uint32_t v[BN];
      {
        uint32_t (&va)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
        uint32_t (&vb)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
        tmem_ld_32x32b_x32_wait(tS, va);
        tmem_ld_32x32b_x32_nowait(tS + 32u, vb);
      }
      if (valid < BN) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i >= valid) v[i] = 0xff800000u;
      }
      // exp2 of one half (keys 32 hf .. 32 hf + 31) against the reference maximum `mr`, packed to bf16 and stored into the P region
      auto half = [&](int hf, float mr) {
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-mr, -mr);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[hf * 32 + 2 * i]), __uint_as_float(v[hf * 32 + 2 * i + 1])), sc2, nm2);
          const float2 e = ((i & 7) < EMU) ? t3_exp2_emu2<2>(t) : make_float2(fast_exp2(t.x), fast_exp2(t.y));
          pk[i] = __byte_perm(__float_as_uint(e.x), __float_as_uint(e.y), 0x7632);   // truncate to bf16; the row sum comes from the MMA
        }
        tmem_st_32x32b_x16(tP + (uint32_t)(hf * 16), pk);
      };
      auto rowmax32 = [&](int hf) {
        float m4[4] = {__uint_as_float(v[hf * 32]), __uint_as_float(v[hf * 32 + 1]), __uint_as_float(v[hf * 32 + 2]), __uint_as_float(v[hf * 32 + 3])};
#pragma unroll
        for (int i = 4; i < 32; i += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) m4[u] = fmaxf(m4[u], __uint_as_float(v[hf * 32 + i + u]));
        }
        return fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      };
      auto second_half_landed = [&]() {
        uint32_t (&vb)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
        tmem_ld_wait_x32(vb);
        tc_fence_before();
        mbar_arrive(&s_free[g]);               // Q.K^T(g, j + 1) may overwrite the scores now
        T3_TR(if (str_) tp[2] = clock64();)
        if (valid < BN) {
#pragma unroll
          for (int i = 32; i < BN; ++i)
            if (i >= valid) v[i] = 0xff800000u;
        }
      };
      if (j == 0) {
        second_half_landed();
        const float mx = fmaxf(rowmax32(0), rowmax32(1)) * p.scale_log2;
        m_ref = (mx == -INFINITY) ? 0.f : mx;
        half(0, m_ref);
        half(1, m_ref);
      } else {
        // OPTIMISTIC maximum: the exponentials start against the running reference at once; the row maximum of this tile is computed
        // beside them (ALU pipe) and only a row that outgrew the reference by 2^8 redoes its tile (rare after the first few tiles)
        const float mxa = rowmax32(0);
        T3_TR(if (str_) tp[3] = clock64();)
        wait_fast(&pv_done[g], (j - 1) & 1);   // P.V(g, j - 1) has retired: P may be rewritten, O rescaled
        T3_TR(if (str_) tp[4] = clock64();)
        tc_fence_after();
        half(0, m_ref);
        second_half_landed();
        const float mx = fmaxf(mxa, rowmax32(1)) * p.scale_log2;
        half(1, m_ref);
        const bool need = mx > m_ref + 8.f;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? mx : m_ref;
          const float f = fast_exp2(m_ref - m_new);
          m_ref = m_new;
#pragma unroll
          for (int c = 0; c < DO / 16; ++c) {
            uint32_t ov[16];
            tmem_ld_32x32b_x16(tO + (uint32_t)(c * 16), ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
            tmem_st_32x32b_x16(tO + (uint32_t)(c * 16), ov);
          }
          half(0, m_ref);
          half(1, m_ref);
        }
      }
      T3_TR(if (str_) tp[5] = clock64();)
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[g]);
      T3_TR(if (str_) tp[6] = clock64();)
      T3_TR(if (p.trace && blockIdx.x == 0 && lane == 0 && j >= 8 && j < 16) p.trace[512 + qd * 32 + g * 8 + (j - 8)] = clock64();)
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int grow = m0 + g * T3_BM + row;
    bf16* orow = p.o + (long long)b * p.o_sb + (long long)grow * p.o_sn + h * D;
    float l_run;                                                 // row sum of P = column D of the accumulator
    {
      uint32_t v8[8];
      tmem_ld_32x32b_x8(tO + (uint32_t)(D & ~7), v8);
      tmem_ld_wait();
      l_run = __uint_as_float(v8[D & 7]);
    }
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    if (p.lse && grow < p.Lq) p.lse[((long long)b * H + h) * p.Lq + grow] = l_run > 0.f ? m_ref + log2f(l_run) : INFINITY;
#pragma unroll
    for (int c = 0; c < DO / 16; ++c) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tO + (uint32_t)(c * 16), v);
      tmem_ld_wait();
      if (grow < p.Lq) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (c * 16 + half * 8 < D) {
            uint4 pk;
            pk.x = pack_bf16(__uint_as_float(v[half * 8 + 0]) * inv, __uint_as_float(v[half * 8 + 1]) * inv);
            pk.y = pack_bf16(__uint_as_float(v[half * 8 + 2]) * inv, __uint_as_float(v[half * 8 + 3]) * inv);
            pk.z = pack_bf16(__uint_as_float(v[half * 8 + 4]) * inv, __uint_as_float(v[half * 8 + 5]) * inv);
            pk.w = pack_bf16(__uint_as_float(v[half * 8 + 6]) * inv, __uint_as_float(v[half * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + c * 16 + half * 8) = pk;
          }
        }
      }
    }
  }
  AF_PDL_TRIGGER_LATE();
  tc_fence_before();
  __syncthreads();
  if (warp == kT3Iss) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Greedy list scheduling of `n3` long and `n2` short blocks per (batch, head) on `sms` SMs, long blocks first (the order of the grid):
// returns the makespan in tile-steps.  A block costs its tile count plus a fixed set-up share.
static double t3_makespan(int bh, int n3, int n2, int sms) {
  double load[1024];
  if (sms > 1024) sms = 1024;
  for (int i = 0; i < sms; ++i) load[i] = 0.0;
  auto place = [&](double c) {
    int best = 0;
    for (int i = 1; i < sms; ++i)
      if (load[i] < load[best]) best = i;
    load[best] += c;
  };
  for (int i = 0; i < bh * n3; ++i) place(3.0 + 0.15);
  for (int i = 0; i < bh * n2; ++i) place(2.0 + 0.15);
  double m = 0.0;
  for (int i = 0; i < sms; ++i) m = load[i] > m ? load[i] : m;
  return m;
}

template <int EMU>
static int launch_t3(const CUtensorMap& tQ, const CUtensorMap& tK, const CUtensorMap& tV, const TaParams& p, int B, int H, cudaStream_t stream) {
  constexpr int smem = T3_G * T3_Q_BYTES + 2 * T3_ST * T3_KV_BYTES + 1024 + 512;
  AF_CONFIG_SMEM((attn_fwd_tcgen05_tri_kernel<EMU>), smem);
  const int n_qt = (p.Lq + T3_BM - 1) / T3_BM;
  T3Units un;
  un.bh = B * H;
  static int forced_n2 = -2, stagger = -1;
  if (forced_n2 == -2) {
    const char* e = getenv("ADAFACE_TRI_N2");          // tuning knob: two-tile units per (batch, head); -1 / unset = chosen by the scheduler model
    forced_n2 = e ? atoi(e) : -1;
    const char* s = getenv("ADAFACE_TRI_STAGGER");     // clocks between the first Q.K^T of consecutive tiles
    stagger = s ? atoi(s) : 0;
  }
  un.stagger = stagger;
  // every split 3 * n3 + 2 * n2 >= n_qt with the smallest cover; pick the one the scheduler model likes best
  int best3 = -1, best2 = 0;
  double best = 0.0;
  for (int n2 = 0; 2 * n2 <= n_qt + 1; ++n2) {
    const int rest = n_qt - 2 * n2;
    if (rest < 0) break;
    const int n3 = (rest + 2) / 3;
    if (3 * n3 + 2 * n2 - n_qt >= 2 && n2 > 0) continue;      // would waste a whole tile slot where a smaller unit fits
    if (forced_n2 >= 0 && n2 != forced_n2) continue;
    const double m = t3_makespan(un.bh, n3, n2, af_num_sms());
    if (best3 < 0 || m < best - 1e-9) { best = m; best3 = n3; best2 = n2; }
  }
  if (best3 < 0) { best3 = (n_qt + 2) / 3; best2 = 0; }
  un.n3 = best3;
  un.n2 = best2;
  const int grid = un.bh * (un.n3 + un.n2);
  AF_CUDA(launch_pdl(2, attn_fwd_tcgen05_tri_kernel<EMU>, dim3(grid), dim3(T3_WARPS * 32), smem, stream, tQ, tK, tV, p, un, H));
  ++g_launch_count;
  AF_CUDA(cudaGetLastError());
  if (p.trace) {      // diagnosis: CTA 0, key tiles 8..15
    static long long hh[1024];
    cudaDeviceSynchronize();
    cudaMemcpy(hh, p.trace, sizeof(hh), cudaMemcpyDeviceToHost);
    const long long t0 = hh[0];
    fprintf(stderr, "tri: n3 %d n2 %d grid %d\n", un.n3, un.n2, grid);
    for (int j = 0; j < 8; ++j) {
      fprintf(stderr, "qk issuer j=%2d:", j + 8);
      for (int g = 0; g < 3; ++g) fprintf(stderr, " g%d s_free wait %6lld..%6lld issued %6lld |", g, hh[j * 16 + g * 3] - t0, hh[j * 16 + g * 3 + 1] - t0, hh[j * 16 + g * 3 + 2] - t0);
      fprintf(stderr, "\npv issuer j=%2d:", j + 8);
      for (int g = 0; g < 3; ++g) fprintf(stderr, " g%d p_full wait %6lld..%6lld issued %6lld |", g, hh[128 + j * 16 + g * 3] - t0, hh[128 + j * 16 + g * 3 + 1] - t0, hh[128 + j * 16 + g * 3 + 2] - t0);
      fprintf(stderr, "\n");
    }
    for (int g = 0; g < 3; ++g)
      for (int j = 0; j < 8; ++j) {
        const long long* tp = hh + 256 + g * 64 + j * 8;
        fprintf(stderr, "softmax g=%d j=%2d: s_full wait %6lld..%6lld  s_free arrived %6lld  max done %6lld  pv_done seen %6lld  exp done %6lld  arrived %6lld  (warps 0..3: %6lld %6lld %6lld %6lld)\n", g, j + 8, tp[0] - t0, tp[1] - t0, tp[2] - t0, tp[3] - t0, tp[4] - t0, tp[5] - t0, tp[6] - t0,
                hh[512 + g * 8 + j] - t0, hh[512 + 32 + g * 8 + j] - t0, hh[512 + 64 + g * 8 + j] - t0, hh[512 + 96 + g * 8 + j] - t0);
      }
  }
  return 0;
}

// d = 40, unmasked, Lq >= 1024, Lk > 128.  q/k/v strides as in attn_fwd_tcgen05.
int attn_fwd_tcgen05_tri(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sh, int64_t k_sn,
                         const void* v, int64_t v_sb, int64_t v_sh, int64_t v_sn, int64_t B, int64_t H, int64_t Lq, int64_t Lk,
                         int64_t drow_q, int64_t drow_kv, const TaParams& p, cudaStream_t stream) {
  static int emu = -1;
  if (emu < 0) {
    const char* e = getenv("ADAFACE_EXP_EMU");       // exp2 pairs of every 8 on the FMA pipe (0..4)
    emu = (e && e[0] >= '0' && e[0] <= '4') ? (e[0] - '0') : 2;
  }
  CUtensorMap tQ, tK, tV;
  if (make_tmap_bf16_heads(&tQ, q, (uint64_t)drow_q, (uint64_t)H, (uint64_t)Lq, (uint64_t)B, (uint64_t)q_sh, (uint64_t)q_sn, (uint64_t)q_sb, T3_BM)) return 3;
  if (make_tmap_bf16_heads(&tK, k, (uint64_t)drow_kv, (uint64_t)H, (uint64_t)Lk, (uint64_t)B, (uint64_t)k_sh, (uint64_t)k_sn, (uint64_t)k_sb, T3_BN)) return 3;
  if (make_tmap_bf16_heads(&tV, v, (uint64_t)drow_kv, (uint64_t)H, (uint64_t)Lk, (uint64_t)B, (uint64_t)v_sh, (uint64_t)v_sn, (uint64_t)v_sb, T3_BN)) return 3;
  const int ib = (int)B, ih = (int)H;
  switch (emu) {
    case 0: return launch_t3<0>(tQ, tK, tV, p, ib, ih, stream);
    case 1: return launch_t3<1>(tQ, tK, tV, p, ib, ih, stream);
    case 3: return launch_t3<3>(tQ, tK, tV, p, ib, ih, stream);
    case 4: return launch_t3<4>(tQ, tK, tV, p, ib, ih, stream);
    default: return launch_t3<2>(tQ, tK, tV, p, ib, ih, stream);
  }
}

}  // namespace adaface

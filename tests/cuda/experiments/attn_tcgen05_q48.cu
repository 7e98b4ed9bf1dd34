// K2, level A (d = 40), generation 3: four query tiles per CTA with 48-KEY tiles and an EARLY-RELEASED score buffer.
//
// ncu on the 64-key four-tile kernel (profiles/r01_ncu_full_attn_quad.txt, source page): 31 % of all warp samples sit on the
// softmax warps' wait for their next score tile -- S and P shared one TMEM region, so Q.K(g, j+1) could only be issued after
// softmax(g, j) had written P(j) and P.V(g, j) had consumed it, and every tile paid the whole tensor round trip (~300 clk) per
// key tile with its exp2 work stopped.  TMEM had no room for a second S (4 tiles x (64 S/P + 48 O) = 448 of 512 columns).
// With 48-key tiles a tile needs 48 (S) + 24 (P, bf16) + 48 (O) = 120 <= 128 columns, so P gets its OWN region and
//   * the softmax warps hand S back as soon as the 48 scores are in registers (s_free) -- the issuer then queues
//     Q.K(g, j+1) at once, and the next scores are ready long before the tile's exp2 / pack work on tile j ends;
//   * P.V(g, j) follows when P(j) is stored (p_full); a P region is rewritten only after its P.V has completed (pv_done).
//   * ONE ISSUER WARP PER TILE.  Timing the 64-key kernel with 0..4 of every 8 exp2 on the FMA pipe moved nothing (333-342 us):
//     neither the XU nor the issue slots bound it -- the single MMA-issuing thread did.  Per key tile it ran 6 mbarrier waits
//     (~90 clk each even when already complete), 28 tcgen05.mma (34-107 clk of issue each, profiles/r01_microbench.md) and 5
//     commits back to back: ~1800 clk, which IS the measured period of a four-tile iteration.  Here warps 16..19 each own one
//     tile's Q.K / P.V stream (3 waits, 6 MMAs, 3 commits per key tile); the tensor pipe orders them, mbarriers with four
//     arrivals release a K/V stage and the epilogue.
// Per 48 keys the tensor pipe now spends 72 + 72 clk (N = 48 for both MMAs, three k16 steps each) instead of 128 + 96 per 64,
// and the dependent chain of a tile is softmax-only.  Everything else is the 64-key kernel's: one elected thread issues every
// tcgen05.mma, K/V tiles arrive by TMA once per 512 queries, row sums come from a ones column in V (column 40 of the
// accumulator), lazy rescale at 2^8, part of the exp2 on the FMA pipe, split-key wave tail.
// Reference arithmetic: F.scaled_dot_product_attention at dalc:321 / ldm attention.py:181-204 (no mask).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "attn_tc_params.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

extern long long g_launch_count;

constexpr int Q48_BM = 128, Q48_BN = 48, Q48_D = 40, Q48_DO = 48, Q48_G = 4, Q48_ST = 4;
constexpr int Q48_Q_BYTES = Q48_BM * 128, Q48_KV_BYTES = Q48_BN * 128;      // 128-byte (64-column) swizzled rows
constexpr int Q48_TMEM_G = 128, Q48_S = 0, Q48_P = 48, Q48_O = 80;          // per tile: S 48 | P 24 | (8 spare) | O 48

__device__ __forceinline__ void tmem_ld_32x32b_x48_wait(uint32_t taddr, uint32_t (&v)[48]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%48];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47}, [%49];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47])
      : "r"(taddr), "r"(taddr + 32u)
      : "memory");
}

// exp2 on the FMA / ALU pipes for a pair of values <= 8: Cody-Waite split with the 1.5 * 2^23 magic constant, minimax
// polynomial of 2^f on [-0.5, 0.5] (degree 3: max rel. error 7.5e-5; degree 2: 1.7e-3, below the 3.9e-3 of the truncating
// bf16 pack that follows), exponent re-inserted with one integer multiply-add.
template <int DEG>
__device__ __forceinline__ float2 q48_exp2_emu2(float2 x) {
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

// EMU = how many of every 8 exp2 pairs go to the FMA-pipe polynomial; DEG = its degree (2 | 3).
// SPLIT (wave tail): the four TMEM slots hold TWO query tiles x TWO halves of the keys, merged in the epilogue.
template <int EMU, int DEG, bool SPLIT>
__global__ void __launch_bounds__((4 * Q48_G + Q48_G) * 32, 1)
attn_fwd_tcgen05_q48_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                            const __grid_constant__ CUtensorMap tmV, const TaParams p, const int unit0, const int H) {
  constexpr int D = Q48_D, DO = Q48_DO, BN = Q48_BN, G = Q48_G, ST = Q48_ST, KT = 3, TG = Q48_TMEM_G;
  // control warps 4G .. 5G-1: warp 4G + g issues every tcgen05.mma of tile slot g; warp 4G also runs the TMA producer
  // (20 warps = 640 threads keep the 96-register budget of the softmax threads; a 21st warp would cap it at 80)
  constexpr int kTma = 4 * G, kMma = 4 * G;
  constexpr int NQ = SPLIT ? G / 2 : G;                           // query tiles per CTA
  constexpr int KH = SPLIT ? 2 : 1;                               // key halves (sub-tiles per ring stage)
  extern __shared__ uint8_t smem_raw_q48[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_q48) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                             // [NQ][128][128 B]
  uint8_t* sK = sQ + NQ * Q48_Q_BYTES;                            // [ST][KH][48][128 B]
  uint8_t* sV = sK + ST * KH * Q48_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ST * KH * Q48_KV_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                                   // [ST]
  uint64_t* kv_empty = kv_full + ST;                              // [ST]
  uint64_t* v_ready = kv_empty + ST;                              // [ST]: the ones column has been written into V stage s
  uint64_t* s_full = v_ready + ST;                                // [G]  issuer -> softmax: S(j) is in TMEM
  uint64_t* s_free = s_full + G;                                  // [G]  softmax -> issuer: S(j) is in registers
  uint64_t* p_full = s_free + G;                                  // [G]  softmax -> issuer: P(j) is in TMEM
  uint64_t* pv_done = p_full + G;                                 // [G]  issuer -> softmax: P.V(j) has completed
  uint64_t* o_full = pv_done + G;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qb = (p.Lq + NQ * Q48_BM - 1) / (NQ * Q48_BM);
  const int unit = unit0 + blockIdx.x;
  const int qb = unit % n_qb, bh = unit / n_qb;
  const int h = bh % H, b = bh / H;
  const int m0 = qb * (NQ * Q48_BM);
  const int n_tiles = ((p.Lk + BN - 1) / BN) / KH;               // iterations; SPLIT: slot half kh covers key tiles [kh * n_tiles, ...)

  if (warp == kTma && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], G);                                 // one commit per tile issuer
      mbar_init(&v_ready[s], 1);
    }
    for (int g = 0; g < G; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 128);
      mbar_init(&p_full[g], 128);
      mbar_init(&pv_done[g], 1);
    }
    mbar_init(o_full, G);
    fence_barrier_init();
  } else if (warp == kMma + 1) {
    tmem_alloc(tmem_slot, G * TG);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // global memory is touched only after the predecessor kernel has completed

  if (warp >= kMma) {
    const int g = warp - kMma;                                    // tile slot whose MMAs this warp's elected lane issues
    const bool leader = elect_one();
    constexpr uint32_t idesc_qk = make_idesc_bf16_f32(Q48_BM, BN, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16_f32(Q48_BM, DO, true);
    const uint32_t aQ = smem_u32(sQ) + (SPLIT ? g >> 1 : g) * Q48_Q_BYTES;
    const uint32_t aK = smem_u32(sK) + (SPLIT ? g & 1 : 0) * Q48_KV_BYTES, aV = smem_u32(sV) + (SPLIT ? g & 1 : 0) * Q48_KV_BYTES;
    const uint32_t tS = tmem_base + (uint32_t)(g * TG + Q48_S), tP = tmem_base + (uint32_t)(g * TG + Q48_P);
    const uint32_t tO = tmem_base + (uint32_t)(g * TG + Q48_O);
    auto issue_qk = [&](int j) {
      const int s = j % ST;
#pragma unroll
      for (int kk = 0; kk < KT; ++kk)
        umma_bf16(tS, make_smem_desc_sw128(aQ + kk * 32), make_smem_desc_sw128(aK + s * KH * Q48_KV_BYTES + kk * 32), idesc_qk,
                  kk > 0 ? 1u : 0u);
      umma_commit(&s_full[g]);
    };
    // one key tile of this slot's MMA stream: next scores as soon as S(j) has been read out, then P.V(j)
    auto mma_step = [&](int j) {
      const int s = j % ST;
      if (j + 1 < n_tiles) {
        mbar_wait(&kv_full[(j + 1) % ST], ((j + 1) / ST) & 1);
        mbar_wait(&s_free[g], j & 1);
        tc_fence_after();
        issue_qk(j + 1);
      }
      mbar_wait(&v_ready[s], (j / ST) & 1);                       // V_j carries its ones column
      mbar_wait(&p_full[g], j & 1);                               // P_j is in TMEM
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < BN / 16; ++k)
        umma_bf16_ts(tO, tP + k * 8, make_smem_desc_sw128_mn(aV + s * KH * Q48_KV_BYTES + k * 2048, Q48_KV_BYTES), idesc_pv,
                     (j | k) != 0 ? 1u : 0u);
      umma_commit(&pv_done[g]);
      umma_commit(&kv_empty[s]);                                  // this slot's reads of stage s (Q.K_j earlier, P.V_j now) are issued
    };
    if (warp == kTma) {
      // TMA producer (elected lane, ST - 1 tiles ahead) + the ONES COLUMN written by all 32 lanes into every V tile that has
      // landed (column D of each V row := 1.0, so column D of the P.V accumulator is the row sum of P) + slot 0's MMA stream
      auto issue_kv = [&](int j) {
        const int s = j % ST;
        mbar_wait(&kv_empty[s], ((j / ST) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], KH * 2 * Q48_KV_BYTES);
#pragma unroll
        for (int kh = 0; kh < KH; ++kh) {
          tma_load_4d(sK + (s * KH + kh) * Q48_KV_BYTES, &tmK, &kv_full[s], 0, h, (j + kh * n_tiles) * BN, b);
          tma_load_4d(sV + (s * KH + kh) * Q48_KV_BYTES, &tmV, &kv_full[s], 0, h, (j + kh * n_tiles) * BN, b);
        }
      };
      if (leader) {
        mbar_arrive_expect_tx(q_full, NQ * Q48_Q_BYTES);
#pragma unroll
        for (int qg = 0; qg < NQ; ++qg) tma_load_4d(sQ + qg * Q48_Q_BYTES, &tmQ, q_full, 0, h, m0 + qg * Q48_BM, b);
        for (int j = 0; j < ST - 1 && j < n_tiles; ++j) issue_kv(j);
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        issue_qk(0);
      }
      // tile t "lands": wait for its TMA, write the ones column, publish v_ready -- done LA tiles AHEAD of slot 0's MMA stream
      // (which shares this thread), so that no slot ever waits on slot 0's softmax for its V tile
      constexpr int LA = ST - 2;
      auto land = [&](int t) {
        const int s = t % ST;
        mbar_wait(&kv_full[s], (t / ST) & 1);
        for (int r = lane; r < KH * BN; r += 32)   // 128B-swizzled tile: element D of row r sits in chunk (D/8) ^ (r & 7)
          *reinterpret_cast<uint16_t*>(sV + s * KH * Q48_KV_BYTES + r * 128 + ((((D >> 3) ^ (r & 7)) << 4) | ((D & 7) << 1))) = 0x3F80;
        fence_proxy_async_smem();
        __syncwarp();
        if (leader) mbar_arrive(&v_ready[s]);
      };
      for (int t = 0; t < LA && t < n_tiles; ++t) land(t);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + LA < n_tiles) land(j + LA);
        if (leader) {
          if (j + ST - 1 < n_tiles) issue_kv(j + ST - 1);       // (stage of tile j - 1: released by all four slots' P.V(j-1))
          mma_step(j);
        }
        __syncwarp();
      }
      if (leader) umma_commit(o_full);
    } else if (leader) {
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) mma_step(j);
      umma_commit(o_full);
    }
  } else {
    const int g = warp >> 2, qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(g * TG);
    float m_ref = -INFINITY;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&s_full[g], j & 1);
      tc_fence_after();
      uint32_t v[BN];
      tmem_ld_32x32b_x48_wait(t_lane + Q48_S, v);
      tc_fence_before();
      mbar_arrive(&s_free[g]);               // S(j) is in registers: the issuer may overwrite it with S(j+1) now
      const int valid = p.Lk - (j + (SPLIT ? (g & 1) * n_tiles : 0)) * BN;
      if (valid < BN) {
#pragma unroll
        for (int i = 0; i < BN; ++i)
          if (i >= valid) v[i] = 0xff800000u;
      }
      float m4[4] = {__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])};
#pragma unroll
      for (int i = 4; i < BN; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) m4[u] = fmaxf(m4[u], __uint_as_float(v[i + u]));
      }
      const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
      if (j == 0) {
        m_ref = (mx == -INFINITY) ? 0.f : mx;
      } else {
        const bool need = mx > m_ref + 8.f;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? mx : m_ref;
          const float f = fast_exp2(m_ref - m_new);
          m_ref = m_new;
          mbar_wait(&pv_done[g], (j - 1) & 1);                    // the accumulator is only touched between two P.V
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < DO / 16; ++c) {
            uint32_t ov[16];
            tmem_ld_32x32b_x16(t_lane + (uint32_t)(Q48_O + c * 16), ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
            tmem_st_32x32b_x16(t_lane + (uint32_t)(Q48_O + c * 16), ov);
          }
        }
      }
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_ref, -m_ref);
      uint32_t pk[BN / 2];
#pragma unroll
      for (int i = 0; i < BN / 2; ++i) {
        const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
        const float2 e = ((i & 7) < EMU) ? q48_exp2_emu2<DEG>(t) : make_float2(fast_exp2(t.x), fast_exp2(t.y));
        pk[i] = __byte_perm(__float_as_uint(e.x), __float_as_uint(e.y), 0x7632);   // truncate to bf16; the row sum comes from the MMA
      }
      if (j > 0) {
        mbar_wait(&pv_done[g], (j - 1) & 1);                      // P.V(j-1) has finished reading the P region
        tc_fence_after();
      }
      {
        uint32_t (&lo)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[0]);
        uint32_t (&hi)[8] = *reinterpret_cast<uint32_t (*)[8]>(&pk[16]);
        tmem_st_32x32b_x16(t_lane + (uint32_t)Q48_P, lo);
        tmem_st_32x32b_x8(t_lane + (uint32_t)(Q48_P + 16), hi);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[g]);
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int qi = SPLIT ? g >> 1 : g;                              // query tile of this TMEM slot
    const int grow = m0 + qi * Q48_BM + row;
    bf16* orow = p.o + (long long)b * p.o_sb + (long long)grow * p.o_sn + h * D;
    if constexpr (SPLIT) {
      // merge the two key halves of a tile: the odd slot hands (m, O[0..DO)) to the even slot through shared memory
      // (the K / V rings are free -- every MMA has completed -- and contiguous: the scratch may run from sK into sV)
      float* xch = reinterpret_cast<float*>(sK) + (size_t)qi * Q48_BM * (DO + 1);
      float acc[DO];
#pragma unroll
      for (int c = 0; c < DO / 16; ++c) {
        uint32_t w[16];
        tmem_ld_32x32b_x16(t_lane + (uint32_t)(Q48_O + c * 16), w);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 16; ++x) acc[c * 16 + x] = __uint_as_float(w[x]);
      }
      if (g & 1) {
#pragma unroll
        for (int x = 0; x < DO; ++x) xch[x * Q48_BM + row] = acc[x];     // column-major: conflict-free
        xch[DO * Q48_BM + row] = m_ref;
      }
      asm volatile("bar.sync %0, 256;" ::"r"(1 + qi) : "memory");
      if (!(g & 1)) {
        const float m1 = xch[DO * Q48_BM + row];
        const float m = fmaxf(m_ref, m1);
        const float f0 = fast_exp2(m_ref - m), f1 = fast_exp2(m1 - m);
#pragma unroll
        for (int x = 0; x < DO; ++x) acc[x] = acc[x] * f0 + xch[x * Q48_BM + row] * f1;
        const float l_run = acc[D];
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        if (p.lse && grow < p.Lq) p.lse[((long long)b * H + h) * p.Lq + grow] = l_run > 0.f ? m + log2f(l_run) : INFINITY;
        if (grow < p.Lq) {
#pragma unroll
          for (int c8 = 0; c8 < D / 8; ++c8) {
            uint4 o4;
            o4.x = pack_bf16(acc[c8 * 8 + 0] * inv, acc[c8 * 8 + 1] * inv);
            o4.y = pack_bf16(acc[c8 * 8 + 2] * inv, acc[c8 * 8 + 3] * inv);
            o4.z = pack_bf16(acc[c8 * 8 + 4] * inv, acc[c8 * 8 + 5] * inv);
            o4.w = pack_bf16(acc[c8 * 8 + 6] * inv, acc[c8 * 8 + 7] * inv);
            *reinterpret_cast<uint4*>(orow + c8 * 8) = o4;
          }
        }
      }
    } else {
      float l_run;                                                 // row sum of P = column D of the accumulator
      {
        uint32_t v8[8];
        tmem_ld_32x32b_x8(t_lane + (uint32_t)(Q48_O + (D & ~7)), v8);
        tmem_ld_wait();
        l_run = __uint_as_float(v8[D & 7]);
      }
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      if (p.lse && grow < p.Lq)
        p.lse[((long long)b * H + h) * p.Lq + grow] = l_run > 0.f ? m_ref + log2f(l_run) : INFINITY;
#pragma unroll
      for (int c = 0; c < DO / 16; ++c) {
        uint32_t w[16];
        tmem_ld_32x32b_x16(t_lane + (uint32_t)(Q48_O + c * 16), w);
        tmem_ld_wait();
        if (grow < p.Lq) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (c * 16 + half * 8 < D) {
              uint4 o4;
              o4.x = pack_bf16(__uint_as_float(w[half * 8 + 0]) * inv, __uint_as_float(w[half * 8 + 1]) * inv);
              o4.y = pack_bf16(__uint_as_float(w[half * 8 + 2]) * inv, __uint_as_float(w[half * 8 + 3]) * inv);
              o4.z = pack_bf16(__uint_as_float(w[half * 8 + 4]) * inv, __uint_as_float(w[half * 8 + 5]) * inv);
              o4.w = pack_bf16(__uint_as_float(w[half * 8 + 6]) * inv, __uint_as_float(w[half * 8 + 7]) * inv);
              *reinterpret_cast<uint4*>(orow + c * 16 + half * 8) = o4;
            }
          }
        }
      }
    }
  }
  pdl_launch_dependents();     // late trigger: the successor's CTAs are scheduled while this one tears down
  tc_fence_before();
  __syncthreads();
  if (warp == kMma + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, G * TG);
  }
}

template <int EMU, int DEG>
static int launch_q48(const CUtensorMap& tQ, const CUtensorMap& tK, const CUtensorMap& tV, const TaParams& p, int B, int H,
                      cudaStream_t stream) {
  constexpr int smem4 = 4 * Q48_Q_BYTES + Q48_ST * 2 * Q48_KV_BYTES + 1024 + 512;
  constexpr int smemS = 2 * Q48_Q_BYTES + Q48_ST * 2 * 2 * Q48_KV_BYTES + 1024 + 512;
  static_assert(2 * Q48_BM * (Q48_DO + 1) * 4 <= 2 * Q48_ST * 2 * Q48_KV_BYTES, "SPLIT merge scratch must fit the K + V rings (contiguous, idle by then)");
  AF_CONFIG_SMEM((attn_fwd_tcgen05_q48_kernel<EMU, DEG, false>), smem4);
  AF_CONFIG_SMEM((attn_fwd_tcgen05_q48_kernel<EMU, DEG, true>), smemS);
  const int n_sm = af_num_sms();
  const int n_qb4 = (p.Lq + 4 * Q48_BM - 1) / (4 * Q48_BM);
  const int units4 = B * H * n_qb4;
  const int n_ktiles = (p.Lk + Q48_BN - 1) / Q48_BN;
  // bulk: whole rounds of four-tile CTAs; tail: the rest as split-key CTAs (two tiles x two key halves), which need an even
  // number of key tiles and whole 256-query units
  int bulk = (units4 / n_sm) * n_sm;
  const bool can_split = (p.Lq % (4 * Q48_BM) == 0) && (n_ktiles % 2 == 0);
  if (!can_split || bulk == 0) bulk = units4;
  if (bulk > 0) {
    AF_CUDA(launch_pdl(2, attn_fwd_tcgen05_q48_kernel<EMU, DEG, false>, dim3(bulk), dim3((4 * Q48_G + Q48_G) * 32), smem4, stream, tQ, tK, tV, p, 0, H));
    ++g_launch_count;
  }
  if (units4 > bulk) {
    AF_CUDA(launch_pdl(2, attn_fwd_tcgen05_q48_kernel<EMU, DEG, true>, dim3(2 * (units4 - bulk)), dim3((4 * Q48_G + Q48_G) * 32), smemS, stream, tQ, tK, tV, p,
                       2 * bulk, H));
    ++g_launch_count;
  }
  AF_CUDA(cudaGetLastError());
  return 0;
}

// d = 40, unmasked, Lq >= 1024, Lk > 128.  Returns -1 when the shape is not eligible.
int attn_fwd_tcgen05_q48(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sh, int64_t k_sn,
                         const void* v, int64_t v_sb, int64_t v_sh, int64_t v_sn, int64_t B, int64_t H, int64_t Lq, int64_t Lk,
                         int64_t drow_q, int64_t drow_kv, const TaParams& p, cudaStream_t stream) {
  static int emu = -1, deg = 3;
  if (emu < 0) {
    const char* e = getenv("ADAFACE_EXP_EMU");       // exp2 pairs of every 8 on the FMA pipe (0..6)
    const char* dg = getenv("ADAFACE_EXP_DEG");      // polynomial degree of the emulated exp2 (2 | 3)
    deg = (dg && dg[0] == '3') ? 3 : 2;
    emu = (e && e[0] >= '0' && e[0] <= '6') ? (e[0] - '0') : 3;
  }
  CUtensorMap tQ, tK, tV;
  if (make_tmap_bf16_heads(&tQ, q, (uint64_t)drow_q, (uint64_t)H, (uint64_t)Lq, (uint64_t)B, (uint64_t)q_sh, (uint64_t)q_sn, (uint64_t)q_sb, Q48_BM)) return 3;
  if (make_tmap_bf16_heads(&tK, k, (uint64_t)drow_kv, (uint64_t)H, (uint64_t)Lk, (uint64_t)B, (uint64_t)k_sh, (uint64_t)k_sn, (uint64_t)k_sb, Q48_BN)) return 3;
  if (make_tmap_bf16_heads(&tV, v, (uint64_t)drow_kv, (uint64_t)H, (uint64_t)Lk, (uint64_t)B, (uint64_t)v_sh, (uint64_t)v_sn, (uint64_t)v_sb, Q48_BN)) return 3;
  const int ib = (int)B, ih = (int)H;
#define Q48_CASE(E, DG) if (emu == E && deg == DG) return launch_q48<E, DG>(tQ, tK, tV, p, ib, ih, stream);
  Q48_CASE(0, 2) Q48_CASE(2, 2) Q48_CASE(3, 2) Q48_CASE(4, 2) Q48_CASE(5, 2) Q48_CASE(2, 3) Q48_CASE(3, 3) Q48_CASE(4, 3)
#undef Q48_CASE
  return launch_q48<3, 2>(tQ, tK, tV, p, ib, ih, stream);
}

}  // namespace adaface

// K5 (tensor-core variant): flash-attention backward on tcgen05 / TMEM for self- / cross-attention (optional key mask) with head
// dims 40 / 80 -- the same recompute-form two-pass scheme as attn_bwd_mma.cu (no atomics, deterministic), with every
// GEMM on the 5th-generation tensor cores:
//
//   dq pass   (CTA = 128 queries = TMEM lanes; loop over 64-key tiles)
//       S  = Q K_j^T,  dP = dO V_j^T            (SS MMAs, M=128, N=64)            -> TMEM [0,64), [64,128)
//       dS = exp2(S scale - lse) * (dP - delta) * scale -> bf16, written over S   (row-per-thread, no shuffles)
//       dQ += dS K_j                            (A from TMEM, K_j read MN-major)  -> TMEM [128, 128+DO)
//   dk/dv pass (CTA = 128 keys = TMEM lanes; loop over 64-query tiles)
//       S^T = K Q_j^T,  dP^T = V dO_j^T         (SS MMAs)                         -> TMEM [0,64), [64,128)
//       P^T -> bf16 over S^T, dS^T -> bf16 over dP^T   (lse_j / delta_j broadcast from shared memory per column)
//       dV += P^T dO_j,  dK += dS^T Q_j         (A from TMEM, dO_j / Q_j MN-major) -> TMEM [128,..), [128+DO,..)
//
// Q / K / V / dO tiles arrive by TMA through 4-D maps {d, head, token, batch} of the caller's [B, L, H*d]-style views
// (zero fill pads d = 40 to the 48-wide MMA K); warp roles as in attn_tcgen05.cu (4 softmax warps, TMA producer,
// MMA issuer on the high warp ids).  P and dS are truncated to bf16 (PRMT instead of F2FP: the XU pipe is what bounds
// these kernels -- one exp2 per score element, as in the forward pass).
// Reference: autograd of F.scaled_dot_product_attention at dalc:321 / ldm attention.py:181-204.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

extern long long g_launch_count;

constexpr int TB_BM = 128;        // TMEM lanes: queries (dq pass) or keys (dk/dv pass)
constexpr int TB_BN = 64;         // streamed tile: keys (dq pass) or queries (dk/dv pass)
constexpr int TB_THREADS = 192;
constexpr int kTbTma = 4, kTbMma = 5;
constexpr int TB_ST = 2;          // streamed-tile pipeline stages

template <int D>
struct TbCfg {
  static constexpr int NA = (D + 63) / 64;
  static constexpr int KT = (D + 15) / 16;
  static constexpr int DO = KT * 16;
  static constexpr int BIG_BYTES = NA * TB_BM * 128;      // a resident 128-row tile
  static constexpr int SMALL_ATOM = TB_BN * 128;          // one 64-row x 64-column atom of a streamed tile
  static constexpr int SMALL_BYTES = NA * SMALL_ATOM;
  static constexpr int TMEM_ACC = 2 * TB_BN;              // accumulators start behind S / dP
};

struct TbParams {
  const float* lse;        // [B, H, Lq] log2 domain
  const float* delta;      // [B, H, Lq]
  bf16 *dq, *dk, *dv;
  long long dq_sb, dq_sn, dk_sb, dk_sn, dv_sb, dv_sn;
  int Lq, Lk, H;
  float scale, scale_log2;
  const uint8_t* key_mask;   // optional [B, Lk] (0 = key masked out, dalc:254-273 img_mask); needs Lk % 64 == 0
};

// ------------------------------------------------------------------------------------------------ dq pass
// BN = keys per streamed tile: 32 keeps S (32) + dP (32) + dQ (48) within 128 TMEM columns => 4 CTAs / SM for d = 40
// (16 softmax warps per SM instead of 8: the pass is bound by exp2 latency, not by the tensor pipe).
template <int D, int BN>
__global__ void __launch_bounds__(TB_THREADS, (D <= 64 && BN <= 32) ? 4 : 2)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, const TbParams p) {
  using Cfg = TbCfg<D>;
  constexpr int NA = Cfg::NA, KT = Cfg::KT, DO = Cfg::DO, ST = TB_ST;
  constexpr int SMALL_ATOM = BN * 128, SMALL_BYTES = NA * SMALL_ATOM, TMEM_ACC = 2 * BN;
  constexpr int TMEM_COLS = (TMEM_ACC + DO <= 128) ? 128 : (TMEM_ACC + DO <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw_bq[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_bq) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                    // [NA][128][128 B]
  uint8_t* sdO = sQ + Cfg::BIG_BYTES;
  uint8_t* sK = sdO + Cfg::BIG_BYTES;                    // [ST][NA][64][128 B]
  uint8_t* sV = sK + ST * SMALL_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ST * SMALL_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                          // [ST]
  uint64_t* kv_empty = kv_full + ST;                     // [ST]
  uint64_t* s_full = kv_empty + ST;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TB_BM, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (p.Lk + BN - 1) / BN;

  if (warp == kTbTma && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
    mbar_init(q_full, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  } else if (warp == kTbMma) {
    tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kTbTma) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * Cfg::BIG_BYTES);
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        tma_load_4d(sQ + a * (TB_BM * 128), &tmQ, q_full, a * 64, h, m0, b);
        tma_load_4d(sdO + a * (TB_BM * 128), &tmdO, q_full, a * 64, h, m0, b);
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % ST;
        mbar_wait(&kv_empty[s], ((j / ST) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * SMALL_BYTES);
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          tma_load_4d(sK + s * SMALL_BYTES + a * SMALL_ATOM, &tmK, &kv_full[s], a * 64, h, j * BN, b);
          tma_load_4d(sV + s * SMALL_BYTES + a * SMALL_ATOM, &tmV, &kv_full[s], a * 64, h, j * BN, b);
        }
      }
    }
  } else if (warp == kTbMma) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(TB_BM, BN, false);
      constexpr uint32_t idesc_acc = make_idesc_bf16_f32(TB_BM, DO, true);
      const uint32_t aQ = smem_u32(sQ), adO = smem_u32(sdO), aK = smem_u32(sK), aV = smem_u32(sV);
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % ST;
        mbar_wait(&kv_full[s], (j / ST) & 1);
        tc_fence_after();
        // S_j and dP_j overwrite the previous tile's operands: tcgen05.mma executes in issue order, and the dQ MMA
        // of tile j-1 (which read dS_{j-1}) was issued before them.
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
          const uint32_t koff = (kk >> 2) * SMALL_ATOM + (kk & 3) * 32, qoff = (kk >> 2) * (TB_BM * 128) + (kk & 3) * 32;
          umma_bf16(tmem_base, make_smem_desc_sw128(aQ + qoff), make_smem_desc_sw128(aK + s * SMALL_BYTES + koff), idesc_s,
                    kk > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
          const uint32_t koff = (kk >> 2) * SMALL_ATOM + (kk & 3) * 32, qoff = (kk >> 2) * (TB_BM * 128) + (kk & 3) * 32;
          umma_bf16(tmem_base + BN, make_smem_desc_sw128(adO + qoff), make_smem_desc_sw128(aV + s * SMALL_BYTES + koff),
                    idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          const uint64_t db = make_smem_desc_sw128_mn(aK + s * SMALL_BYTES + k * 2048, SMALL_ATOM);
          umma_bf16_ts(tmem_base + (uint32_t)TMEM_ACC, tmem_base + (uint32_t)(k * 8), db, idesc_acc, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[s]);
      }
      umma_commit(o_full);
    }
  } else {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const int grow = m0 + row;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qd * 32) << 16);
    const long long stat = ((long long)b * p.H + h) * p.Lq + grow;
    const float lse = grow < p.Lq ? p.lse[stat] : INFINITY;          // +inf => P = 0 for rows past the end
    const float delta = grow < p.Lq ? p.delta[stat] : 0.f;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nl2 = make_float2(-lse, -lse);
    const float2 nd2 = make_float2(-delta, -delta), ss2 = make_float2(p.scale, p.scale);
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int valid = p.Lk - j * BN;
#pragma unroll
      for (int hf = 0; hf < BN / 32; ++hf) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32b_x32_wait(t_lane + (uint32_t)(hf * 32), sv);
        tmem_ld_32x32b_x32_wait(t_lane + (uint32_t)(BN + hf * 32), dv);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 t = __ffma2_rn(make_float2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), sc2, nl2);
          const float2 pr = make_float2(fast_exp2(t.x), fast_exp2(t.y));
          float2 g = __fadd2_rn(make_float2(__uint_as_float(dv[2 * i]), __uint_as_float(dv[2 * i + 1])), nd2);
          g = __fmul2_rn(__fmul2_rn(pr, g), ss2);
          uint32_t gx = __float_as_uint(g.x), gy = __float_as_uint(g.y);
          if (valid < BN) {
            if (hf * 32 + 2 * i >= valid) gx = 0;
            if (hf * 32 + 2 * i + 1 >= valid) gy = 0;
          }
          pk[i] = __byte_perm(gx, gy, 0x7632);
        }
        if (p.key_mask) {      // masked keys: P = 0, hence dS = 0 (32 mask bytes of this half, the same for every row: broadcast loads)
          const uint4* mp = reinterpret_cast<const uint4*>(p.key_mask + (long long)b * p.Lk + j * BN + hf * 32);
          const uint4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
          const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            const uint32_t nz = __vcmpne4(mw[w], 0u);               // 0xff per kept key (4 keys per word)
            pk[2 * w] &= __byte_perm(nz, 0u, 0x1100);               // keys 4w, 4w + 1 -> the two bf16 halves of a packed pair
            pk[2 * w + 1] &= __byte_perm(nz, 0u, 0x3322);           // keys 4w + 2, 4w + 3
          }
        }
        tmem_st_32x32b_x16(t_lane + (uint32_t)(hf * 16), pk);     // columns [16 hf, 16 hf + 16) of S: already consumed
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    bf16* orow = p.dq + (long long)b * p.dq_sb + (long long)grow * p.dq_sn + h * D;
#pragma unroll
    for (int c = 0; c < DO / 16; ++c) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(t_lane + (uint32_t)(TMEM_ACC + c * 16), v);
      tmem_ld_wait();
      if (grow < p.Lq) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (c * 16 + half * 8 < D) {
            uint4 pk;
            pk.x = pack_bf16(__uint_as_float(v[half * 8 + 0]), __uint_as_float(v[half * 8 + 1]));
            pk.y = pack_bf16(__uint_as_float(v[half * 8 + 2]), __uint_as_float(v[half * 8 + 3]));
            pk.z = pack_bf16(__uint_as_float(v[half * 8 + 4]), __uint_as_float(v[half * 8 + 5]));
            pk.w = pack_bf16(__uint_as_float(v[half * 8 + 6]), __uint_as_float(v[half * 8 + 7]));
            *reinterpret_cast<uint4*>(orow + c * 16 + half * 8) = pk;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kTbMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ dk/dv pass
template <int D>
__global__ void __launch_bounds__(TB_THREADS, (D <= 64) ? 2 : 1)
attn_bwd_dkdv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, const TbParams p) {
  using Cfg = TbCfg<D>;
  constexpr int NA = Cfg::NA, KT = Cfg::KT, DO = Cfg::DO, ST = TB_ST;
  constexpr int TMEM_DV = Cfg::TMEM_ACC, TMEM_DK = Cfg::TMEM_ACC + DO;
  constexpr int TMEM_COLS = (TMEM_DK + DO <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw_bk[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_bk) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;                                    // [NA][128][128 B]
  uint8_t* sV = sK + Cfg::BIG_BYTES;
  uint8_t* sQ = sV + Cfg::BIG_BYTES;                     // [ST][NA][64][128 B]
  uint8_t* sdO = sQ + ST * Cfg::SMALL_BYTES;
  float* sStat = reinterpret_cast<float*>(sdO + ST * Cfg::SMALL_BYTES);   // [2 buffers][lse(64) | delta(64)]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + 2 * 2 * TB_BN);
  uint64_t* k_full = bars;
  uint64_t* q_full = bars + 1;                           // [ST]
  uint64_t* q_empty = q_full + ST;                       // [ST]
  uint64_t* s_full = q_empty + ST;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TB_BM, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (p.Lq + TB_BN - 1) / TB_BN;

  if (warp == kTbTma && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
    mbar_init(k_full, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  } else if (warp == kTbMma) {
    tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kTbTma) {
    if (elect_one()) {
      mbar_arrive_expect_tx(k_full, 2 * Cfg::BIG_BYTES);
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        tma_load_4d(sK + a * (TB_BM * 128), &tmK, k_full, a * 64, h, n0, b);
        tma_load_4d(sV + a * (TB_BM * 128), &tmV, k_full, a * 64, h, n0, b);
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % ST;
        mbar_wait(&q_empty[s], ((j / ST) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[s], 2 * Cfg::SMALL_BYTES);
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          tma_load_4d(sQ + s * Cfg::SMALL_BYTES + a * Cfg::SMALL_ATOM, &tmQ, &q_full[s], a * 64, h, j * TB_BN, b);
          tma_load_4d(sdO + s * Cfg::SMALL_BYTES + a * Cfg::SMALL_ATOM, &tmdO, &q_full[s], a * 64, h, j * TB_BN, b);
        }
      }
    }
  } else if (warp == kTbMma) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16_f32(TB_BM, TB_BN, false);
      constexpr uint32_t idesc_acc = make_idesc_bf16_f32(TB_BM, DO, true);
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), adO = smem_u32(sdO);
      mbar_wait(k_full, 0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % ST;
        mbar_wait(&q_full[s], (j / ST) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
          const uint32_t boff = (kk >> 2) * (TB_BM * 128) + (kk & 3) * 32, soff = (kk >> 2) * Cfg::SMALL_ATOM + (kk & 3) * 32;
          umma_bf16(tmem_base, make_smem_desc_sw128(aK + boff), make_smem_desc_sw128(aQ + s * Cfg::SMALL_BYTES + soff), idesc_s,
                    kk > 0 ? 1u : 0u);                                            // S^T = K Q_j^T
        }
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
          const uint32_t boff = (kk >> 2) * (TB_BM * 128) + (kk & 3) * 32, soff = (kk >> 2) * Cfg::SMALL_ATOM + (kk & 3) * 32;
          umma_bf16(tmem_base + TB_BN, make_smem_desc_sw128(aV + boff), make_smem_desc_sw128(adO + s * Cfg::SMALL_BYTES + soff),
                    idesc_s, kk > 0 ? 1u : 0u);                                   // dP^T = V dO_j^T
        }
        umma_commit(s_full);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < TB_BN / 16; ++k) {
          const uint64_t dbo = make_smem_desc_sw128_mn(adO + s * Cfg::SMALL_BYTES + k * 2048, Cfg::SMALL_ATOM);
          umma_bf16_ts(tmem_base + (uint32_t)TMEM_DV, tmem_base + (uint32_t)(k * 8), dbo, idesc_acc, (j | k) != 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < TB_BN / 16; ++k) {
          const uint64_t dbq = make_smem_desc_sw128_mn(aQ + s * Cfg::SMALL_BYTES + k * 2048, Cfg::SMALL_ATOM);
          umma_bf16_ts(tmem_base + (uint32_t)TMEM_DK, tmem_base + (uint32_t)(TB_BN + k * 8), dbq, idesc_acc, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&q_empty[s]);
      }
      umma_commit(o_full);
    }
  } else {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;                      // key row inside the tile
    const int tid = threadIdx.x;                         // 0..127 (softmax warps are warps 0..3)
    const uint32_t t_lane = tmem_base + ((uint32_t)(qd * 32) << 16);
    const long long stat0 = ((long long)b * p.H + h) * p.Lq;
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), ss2 = make_float2(p.scale, p.scale);
    // key mask: this thread's key is the TMEM lane; a masked key has P^T = dS^T = 0 for every query
    uint32_t kmask = 0xffffffffu;
    if (p.key_mask && n0 + row < p.Lk && p.key_mask[(long long)b * p.Lk + n0 + row] == 0) kmask = 0u;
    for (int j = 0; j < n_tiles; ++j) {
      // per-query statistics of this tile -> shared memory (double-buffered: the barrier below is the only sync)
      float* st = sStat + (j & 1) * 2 * TB_BN;
      {
        const int qi = j * TB_BN + (tid & 63);
        float x;
        if (tid < 64) x = qi < p.Lq ? p.lse[stat0 + qi] : INFINITY;              // +inf => P = 0 for queries past the end
        else x = qi < p.Lq ? p.delta[stat0 + qi] : 0.f;
        st[tid] = x;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(s_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32b_x32_wait(t_lane + (uint32_t)(hf * 32), sv);
        tmem_ld_32x32b_x32_wait(t_lane + (uint32_t)(TB_BN + hf * 32), dv);
        uint32_t pp[16], pd[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 l2 = *reinterpret_cast<const float2*>(st + hf * 32 + 2 * i);
          const float2 d2 = *reinterpret_cast<const float2*>(st + TB_BN + hf * 32 + 2 * i);
          const float2 t = __ffma2_rn(make_float2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), sc2,
                                      make_float2(-l2.x, -l2.y));
          const float2 pr = make_float2(fast_exp2(t.x), fast_exp2(t.y));
          float2 g = __fadd2_rn(make_float2(__uint_as_float(dv[2 * i]), __uint_as_float(dv[2 * i + 1])), make_float2(-d2.x, -d2.y));
          g = __fmul2_rn(__fmul2_rn(pr, g), ss2);
          pp[i] = __byte_perm(__float_as_uint(pr.x), __float_as_uint(pr.y), 0x7632) & kmask;
          pd[i] = __byte_perm(__float_as_uint(g.x), __float_as_uint(g.y), 0x7632) & kmask;
        }
        tmem_st_32x32b_x16(t_lane + (uint32_t)(hf * 16), pp);             // P^T over S^T
        tmem_st_32x32b_x16(t_lane + (uint32_t)(TB_BN + hf * 16), pd);     // dS^T over dP^T
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int grow = n0 + row;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      bf16* orow = which ? p.dk + (long long)b * p.dk_sb + (long long)grow * p.dk_sn + h * D
                         : p.dv + (long long)b * p.dv_sb + (long long)grow * p.dv_sn + h * D;
      const int tcol = which ? TMEM_DK : TMEM_DV;
#pragma unroll
      for (int c = 0; c < DO / 16; ++c) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_lane + (uint32_t)(tcol + c * 16), v);
        tmem_ld_wait();
        if (grow < p.Lk) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (c * 16 + half * 8 < D) {
              uint4 pk;
              pk.x = pack_bf16(__uint_as_float(v[half * 8 + 0]), __uint_as_float(v[half * 8 + 1]));
              pk.y = pack_bf16(__uint_as_float(v[half * 8 + 2]), __uint_as_float(v[half * 8 + 3]));
              pk.z = pack_bf16(__uint_as_float(v[half * 8 + 4]), __uint_as_float(v[half * 8 + 5]));
              pk.w = pack_bf16(__uint_as_float(v[half * 8 + 6]), __uint_as_float(v[half * 8 + 7]));
              *reinterpret_cast<uint4*>(orow + c * 16 + half * 8) = pk;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kTbMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int D>
static int launch_bwd_tc(const CUtensorMap& tQb, const CUtensorMap& tKs32, const CUtensorMap& tVs32, const CUtensorMap& tdOb,
                         const CUtensorMap& tQs, const CUtensorMap& tKb, const CUtensorMap& tVb, const CUtensorMap& tdOs,
                         const TbParams& p, int B, int H, cudaStream_t stream) {
  using Cfg = TbCfg<D>;
  constexpr int smem = 2 * Cfg::BIG_BYTES + 2 * TB_ST * Cfg::SMALL_BYTES + 2 * 2 * TB_BN * 4 + 1024 + 128;
  static_assert(smem <= 227 * 1024, "tcgen05 attention backward: shared memory exceeds the SM");
  constexpr int DQ_BN = 64;      // 32 (=> 128 TMEM columns, 4 CTAs / SM at d = 40) measured no faster: 1280 vs 1297 us at B = 8, slower at B = 1
  constexpr int smem_dq = 2 * Cfg::BIG_BYTES + 2 * TB_ST * Cfg::NA * DQ_BN * 128 + 1024 + 128;
  static DevOnce configured;
  const int cfg_dev = af_device();
  if (!configured.done(cfg_dev)) {
    AF_CUDA(cudaFuncSetAttribute((attn_bwd_dq_tc_kernel<D, DQ_BN>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq));
    AF_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.set(cfg_dev);
  }
  attn_bwd_dkdv_tc_kernel<D><<<dim3((p.Lk + TB_BM - 1) / TB_BM, H, B), TB_THREADS, smem, stream>>>(tQs, tKb, tVb, tdOs, p);
  attn_bwd_dq_tc_kernel<D, DQ_BN><<<dim3((p.Lq + TB_BM - 1) / TB_BM, H, B), TB_THREADS, smem_dq, stream>>>(tQb, tKs32, tVs32, tdOb, p);
  AF_CUDA(cudaGetLastError());
  g_launch_count += 2;
  return 0;
}

// Non-causal backward on the tensor cores, with an optional key mask.  Returns -1 when the problem is not eligible (the caller then
// runs the warp-MMA kernels, which handle causal multi-KV, d = 64 / 160, ragged masked lengths and tiny shapes).  delta must
// already hold rowsum(dO o O).
int attn_bwd_tcgen05(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn, const void* v,
                     int64_t v_sb, int64_t v_sn, const void* dout, int64_t do_sb, int64_t do_sn, const float* lse,
                     const float* delta, void* dq, int64_t dq_sb, int64_t dq_sn, void* dk, int64_t dk_sb, int64_t dk_sn, void* dv,
                     int64_t dv_sb, int64_t dv_sn, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t d, float scale,
                     const uint8_t* key_mask, cudaStream_t stream) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("ADAFACE_BWD_TC");          // 1 (default): tcgen05 backward where eligible
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled || !(d == 40 || d == 80) || Lq < 256 || Lk < 64) return -1;
  if (key_mask && Lk % 64 != 0) return -1;      // the dq pass reads the mask in 32-byte pieces per key tile
  CUtensorMap tQb, tKs, tVs, tdOb, tQs, tKb, tVb, tdOs;
  const uint64_t ud = (uint64_t)d, uH = (uint64_t)H, uB = (uint64_t)B;
  if (make_tmap_bf16_heads(&tQb, q, ud, uH, (uint64_t)Lq, uB, ud, (uint64_t)q_sn, (uint64_t)q_sb, TB_BM)) return 3;
  if (make_tmap_bf16_heads(&tdOb, dout, ud, uH, (uint64_t)Lq, uB, ud, (uint64_t)do_sn, (uint64_t)do_sb, TB_BM)) return 3;
  const uint32_t dq_bn = 64;                     // streamed key tile of the dq pass (launch_bwd_tc::DQ_BN)
  if (make_tmap_bf16_heads(&tKs, k, ud, uH, (uint64_t)Lk, uB, ud, (uint64_t)k_sn, (uint64_t)k_sb, dq_bn)) return 3;
  if (make_tmap_bf16_heads(&tVs, v, ud, uH, (uint64_t)Lk, uB, ud, (uint64_t)v_sn, (uint64_t)v_sb, dq_bn)) return 3;
  if (make_tmap_bf16_heads(&tQs, q, ud, uH, (uint64_t)Lq, uB, ud, (uint64_t)q_sn, (uint64_t)q_sb, TB_BN)) return 3;
  if (make_tmap_bf16_heads(&tdOs, dout, ud, uH, (uint64_t)Lq, uB, ud, (uint64_t)do_sn, (uint64_t)do_sb, TB_BN)) return 3;
  if (make_tmap_bf16_heads(&tKb, k, ud, uH, (uint64_t)Lk, uB, ud, (uint64_t)k_sn, (uint64_t)k_sb, TB_BM)) return 3;
  if (make_tmap_bf16_heads(&tVb, v, ud, uH, (uint64_t)Lk, uB, ud, (uint64_t)v_sn, (uint64_t)v_sb, TB_BM)) return 3;
  TbParams p;
  p.lse = lse; p.delta = delta;
  p.dq = (bf16*)dq; p.dk = (bf16*)dk; p.dv = (bf16*)dv;
  p.dq_sb = dq_sb; p.dq_sn = dq_sn; p.dk_sb = dk_sb; p.dk_sn = dk_sn; p.dv_sb = dv_sb; p.dv_sn = dv_sn;
  p.Lq = (int)Lq; p.Lk = (int)Lk; p.H = (int)H;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.key_mask = key_mask;
  return d == 40 ? launch_bwd_tc<40>(tQb, tKs, tVs, tdOb, tQs, tKb, tVb, tdOs, p, (int)B, (int)H, stream)
                 : launch_bwd_tc<80>(tQb, tKs, tVs, tdOb, tQs, tKb, tVb, tdOs, p, (int)B, (int)H, stream);
}

}  // namespace adaface

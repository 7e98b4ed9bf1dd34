// K2 (tensor-bound variant): flash self-attention on tcgen05 / TMEM for SD-1.5 head dims 40 / 80 / 160.
//
// One CTA = 128 queries of one (batch, head); two CTAs co-reside per SM (d = 40) so one CTA's tensor work overlaps the
// other's softmax.  192 threads:
//   warp 0     TMA producer.  Q/K/V tiles come straight from the reference layout [B, L, H*d] through 4-D tensor maps
//              {d, head, token, batch}; the 64-column box is wider than d = 40, and TMA zero-fills the out-of-bounds
//              columns, which is what pads the MMA K dimension to 48 -- no padded copy ever exists in HBM.
//   warp 1     TMEM allocator + the single thread that issues every tcgen05.mma:
//                 S_j  = Q K_j^T      (M=128, N=64, K=16 x ceil(d/16); A, B K-major, 128B swizzle)   -> TMEM S[j&1]
//                 O   += P_j V_j      (M=128, N=ceil16(d), K=16 x 4; A = P from smem, B = V tile read MN-major)
//              QK_{j+1} is issued before PV_j so the tensor pipe works ahead of the softmax warps.
//   warps 2-5  softmax, one query row per thread (no shuffles): tcgen05.ld the 64 scores, lazy-rescaled online
//              softmax (O in TMEM is only rescaled when the row max grows by > 2^8), bf16 P written to shared memory
//              in the swizzled K-major layout, final O / l epilogue.
// Reference arithmetic: F.scaled_dot_product_attention at dalc:321 / ldm attention.py:181-204 (no mask).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "attn_tc_params.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

#ifdef AF_ATTN_TRACE
#define AF_ATTN_TR(...) __VA_ARGS__
#else
#define AF_ATTN_TR(...)
#endif

extern long long g_launch_count;
int attn_fwd_tcgen05_tri(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sh, int64_t k_sn,
                         const void* v, int64_t v_sb, int64_t v_sh, int64_t v_sn, int64_t B, int64_t H, int64_t Lq, int64_t Lk,
                         int64_t drow_q, int64_t drow_kv, const TaParams& p, cudaStream_t stream);

constexpr int TA_BM = 128;
constexpr int TA_BN = 64;
constexpr int TA_THREADS = 192;
// Warp roles.  The SM sub-partition scheduler favours HIGHER warp ids, so the two control warps (TMA producer, MMA
// issuer) sit above the four softmax warps: with the control warps at ids 0/1 the exp-heavy softmax warps starved
// the MMA issuer and every tile waited for its scores.
constexpr int kTmaWarp = 4, kMmaWarp = 5;

template <int D>
struct TaCfg {
  static constexpr int NA = (D + 63) / 64;          // 64-column swizzle atoms per row
  static constexpr int KT = (D + 15) / 16;          // k16 steps of Q K^T
  static constexpr int DO = KT * 16;                // PV MMA N / accumulator columns
  static constexpr int ST = (D <= 64) ? 3 : 2;      // K/V pipeline stages
  static constexpr int Q_BYTES = NA * TA_BM * 128;
  static constexpr int KV_ATOM = TA_BN * 128;       // one 64-key x 64-column atom
  static constexpr int K_BYTES = NA * KV_ATOM;
  static constexpr int V_BYTES = NA * KV_ATOM;
  static constexpr int P_BYTES = TA_BM * 128;       // 128 rows x 64 keys bf16
  static constexpr int TMEM_O = 2 * TA_BN;          // O starts after the two S buffers
  static constexpr int TMEM_P = TMEM_O + DO;        // two bf16 P buffers (32 columns each) when P lives in TMEM
  static constexpr int TMEM_COLS = (TMEM_P + TA_BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = Q_BYTES + ST * (K_BYTES + V_BYTES) + 2 * P_BYTES + 1024 + 256;
};


// exp2 on the FMA/ALU pipes for a pair of values in [-126, 8]: Cody-Waite split with the 1.5*2^23 magic constant,
// degree-3 minimax polynomial of 2^f on [-0.5, 0.5] (max rel. error 7.5e-5, far below bf16's 2^-9), exponent
// re-inserted with one integer shift-add.  Offloads part of the softmax from the 16-op/clk MUFU unit, which is
// what bounds d = 40 attention (one exp per 160 tensor FLOPs).
template <int DEG = 3>
__device__ __forceinline__ float2 exp2_emu2(float2 x) {
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
  } else {      // degree-2 minimax: max rel. error 1.7e-3, below the 3.9e-3 of the truncating bf16 pack that follows
    q = __ffma2_rn(f, make_float2(0.23842894f, 0.23842894f), make_float2(0.7034480f, 0.7034480f));
    q = __ffma2_rn(q, f, make_float2(1.0004431f, 1.0004431f));
  }
  float2 r;
  r.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
  return r;
}

// EMU = how many of every 8 exp2 pairs are evaluated by exp2_emu2 instead of MUFU.EX2 (0..4).
// PT  = P is handed to the P.V MMA through TMEM (tcgen05.st + A-from-TMEM MMA) instead of shared memory: the P
//       round trip was 45% of the kernel's shared-memory traffic (ncu: smem data pipe 72% busy with P in smem).
template <int D, int EMU, bool PT>
__global__ void __launch_bounds__(TA_THREADS, (D <= 64) ? 2 : 1)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const TaParams p) {
  using Cfg = TaCfg<D>;
  constexpr int NA = Cfg::NA, KT = Cfg::KT, DO = Cfg::DO, ST = Cfg::ST;
  extern __shared__ uint8_t smem_raw_ta[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_ta) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;                       // [ST][NA][64 keys][128 B]
  uint8_t* sV = sK + ST * Cfg::K_BYTES;                  // [ST][NA][64 keys][128 B]
  uint8_t* sP = sV + ST * Cfg::V_BYTES;                  // [2][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * Cfg::P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                          // [ST]
  uint64_t* kv_empty = kv_full + ST;                     // [ST]
  uint64_t* s_full = kv_empty + ST;                      // [2]
  uint64_t* s_free = s_full + 2;                         // [2]
  uint64_t* p_full = s_free + 2;                         // [2]
  uint64_t* pv_done = p_full + 2;                        // [2]
  uint64_t* o_full = pv_done + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TA_BM, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (p.Lk + TA_BN - 1) / TA_BN;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&pv_done[i], 1);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  } else if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // global memory is touched only after the predecessor kernel has completed
  AF_PDL_TRIGGER_EARLY();

  if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {   // elect.sync: ptxas then knows one thread is active -> plain R2UR, no waterfall loops
      mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int a = 0; a < NA; ++a) tma_load_4d(sQ + a * (TA_BM * 128), &tmQ, q_full, a * 64, h, m0, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % ST;
        mbar_wait(&kv_empty[s], ((j / ST) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], Cfg::K_BYTES + Cfg::V_BYTES);
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          tma_load_4d(sK + s * Cfg::K_BYTES + a * Cfg::KV_ATOM, &tmK, &kv_full[s], a * 64, h, j * TA_BN, b);
          tma_load_4d(sV + s * Cfg::V_BYTES + a * Cfg::KV_ATOM, &tmV, &kv_full[s], a * 64, h, j * TA_BN, b);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {   // elect.sync: ptxas then knows one thread is active -> plain R2UR, no waterfall loops
      constexpr uint32_t idesc_qk = make_idesc_bf16_f32(TA_BM, TA_BN, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16_f32(TA_BM, DO, true);
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
      auto issue_qk = [&](int j) {
        const int s = j % ST;
        mbar_wait(&kv_full[s], (j / ST) & 1);
        if (j >= 2) mbar_wait(&s_free[j & 1], ((j >> 1) - 1) & 1);    // softmax drained the previous user of S[j&1]
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
          const uint64_t da = make_smem_desc_sw128(aQ + (kk >> 2) * (TA_BM * 128) + (kk & 3) * 32);
          const uint64_t db = make_smem_desc_sw128(aK + s * Cfg::K_BYTES + (kk >> 2) * Cfg::KV_ATOM + (kk & 3) * 32);
          umma_bf16(tmem_base + (uint32_t)((j & 1) * TA_BN), da, db, idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[j & 1]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_qk(j + 1);
        mbar_wait(&p_full[j & 1], (j >> 1) & 1);                      // P_j in smem, O rescaled if it had to be
        tc_fence_after();
        const int s = j % ST;
#pragma unroll
        for (int k = 0; k < TA_BN / 16; ++k) {
          // V tile as the MN-major B operand: 16 keys = two 8-row groups = 2048 B per K step; atoms along d are
          // KV_ATOM bytes apart (LBO).
          const uint64_t db = make_smem_desc_sw128_mn(aV + s * Cfg::V_BYTES + k * 2048, Cfg::KV_ATOM);
          if constexpr (PT) {
            // A = P_j from TMEM: 16 keys = 8 packed columns per K step
            umma_bf16_ts(tmem_base + (uint32_t)Cfg::TMEM_O, tmem_base + (uint32_t)(Cfg::TMEM_P + (j & 1) * (TA_BN / 2) + k * 8),
                         db, idesc_pv, (j | k) != 0 ? 1u : 0u);
          } else {
            const uint64_t da = make_smem_desc_sw128(aP + (j & 1) * Cfg::P_BYTES + k * 32);
            umma_bf16(tmem_base + (uint32_t)Cfg::TMEM_O, da, db, idesc_pv, (j | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&kv_empty[s]);       // K/V stage reusable
        umma_commit(&pv_done[j & 1]);    // O updated, P buffer reusable
      }
      umma_commit(o_full);
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (warps 0..3)
    const int qd = warp & 3;                       // TMEM lane quarter of this warp
    const int row = qd * 32 + lane;                // query row inside the tile
    const uint32_t t_lane = tmem_base + ((uint32_t)(qd * 32) << 16);
    float m_ref = -INFINITY, l_run = 0.f;
    // (A register prefetch of the next tile's scores was measured 1.5x SLOWER: tcgen05.ld is a ~12-cycle operation,
    // so there is no latency to hide and the second 64-register buffer only costs spills.)
    auto tile = [&](const int j, uint32_t (&v)[TA_BN]) {
      const int bsel = j & 1;
      mbar_wait(&s_full[bsel], (j >> 1) & 1);
      tc_fence_after();
      tmem_ld_32x32b_x64_wait(t_lane + (uint32_t)(bsel * TA_BN), v);   // raw q.k dot products of this row
      tc_fence_before();
      mbar_arrive(&s_free[bsel]);                  // S[bsel] may be overwritten by Q K_{j+2}^T
      const int valid = p.Lk - j * TA_BN;          // keys of this tile that exist
      if (valid < TA_BN) {
#pragma unroll
        for (int i = 0; i < TA_BN; ++i)
          if (i >= valid) v[i] = 0xff800000u;      // -inf
      }
      float mx4[4] = {__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])};
#pragma unroll
      for (int i = 4; i < TA_BN; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) mx4[u] = fmaxf(mx4[u], __uint_as_float(v[i + u]));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;   // scale > 0

      if (j == 0) {
        m_ref = (mx == -INFINITY) ? 0.f : mx;
      } else {
        const bool need = mx > m_ref + 8.f;                     // lazy rescale: tolerate P up to 2^8
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1); // rare path: O must be quiescent (P V_{j-1} retired)
          tc_fence_after();
          const float m_new = need ? mx : m_ref;
          const float f = fast_exp2(m_ref - m_new);
          m_ref = m_new;
          l_run *= f;
#pragma unroll
          for (int c = 0; c < DO / 16; ++c) {
            uint32_t ov[16];
            tmem_ld_32x32b_x16(t_lane + (uint32_t)(Cfg::TMEM_O + c * 16), ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
            tmem_st_32x32b_x16(t_lane + (uint32_t)(Cfg::TMEM_O + c * 16), ov);
          }
          tmem_st_wait();
        }
      }
      // P buffer `bsel` was last read by P V_{j-2}: that MMA retired long ago, this wait is (almost) free.  The
      // softmax warps never wait for P V_{j-1} on the common path.
      if (j >= 2) mbar_wait(&pv_done[bsel], ((j - 2) >> 1) & 1);
      // P = exp2(s * scale_log2 - m_ref) (packed fp32x2 FMA), bf16, into the 128B-swizzled K-major tile: the
      // 16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) * 16).
      // fp32 -> bf16 by TRUNCATION (one PRMT per pair on the ALU pipe instead of F2FP on the 16-lane XU pipe that
      // the exponentials already saturate); the row sum is taken over the truncated values, so numerator and
      // denominator stay consistent and the truncation bias cancels in O = sum(p v) / sum(p).
      uint8_t* prow = sP + bsel * Cfg::P_BYTES + row * 128;
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_ref, -m_ref);
      float2 ls[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      uint32_t pk[TA_BN / 2];
#pragma unroll
      for (int c = 0; c < TA_BN / 8; ++c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[c * 8 + 2 * i]), __uint_as_float(v[c * 8 + 2 * i + 1])), sc2, nm2);
          const float2 e = (((c * 4 + i) & 7) < EMU) ? exp2_emu2(t) : make_float2(fast_exp2(t.x), fast_exp2(t.y));
          const uint32_t ex = __float_as_uint(e.x) & 0xffff0000u, ey = __float_as_uint(e.y) & 0xffff0000u;
          ls[i & 1] = __fadd2_rn(ls[i & 1], make_float2(__uint_as_float(ex), __uint_as_float(ey)));
          pk[c * 4 + i] = __byte_perm(ex, ey, 0x7632);          // {lo16 = hi half of e.x, hi16 = hi half of e.y}
        }
        if constexpr (!PT)
          *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) * 16)) = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
      }
      if constexpr (PT) {
        tmem_st_32x32b_x32(t_lane + (uint32_t)(Cfg::TMEM_P + bsel * (TA_BN / 2)), pk);
        tmem_st_wait();
      }
      const float lsum = (ls[0].x + ls[0].y) + (ls[1].x + ls[1].y);
      l_run += lsum;
      if constexpr (!PT) fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core
      tc_fence_before();
      mbar_arrive(&p_full[bsel]);
    };
    {
      uint32_t va[TA_BN];
      for (int j = 0; j < n_tiles; ++j) tile(j, va);
    }
    // ---- epilogue: O / l -> bf16 -> [B, Lq, H*d]
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    const int grow = m0 + row;
    if (p.lse && grow < p.Lq)
      p.lse[((long long)b * gridDim.y + h) * p.Lq + grow] = l_run > 0.f ? m_ref + log2f(l_run) : INFINITY;
    bf16* orow = p.o + (long long)b * p.o_sb + (long long)grow * p.o_sn + h * D;
#pragma unroll
    for (int c = 0; c < DO / 16; ++c) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(t_lane + (uint32_t)(Cfg::TMEM_O + c * 16), v);
      tmem_ld_wait();
      if (grow < p.Lq) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (c * 16 + half * 8 < D) {             // D is a multiple of 8: whole 16-byte chunks
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
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// =============================================================================================
// "Many small CTAs" variant.  The ncu profile of the kernel above shows no saturated unit (XU 56%, tensor 40%,
// issue 50%): with two softmax warps per SM sub-partition the row-per-thread softmax is latency-bound.  This
// variant trades the intra-CTA double buffering for residency: S and P share ONE 64-column TMEM buffer (P_j is
// written over the scores it was computed from), O takes 48 more, so a CTA needs 128 TMEM columns and 49 KB of
// shared memory and FOUR CTAs (16 softmax warps) fit on an SM; the tensor-pipe round trip of one CTA is hidden by
// the other three.  The scores are read in two 32-column passes (max, then exp) to stay within 80 registers.
template <int D, int EMU>
__global__ void __launch_bounds__(TA_THREADS, (D <= 64) ? 4 : 2)
attn_fwd_tcgen05_mc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                           const __grid_constant__ CUtensorMap tmV, const TaParams p) {
  using Cfg = TaCfg<D>;
  constexpr int NA = Cfg::NA, KT = Cfg::KT, DO = Cfg::DO, ST = 2;
  constexpr int TMEM_O = TA_BN;                                   // S/P at column 0, O behind it
  constexpr int TMEM_COLS = (TMEM_O + DO <= 128) ? 128 : 256;
  // DB (d = 80: 64 + 96 + 64 = 224 of the 256 allocated columns): a SECOND score buffer behind O.  Q.K^T(j+1) is issued before the issuer
  // waits for P(j), so the next scores are ready when the softmax warps finish tile j -- the serial chain Q.K^T -> softmax -> P.V of a
  // CTA (2050 clk per 64 keys at d = 80 against ~575 clk of tensor work and ~1100 of softmax) becomes softmax-bound.  K and V tiles then
  // have separate rings: a K stage is free after its Q.K^T, i.e. a whole softmax earlier than the V stage.
  constexpr bool DB = TMEM_O + DO + TA_BN <= TMEM_COLS && TMEM_COLS == 256;
  constexpr int TMEM_S1 = TMEM_O + DO;
  extern __shared__ uint8_t smem_raw_mc[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_mc) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + ST * Cfg::K_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ST * Cfg::V_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                                   // DB: K tiles only
  uint64_t* kv_empty = kv_full + ST;
  uint64_t* s_full = kv_empty + ST;                               // [2] (DB: one per score buffer)
  uint64_t* p_full = s_full + 2;                                  // [2]
  uint64_t* o_full = p_full + 2;
  uint64_t* v_full = o_full + 1;                                  // [ST] DB only
  uint64_t* v_empty = v_full + ST;                                // [ST] DB only
  uint64_t* pv_done = v_empty + ST;                               // DB only: P.V(j) has retired (O may be rescaled)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TA_BM, h = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (p.Lk + TA_BN - 1) / TA_BN;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
    }
    mbar_init(o_full, 1);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  } else if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // global memory is touched only after the predecessor kernel has completed
  AF_PDL_TRIGGER_EARLY();

  if (warp == kTmaWarp) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
      const int c0 = p.wide ? h * D : 0, hh = p.wide ? 0 : h;      // wide maps: whole rows, the head's box starts at column h*d
#pragma unroll
      for (int a = 0; a < NA; ++a) tma_load_4d(sQ + a * (TA_BM * 128), &tmQ, q_full, c0 + a * 64, hh, m0, b);
      if constexpr (DB) {
        auto load_k = [&](int j) {
          const int s = j % ST;
          mbar_wait(&kv_empty[s], ((j / ST) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[s], Cfg::K_BYTES);
#pragma unroll
          for (int a = 0; a < NA; ++a) tma_load_4d(sK + s * Cfg::K_BYTES + a * Cfg::KV_ATOM, &tmK, &kv_full[s], c0 + a * 64, hh, j * TA_BN, b);
        };
        auto load_v = [&](int j) {
          const int s = j % ST;
          mbar_wait(&v_empty[s], ((j / ST) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[s], Cfg::V_BYTES);
#pragma unroll
          for (int a = 0; a < NA; ++a) tma_load_4d(sV + s * Cfg::V_BYTES + a * Cfg::KV_ATOM, &tmV, &v_full[s], c0 + a * 64, hh, j * TA_BN, b);
        };
        // K runs one tile ahead of V: K(j + 1) only waits for Q.K^T(j - 1), V(j) for P.V(j - 2)
        load_k(0);
        for (int j = 0; j < n_tiles; ++j) {
          if (j + 1 < n_tiles) load_k(j + 1);
          load_v(j);
        }
      } else {
        for (int j = 0; j < n_tiles; ++j) {
          const int s = j % ST;
          mbar_wait(&kv_empty[s], ((j / ST) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[s], Cfg::K_BYTES + Cfg::V_BYTES);
#pragma unroll
          for (int a = 0; a < NA; ++a) {
            tma_load_4d(sK + s * Cfg::K_BYTES + a * Cfg::KV_ATOM, &tmK, &kv_full[s], c0 + a * 64, hh, j * TA_BN, b);
            tma_load_4d(sV + s * Cfg::V_BYTES + a * Cfg::KV_ATOM, &tmV, &kv_full[s], c0 + a * 64, hh, j * TA_BN, b);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc_bf16_f32(TA_BM, TA_BN, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16_f32(TA_BM, DO, true);
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
      mbar_wait(q_full, 0);
      if constexpr (DB) {
        auto issue_qk = [&](int j) {      // S_j -> score buffer j & 1 (it overwrites P_{j-2}, whose P.V was issued earlier: in order)
          const int s = j % ST;
          mbar_wait(&kv_full[s], (j / ST) & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < KT; ++kk) {
            const uint64_t da = make_smem_desc_sw128(aQ + (kk >> 2) * (TA_BM * 128) + (kk & 3) * 32);
            const uint64_t db = make_smem_desc_sw128(aK + s * Cfg::K_BYTES + (kk >> 2) * Cfg::KV_ATOM + (kk & 3) * 32);
            umma_bf16(tmem_base + (uint32_t)((j & 1) ? TMEM_S1 : 0), da, db, idesc_qk, kk > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[j & 1]);
          umma_commit(&kv_empty[s]);      // the K stage is free once S_j exists
        };
        issue_qk(0);
        for (int j = 0; j < n_tiles; ++j) {
          const int s = j % ST;
          if (j + 1 < n_tiles) issue_qk(j + 1);
          mbar_wait(&v_full[s], (j / ST) & 1);
          mbar_wait(&p_full[j & 1], (j >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < TA_BN / 16; ++k) {
            const uint64_t db = make_smem_desc_sw128_mn(aV + s * Cfg::V_BYTES + k * 2048, Cfg::KV_ATOM);
            umma_bf16_ts(tmem_base + (uint32_t)TMEM_O, tmem_base + (uint32_t)(((j & 1) ? TMEM_S1 : 0) + k * 8), db, idesc_pv, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&v_empty[s]);
          umma_commit(pv_done);
        }
      } else {
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % ST;
        mbar_wait(&kv_full[s], (j / ST) & 1);
        tc_fence_after();
        // S_j = Q K_j^T.  It overwrites P_{j-1}; tcgen05.mma executes in issue order, so P V_{j-1} has read it.
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
          const uint64_t da = make_smem_desc_sw128(aQ + (kk >> 2) * (TA_BM * 128) + (kk & 3) * 32);
          const uint64_t db = make_smem_desc_sw128(aK + s * Cfg::K_BYTES + (kk >> 2) * Cfg::KV_ATOM + (kk & 3) * 32);
          umma_bf16(tmem_base, da, db, idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[0]);
        mbar_wait(&p_full[0], j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < TA_BN / 16; ++k) {
          const uint64_t db = make_smem_desc_sw128_mn(aV + s * Cfg::V_BYTES + k * 2048, Cfg::KV_ATOM);
          umma_bf16_ts(tmem_base + (uint32_t)TMEM_O, tmem_base + (uint32_t)(k * 8), db, idesc_pv, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[s]);
      }
      }
      umma_commit(o_full);
    }
  } else {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qd * 32) << 16);
    float m_ref = -INFINITY, l_run = 0.f;
    // key mask (dalc:254-273 img_mask; Lk % 64 == 0): the 32 mask bytes of a half tile are the same for every row (broadcast loads);
    // a masked key's score becomes -inf before both passes.  A tile of only masked keys leaves P = exp2(-inf - m_ref) = 0 once a kept key
    // has set the reference; before that (m_ref = 0 from an all-masked first tile) it contributes exactly 0 as well.
    auto mask_half = [&](uint32_t (&v)[32], int j, int hf) {
      const uint4* mp = reinterpret_cast<const uint4*>(p.key_mask + (long long)b * p.Lk + j * TA_BN + hf * 32);
      const uint4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
      const uint32_t mw[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (((mw[i >> 2] >> ((i & 3) * 8)) & 0xffu) == 0) v[i] = 0xff800000u;
    };
    for (int j = 0; j < n_tiles; ++j) {
      // single buffer: s_full also implies P V_{j-1} has retired (commit covers all earlier MMAs); DB: see pv_done below
      const uint32_t t_s = t_lane + (uint32_t)((DB && (j & 1)) ? TMEM_S1 : 0);
      mbar_wait(&s_full[DB ? (j & 1) : 0], DB ? ((j >> 1) & 1) : (j & 1));
      tc_fence_after();
      const int valid = p.Lk - j * TA_BN;
      // ---- pass 1: row max of the raw scores
      float mxr = -INFINITY;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t v[32];
        tmem_ld_32x32b_x32_wait(t_s + (uint32_t)(hf * 32), v);
        if (valid < TA_BN) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (hf * 32 + i >= valid) v[i] = 0xff800000u;
        }
        if (p.key_mask) mask_half(v, j, hf);
        float m4[4] = {__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])};
#pragma unroll
        for (int i = 4; i < 32; i += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) m4[u] = fmaxf(m4[u], __uint_as_float(v[i + u]));
        }
        mxr = fmaxf(mxr, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
      }
      const float mx = mxr * p.scale_log2;
      if (j == 0) {
        m_ref = (mx == -INFINITY) ? 0.f : mx;
      } else {
        const bool need = mx > m_ref + 8.f;
        if (__any_sync(0xffffffffu, need)) {
          if constexpr (DB) {             // Q.K^T(j) was issued BEFORE P.V(j-1): the accumulator is only ours once that MMA has retired
            mbar_wait(pv_done, (j - 1) & 1);
            tc_fence_after();
          }
          const float m_new = need ? mx : m_ref;
          const float f = fast_exp2(m_ref - m_new);
          m_ref = m_new;
          l_run *= f;
#pragma unroll
          for (int c = 0; c < DO / 16; ++c) {
            uint32_t ov[16];
            tmem_ld_32x32b_x16(t_lane + (uint32_t)(TMEM_O + c * 16), ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
            tmem_st_32x32b_x16(t_lane + (uint32_t)(TMEM_O + c * 16), ov);
          }
        }
      }
      // ---- pass 2: P = exp2(s * scale - m_ref), truncated to bf16, written over the scores it came from
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_ref, -m_ref);
      float2 ls[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t v[32];
        tmem_ld_32x32b_x32_wait(t_s + (uint32_t)(hf * 32), v);
        if (valid < TA_BN) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (hf * 32 + i >= valid) v[i] = 0xff800000u;
        }
        if (p.key_mask) mask_half(v, j, hf);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
          const float2 e = ((i & 7) < EMU) ? exp2_emu2(t) : make_float2(fast_exp2(t.x), fast_exp2(t.y));
          const uint32_t ex = __float_as_uint(e.x) & 0xffff0000u, ey = __float_as_uint(e.y) & 0xffff0000u;
          ls[i & 1] = __fadd2_rn(ls[i & 1], make_float2(__uint_as_float(ex), __uint_as_float(ey)));
          pk[i] = __byte_perm(ex, ey, 0x7632);
        }
        tmem_st_32x32b_x16(t_s + (uint32_t)(hf * 16), pk);      // columns [16 hf, 16 hf + 16): already consumed
      }
      tmem_st_wait();
      l_run += (ls[0].x + ls[0].y) + (ls[1].x + ls[1].y);
      tc_fence_before();
      mbar_arrive(&p_full[DB ? (j & 1) : 0]);
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    const int grow = m0 + row;
    if (p.lse && grow < p.Lq)
      p.lse[((long long)b * gridDim.y + h) * p.Lq + grow] = l_run > 0.f ? m_ref + log2f(l_run) : INFINITY;
    bf16* orow = p.o + (long long)b * p.o_sb + (long long)grow * p.o_sn + h * D;
#pragma unroll
    for (int c = 0; c < DO / 16; ++c) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(t_lane + (uint32_t)(TMEM_O + c * 16), v);
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
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =============================================================================================
// PERSISTENT form of the small-CTA kernel for d = 80 (level B).  With one 128-query tile per CTA the 512 tiles of level B (B = 8) are 1.73
// waves of 296 CTA slots, and every CTA pays its own set-up (TMEM allocation, barrier init, descriptor fetch, first loads: several
// microseconds against ~11 us of work for 16 key tiles).  Here 2 x #SM CTAs stay resident and walk the (batch, head, query tile) units with a
// stride of the grid; barrier phases, the K / V rings and the two score buffers simply keep counting across units.  Per unit the pipeline
// is the double-buffered one of attn_fwd_tcgen05_mc_kernel (DB): Q.K^T(j+1) ahead of P.V(j), separate K / V rings, pv_done before a rescale.
template <int D, int EMU>
__global__ void __launch_bounds__(TA_THREADS, 2)
attn_fwd_tcgen05_mcp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                            const __grid_constant__ CUtensorMap tmV, const TaParams p, const int H, const int n_units) {
  using Cfg = TaCfg<D>;
  constexpr int NA = Cfg::NA, KT = Cfg::KT, DO = Cfg::DO, ST = 2;
  constexpr int TMEM_O = TA_BN, TMEM_S1 = TMEM_O + DO, TMEM_COLS = 256;
  static_assert(TMEM_S1 + TA_BN <= TMEM_COLS, "two score buffers and the accumulator must fit 256 TMEM columns");
  extern __shared__ uint8_t smem_raw_mcp[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_mcp) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + ST * Cfg::K_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ST * Cfg::V_BYTES);
  uint64_t* q_full = bars;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;                                    // [ST]
  uint64_t* k_empty = k_full + ST;                                // [ST]
  uint64_t* v_full = k_empty + ST;                                // [ST]
  uint64_t* v_empty = v_full + ST;                                // [ST]
  uint64_t* s_full = v_empty + ST;                                // [2]
  uint64_t* p_full = s_full + 2;                                  // [2]
  uint64_t* pv_done = p_full + 2;
  uint64_t* o_full = pv_done + 1;
  uint64_t* o_free = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = (p.Lq + TA_BM - 1) / TA_BM;
  const int n_tiles = (p.Lk + TA_BN - 1) / TA_BN;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
    }
    mbar_init(pv_done, 1);
    mbar_init(o_full, 1);
    mbar_init(o_free, 128);
    fence_barrier_init();
  } else if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // global memory is touched only after the predecessor kernel has completed
  AF_PDL_TRIGGER_EARLY();

  if (warp == kTmaWarp) {
    if (elect_one()) {
      int tk = 0, tv = 0;      // K / V tiles loaded so far (ring positions keep counting across units)
      for (int u = blockIdx.x, it = 0; u < n_units; u += gridDim.x, ++it) {
        const int bh = u / n_qt, qt = u - bh * n_qt;
        const int b = bh / H, h = bh - b * H;
        const int c0 = p.wide ? h * D : 0, hh = p.wide ? 0 : h;
        if (it > 0) mbar_wait(q_empty, (it - 1) & 1);            // every Q.K^T of the previous unit has read the Q tile
        mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
        for (int a = 0; a < NA; ++a) tma_load_4d(sQ + a * (TA_BM * 128), &tmQ, q_full, c0 + a * 64, hh, qt * TA_BM, b);
        auto load_k = [&](int j) {
          const int s = tk % ST;
          mbar_wait(&k_empty[s], ((tk / ST) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[s], Cfg::K_BYTES);
#pragma unroll
          for (int a = 0; a < NA; ++a) tma_load_4d(sK + s * Cfg::K_BYTES + a * Cfg::KV_ATOM, &tmK, &k_full[s], c0 + a * 64, hh, j * TA_BN, b);
          ++tk;
        };
        auto load_v = [&](int j) {
          const int s = tv % ST;
          mbar_wait(&v_empty[s], ((tv / ST) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[s], Cfg::V_BYTES);
#pragma unroll
          for (int a = 0; a < NA; ++a) tma_load_4d(sV + s * Cfg::V_BYTES + a * Cfg::KV_ATOM, &tmV, &v_full[s], c0 + a * 64, hh, j * TA_BN, b);
          ++tv;
        };
        load_k(0);
        for (int j = 0; j < n_tiles; ++j) {
          if (j + 1 < n_tiles) load_k(j + 1);
          load_v(j);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc_bf16_f32(TA_BM, TA_BN, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16_f32(TA_BM, DO, true);
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
      int tk = 0, tv = 0, g = 0;      // K tiles / V tiles consumed, score tiles issued (global counters)
      for (int u = blockIdx.x, it = 0; u < n_units; u += gridDim.x, ++it) {
        const int g0 = it * n_tiles;
        mbar_wait(q_full, it & 1);
        auto issue_qk = [&](int j) {      // S_j -> score buffer (g0 + j) & 1
          const int s = tk % ST, buf = (g0 + j) & 1;
          mbar_wait(&k_full[s], (tk / ST) & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < KT; ++kk) {
            const uint64_t da = make_smem_desc_sw128(aQ + (kk >> 2) * (TA_BM * 128) + (kk & 3) * 32);
            const uint64_t db = make_smem_desc_sw128(aK + s * Cfg::K_BYTES + (kk >> 2) * Cfg::KV_ATOM + (kk & 3) * 32);
            umma_bf16(tmem_base + (uint32_t)(buf ? TMEM_S1 : 0), da, db, idesc_qk, kk > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[buf]);
          umma_commit(&k_empty[s]);
          ++tk;
          if (j + 1 == n_tiles) umma_commit(q_empty);      // last Q.K^T of the unit: the Q tile may be replaced
        };
        issue_qk(0);
        for (int j = 0; j < n_tiles; ++j) {
          const int s = tv % ST, buf = (g0 + j) & 1;
          if (j + 1 < n_tiles) issue_qk(j + 1);
          mbar_wait(&v_full[s], (tv / ST) & 1);
          mbar_wait(&p_full[buf], ((g0 + j) >> 1) & 1);
          if (j == 0 && it > 0) mbar_wait(o_free, (it - 1) & 1);      // the previous unit's epilogue has read the accumulator
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < TA_BN / 16; ++k) {
            const uint64_t db = make_smem_desc_sw128_mn(aV + s * Cfg::V_BYTES + k * 2048, Cfg::KV_ATOM);
            umma_bf16_ts(tmem_base + (uint32_t)TMEM_O, tmem_base + (uint32_t)((buf ? TMEM_S1 : 0) + k * 8), db, idesc_pv, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&v_empty[s]);
          umma_commit(pv_done);
          ++tv;
        }
        umma_commit(o_full);
        (void)g;
      }
    }
  } else {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qd * 32) << 16);
    for (int u = blockIdx.x, it = 0; u < n_units; u += gridDim.x, ++it) {
      const int bh = u / n_qt, qt = u - bh * n_qt;
      const int b = bh / H, h = bh - b * H;
      const int m0 = qt * TA_BM, g0 = it * n_tiles;
      float m_ref = -INFINITY, l_run = 0.f;
      auto mask_half = [&](uint32_t (&v)[32], int j, int hf) {
        const uint4* mp = reinterpret_cast<const uint4*>(p.key_mask + (long long)b * p.Lk + j * TA_BN + hf * 32);
        const uint4 m0_ = __ldg(mp), m1_ = __ldg(mp + 1);
        const uint32_t mw[8] = {m0_.x, m0_.y, m0_.z, m0_.w, m1_.x, m1_.y, m1_.z, m1_.w};
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (((mw[i >> 2] >> ((i & 3) * 8)) & 0xffu) == 0) v[i] = 0xff800000u;
      };
      for (int j = 0; j < n_tiles; ++j) {
        const int G = g0 + j;
        const uint32_t t_s = t_lane + (uint32_t)((G & 1) ? TMEM_S1 : 0);
        mbar_wait(&s_full[G & 1], (G >> 1) & 1);
        tc_fence_after();
        const int valid = p.Lk - j * TA_BN;
        // the 64 scores of the row are read from TMEM ONCE and stay in registers for the max and the exp2 pass (two CTAs per SM leave 170
        // registers per thread; the four 32-column round trips of the two-pass form were ~700 clk of pure latency per key tile)
        uint32_t v[TA_BN];
        tmem_ld_32x32b_x64_wait(t_s, v);
        if (valid < TA_BN) {
#pragma unroll
          for (int i = 0; i < TA_BN; ++i)
            if (i >= valid) v[i] = 0xff800000u;
        }
        if (p.key_mask) {
          uint32_t (&va)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[0]);
          uint32_t (&vb)[32] = *reinterpret_cast<uint32_t (*)[32]>(&v[32]);
          mask_half(va, j, 0);
          mask_half(vb, j, 1);
        }
        float m4[4] = {__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])};
#pragma unroll
        for (int i = 4; i < TA_BN; i += 4) {
#pragma unroll
          for (int x = 0; x < 4; ++x) m4[x] = fmaxf(m4[x], __uint_as_float(v[i + x]));
        }
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
        if (j == 0) {
          m_ref = (mx == -INFINITY) ? 0.f : mx;
        } else {
          const bool need = mx > m_ref + 8.f;
          if (__any_sync(0xffffffffu, need)) {
            mbar_wait(pv_done, (G - 1) & 1);      // Q.K^T(j) was issued before P.V(j-1): the accumulator is ours once that MMA has retired
            tc_fence_after();
            const float m_new = need ? mx : m_ref;
            const float f = fast_exp2(m_ref - m_new);
            m_ref = m_new;
            l_run *= f;
#pragma unroll
            for (int c = 0; c < DO / 16; ++c) {
              uint32_t ov[16];
              tmem_ld_32x32b_x16(t_lane + (uint32_t)(TMEM_O + c * 16), ov);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
              tmem_st_32x32b_x16(t_lane + (uint32_t)(TMEM_O + c * 16), ov);
            }
          }
        }
        // P = exp2(s * scale - m_ref), truncated to bf16, written over the scores it came from
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_ref, -m_ref);
        float2 ls[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[hf * 32 + 2 * i]), __uint_as_float(v[hf * 32 + 2 * i + 1])), sc2, nm2);
            const float2 e = ((i & 7) < EMU) ? exp2_emu2(t) : make_float2(fast_exp2(t.x), fast_exp2(t.y));
            const uint32_t ex = __float_as_uint(e.x) & 0xffff0000u, ey = __float_as_uint(e.y) & 0xffff0000u;
            ls[i & 1] = __fadd2_rn(ls[i & 1], make_float2(__uint_as_float(ex), __uint_as_float(ey)));
            pk[i] = __byte_perm(ex, ey, 0x7632);
          }
          tmem_st_32x32b_x16(t_s + (uint32_t)(hf * 16), pk);
        }
        tmem_st_wait();
        l_run += (ls[0].x + ls[0].y) + (ls[1].x + ls[1].y);
        tc_fence_before();
        mbar_arrive(&p_full[G & 1]);
      }
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      const int grow = m0 + row;
      if (p.lse && grow < p.Lq)
        p.lse[((long long)b * H + h) * p.Lq + grow] = l_run > 0.f ? m_ref + log2f(l_run) : INFINITY;
      bf16* orow = p.o + (long long)b * p.o_sb + (long long)grow * p.o_sn + h * D;
#pragma unroll
      for (int c = 0; c < DO / 16; ++c) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_lane + (uint32_t)(TMEM_O + c * 16), v);
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
      tc_fence_before();
      mbar_arrive(o_free);
    }
  }
  AF_PDL_TRIGGER_LATE();
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =============================================================================================
// "Quad" variant for d = 40 (level A, half of the whole attention stack's time).  ncu on the many-small-CTAs kernel
// shows XU (exp2) 63 % + tensor phases that do not overlap: the four co-resident CTAs phase-lock -- they share the XU
// pipe, so they finish their softmax together, then queue their P.V / Q.K MMAs on the tensor pipe together while the
// XU idles (measured 830 clk per CTA tile against 512 clk of exp2 work).  This kernel puts the four 128-query tiles
// into ONE CTA (512 consecutive queries of one (batch, head), 16 softmax warps + TMA + MMA warp, all 512 TMEM columns:
// S/P 64 + O 48 per tile) so that ONE thread issues every tcgen05.mma in a fixed round-robin:
//     P.V(g, j), Q.K(g, j+1)  for g = 0..3,  each gated by that tile's own p_full barrier.
// A tile's next scores are produced right after its P.V while the other three tiles are still in their softmax, so
// the tensor work of one tile always runs under the exp2 work of the others, and every K/V tile is fetched once per
// 512 queries instead of once per 128 (4x less TMA / L2 traffic).
// (Tried on top of this and dropped: THREE tiles per CTA with TWO threads per query row -- 24 softmax warps, scores read
// from TMEM once, half-row maxima exchanged through shared memory + 64-thread named barriers: correct, but 418 us vs
// 372 us; the extra hand-off per tile costs more than the shorter per-warp chain saves.)
// Wave quantisation: a CTA now lasts 4x longer, so 512 units on 148 SMs would leave 80 SMs idle for a whole CTA
// lifetime in the 4th round (measured: 387 us, no better than the small-CTA kernel).  The launcher therefore runs
// floor(units / 148) * 148 four-tile CTAs and covers the remainder with TWO-tile CTAs (G = 2, 256 TMEM columns)
// in a second launch, which finish in less than half the time of a four-tile CTA.
constexpr int TQ_G = 4;                       // query tiles per CTA (bulk launch)
constexpr int TQ_ST = 3;                      // K/V ring: tiles j and j+1 are live at once, j+2 in flight

// EMU = how many of every 8 exp2 pairs go to the FMA-pipe polynomial (exp2_emu2) instead of MUFU.EX2.
// SPLIT (wave tail, G = 4): the four TMEM slots hold TWO query tiles x TWO halves of the keys (slot g = tile g / 2, key
// half g % 2); each slot runs the same online softmax over its half and the two partial results of a tile are merged
// in the epilogue through shared memory (O = 2^(m0-m) O0 + 2^(m1-m) O1 -- the row sum rides along as column D).  A
// tail CTA then has the full 16 softmax warps for half as long, instead of 8 warps for the whole key range.
// MASK (key mask, dalc:254-273 img_mask): the head dim is zero-padded from D to the 64-wide MMA K, so column D of Q and K is free:
// the TMA warp sets Q[:, D] = 1 once and K[j, D] = 0 / -65536 (kept / masked key) in every K tile it lands -- next to the ones
// column it already writes into V -- and the tensor core adds the mask bias to the scores: S = q.k + 1 * bias.  The softmax warps
// run the unmasked code (a masked key's exp2 underflows to exactly 0; a tile of only masked keys is erased by the next rescale,
// whose factor 2^(-65536 scale) is 0); the issuer waits for the patched tiles (q_ready, v_ready) instead of the raw TMA barriers.
// NI = MMA-issuing warps (1 | 2).  Traced (-DAF_ATTN_TRACE build, ADAFACE_ATTN_TRACE=1): per 64-key step (2150 clk) the single
// issuer spends ~300 clk in two K/V barrier waits, then per tile a p_full wait (>= 100 clk even when complete) and 130-260 clk for
// 4 P.V + 3 Q.K issues + commit; a softmax warp works ~1100 clk (TMEM read + max 255, exp2 + pack 840) and waits ~900 clk for its
// next scores.  Giving each ping-pong half its own issuer (NI = 2) halves the issuer's chain but does NOT shorten the step (297.0 vs
// 295.9 us): the period is one tile's own dependency chain (softmax -> P.V / Q.K -> hand-offs) with the MUFU pipe 71 % busy.
template <int D, int G, int EMU, bool SPLIT, bool MASK = false, int NI = 1>
__global__ void __launch_bounds__((4 * G + 1 + NI) * 32, 1)
attn_fwd_tcgen05_quad_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                             const __grid_constant__ CUtensorMap tmV, const TaParams p, const int unit0, const int H) {
  using Cfg = TaCfg<D>;
  constexpr int NA = Cfg::NA, KT = Cfg::KT, DO = Cfg::DO, ST = TQ_ST;
  constexpr int TMEM_G = 128, TMEM_O = TA_BN;                     // per tile: S/P at +0, O at +64
  static_assert(TMEM_O + DO <= TMEM_G && NA == 1 && D < DO, "quad kernel: head dim must fit 128 TMEM columns per tile and leave a pad column for the row sums");
  constexpr int kTma = 4 * G, kMma = 4 * G + 1;
  constexpr int NQ = SPLIT ? G / 2 : G;                           // query tiles per CTA
  constexpr int KH = SPLIT ? 2 : 1;                               // key halves (sub-tiles per ring stage)
  static_assert(!SPLIT || G == 4, "SPLIT mode is built for four TMEM slots");
  extern __shared__ uint8_t smem_raw_tq[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_tq) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                             // [NQ][128][128 B]
  uint8_t* sK = sQ + NQ * Cfg::Q_BYTES;                           // [ST][KH][64][128 B]
  uint8_t* sV = sK + ST * KH * Cfg::K_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ST * KH * Cfg::V_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                                   // [ST]
  uint64_t* kv_empty = kv_full + ST;                              // [ST]
  uint64_t* s_full = kv_empty + ST;                               // [G]
  uint64_t* p_full = s_full + G;                                  // [G]
  uint64_t* o_full = p_full + G;
  uint64_t* v_ready = o_full + 1;                                 // [ST]: the ones column has been written into V stage s
  uint64_t* q_ready = v_ready + ST;                               // MASK: column D of the Q tiles has been set to 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // unit = (batch, head, block of G query tiles), linearised with the query block fastest
  const int n_qb = (p.Lq + NQ * TA_BM - 1) / (NQ * TA_BM);
  const int unit = unit0 + blockIdx.x;
  const int qb = unit % n_qb, bh = unit / n_qb;
  const int h = bh % H, b = bh / H;
  const int m0 = qb * (NQ * TA_BM);
  const int n_tiles = ((p.Lk + TA_BN - 1) / TA_BN) / KH;         // iterations; SPLIT: slot half kh covers key tiles [kh * n_tiles, ...)

  if (warp == kTma && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ST; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], NI);
      mbar_init(&v_ready[s], 1);
    }
    for (int g = 0; g < G; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 128);
    }
    mbar_init(o_full, NI);
    mbar_init(q_ready, 1);
    fence_barrier_init();
  } else if (warp == kMma) {
    tmem_alloc(tmem_slot, G * TMEM_G);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // global memory is touched only after the predecessor kernel has completed
  AF_PDL_TRIGGER_EARLY();

  if (warp == kTma) {
    // The whole warp runs this loop: one elected lane issues the TMA loads (two tiles ahead), then all 32 lanes write
    // the ONES COLUMN into the V tile that has just landed: column D of every V row := 1.0, so that column D of the
    // P.V accumulator is the row sum of P -- the softmax warps then neither mask-truncate nor add up their probabilities.
    const bool leader = elect_one();
    auto issue_kv = [&](int j) {
      const int s = j % ST;
      mbar_wait(&kv_empty[s], ((j / ST) & 1) ^ 1);
      mbar_arrive_expect_tx(&kv_full[s], KH * (Cfg::K_BYTES + Cfg::V_BYTES));
#pragma unroll
      for (int kh = 0; kh < KH; ++kh) {
        tma_load_4d(sK + (s * KH + kh) * Cfg::K_BYTES, &tmK, &kv_full[s], 0, h, (j + kh * n_tiles) * TA_BN, b);
        tma_load_4d(sV + (s * KH + kh) * Cfg::V_BYTES, &tmV, &kv_full[s], 0, h, (j + kh * n_tiles) * TA_BN, b);
      }
    };
    if (leader) {
      mbar_arrive_expect_tx(q_full, NQ * Cfg::Q_BYTES);
#pragma unroll
      for (int g = 0; g < NQ; ++g) tma_load_4d(sQ + g * Cfg::Q_BYTES, &tmQ, q_full, 0, h, m0 + g * TA_BM, b);
      issue_kv(0);
      if (n_tiles > 1) issue_kv(1);
    }
    // MASK: bias of this lane's K rows (r = lane, lane + 32, ...) of tile j, fetched one tile ahead
    const uint8_t* mrow = MASK ? p.key_mask + (long long)b * p.Lk : nullptr;
    auto mask_bias = [&](int j, int r) -> uint16_t {
      const int key = (j + (r / TA_BN) * n_tiles) * TA_BN + (r % TA_BN);
      return (key < p.Lk && __ldg(mrow + key) == 0) ? (uint16_t)0xC780 : (uint16_t)0;      // bf16 -65536 | 0
    };
    uint16_t kb[KH * TA_BN / 32];
    if constexpr (MASK) {
#pragma unroll
      for (int i = 0; i < KH * TA_BN / 32; ++i) kb[i] = mask_bias(0, lane + 32 * i);
      mbar_wait(q_full, 0);
      for (int r = lane; r < NQ * TA_BM; r += 32)
        *reinterpret_cast<uint16_t*>(sQ + r * 128 + ((((D >> 3) ^ (r & 7)) << 4) | ((D & 7) << 1))) = 0x3F80;
      fence_proxy_async_smem();
      __syncwarp();
      if (leader) mbar_arrive(q_ready);
    }
    auto patch = [&](int j) {
      const int s = j % ST;
      mbar_wait(&kv_full[s], (j / ST) & 1);
#pragma unroll
      for (int r = lane; r < KH * TA_BN; r += 32)  // 128B-swizzled tile: element D of row r sits in chunk (D/8) ^ (r & 7)
        *reinterpret_cast<uint16_t*>(sV + s * KH * Cfg::V_BYTES + r * 128 + ((((D >> 3) ^ (r & 7)) << 4) | ((D & 7) << 1))) = 0x3F80;
      if constexpr (MASK) {
#pragma unroll
        for (int i = 0; i < KH * TA_BN / 32; ++i) {
          const int r = lane + 32 * i;
          *reinterpret_cast<uint16_t*>(sK + s * KH * Cfg::K_BYTES + r * 128 + ((((D >> 3) ^ (r & 7)) << 4) | ((D & 7) << 1))) = kb[i];
        }
        if (j + 1 < n_tiles) {
#pragma unroll
          for (int i = 0; i < KH * TA_BN / 32; ++i) kb[i] = mask_bias(j + 1, lane + 32 * i);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (leader) mbar_arrive(&v_ready[s]);
    };
    if constexpr (MASK) {
      // the issuer needs the PATCHED K tile j + 1 while it works on tile j: patch one tile ahead of the load that may block on a
      // ring slot (tile j + 2 waits for tile j - 1 to be consumed)
      // Two independent event streams -- "ring slot free -> issue the next load" and "tile landed -> patch it" -- served by polling, whichever is
      // ready first.  (Traced: with the fixed order patch(j + 1); issue_kv(j + 2) the load of tile j + 2 was only issued after tile j + 1 had
      // LANDED, one load in flight at a time, and the issuer waited ~1800 clk per step for its next patched K tile.)
      int next_load = n_tiles > 1 ? 2 : 1, next_patch = 0;
      while (next_patch < n_tiles) {
        int did = 0;
        if (next_load < n_tiles) {
          int ok = 0;
          if (leader) {
            uint32_t t;
            asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                         : "=r"(t) : "r"(smem_u32(&kv_empty[next_load % ST])), "r"((uint32_t)(((next_load / ST) & 1) ^ 1)) : "memory");
            ok = (int)t;
            if (ok) issue_kv(next_load);
          }
          ok = __any_sync(0xffffffffu, ok);
          if (ok) { ++next_load; did = 1; }
        }
        if (next_patch < next_load) {
          int ok = 0;
          if (leader) {
            uint32_t t;
            asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                         : "=r"(t) : "r"(smem_u32(&kv_full[next_patch % ST])), "r"((uint32_t)((next_patch / ST) & 1)) : "memory");
            ok = (int)t;
          }
          ok = __any_sync(0xffffffffu, ok);
          if (ok) { patch(next_patch); ++next_patch; did = 1; }
        }
        if (!did) __nanosleep(32);
      }
    } else {
      for (int j = 0; j < n_tiles; ++j) {
        patch(j);
        if (leader && j + 2 < n_tiles) issue_kv(j + 2);
        __syncwarp();
      }
    }
  } else if (warp >= kMma) {
    if (elect_one()) {
      static_assert(NI == 1 || (NI == 2 && G % 2 == 0), "one issuer, or one per ping-pong half");
      constexpr int GPI = G / NI;                                 // tiles per issuer
      const int g_lo = (warp - kMma) * GPI;
      constexpr uint32_t idesc_qk = make_idesc_bf16_f32(TA_BM, TA_BN, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16_f32(TA_BM, DO, true);
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
      auto issue_qk = [&](int g, int j) {
        const int s = j % ST;
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
          const uint64_t da = make_smem_desc_sw128(aQ + (SPLIT ? g >> 1 : g) * Cfg::Q_BYTES + (kk & 3) * 32);
          const uint64_t db = make_smem_desc_sw128(aK + (s * KH + (SPLIT ? g & 1 : 0)) * Cfg::K_BYTES + (kk & 3) * 32);
          umma_bf16(tmem_base + (uint32_t)(g * TMEM_G), da, db, idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[g]);
      };
      // issuer-side wait: a non-blocking test first (traced: the suspending try_wait of mbar_wait needs ~130 clk even for a completed phase,
      // six of them per 64-key step were 800 of the issuer's 2150 clk)
      auto iss_wait = [&](uint64_t* bar, uint32_t parity) {
#ifndef AF_ATTN_ISSUER_SLOW_WAIT
        uint32_t ok;
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
#endif
        mbar_wait(bar, parity);
      };
      mbar_wait(MASK ? q_ready : q_full, 0);
      mbar_wait(MASK ? &v_ready[0] : &kv_full[0], 0);
      tc_fence_after();
      // PING-PONG: the tiles form two halves; the second half's first scores are only produced once the first half
      // has finished its first softmax.  From then on the blocking round-robin below keeps the halves half a period
      // apart: while one half exponentiates (XU), the other half's P.V / Q.K run on the tensor pipe.
      if constexpr (NI == 1) {
#pragma unroll
        for (int g = 0; g < G / 2; ++g) issue_qk(g, 0);
#pragma unroll
        for (int g = 0; g < G / 2; ++g) mbar_wait(&p_full[g], 0);
#pragma unroll
        for (int g = G / 2; g < G; ++g) issue_qk(g, 0);
      } else {
        if (g_lo != 0) {          // second issuer: starts when the first half's P_0 exist (their next phase is > 1000 clk away)
#pragma unroll
          for (int g = 0; g < GPI; ++g) mbar_wait(&p_full[g], 0);
        }
#pragma unroll
        for (int g = 0; g < GPI; ++g) issue_qk(g_lo + g, 0);
      }
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % ST;
        const bool more = j + 1 < n_tiles;
        AF_ATTN_TR(const bool tr = p.trace && blockIdx.x == 0 && g_lo == 0 && j >= 8 && j < 16; if (tr) p.trace[(j - 8) * 16] = clock64();)
        iss_wait(&v_ready[s], (j / ST) & 1);                      // V_j carries its ones column
        if (more) iss_wait(MASK ? &v_ready[(j + 1) % ST] : &kv_full[(j + 1) % ST], ((j + 1) / ST) & 1);
        tc_fence_after();
        AF_ATTN_TR(if (tr) p.trace[(j - 8) * 16 + 1] = clock64();)
#pragma unroll
        for (int gi = 0; gi < GPI; ++gi) {
          const int g = g_lo + gi;
          AF_ATTN_TR(if (tr) p.trace[(j - 8) * 16 + 2 + gi * 3] = clock64();)
          iss_wait(&p_full[g], j & 1);                            // tile g's P_j is in TMEM
          AF_ATTN_TR(if (tr) p.trace[(j - 8) * 16 + 3 + gi * 3] = clock64();)
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < TA_BN / 16; ++k) {
            const uint64_t db = make_smem_desc_sw128_mn(aV + (s * KH + (SPLIT ? g & 1 : 0)) * Cfg::V_BYTES + k * 2048, Cfg::KV_ATOM);
            umma_bf16_ts(tmem_base + (uint32_t)(g * TMEM_G + TMEM_O), tmem_base + (uint32_t)(g * TMEM_G + k * 8), db, idesc_pv,
                         (j | k) != 0 ? 1u : 0u);
          }
          if (more) issue_qk(g, j + 1);                           // overwrites P_j: executes after the P.V above (in order)
          AF_ATTN_TR(if (tr) p.trace[(j - 8) * 16 + 4 + gi * 3] = clock64();)
        }
        umma_commit(&kv_empty[s]);                                // this issuer's P.V_j have been issued
      }
      umma_commit(o_full);
    }
  } else {
    const int g = warp >> 2, qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(g * TMEM_G);
    float m_ref = -INFINITY;
    for (int j = 0; j < n_tiles; ++j) {
      AF_ATTN_TR(const bool str_ = p.trace && blockIdx.x == 0 && lane == 0 && qd == 0 && j >= 8 && j < 16; long long* tp = p.trace + 256 + g * 64 + (j - 8) * 8;
                 if (str_) tp[0] = clock64();)
#ifndef AF_ATTN_OPTIMISTIC
      mbar_wait(&s_full[g], j & 1);          // also implies P V_{j-1} of this tile has retired
      AF_ATTN_TR(if (str_) tp[1] = clock64();)
      tc_fence_after();
      const int valid = p.Lk - (j + (SPLIT ? (g & 1) * n_tiles : 0)) * TA_BN;
      // the 64 scores of this row are read from TMEM ONCE and stay in registers for both the max and the exp pass
      // (one CTA per SM: 112 registers per thread are available)
      uint32_t v[TA_BN];
      tmem_ld_32x32b_x64_wait(t_lane, v);
      if (valid < TA_BN) {
#pragma unroll
        for (int i = 0; i < TA_BN; ++i)
          if (i >= valid) v[i] = 0xff800000u;
      }
      float m4[4] = {__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])};
#pragma unroll
      for (int i = 4; i < TA_BN; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) m4[u] = fmaxf(m4[u], __uint_as_float(v[i + u]));
      }
      const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
      AF_ATTN_TR(if (str_) tp[2] = clock64() + (mx > 1e30f ? 1 : 0);)       // (after the row maximum: the TMEM load has landed)
      if (j == 0) {
        m_ref = (mx == -INFINITY) ? 0.f : mx;
      } else {
        const bool need = mx > m_ref + 8.f;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? mx : m_ref;
          const float f = fast_exp2(m_ref - m_new);
          m_ref = m_new;
#pragma unroll
          for (int c = 0; c < DO / 16; ++c) {
            uint32_t ov[16];
            tmem_ld_32x32b_x16(t_lane + (uint32_t)(TMEM_O + c * 16), ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
            tmem_st_32x32b_x16(t_lane + (uint32_t)(TMEM_O + c * 16), ov);
          }
        }
      }
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_ref, -m_ref);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[hf * 32 + 2 * i]), __uint_as_float(v[hf * 32 + 2 * i + 1])), sc2, nm2);
          // EMU 0..4: that many of every 8 pairs on the degree-3 polynomial; EMU 5..7: 2..4 pairs on the degree-2 one
          constexpr int EMU_N = EMU <= 4 ? EMU : EMU - 3, EMU_DEG = EMU <= 4 ? 3 : 2;
          const float2 e = ((i & 7) < EMU_N) ? exp2_emu2<EMU_DEG>(t) : make_float2(fast_exp2(t.x), fast_exp2(t.y));
          pk[i] = __byte_perm(__float_as_uint(e.x), __float_as_uint(e.y), 0x7632);   // truncate to bf16; the row sum comes from the MMA
        }
        tmem_st_32x32b_x16(t_lane + (uint32_t)(hf * 16), pk);
      }
#else
      // (-DAF_ATTN_OPTIMISTIC; measured SLOWER, 327 vs 296 us: the scores' wait shrinks, so 2.6 instead of 2 warps per scheduler sit in their exp2
      // phase at once and each takes 1550 instead of 1000 clk -- profiles/r02_attn_trace_quad_optimistic.txt.  Kept as a record.)
      // OPTIMISTIC softmax step.  The reference maximum m_ref sits 7 octaves ABOVE the largest score seen so far, so every probability is
      // <= 2^-7 until a score outgrows the old maximum by 2^8 -- exactly then some bf16 P has its top exponent bit set (P >= 2).  The
      // exponentials therefore start against the running reference as soon as the scores are in registers; instead of a row-maximum
      // pass in front of them (30 FMNMX3 on the critical path) the packed P words are OR-ed together (LOP3, beside the MUFU work) and one
      // bit test decides whether the tile has to be redone against a new reference (first tile, then rare).
      {
        uint32_t ok;      // non-blocking test first: a completed phase answers in tens of clocks, the suspending try_wait was traced at ~130
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&s_full[g])), "r"((uint32_t)(j & 1)) : "memory");
        if (!ok) mbar_wait(&s_full[g], j & 1);          // also implies P V_{j-1} of this tile has retired
      }
      AF_ATTN_TR(if (str_) tp[1] = clock64();)
      tc_fence_after();
      const int valid = p.Lk - (j + (SPLIT ? (g & 1) * n_tiles : 0)) * TA_BN;
      uint32_t v[TA_BN];
      tmem_ld_32x32b_x64_wait(t_lane, v);
      if (valid < TA_BN) {
#pragma unroll
        for (int i = 0; i < TA_BN; ++i)
          if (i >= valid) v[i] = 0xff800000u;
      }
      AF_ATTN_TR(if (str_) tp[2] = clock64();)
      constexpr int EMU_N = EMU <= 4 ? EMU : EMU - 3, EMU_DEG = EMU <= 4 ? 3 : 2;
      auto exp_tile = [&](float mr) -> uint32_t {      // P = 2^(s * scale - mr) -> bf16 -> TMEM (over the scores); returns the OR of the packed words
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-mr, -mr);
        uint32_t any = 0;
        float mxe = -INFINITY;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float2 t = __ffma2_rn(make_float2(__uint_as_float(v[hf * 32 + 2 * i]), __uint_as_float(v[hf * 32 + 2 * i + 1])), sc2, nm2);
            float2 e;
            if ((i & 7) < EMU_N) {
              mxe = fmaxf(mxe, fmaxf(t.x, t.y));      // the polynomial path wraps its exponent on overflow: its arguments are checked directly
              e = exp2_emu2<EMU_DEG>(t);
            } else {
              e = make_float2(fast_exp2(t.x), fast_exp2(t.y));
            }
            pk[i] = __byte_perm(__float_as_uint(e.x), __float_as_uint(e.y), 0x7632);   // truncate to bf16; the row sum comes from the MMA
          }
#pragma unroll
          for (int i = 0; i < 16; i += 2) any |= pk[i] | pk[i + 1];
          tmem_st_32x32b_x16(t_lane + (uint32_t)(hf * 16), pk);
        }
        return any | (mxe >= 1.f ? 0x4000u : 0u);
      };
      bool redo = j == 0;
      if (j > 0) redo = (exp_tile(m_ref) & 0x40004000u) != 0;      // some P >= 2 (or inf / NaN)
      if (__any_sync(0xffffffffu, redo)) {
        float m4[4] = {__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])};
#pragma unroll
        for (int i = 4; i < TA_BN; i += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) m4[u] = fmaxf(m4[u], __uint_as_float(v[i + u]));
        }
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2;
        if (j == 0) {
          m_ref = (mx == -INFINITY) ? 0.f : mx + 7.f;
        } else {
          const float m_new = redo ? mx + 7.f : m_ref;      // rows that stayed below the threshold keep their reference (factor 1)
          const float f = fast_exp2(m_ref - m_new);
          m_ref = m_new;
#pragma unroll
          for (int c = 0; c < DO / 16; ++c) {
            uint32_t ov[16];
            tmem_ld_32x32b_x16(t_lane + (uint32_t)(TMEM_O + c * 16), ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
            tmem_st_32x32b_x16(t_lane + (uint32_t)(TMEM_O + c * 16), ov);
          }
        }
        exp_tile(m_ref);
      }
#endif
      AF_ATTN_TR(if (str_) tp[3] = clock64();)
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[g]);
      AF_ATTN_TR(if (str_) tp[4] = clock64();)
      AF_ATTN_TR(if (p.trace && blockIdx.x == 0 && lane == 0 && j >= 8 && j < 16) p.trace[512 + qd * 32 + g * 8 + (j - 8)] = clock64();)
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const int qi = SPLIT ? g >> 1 : g;                              // query tile of this TMEM slot
    const int grow = m0 + qi * TA_BM + row;
    bf16* orow = p.o + (long long)b * p.o_sb + (long long)grow * p.o_sn + h * D;
    if constexpr (SPLIT) {
      // merge the two key halves of a tile: the odd slot hands (m, O[0..DO)) to the even slot through shared memory
      // (the K / V ring is free: every MMA has completed)
      float* xch = reinterpret_cast<float*>(sK) + (size_t)qi * TA_BM * (DO + 1);
      float acc[DO];
#pragma unroll
      for (int c = 0; c < DO / 16; ++c) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_lane + (uint32_t)(TMEM_O + c * 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 16; ++x) acc[c * 16 + x] = __uint_as_float(v[x]);
      }
      if (g & 1) {
#pragma unroll
        for (int x = 0; x < DO; ++x) xch[x * TA_BM + row] = acc[x];     // column-major: conflict-free
        xch[DO * TA_BM + row] = m_ref;
      }
      asm volatile("bar.sync %0, 256;" ::"r"(1 + qi) : "memory");
      if (!(g & 1)) {
        const float m1 = xch[DO * TA_BM + row];
        const float m = fmaxf(m_ref, m1);
        const float f0 = fast_exp2(m_ref - m), f1 = fast_exp2(m1 - m);
#pragma unroll
        for (int x = 0; x < DO; ++x) acc[x] = acc[x] * f0 + xch[x * TA_BM + row] * f1;
        const float l_run = acc[D];
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        if (p.lse && grow < p.Lq) p.lse[((long long)b * H + h) * p.Lq + grow] = l_run > 0.f ? m + log2f(l_run) : INFINITY;
        if (grow < p.Lq) {
#pragma unroll
          for (int c8 = 0; c8 < D / 8; ++c8) {
            uint4 pk;
            pk.x = pack_bf16(acc[c8 * 8 + 0] * inv, acc[c8 * 8 + 1] * inv);
            pk.y = pack_bf16(acc[c8 * 8 + 2] * inv, acc[c8 * 8 + 3] * inv);
            pk.z = pack_bf16(acc[c8 * 8 + 4] * inv, acc[c8 * 8 + 5] * inv);
            pk.w = pack_bf16(acc[c8 * 8 + 6] * inv, acc[c8 * 8 + 7] * inv);
            *reinterpret_cast<uint4*>(orow + c8 * 8) = pk;
          }
        }
      }
    } else {
      float l_run;                                                 // row sum of P = column D of the accumulator
      {
        uint32_t v8[8];
        tmem_ld_32x32b_x8(t_lane + (uint32_t)(TMEM_O + (D & ~7)), v8);
        tmem_ld_wait();
        l_run = __uint_as_float(v8[D & 7]);
      }
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      if (p.lse && grow < p.Lq)
        p.lse[((long long)b * H + h) * p.Lq + grow] = l_run > 0.f ? m_ref + log2f(l_run) : INFINITY;
#pragma unroll
      for (int c = 0; c < DO / 16; ++c) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_lane + (uint32_t)(TMEM_O + c * 16), v);
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
  }
  AF_PDL_TRIGGER_LATE();
  tc_fence_before();
  __syncthreads();
  if (warp == kMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, G * TMEM_G);
  }
}

// tQ4 / tQ2: Q maps are identical (128-row boxes); the tail launch only changes the unit size.
template <int D, int EMU, bool MASK = false, int NI = 1>
static int launch_ta_quad(const CUtensorMap& tQ, const CUtensorMap& tK, const CUtensorMap& tV, const TaParams& p, int B, int H,
                          cudaStream_t stream) {
  using Cfg = TaCfg<D>;
  constexpr int smem4 = 4 * Cfg::Q_BYTES + TQ_ST * (Cfg::K_BYTES + Cfg::V_BYTES) + 1024 + 256;
  constexpr int smem2 = 2 * Cfg::Q_BYTES + TQ_ST * (Cfg::K_BYTES + Cfg::V_BYTES) + 1024 + 256;
  constexpr int smemS = 2 * Cfg::Q_BYTES + TQ_ST * 2 * (Cfg::K_BYTES + Cfg::V_BYTES) + 1024 + 256;
  static_assert(2 * TA_BM * (Cfg::DO + 1) * 4 <= TQ_ST * 2 * (Cfg::K_BYTES + Cfg::V_BYTES), "SPLIT merge scratch must fit the K/V ring");
  static DevOnce configured;
  const int cfg_dev = af_device();
  if (!configured.done(cfg_dev)) {
    AF_CUDA(cudaFuncSetAttribute((attn_fwd_tcgen05_quad_kernel<D, 4, EMU, false, MASK, NI>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem4));
    AF_CUDA(cudaFuncSetAttribute((attn_fwd_tcgen05_quad_kernel<D, 2, EMU, false, MASK, NI>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
    AF_CUDA(cudaFuncSetAttribute((attn_fwd_tcgen05_quad_kernel<D, 4, EMU, true, MASK, NI>), cudaFuncAttributeMaxDynamicSharedMemorySize, smemS));
    configured.set(cfg_dev);
  }
  static int tail_mode = -1;
  if (tail_mode < 0) {
    const char* e = getenv("ADAFACE_QUAD_TAIL");     // 1 (default): two tiles x two key halves per tail CTA; 2: two-tile CTAs
    tail_mode = (e && e[0] == '2') ? 2 : 1;
  }
  const int n_sm = af_num_sms();
  const int n_qb4 = (p.Lq + 4 * TA_BM - 1) / (4 * TA_BM);
  const int units4 = B * H * n_qb4;
  const int n_ktiles = (p.Lk + TA_BN - 1) / TA_BN;
  // bulk: whole rounds of four-tile CTAs; tail: the rest as two-tile CTAs (only when Lq splits evenly into them)
  int bulk = (units4 / n_sm) * n_sm;
  if (p.Lq % (4 * TA_BM) != 0 || bulk == 0) bulk = units4;
  if (bulk > 0) {
    AF_CUDA(launch_pdl(2, attn_fwd_tcgen05_quad_kernel<D, 4, EMU, false, MASK, NI>, dim3(bulk), dim3((17 + NI) * 32), smem4, stream, tQ, tK, tV, p, 0, H));
    ++g_launch_count;
  }
  if (units4 > bulk) {
    if (tail_mode == 1 && n_ktiles % 2 == 0 && p.Lk % TA_BN == 0)
      AF_CUDA(launch_pdl(2, attn_fwd_tcgen05_quad_kernel<D, 4, EMU, true, MASK, NI>, dim3(2 * (units4 - bulk)), dim3((17 + NI) * 32), smemS, stream, tQ, tK, tV, p, 2 * bulk, H));
    else
      AF_CUDA(launch_pdl(2, attn_fwd_tcgen05_quad_kernel<D, 2, EMU, false, MASK, NI>, dim3(2 * (units4 - bulk)), dim3((9 + NI) * 32), smem2, stream, tQ, tK, tV, p, 2 * bulk, H));
    ++g_launch_count;
  }
  AF_CUDA(cudaGetLastError());
  if (p.trace) {      // diagnosis: CTA 0 of the bulk launch, key tiles 8..15
    static long long h[1024];
    cudaDeviceSynchronize();
    cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
    const long long t0 = h[0];
    for (int j = 0; j < 8; ++j) {
      fprintf(stderr, "issuer j=%2d: kv wait %6lld..%6lld |", j + 8, h[j * 16] - t0, h[j * 16 + 1] - t0);
      for (int g = 0; g < 4; ++g) fprintf(stderr, " g%d p_full wait %6lld..%6lld issued %6lld |", g, h[j * 16 + 2 + g * 3] - t0, h[j * 16 + 3 + g * 3] - t0, h[j * 16 + 4 + g * 3] - t0);
      fprintf(stderr, "\n");
    }
    for (int g = 0; g < 4; ++g)
      for (int j = 0; j < 8; ++j) {
        const long long* tp = h + 256 + g * 64 + j * 8;
        fprintf(stderr, "softmax g=%d j=%2d: s_full wait %6lld..%6lld  max done %6lld  exp+pack done %6lld  arrived %6lld  (warps 0..3: %6lld %6lld %6lld %6lld)\n", g, j + 8, tp[0] - t0, tp[1] - t0, tp[2] - t0, tp[3] - t0, tp[4] - t0,
                h[512 + g * 8 + j] - t0, h[512 + 32 + g * 8 + j] - t0, h[512 + 64 + g * 8 + j] - t0, h[512 + 96 + g * 8 + j] - t0);
      }
  }
  return 0;
}

template <int D, int EMU>
static int launch_ta_mc(const CUtensorMap& tQ, const CUtensorMap& tK, const CUtensorMap& tV, const TaParams& p, int B, int H,
                        cudaStream_t stream) {
  using Cfg = TaCfg<D>;
  constexpr int smem = Cfg::Q_BYTES + 2 * (Cfg::K_BYTES + Cfg::V_BYTES) + 1024 + 128;
  static DevOnce configured;
  const int cfg_dev = af_device();
  if (!configured.done(cfg_dev)) {
    AF_CUDA(cudaFuncSetAttribute(attn_fwd_tcgen05_mc_kernel<D, EMU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.set(cfg_dev);
  }
  dim3 grid((p.Lq + TA_BM - 1) / TA_BM, H, B);
  AF_CUDA(launch_pdl(2, attn_fwd_tcgen05_mc_kernel<D, EMU>, grid, dim3(TA_THREADS), smem, stream, tQ, tK, tV, p));
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

template <int D, int EMU>
static int launch_ta_mcp(const CUtensorMap& tQ, const CUtensorMap& tK, const CUtensorMap& tV, const TaParams& p, int B, int H,
                         cudaStream_t stream) {
  using Cfg = TaCfg<D>;
  constexpr int smem = Cfg::Q_BYTES + 2 * (Cfg::K_BYTES + Cfg::V_BYTES) + 1024 + 256;
  AF_CONFIG_SMEM((attn_fwd_tcgen05_mcp_kernel<D, EMU>), smem);
  const int n_units = B * H * ((p.Lq + TA_BM - 1) / TA_BM);
  int grid = 2 * af_num_sms();
  if (grid > n_units) grid = n_units;
  AF_CUDA(launch_pdl(2, attn_fwd_tcgen05_mcp_kernel<D, EMU>, dim3(grid), dim3(TA_THREADS), smem, stream, tQ, tK, tV, p, H, n_units));
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

template <int D, int EMU, bool PT>
static int launch_ta(const CUtensorMap& tQ, const CUtensorMap& tK, const CUtensorMap& tV, const TaParams& p, int B, int H,
                     cudaStream_t stream) {
  using Cfg = TaCfg<D>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "tcgen05 attention: shared memory exceeds the SM");
  static DevOnce configured;
  const int cfg_dev = af_device();
  if (!configured.done(cfg_dev)) {
    AF_CUDA(cudaFuncSetAttribute(attn_fwd_tcgen05_kernel<D, EMU, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured.set(cfg_dev);
  }
  dim3 grid((p.Lq + TA_BM - 1) / TA_BM, H, B);
  AF_CUDA(launch_pdl(2, attn_fwd_tcgen05_kernel<D, EMU, PT>, grid, dim3(TA_THREADS), Cfg::SMEM_BYTES, stream, tQ, tK, tV, p));
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

// =============================================================================================
// Short-context variant (cross-attention over the 77 / 97-token prompt, Lk <= 128): HBM-bound, so built to stream.
// With one CTA per 128-query tile the kernel above spends most of a CTA's life in set-up (TMEM allocation, barrier
// init, descriptor fetch, first-TMA latency: ~9 us per wave for ~1 us of work).  Here CTAs are PERSISTENT: each owns
// a contiguous range of (batch, head, query-tile) units, keeps K and V of the current (batch, head) in shared memory
// (loaded once, all NK = ceil16(Lk) keys as ONE tile), double-buffers the Q tiles through TMA, and runs an EXACT
// single-pass softmax per tile (every key is present: no online rescaling).  S (NK fp32 columns) and P (bf16, written
// over S) share one TMEM region, O sits behind it; 128 TMEM columns when NK + DO <= 128 (d = 40, 77 keys: 4 CTAs / SM).
struct TcParams {
  bf16* o;
  long long o_sb, o_sn;
  int Lq, Lk, H, NK, n_qtiles, n_units, tmem_cols;
  int q_wide;              // Q map spans whole [B, Lq, H*d] rows: the box of head h starts at column h * d (see launch_tc_cross)
  float scale_log2;
  float* lse;
  long long* trace;        // diagnosis only (-DAF_ATTN_TRACE build, env ADAFACE_ATTN_TRACE): CTA 0 stamps its hand-offs
};

// NKT = compile-time number of staged keys (80: the 77-token prompt; 128: anything up to 128, e.g. 97 in training):
// the score loops unroll completely and only the chunk that straddles Lk pays for masking.
#ifndef AF_TC_EMU
#define AF_TC_EMU 0      // measured at level A: 36.5 us (0) / 36.1 (2) / 35.9 (3) per cross block -- not MUFU-bound, left off
#endif
constexpr int TC_EMU = AF_TC_EMU;
template <int D, int NKT>
__global__ void __launch_bounds__(TA_THREADS, (D <= 64) ? 4 : 2)
attn_cross_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const TcParams p) {
  using Cfg = TaCfg<D>;
  constexpr int NA = Cfg::NA, KT = Cfg::KT;
  constexpr int DO = (D % 16 == 0) ? D + 16 : Cfg::DO;      // accumulator width incl. a pad column D that receives the row sums
  constexpr int NK = NKT;                                  // staged keys (multiple of 16, <= 128)
  constexpr int KV_ATOM = NK * 128;                        // bytes of one 64-column atom of the K / V tile
  constexpr int TMEM_O = NK;                               // O behind the S / P region
  extern __shared__ uint8_t smem_raw_tc[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_tc) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                      // [2][NA][128 rows][128 B]
  uint8_t* sK = sQ + 2 * Cfg::Q_BYTES;                     // [NA][NK][128 B]
  uint8_t* sV = sK + NA * KV_ATOM;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NA * KV_ATOM);
  uint64_t* q_full = bars;                                 // [2]
  uint64_t* q_empty = bars + 2;                            // [2]
  uint64_t* kv_full = bars + 4;
  uint64_t* kv_empty = bars + 5;
  uint64_t* s_full = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* o_full = bars + 8;
  uint64_t* o_free = bars + 9;
  uint64_t* v_ready = bars + 10;                           // K / V of the current (batch, head) landed AND V carries its ones column
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = (int)((long long)blockIdx.x * p.n_units / gridDim.x);
  const int u1 = (int)((long long)(blockIdx.x + 1) * p.n_units / gridDim.x);

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 1);
    mbar_init(v_ready, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(o_free, 128);
    fence_barrier_init();
  } else if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // global memory is touched only after the predecessor kernel has completed
  AF_PDL_TRIGGER_EARLY();

  if (warp == kTmaWarp) {
    // whole warp: one elected lane issues the TMA loads; when a new (batch, head) arrives all lanes write the ONES
    // COLUMN (column D of every V row := 1.0) so that column D of the P.V accumulator is the row sum of P
    const bool leader = elect_one();
    int prev_bh = -1, kv_loads = 0;
    for (int u = u0, i = 0; u < u1; ++u, ++i) {
      const int bh = u / p.n_qtiles, qt = u - bh * p.n_qtiles;
      const int b = bh / p.H, h = bh - b * p.H;
      const bool new_kv = bh != prev_bh;
      if (leader) {
        if (new_kv) {
          if (kv_loads > 0) mbar_wait(kv_empty, (kv_loads - 1) & 1);   // every MMA that read the old K / V has retired
          mbar_arrive_expect_tx(kv_full, 2 * NA * KV_ATOM);
#pragma unroll
          for (int a = 0; a < NA; ++a) {
            tma_load_4d(sK + a * KV_ATOM, &tmK, kv_full, a * 64, h, 0, b);
            tma_load_4d(sV + a * KV_ATOM, &tmV, kv_full, a * 64, h, 0, b);
          }
        }
        const int s = i & 1;
        if (i >= 2) mbar_wait(&q_empty[s], ((i >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&q_full[s], Cfg::Q_BYTES);
#pragma unroll
        for (int a = 0; a < NA; ++a)
          tma_load_4d(sQ + s * Cfg::Q_BYTES + a * (TA_BM * 128), &tmQ, &q_full[s], p.q_wide ? h * D + a * 64 : a * 64, p.q_wide ? 0 : h, qt * TA_BM, b);
      }
      if (new_kv) {
        mbar_wait(kv_full, kv_loads & 1);
        uint8_t* atom = sV + (D >> 6) * KV_ATOM;             // the 64-column atom that holds column D
        for (int r = lane; r < NK; r += 32)
          *reinterpret_cast<uint16_t*>(atom + r * 128 + (((((D & 63) >> 3) ^ (r & 7)) << 4) | ((D & 7) << 1))) = 0x3F80;
        fence_proxy_async_smem();
        __syncwarp();
        if (leader) mbar_arrive(v_ready);
        ++kv_loads;
        prev_bh = bh;
      }
      __syncwarp();
    }
  } else if (warp == kMmaWarp) {
    if (elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc_bf16_f32(TA_BM, NK, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16_f32(TA_BM, DO, true);
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
      auto tc_wait = [&](uint64_t* bar, uint32_t parity) {      // non-blocking test first: the suspending try_wait costs ~130 clk even on a completed phase
        uint32_t ok;
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!ok) mbar_wait(bar, parity);
      };
      int prev_bh = -1, kv_loads = 0;
      for (int u = u0, i = 0; u < u1; ++u, ++i) {
        const int bh = u / p.n_qtiles;
        if (bh != prev_bh) {
          mbar_wait(v_ready, kv_loads & 1);
          ++kv_loads;
          prev_bh = bh;
        }
        const int s = i & 1;
        AF_ATTN_TR(const bool tr = p.trace && blockIdx.x == 0 && i < 4; if (tr) p.trace[i * 8] = clock64();)
        tc_wait(&q_full[s], (i >> 1) & 1);
        AF_ATTN_TR(if (tr) p.trace[i * 8 + 1] = clock64();)
        if (i >= 1) tc_wait(o_free, (i - 1) & 1);            // the softmax warps have drained S / P / O of the previous unit
        AF_ATTN_TR(if (tr) p.trace[i * 8 + 2] = clock64();)
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
          const uint64_t da = make_smem_desc_sw128(aQ + s * Cfg::Q_BYTES + (kk >> 2) * (TA_BM * 128) + (kk & 3) * 32);
          const uint64_t db = make_smem_desc_sw128(aK + (kk >> 2) * KV_ATOM + (kk & 3) * 32);
          umma_bf16(tmem_base, da, db, idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        umma_commit(&q_empty[s]);                            // the Q tile is free once S has been produced
        AF_ATTN_TR(if (tr) p.trace[i * 8 + 3] = clock64();)
        tc_wait(p_full, i & 1);
        AF_ATTN_TR(if (tr) p.trace[i * 8 + 4] = clock64();)
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < NK / 16; ++k) {
          const uint64_t db = make_smem_desc_sw128_mn(aV + k * 2048, (uint32_t)KV_ATOM);
          umma_bf16_ts(tmem_base + (uint32_t)TMEM_O, tmem_base + (uint32_t)(k * 8), db, idesc_pv, k != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        AF_ATTN_TR(if (tr) p.trace[i * 8 + 5] = clock64();)
        const int next_bh = (u + 1 < u1) ? (u + 1) / p.n_qtiles : -1;
        if (next_bh != bh) umma_commit(kv_empty);
      }
    }
  } else {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(qd * 32) << 16);
    for (int u = u0, i = 0; u < u1; ++u, ++i) {
      const int bh = u / p.n_qtiles, qt = u - bh * p.n_qtiles;
      const int b = bh / p.H, h = bh - b * p.H;
      AF_ATTN_TR(const bool str_ = p.trace && blockIdx.x == 0 && warp == 0 && lane == 0 && i < 4; long long* tp = p.trace + 64 + i * 8; if (str_) tp[0] = clock64();)
      mbar_wait(s_full, i & 1);
      AF_ATTN_TR(if (str_) tp[1] = clock64();)
      tc_fence_after();
      // ---- pass 1: row max of the raw scores (keys >= Lk are zero rows of K: masked out).  The scores are read in the widest chunks
      //      the register budget allows (32 columns; 80 registers per thread at four CTAs per SM): every tcgen05.ld round trip costs ~170 clk
      //      there (traced: ten 16-column round trips were 2000 of a tile's 4600 clk).
      float mxr = -INFINITY;
      auto chunk_max = [&](auto& v, int c0) {
        constexpr int W = sizeof(v) / sizeof(v[0]);
        if (c0 + W > p.Lk) {
#pragma unroll
          for (int j = 0; j < W; ++j)
            if (c0 + j >= p.Lk) v[j] = 0xff800000u;
        }
        float m4[4] = {__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3])};
#pragma unroll
        for (int j = 4; j < W; j += 4) {
#pragma unroll
          for (int x = 0; x < 4; ++x) m4[x] = fmaxf(m4[x], __uint_as_float(v[j + x]));
        }
        mxr = fmaxf(mxr, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
      };
#pragma unroll
      for (int c = 0; c + 32 <= NK; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32_wait(t_lane + (uint32_t)c, v);
        chunk_max(v, c);
      }
      if constexpr (NK % 32 != 0) {
        static_assert(NK % 32 == 16, "staged keys: 80 or 128");
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_lane + (uint32_t)(NK - 16), v);
        tmem_ld_wait();
        chunk_max(v, NK - 16);
      }
      const float m_ref = mxr * p.scale_log2;
      AF_ATTN_TR(if (str_) tp[2] = clock64() + (m_ref > 1e30f ? 1 : 0);)
      // ---- pass 2: P = exp2(s * scale - max) truncated to bf16, written over the scores it came from (32-column chunks; the tail 16)
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_ref, -m_ref);
      auto chunk_exp = [&](auto& v, auto& pk, int c0) {
        constexpr int W = sizeof(v) / sizeof(v[0]);
        if (c0 + W > p.Lk) {
#pragma unroll
          for (int j = 0; j < W; ++j)
            if (c0 + j >= p.Lk) v[j] = 0xff800000u;
        }
#pragma unroll
        for (int j = 0; j < W / 2; ++j) {
          const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), sc2, nm2);
          // (TC_EMU of every 8 pairs on the FMA-pipe polynomial: A/B knob, see AF_TC_EMU)
          const float2 e = ((j & 7) < TC_EMU) ? exp2_emu2<2>(t) : make_float2(fast_exp2(t.x), fast_exp2(t.y));
          pk[j] = __byte_perm(__float_as_uint(e.x), __float_as_uint(e.y), 0x7632);   // truncate; the row sum comes from the MMA
        }
      };
#pragma unroll
      for (int c = 0; c + 32 <= NK; c += 32) {
        uint32_t v[32], pk[16];
        tmem_ld_32x32b_x32_wait(t_lane + (uint32_t)c, v);
        chunk_exp(v, pk, c);
        tmem_st_32x32b_x16(t_lane + (uint32_t)(c >> 1), pk);      // columns [c/2, c/2 + 16): already consumed
      }
      if constexpr (NK % 32 != 0) {
        uint32_t v[16], pk[8];
        tmem_ld_32x32b_x16(t_lane + (uint32_t)(NK - 16), v);
        tmem_ld_wait();
        chunk_exp(v, pk, NK - 16);
        tmem_st_32x32b_x8(t_lane + (uint32_t)((NK - 16) >> 1), pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
      AF_ATTN_TR(if (str_) tp[3] = clock64();)
      // ---- epilogue: O / l -> bf16 -> [B, Lq, H*d]
      mbar_wait(o_full, i & 1);
      AF_ATTN_TR(if (str_) tp[4] = clock64();)
      tc_fence_after();
      // the whole accumulator row in ONE round trip (32-column loads, one wait): row sum of P = column D
      uint32_t acc[DO];
#pragma unroll
      for (int c = 0; c + 32 <= DO; c += 32) {
        uint32_t (&blk)[32] = *reinterpret_cast<uint32_t (*)[32]>(&acc[c]);
        tmem_ld_32x32b_x32_nowait(t_lane + (uint32_t)(TMEM_O + c), blk);
      }
      if constexpr (DO % 32 != 0) {
        uint32_t (&blk)[16] = *reinterpret_cast<uint32_t (*)[16]>(&acc[DO - 16]);
        tmem_ld_32x32b_x16(t_lane + (uint32_t)(TMEM_O + DO - 16), blk);
      }
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < DO; ++c) asm volatile("" : "+r"(acc[c]));      // nothing is consumed before the wait
      const float l_run = __uint_as_float(acc[D]);
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      const int grow = qt * TA_BM + row;
      if (p.lse && grow < p.Lq)
        p.lse[((long long)b * p.H + h) * p.Lq + grow] = l_run > 0.f ? m_ref + log2f(l_run) : INFINITY;
      bf16* orow = p.o + (long long)b * p.o_sb + (long long)grow * p.o_sn + h * D;
      if (grow < p.Lq) {
#pragma unroll
        for (int c8 = 0; c8 < D / 8; ++c8) {
          uint4 pk;
          pk.x = pack_bf16(__uint_as_float(acc[c8 * 8 + 0]) * inv, __uint_as_float(acc[c8 * 8 + 1]) * inv);
          pk.y = pack_bf16(__uint_as_float(acc[c8 * 8 + 2]) * inv, __uint_as_float(acc[c8 * 8 + 3]) * inv);
          pk.z = pack_bf16(__uint_as_float(acc[c8 * 8 + 4]) * inv, __uint_as_float(acc[c8 * 8 + 5]) * inv);
          pk.w = pack_bf16(__uint_as_float(acc[c8 * 8 + 6]) * inv, __uint_as_float(acc[c8 * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c8 * 8) = pk;
        }
      }
      tc_fence_before();
      mbar_arrive(o_free);
      AF_ATTN_TR(if (str_) tp[5] = clock64();)
    }
  }
  AF_PDL_TRIGGER_LATE();
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

template <int D, int NKT>
static int launch_tc_cross(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sh,
                           int64_t k_sn, const void* v, int64_t v_sb, int64_t v_sh, int64_t v_sn, void* o, int64_t o_sb,
                           int64_t o_sn, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t drow_q, int64_t drow_kv,
                           float scale, float* lse, cudaStream_t stream) {
  using Cfg = TaCfg<D>;
  constexpr int NK = NKT;
  CUtensorMap tQ, tK, tV;
  // Q in the reference layout [B, Lq, H*d] (heads side by side in a row): a box over ONE head's d columns makes the TMA unit fetch 80-byte
  // row pieces and fill the rest of the 128-byte swizzle row itself -- 16 useful B/clk/SM (profiles/r01_microbench.md), and the traced
  // kernel waited ~950 clk per tile for its Q.  Instead the map spans whole rows and the box of head h starts at column h*d and is 64
  // columns wide: full 128-byte rows at 73 B/clk/SM.  Columns d..63 of the tile then hold the NEXT head's values (zeros past the row end);
  // they meet K's columns d..63, which ARE zero (K keeps its per-head map with out-of-bounds fill), so the scores are unchanged.
  static int wide_on = -1;
  if (wide_on < 0) {
    const char* e = getenv("ADAFACE_CROSS_QWIDE");      // A/B switch (default on)
    wide_on = (e && e[0] == '0') ? 0 : 1;
  }
  const bool q_wide = wide_on && q_sh == (int64_t)D && drow_q == (int64_t)D && q_sn >= H * (int64_t)D;
  if (q_wide) {
    if (make_tmap_bf16_heads(&tQ, q, (uint64_t)(H * D), 1, (uint64_t)Lq, (uint64_t)B, (uint64_t)(H * D), (uint64_t)q_sn, (uint64_t)q_sb, TA_BM)) return 3;
  } else {
    if (make_tmap_bf16_heads(&tQ, q, (uint64_t)drow_q, (uint64_t)H, (uint64_t)Lq, (uint64_t)B, (uint64_t)q_sh, (uint64_t)q_sn, (uint64_t)q_sb, TA_BM)) return 3;
  }
  if (make_tmap_bf16_heads(&tK, k, (uint64_t)drow_kv, (uint64_t)H, (uint64_t)Lk, (uint64_t)B, (uint64_t)k_sh, (uint64_t)k_sn, (uint64_t)k_sb, (uint32_t)NK)) return 3;
  if (make_tmap_bf16_heads(&tV, v, (uint64_t)drow_kv, (uint64_t)H, (uint64_t)Lk, (uint64_t)B, (uint64_t)v_sh, (uint64_t)v_sn, (uint64_t)v_sb, (uint32_t)NK)) return 3;
  TcParams p;
  p.q_wide = q_wide ? 1 : 0;
  p.o = (bf16*)o; p.o_sb = o_sb; p.o_sn = o_sn;
  p.Lq = (int)Lq; p.Lk = (int)Lk; p.H = (int)H; p.NK = NK;
  p.n_qtiles = (int)((Lq + TA_BM - 1) / TA_BM);
  p.n_units = (int)(B * H) * p.n_qtiles;
  constexpr int DO_S = (D % 16 == 0) ? D + 16 : Cfg::DO;
  p.tmem_cols = (NK + DO_S <= 128) ? 128 : 256;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse;
  p.trace = nullptr;
  if (getenv("ADAFACE_ATTN_TRACE")) {
    static long long* tbuf = nullptr;
    if (!tbuf) cudaMalloc(&tbuf, 1024 * 8);
    cudaMemset(tbuf, 0, 1024 * 8);
    p.trace = tbuf;
  }
  const int smem = 2 * Cfg::Q_BYTES + 2 * Cfg::NA * NK * 128 + 1024 + 128;
  static int configured[AF_MAX_DEV] = {0};                 // per device
  const int cfg_dev = af_device();
  if (smem > configured[cfg_dev]) {
    AF_CUDA(cudaFuncSetAttribute((attn_cross_tc_kernel<D, NKT>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[cfg_dev] = smem;
  }
  int per_sm = 512 / p.tmem_cols;                           // TMEM columns bound the residency
  const int by_smem = (227 * 1024) / (smem + 1024);
  if (per_sm > by_smem) per_sm = by_smem;
  if (per_sm > ((D <= 64) ? 4 : 2)) per_sm = (D <= 64) ? 4 : 2;
  if (per_sm < 1) per_sm = 1;
  int grid = af_num_sms() * per_sm;
  if (grid > p.n_units) grid = p.n_units;
  AF_CUDA(launch_pdl(2, attn_cross_tc_kernel<D, NKT>, dim3(grid), dim3(TA_THREADS), smem, stream, tQ, tK, tV, p));
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  if (p.trace) {      // diagnosis: CTA 0, its first four tiles
    static long long h[1024];
    cudaDeviceSynchronize();
    cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
    const long long t0 = h[0];
    fprintf(stderr, "cross: grid %d, units %d\n", grid, p.n_units);
    for (int i = 0; i < 4; ++i)
      fprintf(stderr, "issuer tile %d: q_full wait %6lld..%6lld  o_free seen %6lld  QK issued %6lld  p_full seen %6lld  PV issued %6lld\n", i, h[i * 8] - t0, h[i * 8 + 1] - t0,
              h[i * 8 + 2] - t0, h[i * 8 + 3] - t0, h[i * 8 + 4] - t0, h[i * 8 + 5] - t0);
    for (int i = 0; i < 4; ++i)
      fprintf(stderr, "softmax tile %d: s_full wait %6lld..%6lld  max done %6lld  P arrived %6lld  o_full seen %6lld  stored %6lld\n", i, h[64 + i * 8] - t0, h[64 + i * 8 + 1] - t0,
              h[64 + i * 8 + 2] - t0, h[64 + i * 8 + 3] - t0, h[64 + i * 8 + 4] - t0, h[64 + i * 8 + 5] - t0);
  }
  return 0;
}

// Unmasked attention on the tensor-core path.  Returns -1 when the problem is not eligible (caller falls through to
// the warp-MMA kernel, which handles masks, causal multi-KV and tiny shapes), 0 on success, > 0 on error.
// q/k/v element strides: batch (sb), head (sh), token (sn).  Reference layout [B, L, H*d]: sh = d, sn = H*d.
int attn_fwd_tcgen05(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sh,
                     int64_t k_sn, const void* v, int64_t v_sb, int64_t v_sh, int64_t v_sn, void* o, int64_t o_sb,
                     int64_t o_sn, int64_t B, int64_t H, int64_t Lq, int64_t Lk, int64_t d, int64_t drow_q, int64_t drow_kv,
                     float scale, float* lse, const uint8_t* key_mask, cudaStream_t stream) {
  if (!(d == 40 || d == 80 || d == 160)) return -1;
  // a key mask runs on the four-tile kernel only (d = 40, level A: the case that costs 3.5x on the warp-MMA kernel)
  const bool mask_quad = key_mask && d == 40 && drow_q == 40 && drow_kv == 40 && Lq >= 2 * TQ_G * TA_BM && Lk > 128;
  // other masked shapes (level B: d = 80; short d = 40 maps): the small-CTA kernel masks the scores in registers
  const bool mask_mc = key_mask && !mask_quad && (d == 40 || d == 80) && Lk % 64 == 0 && Lk > 128 && Lq >= 128;
  if (key_mask && !mask_quad && !mask_mc) return -1;
  // drow_* = elements that exist in a row: d, or the zero-padded width of a head-major buffer
  if (drow_q < d) drow_q = d;
  if (drow_kv < d) drow_kv = d;
  static int tc_cross = -1;
  if (tc_cross < 0) {
    const char* e = getenv("ADAFACE_CROSS_TC");       // 1 (default): persistent short-context kernel for Lk <= 128
    tc_cross = (e && e[0] == '0') ? 0 : 1;
  }
  if (tc_cross && Lk <= 128 && Lq >= 512 && (d == 40 || d == 80)) {
#define AF_TC_ARGS q, q_sb, q_sh, q_sn, k, k_sb, k_sh, k_sn, v, v_sb, v_sh, v_sn, o, o_sb, o_sn, B, H, Lq, Lk, drow_q, drow_kv, scale, lse, stream
    if (d == 40) return Lk <= 80 ? launch_tc_cross<40, 80>(AF_TC_ARGS) : launch_tc_cross<40, 128>(AF_TC_ARGS);
    return Lk <= 80 ? launch_tc_cross<80, 80>(AF_TC_ARGS) : launch_tc_cross<80, 128>(AF_TC_ARGS);
#undef AF_TC_ARGS
  }
  CUtensorMap tQ, tK, tV;
  if (make_tmap_bf16_heads(&tQ, q, (uint64_t)drow_q, (uint64_t)H, (uint64_t)Lq, (uint64_t)B, (uint64_t)q_sh, (uint64_t)q_sn, (uint64_t)q_sb, TA_BM)) return 3;
  if (make_tmap_bf16_heads(&tK, k, (uint64_t)drow_kv, (uint64_t)H, (uint64_t)Lk, (uint64_t)B, (uint64_t)k_sh, (uint64_t)k_sn, (uint64_t)k_sb, TA_BN)) return 3;
  if (make_tmap_bf16_heads(&tV, v, (uint64_t)drow_kv, (uint64_t)H, (uint64_t)Lk, (uint64_t)B, (uint64_t)v_sh, (uint64_t)v_sn, (uint64_t)v_sb, TA_BN)) return 3;
  TaParams p;
  p.o = (bf16*)o;
  p.o_sb = o_sb;
  p.o_sn = o_sn;
  p.Lq = (int)Lq;
  p.Lk = (int)Lk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = lse;
  p.key_mask = key_mask;
  p.wide = 0;
  p.trace = nullptr;
  if (getenv("ADAFACE_ATTN_TRACE")) {
    static long long* tbuf = nullptr;
    if (!tbuf) cudaMalloc(&tbuf, 1024 * 8);
    cudaMemset(tbuf, 0, 1024 * 8);
    p.trace = tbuf;
  }
  // d = 80 in the reference layout [B, L, H*d] (also the q / k / v column blocks of a fused projection buffer): per-head boxes make the TMA unit
  // fetch 160-byte row pieces and fill the rest of its two 128-byte swizzle rows itself (16 useful B/clk/SM in the microbenchmark; suspected
  // to bound the small-CTA kernel at level B -- 2 CTAs x 20 KB of K / V per 64-key step -- but the A/B below says it does not).  Maps over whole rows with the head's box starting at column
  // h*d move full 128-byte rows; 80 = 5 x 16, so neither Q.K^T (five k16 steps) nor P.V (N = 80) ever touches the neighbour head's columns
  // that ride along in the second swizzle atom.
  CUtensorMap tQw, tKw, tVw;
  auto wide80 = [&]() -> bool {
    static int on = -1;
    if (on < 0) {
      const char* e = getenv("ADAFACE_ATTN_WIDE80");      // A/B switch, default OFF: measured no change at level B (48.1 us either way)
      on = (e && e[0] == '1') ? 1 : 0;
    }
    if (!on || d != 80 || q_sh != 80 || k_sh != 80 || v_sh != 80 || drow_q != 80 || drow_kv != 80) return false;
    const uint64_t W = (uint64_t)(H * 80);
    if ((uint64_t)q_sn < W || (uint64_t)k_sn < W || (uint64_t)v_sn < W) return false;
    if (make_tmap_bf16_heads(&tQw, q, W, 1, (uint64_t)Lq, (uint64_t)B, W, (uint64_t)q_sn, (uint64_t)q_sb, TA_BM)) return false;
    if (make_tmap_bf16_heads(&tKw, k, W, 1, (uint64_t)Lk, (uint64_t)B, W, (uint64_t)k_sn, (uint64_t)k_sb, TA_BN)) return false;
    if (make_tmap_bf16_heads(&tVw, v, W, 1, (uint64_t)Lk, (uint64_t)B, W, (uint64_t)v_sn, (uint64_t)v_sb, TA_BN)) return false;
    p.wide = 1;
    return true;
  };
  static int emu = -1, psmem = 0, mc = 1;
  static bool emu_set = false;
  if (emu < 0) {
    const char* m = getenv("ADAFACE_ATTN_MC");      // 1 (default): many-small-CTAs kernel for d = 40 / 80
    mc = (m && m[0] == '0') ? 0 : 1;
    const char* e = getenv("ADAFACE_EXP_EMU");      // tuning knob: exp2 pairs of every 8 moved off the MUFU unit
    emu = (e && e[0] >= '0' && e[0] <= '7') ? (e[0] - '0') : 0;
    emu_set = e != nullptr;
    const char* ps = getenv("ADAFACE_P_SMEM");      // debugging aid: hand P to the MMA through shared memory
    psmem = (ps && ps[0] == '1') ? 1 : 0;
  }
  const int ib = (int)B, ih = (int)H;
  static int quad = -1;
  if (quad < 0) {
    const char* e = getenv("ADAFACE_ATTN_QUAD");    // 1 (default): four query tiles per CTA, one in-order MMA issuer (d = 40)
    quad = (e && e[0] == '0') ? 0 : 1;
  }
  if (mask_quad) return (quad && mc && !psmem) ? launch_ta_quad<40, 5, true>(tQ, tK, tV, p, ib, ih, stream) : -1;
  if (mask_mc) {
    if (!mc || psmem) return -1;
    if (d == 80 && wide80()) return launch_ta_mc<80, 0>(tQw, tKw, tVw, p, ib, ih, stream);
    return d == 40 ? launch_ta_mc<40, 2>(tQ, tK, tV, p, ib, ih, stream) : launch_ta_mc<80, 0>(tQ, tK, tV, p, ib, ih, stream);
  }
  {
    static int tri = -1;
    if (tri < 0) {
      const char* e = getenv("ADAFACE_ATTN_TRI");     // 1: three-tile kernel with decoupled S / P regions and Q in TMEM (attn_tcgen05_tri.cu)
      tri = (e && e[0] == '1') ? 1 : 0;
    }
    if (tri && d == 40 && Lq >= 1024 && Lk > 128)
      return attn_fwd_tcgen05_tri(q, q_sb, q_sh, q_sn, k, k_sb, k_sh, k_sn, v, v_sb, v_sh, v_sn, B, H, Lq, Lk, drow_q, drow_kv, p, stream);
  }
  if (quad && mc && !psmem && d == 40 && Lq >= 2 * TQ_G * TA_BM && Lk > 128) {
    static int issuers = -1;
    if (issuers < 0) {
      const char* e = getenv("ADAFACE_ATTN_ISSUERS");      // A/B switch: 1 (default) = single MMA-issuing thread, 2 = one per ping-pong half
      issuers = (e && e[0] == '2') ? 2 : 1;                // (measured: 297.0 us with two, 295.9 us with one -- the issuer is not the bound)
    }
    if (issuers == 2 && !emu_set) return launch_ta_quad<40, 5, false, 2>(tQ, tK, tV, p, ib, ih, stream);
    switch (emu_set ? emu : 5) {      // default: 2 of every 8 exp2 pairs on the FMA pipe, degree-2 polynomial (round 2, us: EMU 2: 298.0, 3: 311.3,
                                      // 5 (= 2 pairs, degree 2): 294.9, 6: 306.2, 7: 310.3; round 1: 366 / 349 / 333 / 334 / 342 for 0..4)
      case 0: return launch_ta_quad<40, 0>(tQ, tK, tV, p, ib, ih, stream);
      case 1: return launch_ta_quad<40, 1>(tQ, tK, tV, p, ib, ih, stream);
      case 2: return launch_ta_quad<40, 2>(tQ, tK, tV, p, ib, ih, stream);
      case 3: return launch_ta_quad<40, 3>(tQ, tK, tV, p, ib, ih, stream);
      case 4: return launch_ta_quad<40, 4>(tQ, tK, tV, p, ib, ih, stream);
      case 5: return launch_ta_quad<40, 5>(tQ, tK, tV, p, ib, ih, stream);
      case 6: return launch_ta_quad<40, 6>(tQ, tK, tV, p, ib, ih, stream);
      default: return launch_ta_quad<40, 7>(tQ, tK, tV, p, ib, ih, stream);
    }
  }
  if (mc && !psmem && d == 40) {
    switch (emu) {
      case 0: return launch_ta_mc<40, 0>(tQ, tK, tV, p, ib, ih, stream);
      case 1: return launch_ta_mc<40, 1>(tQ, tK, tV, p, ib, ih, stream);
      case 2: return launch_ta_mc<40, 2>(tQ, tK, tV, p, ib, ih, stream);
      case 3: return launch_ta_mc<40, 3>(tQ, tK, tV, p, ib, ih, stream);
      default: return launch_ta_mc<40, 4>(tQ, tK, tV, p, ib, ih, stream);
    }
  }
  static int mcp = -1;
  if (mcp < 0) {
    const char* e = getenv("ADAFACE_ATTN_MCP");       // 1: persistent small-CTA kernel for d = 80 (off by default: verified on the attention
    mcp = (e && e[0] == '1') ? 1 : 0;                  // tests only when the round's GPU budget ran out)
  }
  if (mc && !psmem && d == 80 && mcp && Lk > 128) {
    if (wide80()) return launch_ta_mcp<80, 0>(tQw, tKw, tVw, p, ib, ih, stream);
    return launch_ta_mcp<80, 0>(tQ, tK, tV, p, ib, ih, stream);
  }
  if (mc && !psmem && d == 80) {
    if (wide80()) return launch_ta_mc<80, 0>(tQw, tKw, tVw, p, ib, ih, stream);
    return emu ? launch_ta_mc<80, 2>(tQ, tK, tV, p, ib, ih, stream) : launch_ta_mc<80, 0>(tQ, tK, tV, p, ib, ih, stream);
  }
  switch (d) {
    case 40:
      if (psmem) return launch_ta<40, 0, false>(tQ, tK, tV, p, ib, ih, stream);
      switch (emu) {
        case 0: return launch_ta<40, 0, true>(tQ, tK, tV, p, ib, ih, stream);
        case 1: return launch_ta<40, 1, true>(tQ, tK, tV, p, ib, ih, stream);
        case 2: return launch_ta<40, 2, true>(tQ, tK, tV, p, ib, ih, stream);
        case 3: return launch_ta<40, 3, true>(tQ, tK, tV, p, ib, ih, stream);
        default: return launch_ta<40, 4, true>(tQ, tK, tV, p, ib, ih, stream);
      }
    case 80:
      if (psmem) return launch_ta<80, 0, false>(tQ, tK, tV, p, ib, ih, stream);
      return emu ? launch_ta<80, 2, true>(tQ, tK, tV, p, ib, ih, stream) : launch_ta<80, 0, true>(tQ, tK, tV, p, ib, ih, stream);
    case 160:
      if (psmem) return launch_ta<160, 0, false>(tQ, tK, tV, p, ib, ih, stream);
      return launch_ta<160, 0, true>(tQ, tK, tV, p, ib, ih, stream);
  }
  return -1;
}

}  // namespace adaface

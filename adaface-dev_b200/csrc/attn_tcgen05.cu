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

#include "common.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

extern long long g_launch_count;

constexpr int TA_BM = 128;
constexpr int TA_BN = 64;
constexpr int TA_THREADS = 192;

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
  static constexpr int TMEM_COLS = (TMEM_O + DO <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = Q_BYTES + ST * (K_BYTES + V_BYTES) + 2 * P_BYTES + 1024 + 256;
};

struct TaParams {
  bf16* o;
  long long o_sb, o_sn;
  int Lq, Lk;
  float scale_log2;
};

template <int D>
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

  if (warp == 0 && lane == 0) {
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
  } else if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
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
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
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
          const uint64_t da = make_smem_desc_sw128(aP + (j & 1) * Cfg::P_BYTES + k * 32);
          // V tile as the MN-major B operand: 16 keys = two 8-row groups = 2048 B per K step; atoms along d are
          // KV_ATOM bytes apart (LBO).
          const uint64_t db = make_smem_desc_sw128_mn(aV + s * Cfg::V_BYTES + k * 2048, Cfg::KV_ATOM);
          umma_bf16(tmem_base + (uint32_t)Cfg::TMEM_O, da, db, idesc_pv, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[s]);       // K/V stage reusable
        umma_commit(&pv_done[j & 1]);    // O updated, P buffer reusable
      }
      umma_commit(o_full);
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (warps 2..5)
    const int qd = warp & 3;                       // TMEM lane quarter of this warp
    const int row = qd * 32 + lane;                // query row inside the tile
    const uint32_t t_lane = tmem_base + ((uint32_t)(qd * 32) << 16);
    float m_ref = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int bsel = j & 1;
      mbar_wait(&s_full[bsel], (j >> 1) & 1);
      tc_fence_after();
      float x[TA_BN];
#pragma unroll
      for (int c = 0; c < TA_BN / 16; ++c) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_lane + (uint32_t)(bsel * TA_BN + c * 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) x[c * 16 + i] = __uint_as_float(v[i]) * p.scale_log2;
      }
      tc_fence_before();
      mbar_arrive(&s_free[bsel]);                  // S[bsel] may be overwritten by Q K_{j+2}^T
      const int valid = p.Lk - j * TA_BN;          // keys of this tile that exist
      if (valid < TA_BN) {
#pragma unroll
        for (int i = 0; i < TA_BN; ++i)
          if (i >= valid) x[i] = -INFINITY;
      }
      float mx = x[0];
#pragma unroll
      for (int i = 1; i < TA_BN; ++i) mx = fmaxf(mx, x[i]);

      if (j == 0) {
        m_ref = (mx == -INFINITY) ? 0.f : mx;
      } else {
        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);   // O and P[bsel] are quiescent
        tc_fence_after();
        const bool need = mx > m_ref + 8.f;                     // lazy rescale: tolerate P up to 2^8
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? mx : m_ref;
          const float f = fast_exp2(m_ref - m_new);
          m_ref = m_new;
          l_run *= f;
#pragma unroll
          for (int c = 0; c < DO / 16; ++c) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_lane + (uint32_t)(Cfg::TMEM_O + c * 16), v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            tmem_st_32x32b_x16(t_lane + (uint32_t)(Cfg::TMEM_O + c * 16), v);
          }
          tmem_st_wait();
        }
      }
      // P = exp2(x - m_ref), bf16, into the 128B-swizzled K-major tile: 16-byte chunk c of row r lives at
      // r*128 + ((c ^ (r & 7)) * 16).
      uint8_t* prow = sP + bsel * Cfg::P_BYTES + row * 128;
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < TA_BN / 8; ++c) {
        float e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          e[i] = fast_exp2(x[c * 8 + i] - m_ref);
          lsum += e[i];
        }
        uint4 pk = make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7]));
        *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) * 16)) = pk;
      }
      l_run += lsum;
      fence_proxy_async_smem();                    // generic-proxy smem writes -> visible to the tensor core
      tc_fence_before();
      mbar_arrive(&p_full[bsel]);
    }
    // ---- epilogue: O / l -> bf16 -> [B, Lq, H*d]
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
    const int grow = m0 + row;
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
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int D>
static int launch_ta(const CUtensorMap& tQ, const CUtensorMap& tK, const CUtensorMap& tV, const TaParams& p, int B, int H,
                     cudaStream_t stream) {
  using Cfg = TaCfg<D>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "tcgen05 attention: shared memory exceeds the SM");
  static bool configured = false;
  if (!configured) {
    AF_CUDA(cudaFuncSetAttribute(attn_fwd_tcgen05_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  dim3 grid((p.Lq + TA_BM - 1) / TA_BM, H, B);
  attn_fwd_tcgen05_kernel<D><<<grid, TA_THREADS, Cfg::SMEM_BYTES, stream>>>(tQ, tK, tV, p);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

// Unmasked attention on the tensor-core path.  Returns -1 when the problem is not eligible (caller falls through to
// the warp-MMA kernel, which handles masks, causal multi-KV and tiny shapes), 0 on success, > 0 on error.
int attn_fwd_tcgen05(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn, const void* v,
                     int64_t v_sb, int64_t v_sn, void* o, int64_t o_sb, int64_t o_sn, int64_t B, int64_t H, int64_t Lq,
                     int64_t Lk, int64_t d, float scale, cudaStream_t stream) {
  if (!(d == 40 || d == 80 || d == 160)) return -1;
  CUtensorMap tQ, tK, tV;
  if (make_tmap_bf16_heads(&tQ, q, (uint64_t)d, (uint64_t)H, (uint64_t)Lq, (uint64_t)B, (uint64_t)q_sn, (uint64_t)q_sb, TA_BM)) return 3;
  if (make_tmap_bf16_heads(&tK, k, (uint64_t)d, (uint64_t)H, (uint64_t)Lk, (uint64_t)B, (uint64_t)k_sn, (uint64_t)k_sb, TA_BN)) return 3;
  if (make_tmap_bf16_heads(&tV, v, (uint64_t)d, (uint64_t)H, (uint64_t)Lk, (uint64_t)B, (uint64_t)v_sn, (uint64_t)v_sb, TA_BN)) return 3;
  TaParams p;
  p.o = (bf16*)o;
  p.o_sb = o_sb;
  p.o_sn = o_sn;
  p.Lq = (int)Lq;
  p.Lk = (int)Lk;
  p.scale_log2 = scale * 1.4426950408889634f;
  switch (d) {
    case 40: return launch_ta<40>(tQ, tK, tV, p, (int)B, (int)H, stream);
    case 80: return launch_ta<80>(tQ, tK, tV, p, (int)B, (int)H, stream);
    case 160: return launch_ta<160>(tQ, tK, tV, p, (int)B, (int)H, stream);
  }
  return -1;
}

}  // namespace adaface

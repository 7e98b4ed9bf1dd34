// K2: flash-style attention on warp-level tensor-core MMA (mma.sync m16n8k16, bf16 -> fp32).
//
//  * attn_fwd_kernel<D>: online-softmax attention for SD-1.5 head dims 40/80/160 (+64 for the CLIP-shaped
//    encoder): self-attention with the optional img_mask key mask (dalc:254-273), masked cross-attention
//    and CLIPAttentionMKV's causal multi-K/V attention (arc2face_models.py:170-217).
//  (K3, the short-context cross-attention / capture kernel, lives in attn_cross_stream.cu.)
//
// Head dim 40 is padded to 48 only in shared memory (zero columns); HBM layouts stay those of the
// reference: [B, L, H*d] with heads interleaved.
#include <math.h>
#include <stdlib.h>

#include "attn_common.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

extern long long g_launch_count;

struct AttnParams {
  const bf16 *q, *k, *v;
  bf16* o;
  long long q_sb, q_sn, k_sb, k_sn, v_sb, v_sn, o_sb, o_sn;
  int B, H, Lq, Lk;
  const uint8_t* key_mask;
  int causal_mult;
  float scale_log2;
  float* lse;   // optional [B, H, Lq]: log2(sum_j exp2(s_ij)) for the backward pass (+inf for an empty row)
};

template <int D>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(const AttnParams p) {
  using A = AttDims<D>;
  constexpr int LD = A::LD, KT = A::KT, NT_O = A::NT_O, NT_S = ATT_BN / 8;
  extern __shared__ __align__(16) uint8_t smem_att[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_att);
  bf16* sK = sQ + ATT_BM * LD;                // [2][BN][LD]
  bf16* sV = sK + 2 * ATT_BN * LD;            // [2][BN][LD]
  uint8_t* sValid = reinterpret_cast<uint8_t*>(sV + 2 * ATT_BN * LD);   // [2][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.x * ATT_BM, h = blockIdx.y, b = blockIdx.z;
  const bf16* gq = p.q + (long long)b * p.q_sb + h * D;
  const bf16* gk = p.k + (long long)b * p.k_sb + h * D;
  const bf16* gv = p.v + (long long)b * p.v_sb + h * D;

  int n_tiles = (p.Lk + ATT_BN - 1) / ATT_BN;
  if (p.causal_mult > 0) {
    const long long last_key = (long long)min(p.Lq, m0 + ATT_BM) * p.causal_mult;   // exclusive
    n_tiles = min(n_tiles, (int)((last_key + ATT_BN - 1) / ATT_BN));
  }

  auto load_kv = [&](int tile, int buf) {
    load_rows<D>(sK + buf * ATT_BN * LD, gk, p.k_sn, tile * ATT_BN, p.Lk, ATT_BN, p.causal_mult, p.H * D);
    load_rows<D>(sV + buf * ATT_BN * LD, gv, p.v_sn, tile * ATT_BN, p.Lk, ATT_BN, p.causal_mult, p.H * D);
    if (threadIdx.x < ATT_BN) {
      const int j = tile * ATT_BN + threadIdx.x;
      uint8_t ok = j < p.Lk;
      if (ok && p.key_mask) ok = p.key_mask[(long long)b * p.Lk + j] != 0;
      sValid[buf * ATT_BN + threadIdx.x] = ok;
    }
  };

  zero_pad_cols<D>(sQ, ATT_BM);
  zero_pad_cols<D>(sK, 2 * ATT_BN);
  load_rows<D>(sQ, gq, p.q_sn, m0, p.Lq, ATT_BM);
  load_kv(0, 0);
  cp_async_commit();

  uint32_t qf[KT][4];
  float acc_o[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i) acc_o[i][0] = acc_o[i][1] = acc_o[i][2] = acc_o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int r_lo = m0 + warp * 16 + g;   // query row of c0/c1; c2/c3 belong to r_lo + 8

  for (int tile = 0; tile < n_tiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < n_tiles) {
      load_kv(tile + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (tile == 0) {
#pragma unroll
      for (int kk = 0; kk < KT; ++kk)
        ldsm_x4(smem_u32(sQ + (warp * 16 + (lane & 15)) * LD + kk * 16 + (lane >> 4) * 8), qf[kk][0], qf[kk][1],
                qf[kk][2], qf[kk][3]);
    }
    const bf16* tK = sK + buf * ATT_BN * LD;
    const bf16* tV = sV + buf * ATT_BN * LD;
    const uint8_t* tValid = sValid + buf * ATT_BN;

    float acc_s[NT_S][4];
#pragma unroll
    for (int i = 0; i < NT_S; ++i) acc_s[i][0] = acc_s[i][1] = acc_s[i][2] = acc_s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
#pragma unroll
      for (int np = 0; np < NT_S / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(tK + (np * 16 + (lane >> 4) * 8 + (lane & 7)) * LD + kk * 16 + ((lane >> 3) & 1) * 8), b0, b1,
                b2, b3);
        mma_bf16_16816(acc_s[2 * np], qf[kk], b0, b1);
        mma_bf16_16816(acc_s[2 * np + 1], qf[kk], b2, b3);
      }
    }
    // scale + masks, running max
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int cl = nt * 8 + 2 * t + (e & 1);
        const int r = r_lo + (e >> 1) * 8;
        bool ok = tValid[cl] != 0;
        if (p.causal_mult > 0) ok = ok && ((long long)(tile * ATT_BN + cl) < (long long)(r + 1) * p.causal_mult);
        const float s = ok ? acc_s[nt][e] * p.scale_log2 : -INFINITY;
        acc_s[nt][e] = s;
        mx[e >> 1] = fmaxf(mx[e >> 1], s);
      }
    }
    float corr[2], m_use[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
      const float m_new = fmaxf(m_run[i], mx[i]);
      m_use[i] = (m_new == -INFINITY) ? 0.f : m_new;
      corr[i] = fast_exp2(m_run[i] - m_use[i]);   // m_run = -inf -> 0
      m_run[i] = m_new;
      l_run[i] *= corr[i];
    }
#pragma unroll
    for (int nt = 0; nt < NT_S; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = fast_exp2(acc_s[nt][e] - m_use[e >> 1]);
        acc_s[nt][e] = pv;
        l_run[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int i = 0; i < NT_O; ++i) {
      acc_o[i][0] *= corr[0];
      acc_o[i][1] *= corr[0];
      acc_o[i][2] *= corr[1];
      acc_o[i][3] *= corr[1];
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < ATT_BN / 16; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16(acc_s[2 * kk][0], acc_s[2 * kk][1]);
      a[1] = pack_bf16(acc_s[2 * kk][2], acc_s[2 * kk][3]);
      a[2] = pack_bf16(acc_s[2 * kk + 1][0], acc_s[2 * kk + 1][1]);
      a[3] = pack_bf16(acc_s[2 * kk + 1][2], acc_s[2 * kk + 1][3]);
      const bf16* vrow = tV + (kk * 16 + (lane & 15)) * LD;
#pragma unroll
      for (int nt = 0; nt + 1 < NT_O; nt += 2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(smem_u32(vrow + nt * 8 + (lane >> 4) * 8), b0, b1, b2, b3);
        mma_bf16_16816(acc_o[nt], a, b0, b1);
        mma_bf16_16816(acc_o[nt + 1], a, b2, b3);
      }
      if (NT_O & 1) {
        uint32_t b0, b1;
        ldsm_x2_trans(smem_u32(vrow + (NT_O - 1) * 8), b0, b1);
        mma_bf16_16816(acc_o[NT_O - 1], a, b0, b1);
      }
    }
    __syncthreads();   // everyone done with `buf` before the next iteration's prefetch overwrites it
  }

  // finalize: O / l, stage through this warp's rows of sQ, coalesced 16-byte stores
  float inv[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float l = l_run[i];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    inv[i] = l > 0.f ? 1.f / l : 0.f;
    const int r = r_lo + i * 8;
    if (p.lse && t == 0 && r < p.Lq)
      p.lse[((long long)b * p.H + h) * p.Lq + r] = l > 0.f ? m_run[i] + log2f(l) : INFINITY;
  }
  bf16* sO = sQ + warp * 16 * LD;
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < NT_O; ++nt) {
    *reinterpret_cast<uint32_t*>(sO + g * LD + nt * 8 + 2 * t) = pack_bf16(acc_o[nt][0] * inv[0], acc_o[nt][1] * inv[0]);
    *reinterpret_cast<uint32_t*>(sO + (g + 8) * LD + nt * 8 + 2 * t) =
        pack_bf16(acc_o[nt][2] * inv[1], acc_o[nt][3] * inv[1]);
  }
  __syncwarp();
  bf16* go = p.o + (long long)b * p.o_sb + h * D;
  for (int c = lane; c < 16 * A::CH; c += 32) {
    const int r = c / A::CH, ch = c - r * A::CH;
    const int gr = m0 + warp * 16 + r;
    if (gr < p.Lq) *reinterpret_cast<uint4*>(go + (long long)gr * p.o_sn + ch * 8) = *reinterpret_cast<const uint4*>(sO + r * LD + ch * 8);
  }
}

template <int D>
static int launch_attn(const AttnParams& p, cudaStream_t stream) {
  using A = AttDims<D>;
  constexpr int smem = (ATT_BM + 4 * ATT_BN) * A::LD * 2 + 2 * ATT_BN;
  static DevOnce configured;
  const int cfg_dev = af_device();
  if (!configured.done(cfg_dev)) {
    AF_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.set(cfg_dev);
  }
  dim3 grid((p.Lq + ATT_BM - 1) / ATT_BM, p.H, p.B);
  attn_fwd_kernel<D><<<grid, ATT_THREADS, smem, stream>>>(p);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

static int check_view(const char* what, const void* ptr, int64_t sb, int64_t sn, int64_t d) {
  AF_CHECK(ptr != nullptr, "attention: null %s", what);
  AF_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && sb % 8 == 0 && sn % 8 == 0 && d % 8 == 0,
           "attention: %s must be 16-byte aligned with strides / head dim multiples of 8 (sb=%lld sn=%lld d=%lld)", what,
           (long long)sb, (long long)sn, (long long)d);
  return 0;
}

int attn_fwd_tcgen05(const void*, int64_t, int64_t, int64_t, const void*, int64_t, int64_t, int64_t, const void*, int64_t,
                     int64_t, int64_t, void*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                     float, float*, const uint8_t*, cudaStream_t);

static bool legacy_attention_forced() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ADAFACE_ATTN_LEGACY");   // debugging aid: route everything to the warp-MMA kernel
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

int attn_fwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn, const void* v,
             int64_t v_sb, int64_t v_sn, void* o, int64_t o_sb, int64_t o_sn, int64_t B, int64_t H, int64_t Lq,
             int64_t Lk, int64_t d, const uint8_t* key_mask, int causal_mult, float scale, float* lse,
             cudaStream_t stream) {
  if (check_view("q", q, q_sb, q_sn, d) || check_view("k", k, k_sb, k_sn, d) || check_view("v", v, v_sb, v_sn, d) ||
      check_view("o", o, o_sb, o_sn, d))
    return 1;
  AF_CHECK(B > 0 && H > 0 && Lq > 0 && Lk > 0, "attn_fwd: empty problem B=%lld H=%lld Lq=%lld Lk=%lld", (long long)B,
           (long long)H, (long long)Lq, (long long)Lk);
  AF_CHECK(B <= 65535 && H <= 65535, "attn_fwd: B/H exceed grid limits");
  AF_CHECK(causal_mult >= 0, "attn_fwd: causal_mult must be >= 0");
  if (causal_mult == 0 && !legacy_attention_forced()) {
    // tcgen05 / TMEM kernels (attn_tcgen05.cu): unmasked attention, and key masks at d = 40 / long sequences (level A)
    const int rc = attn_fwd_tcgen05(q, q_sb, d, q_sn, k, k_sb, d, k_sn, v, v_sb, d, v_sn, o, o_sb, o_sn, B, H, Lq, Lk, d, d, d, scale, lse, key_mask, stream);
    if (rc >= 0) return rc;
  }
  AttnParams p;
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.o = (bf16*)o;
  p.q_sb = q_sb; p.q_sn = q_sn; p.k_sb = k_sb; p.k_sn = k_sn; p.v_sb = v_sb; p.v_sn = v_sn; p.o_sb = o_sb; p.o_sn = o_sn;
  p.B = (int)B; p.H = (int)H; p.Lq = (int)Lq; p.Lk = (int)Lk;
  p.key_mask = key_mask;
  p.causal_mult = causal_mult;
  p.scale_log2 = scale * LOG2E;
  p.lse = lse;
  switch (d) {
    case 40: return launch_attn<40>(p, stream);
    case 64: return launch_attn<64>(p, stream);
    case 80: return launch_attn<80>(p, stream);
    case 160: return launch_attn<160>(p, stream);
  }
  set_error("attn_fwd: unsupported head dim %lld (supported: 40, 64, 80, 160)", (long long)d);
  return 1;
}

}  // namespace adaface

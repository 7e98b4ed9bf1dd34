// K3: cross-attention over a short context (S <= 128 keys) as an HBM-STREAMING kernel: the fast path of dalc:321 for
// the 77-token prompt and the slow SDPA of dalc:79-139 with capture / normalize / mix.
//
// Roofline: AI = S FLOP/B (77 << the ~250 FLOP/B ridge), so the kernel is bound by reading Q and writing O (+ the
// captured [B,H,Lq,S] fp32 maps).  Design for that:
//   * PERSISTENT CTAs: a CTA owns one (batch, head) [mix: the (sc, mc) pair] and a contiguous chunk of 64-query tiles.
//     K and V (77 x d) are staged ONCE per CTA in shared memory -- converted to bf16 hi (+ lo) parts when q/k/v arrive
//     in fp32 -- together with the per-column normalize terms; the grid is sized to the number of resident CTAs
//     (occupancy query x 148 SMs), so staging is amortised over the whole chunk.
//   * WARP-DECOUPLED pipeline: each of the 4 warps owns 16 query rows of every tile and never synchronises with the
//     others after the staging phase.  The next tile's Q rows are prefetched into REGISTERS (global loads in flight
//     across the whole tile computation) and converted / stored to the warp's private smem slab at the top of the
//     next iteration; with 2-6 CTAs per SM that keeps tens of KB of loads in flight per SM.
//   * exact softmax over the <= 128 keys in registers; probabilities / edited scores / subject columns go through a
//     warp-private fp32 staging slab and leave as one contiguous burst per 16 rows (16-byte vectors when aligned).
//   * fp32 inputs (capture path): q.k as hi.hi + lo.hi + hi.lo (~2^-17 relative), which keeps captured probabilities
//     within 1e-3 of the fp32 reference (plain bf16 q/k: ~3e-3).
// Tensor work is tiny (0.8 GFLOP per sample and layer), so warp-level mma.sync is the right tool; a tcgen05 pipeline
// only adds TMEM / barrier set-up latency per tile here (measured: 31.6 us vs this kernel at B = 8, level A).
#include <math.h>

#include "attn_common.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

extern long long g_launch_count;

struct CapParams {
  const void *q, *k, *v;   // bf16, or fp32 when the kernel is instantiated with F32IN
  bf16* o;
  long long q_sb, q_sn, k_sb, k_sn, v_sb, v_sn, o_sb, o_sn;
  int B, H, Lq, S;
  float scale;
  float* prob;
  float* score;
  float* prob_subj;
  const int32_t* subj_cols;
  int n_subj;
  const uint8_t* col_flag;
  const float* qmean;      // [B, H*d]
  const float* ca_scale;   // device scalar (cross_attn_scale_factor) or null = 1
  int mix;
  int tiles_per_cta;
  // ---- fused capture consumers: reductions of the probability map computed where the map is in registers,
  //      so that the [B,H,Lq,S] fp32 map itself need not be written (prob / score may be NULL)
  const uint8_t* sum_flag;   // [B, S]: columns summed into subj_sum (the instance's subject tokens)
  float* subj_sum;           // [B, H, Lq] = sum over flagged columns of prob      (ldm/util.py:1862-1868, do_sum=True)
  const float* ref_prob;     // [B, H, Lq, S]: probabilities of the reference instance (sc_rep, detached; ldm/util.py:2084-2089)
  float* sq_part;            // [B, H, sq_slots, 4]: per-(CTA, warp) partial sums of (prob - ref_prob)^2; zero-initialised by the caller
  int sq_slots;
};

// NTS = n8 key tiles held in registers (10 -> up to 80 keys, 16 -> up to 128); KROWS = NTS * 8 staged key rows.
template <int D, int NTS, bool MIX, bool F32IN>
__global__ void __launch_bounds__(ATT_THREADS) attn_cross_stream_kernel(const CapParams p) {
  using A = AttDims<D>;
  constexpr int LD = A::LD, KT = A::KT, NT_O = A::NT_O, KROWS = NTS * 8, NI = MIX ? 2 : 1, NP = F32IN ? 2 : 1;
  constexpr int QSLAB = NP * NI * 16 * LD;                       // one warp's private Q slab (elements)
  constexpr int NCH = 16 * (F32IN ? D / 4 : D / 8);              // 16-byte chunks of one warp's 16 Q rows
  constexpr int PER = (NCH + 31) / 32;                           // ... per lane
  constexpr bool PREFETCH = (NI * PER <= 10);                    // register budget; d = 160 fp32 loads in place
  extern __shared__ __align__(16) uint8_t smem_cs[];
  bf16* sK = reinterpret_cast<bf16*>(smem_cs);                   // [NP][NI][KROWS][LD]      (NP: hi / lo parts)
  bf16* sKlo = sK + NI * KROWS * LD;
  bf16* sV = sK + NP * NI * KROWS * LD;                          // [NI][KROWS][LD]
  bf16* sQ = sV + NI * KROWS * LD;                               // [4 warps][NP][NI][16][LD]
  float* sColMean = reinterpret_cast<float*>(sQ + 4 * QSLAB);    // [KROWS]
  uint8_t* sFlag = reinterpret_cast<uint8_t*>(sColMean + KROWS); // [KROWS]
  uint8_t* sSumFlag = sFlag + KROWS;                             // [KROWS]
  float* sStage = reinterpret_cast<float*>(sSumFlag + KROWS);    // [4 warps][16][S] fp32, only when maps are captured

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int chunk = blockIdx.x, h = blockIdx.y, b0 = blockIdx.z;
  const int S = p.S;
  const int nts = (S + 7) / 8;               // n8 tiles that hold real keys
  const int half = p.B / 2;
  const int n_tiles_all = (p.Lq + ATT_BM - 1) / ATT_BM;
  const int tile_begin = chunk * p.tiles_per_cta, tile_end = min(n_tiles_all, tile_begin + p.tiles_per_cta);
  if (tile_begin >= tile_end) return;

  // ---- stage K / V once per CTA
  zero_pad_cols<D>(sK, (NP + 1) * NI * KROWS);
  zero_pad_cols<D>(sQ, 4 * NP * NI * 16);
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int b = b0 + i * half;
    if constexpr (F32IN) {
      load_rows_f32_split<D>(sK + i * KROWS * LD, sKlo + i * KROWS * LD, (const float*)p.k + (long long)b * p.k_sb + h * D,
                             p.k_sn, 0, S, KROWS);
      load_rows_f32_split<D>(sV + i * KROWS * LD, nullptr, (const float*)p.v + (long long)b * p.v_sb + h * D, p.v_sn, 0, S,
                             KROWS);
    } else {
      load_rows<D>(sK + i * KROWS * LD, (const bf16*)p.k + (long long)b * p.k_sb + h * D, p.k_sn, 0, S, KROWS);
      load_rows<D>(sV + i * KROWS * LD, (const bf16*)p.v + (long long)b * p.v_sb + h * D, p.v_sn, 0, S, KROWS);
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  // normalize: per-column mean over the queries = scale * qmean . k_j  (dalc:123-126)
  if (threadIdx.x < KROWS) {
    const int j = threadIdx.x;
    float cm = 0.f;
    uint8_t fl = 0;
    if (!MIX && p.col_flag && j < S && p.col_flag[(long long)b0 * S + j]) {
      fl = 1;
      const float* qm = p.qmean + ((long long)b0 * p.H + h) * D;
      const bf16* kr = sK + j * LD;
      const bf16* kl = sKlo + j * LD;
#pragma unroll 8
      for (int dd = 0; dd < D; ++dd) {
        float kv = __bfloat162float(kr[dd]);
        if constexpr (F32IN) kv += __bfloat162float(kl[dd]);
        cm += qm[dd] * kv;
      }
      cm *= p.scale;
    }
    sColMean[j] = cm;
    sFlag[j] = fl;
    sSumFlag[j] = (!MIX && p.sum_flag && j < S && p.sum_flag[(long long)b0 * S + j]) ? 1 : 0;
  }
  __syncthreads();        // the last block-wide barrier: from here on every warp runs its own pipeline

  bf16* myQ = sQ + warp * QSLAB;             // [NP][NI][16][LD]
  bf16* myQlo = myQ + NI * 16 * LD;
  float* st = sStage + warp * 16 * S;
  const float sc = MIX ? 0.5f * p.scale : p.scale;   // (score_sc + score_mc) / 2, dalc:117
  const float ca_scale = p.ca_scale ? __ldg(p.ca_scale) : 1.f;

  // normalize flags of this thread's 2 * NTS score columns as one bitmask (bit 2 nt + (e & 1))
  uint32_t flag_bits = 0;
#pragma unroll
  for (int nt = 0; nt < NTS; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e)
      if (sFlag[nt * 8 + 2 * t + e]) flag_bits |= 1u << (2 * nt + e);
  }

  uint32_t sum_bits = 0;      // columns of this thread that belong to the subject-column sum
#pragma unroll
  for (int nt = 0; nt < NTS; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e)
      if (sSumFlag[nt * 8 + 2 * t + e]) sum_bits |= 1u << (2 * nt + e);
  }
  float sq_acc = 0.f;         // running sum of (prob - ref_prob)^2 over this warp's rows of every tile of the CTA

  // ---- Q rows of one tile: global -> registers (prefetch) -> bf16 hi/lo slab
  uint4 qreg[NI][PER];
  auto q_row_ptr = [&](int i, int row) -> const uint8_t* {
    const long long b = b0 + i * half;
    return reinterpret_cast<const uint8_t*>(p.q) + ((b * p.q_sb + (long long)row * p.q_sn + h * D) << (F32IN ? 2 : 1));
  };
  auto q_fetch = [&](int row0) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
#pragma unroll
      for (int x = 0; x < PER; ++x) {
        const int c = lane + 32 * x;
        constexpr int CPR = NCH / 16;        // chunks per row
        const int r = c / CPR, cc = c - r * CPR;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (c < NCH && row0 + r < p.Lq) v = __ldg(reinterpret_cast<const uint4*>(q_row_ptr(i, row0 + r)) + cc);
        qreg[i][x] = v;
      }
    }
  };
  auto q_commit = [&]() {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
#pragma unroll
      for (int x = 0; x < PER; ++x) {
        const int c = lane + 32 * x;
        constexpr int CPR = NCH / 16;
        const int r = c / CPR, cc = c - r * CPR;
        if (c < NCH) {
          const uint4 v = qreg[i][x];
          if constexpr (F32IN) {
            const float4 f = *reinterpret_cast<const float4*>(&v);
            const __nv_bfloat162 h01 = __floats2bfloat162_rn(f.x, f.y), h23 = __floats2bfloat162_rn(f.z, f.w);
            uint2 hv, lv;
            hv.x = *reinterpret_cast<const uint32_t*>(&h01);
            hv.y = *reinterpret_cast<const uint32_t*>(&h23);
            lv.x = pack_bf16(f.x - __bfloat162float(h01.x), f.y - __bfloat162float(h01.y));
            lv.y = pack_bf16(f.z - __bfloat162float(h23.x), f.w - __bfloat162float(h23.y));
            *reinterpret_cast<uint2*>(myQ + (i * 16 + r) * LD + cc * 4) = hv;
            *reinterpret_cast<uint2*>(myQlo + (i * 16 + r) * LD + cc * 4) = lv;
          } else {
            *reinterpret_cast<uint4*>(myQ + (i * 16 + r) * LD + cc * 8) = v;
          }
        }
      }
    }
  };

  if constexpr (PREFETCH) q_fetch(tile_begin * ATT_BM + warp * 16);

  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int row0 = tile * ATT_BM + warp * 16;
    const int nrows = max(0, min(16, p.Lq - row0));
    __syncwarp();                            // previous iteration's reads of the slab (O staging) are done
    if constexpr (PREFETCH) {
      q_commit();
      if (tile + 1 < tile_end) q_fetch(row0 + ATT_BM);       // in flight during this tile's math and stores
    } else {
      // in-place load (registers could not hold a prefetched tile): same conversion, no overlap
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        for (int c = lane; c < NCH; c += 32) {
          constexpr int CPR = NCH / 16;
          const int r = c / CPR, cc = c - r * CPR;
          uint4 v = make_uint4(0, 0, 0, 0);
          if (row0 + r < p.Lq) v = __ldg(reinterpret_cast<const uint4*>(q_row_ptr(i, row0 + r)) + cc);
          if constexpr (F32IN) {
            const float4 f = *reinterpret_cast<const float4*>(&v);
            const __nv_bfloat162 h01 = __floats2bfloat162_rn(f.x, f.y), h23 = __floats2bfloat162_rn(f.z, f.w);
            uint2 hv, lv;
            hv.x = *reinterpret_cast<const uint32_t*>(&h01);
            hv.y = *reinterpret_cast<const uint32_t*>(&h23);
            lv.x = pack_bf16(f.x - __bfloat162float(h01.x), f.y - __bfloat162float(h01.y));
            lv.y = pack_bf16(f.z - __bfloat162float(h23.x), f.w - __bfloat162float(h23.y));
            *reinterpret_cast<uint2*>(myQ + (i * 16 + r) * LD + cc * 4) = hv;
            *reinterpret_cast<uint2*>(myQlo + (i * 16 + r) * LD + cc * 4) = lv;
          } else {
            *reinterpret_cast<uint4*>(myQ + (i * 16 + r) * LD + cc * 8) = v;
          }
        }
      }
    }
    __syncwarp();
    if (nrows == 0) continue;

    // ---- scores
    float acc_s[NTS][4];
#pragma unroll
    for (int i = 0; i < NTS; ++i) acc_s[i][0] = acc_s[i][1] = acc_s[i][2] = acc_s[i][3] = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) {
        const int q_off = (i * 16 + (lane & 15)) * LD + kk * 16 + (lane >> 4) * 8;
        uint32_t qf[4], ql[4];
        ldsm_x4(smem_u32(myQ + q_off), qf[0], qf[1], qf[2], qf[3]);
        if constexpr (F32IN) ldsm_x4(smem_u32(myQlo + q_off), ql[0], ql[1], ql[2], ql[3]);
        // three passes (hi.hi, lo.hi, hi.lo) over ALL key tiles: consecutive MMAs hit different accumulators, so the
        // tensor pipe never waits on the previous MMA's result
#pragma unroll
        for (int pass = 0; pass < (F32IN ? 3 : 1); ++pass) {
          const bf16* kbase = pass == 2 ? sKlo : sK;
          const uint32_t (&af)[4] = pass == 1 ? ql : qf;
#pragma unroll
          for (int np = 0; np < NTS / 2; ++np) {
            if (2 * np < nts) {
              const int k_off = i * KROWS * LD + (np * 16 + (lane >> 4) * 8 + (lane & 7)) * LD + kk * 16 + ((lane >> 3) & 1) * 8;
              uint32_t r0, r1, r2, r3;
              ldsm_x4(smem_u32(kbase + k_off), r0, r1, r2, r3);
              mma_bf16_16816(acc_s[2 * np], af, r0, r1);
              mma_bf16_16816(acc_s[2 * np + 1], af, r2, r3);
            }
          }
        }
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NTS; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = nt * 8 + 2 * t + (e & 1);
        float s = -INFINITY;
        if (c < S) {
          s = acc_s[nt][e] * sc;
          if (flag_bits & (1u << (2 * nt + (e & 1)))) s = (s - sColMean[c]) * ca_scale;   // dalc:126-130
        }
        acc_s[nt][e] = s;
        mx[e >> 1] = fmaxf(mx[e >> 1], s);
      }
    }

    auto stage_and_store = [&](float* gbase, bool with_subj) {
      // registers -> smem [16][S] -> one contiguous global burst per instance (16-byte vectors: 16 * S floats start
      // at a multiple of 64 bytes)
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = nt * 8 + 2 * t + (e & 1);
          if (c < S) st[(g + (e >> 1) * 8) * S + c] = acc_s[nt][e];
        }
      }
      __syncwarp();
      for (int i = 0; i < NI; ++i) {
        const int b = b0 + i * half;
        if (gbase) {
          float* dst = gbase + (((long long)b * p.H + h) * p.Lq + row0) * S;
          const int n = nrows * S;
          if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            for (int x = lane; x < n / 4; x += 32) reinterpret_cast<float4*>(dst)[x] = reinterpret_cast<const float4*>(st)[x];
          } else {
            for (int x = lane; x < n; x += 32) dst[x] = st[x];
          }
        }
        if (with_subj && p.prob_subj) {
          float* dst = p.prob_subj + (((long long)b * p.H + h) * p.Lq + row0) * p.n_subj;
          const int32_t* cols = p.subj_cols + (long long)b * p.n_subj;
          for (int x = lane; x < nrows * p.n_subj; x += 32) {
            const int r = x / p.n_subj, jj = x - r * p.n_subj;
            const int c = cols[jj];
            dst[x] = (c >= 0 && c < S) ? st[r * S + c] : 0.f;
          }
        }
      }
    };
    if (p.score) stage_and_store(p.score, false);   // edited score (dalc:139, quirk 3)

    // ---- exact softmax over the S keys
    float inv[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < NTS; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = exp2f((acc_s[nt][e] - mx[e >> 1]) * LOG2E);
        acc_s[nt][e] = pv;
        sum[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 1);
      sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 2);
      inv[i] = 1.f / sum[i];
    }
#pragma unroll
    for (int nt = 0; nt < NTS; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) acc_s[nt][e] *= inv[e >> 1];
    }
    if (p.prob || p.prob_subj) stage_and_store(p.prob, true);
    if constexpr (!MIX) {
      // ---- fused consumers: the probabilities are in registers -- reduce them here instead of writing the map
      if (p.subj_sum) {
        float ss[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < NTS; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (sum_bits & (1u << (2 * nt + (e & 1)))) ss[e >> 1] += acc_s[nt][e];
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          ss[i] += __shfl_xor_sync(0xffffffffu, ss[i], 1);
          ss[i] += __shfl_xor_sync(0xffffffffu, ss[i], 2);
        }
        if (t == 0) {
          float* dst = p.subj_sum + ((long long)b0 * p.H + h) * p.Lq + row0;
          if (g < nrows) dst[g] = ss[0];
          if (g + 8 < nrows) dst[g + 8] = ss[1];
        }
      }
      if (p.ref_prob) {
        // the reference rows of this warp tile are one contiguous block of nrows * S floats: bring it in with coalesced
        // 16-byte loads through the warp's staging slab (per-thread 4-byte loads at stride S ran at 1/3 of the speed)
        const float* ref = p.ref_prob + (((long long)b0 * p.H + h) * p.Lq + row0) * S;
        const int n = nrows * S;
        __syncwarp();                          // the slab's previous contents (prob staging) have been stored
        if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(ref) & 15) == 0) {
          for (int x = lane; x < n / 4; x += 32) reinterpret_cast<float4*>(st)[x] = __ldg(reinterpret_cast<const float4*>(ref) + x);
        } else {
          for (int x = lane; x < n; x += 32) st[x] = __ldg(ref + x);
        }
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < NTS; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = nt * 8 + 2 * t + (e & 1), r = g + (e >> 1) * 8;
            if (c < S && r < nrows) {
              const float df = acc_s[nt][e] - st[r * S + c];
              sq_acc += df * df;
            }
          }
        }
      }
    }

    // ---- O = P V per instance
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      float acc_o[NT_O][4];
#pragma unroll
      for (int x = 0; x < NT_O; ++x) acc_o[x][0] = acc_o[x][1] = acc_o[x][2] = acc_o[x][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < NTS / 2; ++kk) {
        if (2 * kk < nts) {
          uint32_t a[4];
          a[0] = pack_bf16(acc_s[2 * kk][0], acc_s[2 * kk][1]);
          a[1] = pack_bf16(acc_s[2 * kk][2], acc_s[2 * kk][3]);
          a[2] = pack_bf16(acc_s[2 * kk + 1][0], acc_s[2 * kk + 1][1]);
          a[3] = pack_bf16(acc_s[2 * kk + 1][2], acc_s[2 * kk + 1][3]);
          const bf16* vrow = sV + i * KROWS * LD + (kk * 16 + (lane & 15)) * LD;
#pragma unroll
          for (int nt = 0; nt + 1 < NT_O; nt += 2) {
            uint32_t r0, r1, r2, r3;
            ldsm_x4_trans(smem_u32(vrow + nt * 8 + (lane >> 4) * 8), r0, r1, r2, r3);
            mma_bf16_16816(acc_o[nt], a, r0, r1);
            mma_bf16_16816(acc_o[nt + 1], a, r2, r3);
          }
          if (NT_O & 1) {
            uint32_t r0, r1;
            ldsm_x2_trans(smem_u32(vrow + (NT_O - 1) * 8), r0, r1);
            mma_bf16_16816(acc_o[NT_O - 1], a, r0, r1);
          }
        }
      }
      bf16* sO = myQ + i * 16 * LD;            // this instance's Q rows: no longer needed
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < NT_O; ++nt) {
        *reinterpret_cast<uint32_t*>(sO + g * LD + nt * 8 + 2 * t) = pack_bf16(acc_o[nt][0], acc_o[nt][1]);
        *reinterpret_cast<uint32_t*>(sO + (g + 8) * LD + nt * 8 + 2 * t) = pack_bf16(acc_o[nt][2], acc_o[nt][3]);
      }
      __syncwarp();
      const int b = b0 + i * half;
      bf16* go = p.o + (long long)b * p.o_sb + h * D;
      for (int c = lane; c < nrows * A::CH; c += 32) {
        const int r = c / A::CH, ch = c - r * A::CH;
        *reinterpret_cast<uint4*>(go + (long long)(row0 + r) * p.o_sn + ch * 8) = *reinterpret_cast<const uint4*>(sO + r * LD + ch * 8);
      }
    }
  }
  if constexpr (!MIX) {
    if (p.sq_part) {          // deterministic: one slot per (batch, head, chunk, warp), summed by the caller in a fixed order
      sq_acc = warp_sum(sq_acc);
      if (lane == 0 && chunk < p.sq_slots) p.sq_part[((((long long)b0 * p.H + h) * p.sq_slots) + chunk) * 4 + warp] = sq_acc;
    }
  }
}

template <int D, int NTS, bool MIX, bool F32IN>
static int launch_cs(CapParams& p, cudaStream_t stream) {
  using A = AttDims<D>;
  constexpr int NI = MIX ? 2 : 1, NP = F32IN ? 2 : 1, KROWS = NTS * 8;
  const bool maps = p.prob || p.score || p.prob_subj || p.ref_prob;      // needs the per-warp fp32 staging slab
  const int smem = ((NP + 1) * NI * KROWS + 4 * NP * NI * 16) * A::LD * 2 + KROWS * 4 + 2 * KROWS + (maps ? 4 * 16 * p.S * 4 : 0) + 16;
  AF_CHECK(smem <= 227 * 1024, "attn_cross_stream: %d bytes of shared memory exceed the SM", smem);
  static int configured_smem[AF_MAX_DEV] = {0}, ctas_per_sm[AF_MAX_DEV][2] = {{0, 0}};      // per device
  const int cfg_dev = af_device();
  auto kern = attn_cross_stream_kernel<D, NTS, MIX, F32IN>;
  if (smem > configured_smem[cfg_dev]) {
    AF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem[cfg_dev] = smem;
  }
  int& occ = ctas_per_sm[cfg_dev][maps ? 1 : 0];
  if (occ == 0) {
    AF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, ATT_THREADS, smem));
    if (occ < 1) occ = 1;
  }
  // one wave of resident CTAs: chunks per (batch, head) so that the grid just fills the machine
  const int nb = MIX ? p.B / 2 : p.B;
  const int tiles = (p.Lq + ATT_BM - 1) / ATT_BM;
  int chunks = (af_num_sms() * occ + nb * p.H - 1) / (nb * p.H);
  if (chunks > tiles) chunks = tiles;
  if (chunks < 1) chunks = 1;
  p.tiles_per_cta = (tiles + chunks - 1) / chunks;
  chunks = (tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  AF_CHECK(!p.sq_part || chunks <= p.sq_slots, "attn_cross_stream: sq_part has %d slots per (batch, head), %d needed", p.sq_slots, chunks);
  kern<<<dim3(chunks, p.H, nb), ATT_THREADS, smem, stream>>>(p);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

template <int D, bool MIX, bool F32IN>
static int launch_cs_keys(CapParams& p, cudaStream_t stream) {
  return p.S <= 80 ? launch_cs<D, 10, MIX, F32IN>(p, stream) : launch_cs<D, 16, MIX, F32IN>(p, stream);
}

static int check_view_cs(const char* what, const void* ptr, int64_t sb, int64_t sn, int64_t d, int64_t al) {
  AF_CHECK(ptr != nullptr, "attention: null %s", what);
  AF_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && sb % al == 0 && sn % al == 0 && d % 8 == 0,
           "attention: %s must be 16-byte aligned with strides multiples of %lld elements (sb=%lld sn=%lld d=%lld)", what,
           (long long)al, (long long)sb, (long long)sn, (long long)d);
  return 0;
}

int attn_cross_capture_fwd(const void* q, int64_t q_sb, int64_t q_sn, const void* k, int64_t k_sb, int64_t k_sn,
                           const void* v, int64_t v_sb, int64_t v_sn, void* o, int64_t o_sb, int64_t o_sn, int64_t B,
                           int64_t H, int64_t Lq, int64_t S, int64_t d, float scale, float* prob, float* score,
                           float* prob_subj, const int32_t* subj_cols, int64_t n_subj, const uint8_t* col_flag,
                           const float* qmean, const float* ca_scale, int mix, int in_dtype, cudaStream_t stream,
                           const uint8_t* sum_flag, float* subj_sum, const float* ref_prob, float* sq_part, int64_t sq_slots) {
  AF_CHECK(in_dtype == ADAFACE_BF16 || in_dtype == ADAFACE_F32, "attn_cross_capture_fwd: bad in_dtype %d", in_dtype);
  AF_CHECK(!(mix && (subj_sum || ref_prob)), "attn_cross_capture_fwd: the fused consumers are not defined for mix_attn_mats_in_batch");
  AF_CHECK(!subj_sum == !sum_flag, "attn_cross_capture_fwd: subj_sum and sum_flag go together");
  AF_CHECK(!ref_prob == !sq_part && (!sq_part || sq_slots > 0), "attn_cross_capture_fwd: ref_prob, sq_part and sq_slots go together");
  const int64_t al = in_dtype == ADAFACE_F32 ? 4 : 8;
  if (check_view_cs("q", q, q_sb, q_sn, d, al) || check_view_cs("k", k, k_sb, k_sn, d, al) ||
      check_view_cs("v", v, v_sb, v_sn, d, al) || check_view_cs("o", o, o_sb, o_sn, d, 8))
    return 1;
  AF_CHECK(B > 0 && H > 0 && Lq > 0 && S > 0, "attn_cross_capture_fwd: empty problem");
  AF_CHECK(S <= 128, "attn_cross_capture_fwd: context length %lld exceeds 128 keys", (long long)S);
  AF_CHECK(B <= 65535 && H <= 65535, "attn_cross_capture_fwd: B/H exceed grid limits");
  AF_CHECK(!(mix && (B % 2)), "mix_attn_mats_in_batch needs an even batch [sc.., mc..] (dalc:113), got B=%lld",
           (long long)B);
  AF_CHECK(!(mix && col_flag), "normalize and mix are mutually exclusive (dalc:108-119: mix wins)");
  AF_CHECK(!(col_flag && !qmean), "normalize_cross_attn needs qmean (and subj_indices, dalc:120)");
  AF_CHECK(!(prob_subj && (!subj_cols || n_subj <= 0)), "prob_subj needs subj_cols / n_subj");
  CapParams p;
  p.q = q; p.k = k; p.v = v; p.o = (bf16*)o;
  p.q_sb = q_sb; p.q_sn = q_sn; p.k_sb = k_sb; p.k_sn = k_sn; p.v_sb = v_sb; p.v_sn = v_sn; p.o_sb = o_sb; p.o_sn = o_sn;
  p.B = (int)B; p.H = (int)H; p.Lq = (int)Lq; p.S = (int)S;
  p.scale = scale;
  p.prob = prob; p.score = score; p.prob_subj = prob_subj; p.subj_cols = subj_cols; p.n_subj = (int)n_subj;
  p.col_flag = col_flag; p.qmean = qmean; p.ca_scale = ca_scale; p.mix = mix;
  p.tiles_per_cta = 1;
  p.sum_flag = sum_flag; p.subj_sum = subj_sum; p.ref_prob = ref_prob; p.sq_part = sq_part; p.sq_slots = (int)sq_slots;
  const bool f32 = in_dtype == ADAFACE_F32;
  if (mix) {
    switch (d) {
      case 40: return f32 ? launch_cs_keys<40, true, true>(p, stream) : launch_cs_keys<40, true, false>(p, stream);
      case 80: return f32 ? launch_cs_keys<80, true, true>(p, stream) : launch_cs_keys<80, true, false>(p, stream);
    }
    set_error("attn_cross_capture_fwd: mix supports head dims 40 and 80, got %lld", (long long)d);
    return 1;
  }
  switch (d) {
    case 40: return f32 ? launch_cs_keys<40, false, true>(p, stream) : launch_cs_keys<40, false, false>(p, stream);
    case 80: return f32 ? launch_cs_keys<80, false, true>(p, stream) : launch_cs_keys<80, false, false>(p, stream);
    case 160: return f32 ? launch_cs_keys<160, false, true>(p, stream) : launch_cs_keys<160, false, false>(p, stream);
  }
  set_error("attn_cross_capture_fwd: unsupported head dim %lld (supported: 40, 80, 160)", (long long)d);
  return 1;
}

}  // namespace adaface

// Shared pieces of the warp-MMA attention kernels (forward: attn_mma.cu, backward: attn_bwd_mma.cu).
#pragma once
#include "common.cuh"

namespace adaface {

constexpr int ATT_BM = 64;       // queries per CTA (4 warps x 16 rows)
constexpr int ATT_BN = 64;       // keys per pipeline stage
constexpr int ATT_THREADS = 128;
constexpr float LOG2E = 1.4426950408889634f;

template <int D>
struct AttDims {
  static constexpr int DP = (D + 15) / 16 * 16;   // MMA-K padded head dim
  static constexpr int LD = DP + 8;               // smem row pitch (elements): conflict-free ldmatrix
  static constexpr int KT = DP / 16;              // k16 steps of Q.K^T
  static constexpr int NT_O = D / 8;              // n8 tiles of the output
  static constexpr int CH = D / 8;                // 16-byte chunks per row in HBM
};

// rows [row0, row0+rows) of a [L, H*d] head slice -> smem tile (zero fill beyond L)
// mult > 1: row j is sub-key (j % mult) of token (j / mult), the sub-keys of a token being `sub` elements apart.
template <int D>
__device__ __forceinline__ void load_rows(bf16* s, const bf16* g, long long stride_n, int row0, int L, int rows,
                                          int mult = 1, int sub = 0) {
  constexpr int CH = AttDims<D>::CH, LD = AttDims<D>::LD;
  for (int c = threadIdx.x; c < rows * CH; c += blockDim.x) {
    const int r = c / CH, ch = c - r * CH;
    const int gr = row0 + r;
    const bool ok = gr < L;
    const int j = ok ? gr : 0;
    const long long off = mult > 1 ? (long long)(j / mult) * stride_n + (long long)(j % mult) * sub : (long long)j * stride_n;
    cp_async_16(s + r * LD + ch * 8, g + off + ch * 8, ok);
  }
}
template <int D>
__device__ __forceinline__ void zero_pad_cols(bf16* s, int rows) {
  constexpr int DP = AttDims<D>::DP, LD = AttDims<D>::LD;
  if (DP == D) return;
  for (int r = threadIdx.x; r < rows; r += blockDim.x)
    *reinterpret_cast<uint4*>(s + r * LD + D) = make_uint4(0, 0, 0, 0);   // DP - D == 8 elements
}

// fp32 rows -> bf16 hi (+ lo = bf16(x - hi)) tiles: q.k is then evaluated as hi.hi + lo.hi + hi.lo, i.e. to ~2^-17
// relative, so that captured probabilities meet the 1e-3 bar (plain bf16 q/k give ~3e-3).
template <int D>
__device__ __forceinline__ void load_rows_f32_split(bf16* s_hi, bf16* s_lo, const float* g, long long stride_n, int row0,
                                                    int L, int rows) {
  constexpr int LD = AttDims<D>::LD, C4 = D / 4;
  for (int c = threadIdx.x; c < rows * C4; c += blockDim.x) {
    const int r = c / C4, c4 = c - r * C4;
    const int gr = row0 + r;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < L) x = *reinterpret_cast<const float4*>(g + (long long)gr * stride_n + c4 * 4);
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(x.x, x.y), h23 = __floats2bfloat162_rn(x.z, x.w);
    uint2 hv;
    hv.x = *reinterpret_cast<const uint32_t*>(&h01);
    hv.y = *reinterpret_cast<const uint32_t*>(&h23);
    *reinterpret_cast<uint2*>(s_hi + r * LD + c4 * 4) = hv;
    if (s_lo) {
      uint2 lv;
      lv.x = pack_bf16(x.x - __bfloat162float(h01.x), x.y - __bfloat162float(h01.y));
      lv.y = pack_bf16(x.z - __bfloat162float(h23.x), x.w - __bfloat162float(h23.y));
      *reinterpret_cast<uint2*>(s_lo + r * LD + c4 * 4) = lv;
    }
  }
}

}  // namespace adaface

// Parameters shared by the tcgen05 attention kernels (attn_tcgen05.cu, attn_tcgen05_q48.cu).
#pragma once
#include "common.cuh"

namespace adaface {

struct TaParams {
  bf16* o;
  long long o_sb, o_sn;
  int Lq, Lk;
  float scale_log2;
  float* lse;              // optional [B, H, Lq]: log2-domain log-sum-exp of each row, kept for the backward pass
  const uint8_t* key_mask; // optional [B, Lk]: 0 = key masked out (four-tile kernel only; every row keeps >= 1 key)
  int wide;                // small-CTA kernel, d = 80: the Q / K / V maps span whole [B, L, H*d] rows and head h's box starts at column h*d
  long long* trace;        // diagnosis only (env ADAFACE_ATTN_TRACE): CTA 0 of the four-tile kernel stamps its hand-offs
};

}  // namespace adaface

// K1: projection GEMM on tcgen05 / TMEM, operands fed by TMA, LoRA/DoRA folded into the same tile.
//
//   Y[M,N] = act( colscale[N] * ( X[M,K] W[N,K]^T + T[M,R] Bs[N,R]^T ) + bias[N] ) + residual[M,N]
//
// Persistent, warp-specialised kernel: 128 x BN output tiles, STAGES-deep TMA ring, two TMEM accumulators so the
// epilogue of one tile overlaps the main loop of the next (details at the kernel).
// Reference arithmetic: nn.Linear / peft lora.Linear(+DoRA) call sites dalc:235-249, 280-288, 328-331
// (SURVEY.md 8a rows A1, A4); ldm/modules/attention.py:31-58, 156-164; HF CLIP q/k/v/out_proj, fc1, fc2.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/adaface_b200.h"

namespace adaface {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;          // 64 bf16 = 128 B = one swizzle span
constexpr int GEMM_THREADS = 352;     // 8 epilogue warps + TMA warp + MMA warp + store warp
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
  // implicit-GEMM 3x3 convolution (CONV instantiation only): the A operand is the NHWC activation itself, one TMA box
  // of `cv_tile_rows` output pixels (whole image rows) per (tap, 64-channel chunk), shifted by the tap offset; the zero
  // padding of the convolution is TMA's out-of-bounds fill.
  int cv_kc;          // 64-channel chunks per tap (K loop = 9 taps x cv_kc)
  int cv_W, cv_HW;    // OUTPUT width, pixels per output image
  int cv_img_rows;    // output pixels of one image covered by a tile (W * image rows per tile)
  int cv_tpi;         // tiles per image (cv_imgs == 1)
  int cv_imgs;        // images per tile (> 1 only when a whole image is smaller than 128 pixels)
  int cv_stride;      // 1 | 2
  int num_m;          // number of M tiles (GEMM: ceil(M / 128))
  const float* rowbias;   // fp32 [images, N] added per output image (the ResBlock's time-embedding term), or NULL
  // split-K (CONV only): a small-M convolution (levels C / D: 4-16 M tiles, K up to 23040) would leave most SMs idle and
  // stream its K range through a handful of CTAs; `splits` CTAs share one output tile instead, each accumulating
  // `kb_per_split` K blocks and storing its raw fp32 partial tile to splitk_ws[split][M][N]; conv_splitk_reduce_kernel
  // then sums the partials in fixed order and applies the epilogue terms.
  int splits, kb_per_split;
  float* splitk_ws;
  int dbg;            // diagnosis only (env ADAFACE_GEMM_DBG): 1 = epilogue skips the global stores, 2 = epilogue skips everything
  int tma_store;      // FULL bf16 chunks leave through TMA bulk stores issued by the store warp (tmY: plain row-major output)
  long long* trace;   // diagnosis only (env ADAFACE_GEMM_TRACE): CTA 0 records clock64 stamps of its producer / MMA / epilogue hand-offs
};

// Tensor maps of the CONV A operand: one for stride 1; one per input parity (py, px) for stride 2.
struct ConvMaps {
  CUtensorMap m[4];
};
struct NoConvMaps {};
template <bool CONV>
struct ConvArg {
  using type = NoConvMaps;
};
template <>
struct ConvArg<true> {
  using type = ConvMaps;
};

// BRES = W-STATIONARY schedule for short-K projections (K <= 320: level A).  The operand feed, not the tensor pipe, bounds these
// GEMMs: a 128 x BN tile pulls (128 + BN) x 128 bytes per 64-wide K block out of L2 for 2 BN tensor cycles (BN = 192: 107 B/clk per
// SM, ~30 TB/s over the chip; traced: the MMA thread waits on the full barriers, 470 clk per K block against the 384 clk floor, 680
// with the epilogue's traffic on top).  Here a CTA is pinned to ONE column block: its [BN, K] weight slice (<= 5 swizzled slabs) is
// loaded once and stays in shared memory, the ring carries only A tiles (16 KB per K block) and the CTA walks down the M tiles of
// its column block: 43 B/clk per SM at BN = 192.
constexpr int BRES_KB = 5;            // resident K blocks (K <= 320)
template <int BN, bool BRES = false>
struct GemmCfg {
  static constexpr int B_STAGE_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = BRES ? A_STAGE_BYTES : A_STAGE_BYTES + B_STAGE_BYTES;      // ring stage
  static constexpr int RESIDENT_BYTES = BRES ? BRES_KB * B_STAGE_BYTES : 0;
  static constexpr int ST_NS = (BN == 192 && !BRES) ? 4 : 2;   // output slabs (128 rows x 32 columns bf16, 8 KB) per epilogue half: what the ring leaves free
  static constexpr int STORE_STAGE_BYTES = 2 * ST_NS * 8192;   // (>= 4 KB per warp for the non-TMA path)
  static constexpr int RING_BUDGET = (BRES ? 227 * 1024 - 1536 - STORE_STAGE_BYTES - RESIDENT_BYTES : 192 * 1024);
  static constexpr int STAGES = (RING_BUDGET / STAGE_BYTES) > 8 ? 8 : (RING_BUDGET / STAGE_BYTES);   // 64:8 128:6 160:5 192:4 256:4; BRES 128:7 160:5 192:4
  static constexpr int TMEM_COLS = 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;   // two accumulator buffers
  static constexpr int SMEM_BYTES = RESIDENT_BYTES + STAGES * STAGE_BYTES + STORE_STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers*/;
};

// Exact (erf) GELU, F.gelu's default (ldm/modules/attention.py:31-41 GEGLU).  The GEGLU epilogue evaluates it 32 times per chunk and
// is instruction-bound (erff(): ~30 instructions and a MUFU each; an Abramowitz-Stegun form with a reciprocal and an exp2 was
// SLOWER: two quarter-rate MUFU ops per element).  Here erf(z), z = |x| / sqrt 2 clamped to 3, is an odd degree-17 polynomial
// (least-squares fit on Chebyshev nodes, max |error| 2e-5 in fp32 -- two orders below the bf16 rounding of the product that
// follows): 8 FMAs on the FMA pipe, no MUFU, no branch.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fminf(fabsf(x) * 0.70710678118654752f, 3.0f);
  const float u = z * z;
  float p = fmaf(3.912539625616773e-08f, u, -1.883036475192057e-06f);
  p = fmaf(p, u, 4.008835821878165e-05f);
  p = fmaf(p, u, -0.000502921873703599f);
  p = fmaf(p, u, 0.004196857567876577f);
  p = fmaf(p, u, -0.024998901411890984f);
  p = fmaf(p, u, 0.11093080043792725f);
  p = fmaf(p, u, -0.3752196431159973f);
  p = fmaf(p, u, 1.1282505989074707f);
  const float erf_abs = fminf(p * z, 1.f);                    // erf(|x| / sqrt 2)
  const float h = 0.5f * x;
  return fmaf(fabsf(h), erf_abs, h);                          // 0.5 x (1 + sign(x) erf(|x| / sqrt 2))
}

// Epilogue of one 32-column chunk of one row: scale / bias / activation / residual / convert / store.
// `c0` = first tile-local column of the chunk, `g` = the matching gate values when ACT == GEGLU.
// FULL = the whole chunk lies inside N: no per-element bounds predicates (the common case; the epilogue warps run
// one per scheduler slot with little ILP, so every instruction removed here is ~4 cycles of the critical path).
// bf16 stores of a FULL chunk go through a warp-private shared-memory slab: a thread owns one ROW of the accumulator
// (TMEM lane), so storing straight from registers makes every warp-level store touch 32 rows x 16 bytes = 32
// half-filled sectors (ncu: 32 sectors / request, 2x the output bytes in L1 sector writes, l1tex the busiest unit).
// After the XOR-swizzled transposition each store instruction writes 8 rows x 64 contiguous bytes = 16 full sectors.
__device__ __forceinline__ void stage_row64(uint8_t* slab, int lane, const float (&f)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(slab + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
        make_uint4(pack_bf16(f[j * 8 + 0], f[j * 8 + 1]), pack_bf16(f[j * 8 + 2], f[j * 8 + 3]), pack_bf16(f[j * 8 + 4], f[j * 8 + 5]),
                   pack_bf16(f[j * 8 + 6], f[j * 8 + 7]));
  __syncwarp();
}
__device__ __forceinline__ uint4 unstage_piece(const uint8_t* slab, int r, int pc) {
  return *reinterpret_cast<const uint4*>(slab + r * 64 + ((pc ^ ((r >> 1) & 3)) << 4));
}

template <int BN, bool FULL, int ST_NS>
__device__ __forceinline__ void epilogue_chunk32(const GemmEpilogue& ep, const uint32_t (&v)[32], const uint32_t (&g)[32], int row,
                                                 bool row_ok, int row_end, const float* rowbias, int n_blk, int c0, uint8_t* slab,
                                                 int lane, uint8_t* slab_half, uint64_t* st_full, uint64_t* st_free, int qrow, uint32_t& st_cnt) {
  const int n0 = n_blk * BN;
  const bool geglu = ep.act == ADAFACE_ACT_GEGLU;
  const int col0 = n0 + c0;
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  // per-column vectors: a FULL chunk reads its 32 floats as eight 16-byte loads (the scalar form issued 32 loads per vector and
  // chunk -- a third of the general epilogue's instructions)
  auto vec_op = [&](float (&x)[32], const float* vec, int cbase, bool mul) {
    if (FULL && (reinterpret_cast<uintptr_t>(vec + cbase) & 15) == 0) {
      const float4* v4 = reinterpret_cast<const float4*>(vec + cbase);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = __ldg(v4 + j);
        if (mul) {
          x[4 * j] *= b.x; x[4 * j + 1] *= b.y; x[4 * j + 2] *= b.z; x[4 * j + 3] *= b.w;
        } else {
          x[4 * j] += b.x; x[4 * j + 1] += b.y; x[4 * j + 2] += b.z; x[4 * j + 3] += b.w;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (FULL || cbase + j < ep.N) x[j] = mul ? x[j] * __ldg(vec + cbase + j) : x[j] + __ldg(vec + cbase + j);
    }
  };
  auto affine = [&](float (&x)[32], int cbase) {
    if (ep.colscale) vec_op(x, ep.colscale, cbase, true);
    if (ep.bias) vec_op(x, ep.bias, cbase, false);
  };
  affine(f, col0);
  if (rowbias) vec_op(f, rowbias, col0, false);
  int out_col0 = col0;
  int out_n = ep.N;
  if (geglu) {
    float gt[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) gt[j] = __uint_as_float(g[j]);
    affine(gt, col0 + BN / 2);
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] *= gelu_erf(gt[j]);
    out_col0 = n_blk * (BN / 2) + c0;
    out_n = ep.N / 2;
  } else if (ep.act == ADAFACE_ACT_QUICK_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = f[j] / (1.f + __expf(-1.702f * f[j]));
  }
  if (ep.dbg & 1) return;
  const bool staged = FULL && !ep.y_f32 && (ep.hs_d > 0 || ((ep.ldy & 7) == 0 && (out_col0 & 7) == 0 && (reinterpret_cast<uintptr_t>(ep.y) & 15) == 0));   // warp-uniform
  if (!row_ok && !staged) return;
  if (ep.residual && row_ok) {
    if (ep.res_f32) {
      const float* r = reinterpret_cast<const float*>(ep.residual) + (long long)row * ep.ldr + out_col0;
      if (FULL && (reinterpret_cast<uintptr_t>(r) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = *reinterpret_cast<const float4*>(r + 4 * j);
          f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (FULL || out_col0 + j < out_n) f[j] += r[j];
      }
    } else {
      const bf16* r = reinterpret_cast<const bf16*>(ep.residual) + (long long)row * ep.ldr + out_col0;
      if (FULL && (reinterpret_cast<uintptr_t>(r) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 b = *reinterpret_cast<const uint4*>(r + 8 * j);
          const uint32_t w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            f[8 * j + 2 * u] += __uint_as_float(w[u] << 16);
            f[8 * j + 2 * u + 1] += __uint_as_float(w[u] & 0xffff0000u);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (FULL || out_col0 + j < out_n) f[j] += __bfloat162float(r[j]);
      }
    }
  }
  if (ep.y_f32) {
    float* y = reinterpret_cast<float*>(ep.y) + (long long)row * ep.ldy + out_col0;
    if (FULL && ((reinterpret_cast<uintptr_t>(y) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(y + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (FULL || out_col0 + j < out_n) y[j] = f[j];
    }
  } else if (staged && ep.tma_store) {
    // The four warps of this epilogue half (TMEM lane quarters 0..3) drop their 32 rows of the chunk into ONE 128-row slab in the
    // 64-byte-swizzle layout and arrive on its `full` barrier; the STORE WARP turns the slab into one 8 KB TMA bulk store.
    // (Measured: a thread that issues cp.async.bulk.tensor stores is throttled to the store engine's 32 B/clk/SM -- with each
    // epilogue warp issuing its own 2 KB boxes the warps spent 2/3 of a chunk blocked in the issue: 1150 clk per chunk against
    // 390 clk of TMEM read + conversion.  Issued from a dedicated thread, the engine drains under the next chunk's work.)
    const uint32_t seq = st_cnt++;                   // == this half's count of staged chunks (its four warps stage the same chunks)
    const int slot = (int)(seq % ST_NS);
    const bool xtr = ep.trace && blockIdx.x == 0 && threadIdx.x == 0 && seq < 12;
    if (xtr) ep.trace[1800 + seq * 5] = clock64();
    if (seq >= ST_NS) mbar_wait(&st_free[slot], ((seq / ST_NS) - 1) & 1);      // the store that read this slab has finished reading
    if (xtr) ep.trace[1800 + seq * 5 + 1] = clock64();
    uint8_t* sl = slab_half + slot * 8192 + qrow * 64;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<uint4*>(sl + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
          make_uint4(pack_bf16(f[j * 8 + 0], f[j * 8 + 1]), pack_bf16(f[j * 8 + 2], f[j * 8 + 3]), pack_bf16(f[j * 8 + 4], f[j * 8 + 5]),
                     pack_bf16(f[j * 8 + 6], f[j * 8 + 7]));
    if (xtr) ep.trace[1800 + seq * 5 + 2] = clock64();
    fence_proxy_async_smem();                        // every writer orders its generic-proxy writes before the async-proxy read
    if (xtr) ep.trace[1800 + seq * 5 + 3] = clock64();
    __syncwarp();
    if (lane == 0) mbar_arrive(&st_full[slot]);
    if (xtr) ep.trace[1800 + seq * 5 + 4] = clock64();
  } else if (staged) {
    stage_row64(slab, lane, f);
    const int row_base = row - lane;                 // first row of this warp's 32-row slice
    const int pc = lane & 3;                         // 16-byte piece = 8 output columns
    const int col = out_col0 + pc * 8;
    int which = 0, hh = 0, dd = 0;
    if (ep.hs_d > 0) {                               // 8-column groups never straddle a head (hs_d % 8 == 0)
      which = col / ep.hs_C;
      const int rem = col - which * ep.hs_C;
      hh = rem / ep.hs_d;
      dd = rem - hh * ep.hs_d;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = i * 8 + (lane >> 2);
      const int grow = row_base + r;
      if (grow < row_end) {
        bf16* dst;
        if (ep.hs_d > 0) {
          const int bb = grow / ep.hs_rows, nn = grow - bb * ep.hs_rows;
          dst = reinterpret_cast<bf16*>(ep.y) + ((((long long)which * ep.hs_B + bb) * ep.hs_H + hh) * ep.hs_rows + nn) * ep.hs_dpad + dd;
        } else {
          dst = reinterpret_cast<bf16*>(ep.y) + (long long)grow * ep.ldy + col;
        }
        *reinterpret_cast<uint4*>(dst) = unstage_piece(slab, r, pc);
      }
    }
    __syncwarp();                                    // the slab is rewritten by this warp's next chunk
  } else if (ep.hs_d > 0) {
    const int bb = row / ep.hs_rows, nn = row - bb * ep.hs_rows;
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      const int col = out_col0 + g8 * 8;             // 8-column groups never straddle a head (hs_d % 8 == 0)
      if (FULL || col < out_n) {
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
    if (FULL && ((reinterpret_cast<uintptr_t>(y) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 8)
        *reinterpret_cast<uint4*>(y + j) = make_uint4(pack_bf16(f[j], f[j + 1]), pack_bf16(f[j + 2], f[j + 3]), pack_bf16(f[j + 4], f[j + 5]),
                                                      pack_bf16(f[j + 6], f[j + 7]));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (FULL || out_col0 + j < out_n) y[j] = __float2bfloat16(f[j]);
    }
  }
}

// Persistent kernel: grid = min(#tiles, #SMs); CTA c owns tiles c, c + grid, ...  (n fastest, so CTAs running
// together share their A rows through L2).  Warp roles (320 threads; the control warps take the HIGH warp ids,
// which the sub-partition scheduler favours):
//   warps 0-7  epilogue: tcgen05.ld the fp32 accumulator (TMEM lane quarter = warp id & 3; warps w and w+4 take
//              alternate 32-column chunks), DoRA column scale, bias, activation, residual, convert, store.  Runs one
//              tile behind the MMA warp (two TMEM accumulators).  Two warps per scheduler because a lone epilogue
//              warp runs at ~0.25 IPC (measured: the epilogue, not the main loop, bounded the first version).
//   warp 8     TMA producer: 64-wide K slabs of X/W (then of T/Bs: the rank-R LoRA tail continues the same
//              accumulation) into a STAGES-deep 128B-swizzled ring that runs across tile boundaries.
//   warp 9     TMEM allocator + the single thread that issues tcgen05.mma (M=128, N=BN, K=16).
//
// CONV = implicit-GEMM 3x3 convolution over an NHWC activation (openaimodel.py:164-277 ResBlock / Downsample / Upsample
// convolutions): M = output pixels, N = output channels, K = 9 taps x input channels.  Only the
// producer differs: the A stage of (tap, chunk) is a 4-D TMA box {64 channels, W, image rows, images} of the activation
// displaced by the tap offset -- rows / columns outside the image arrive as zeros (the padding) -- and lands in shared
// memory exactly like a [128, 64] K-major tile, so the MMA and epilogue code is the GEMM's.  A stride-2 convolution
// reads through four tensor maps, one per input parity (y & 1, x & 1), each a stride-2 view of the activation.
template <int BN, bool CONV, bool LEAN = false, bool BRES = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tn_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                           const __grid_constant__ CUtensorMap tmB,
                                                                           const __grid_constant__ CUtensorMap tmA2,
                                                                           const __grid_constant__ CUtensorMap tmB2,
                                                                           const __grid_constant__ typename ConvArg<CONV>::type cmaps,
                                                                           const __grid_constant__ CUtensorMap tmY,
                                                                           const GemmEpilogue ep) {
  using Cfg = GemmCfg<BN, BRES>;
  static_assert(!(BRES && CONV), "the W-stationary schedule is built for the plain projection GEMM");
  static_assert(Cfg::STAGES >= 3 && Cfg::SMEM_BYTES <= 227 * 1024, "GEMM shared-memory plan");
  constexpr int STAGES = Cfg::STAGES, ST_NS = Cfg::ST_NS;
  constexpr int kTma = 8, kMma = 9, kStore = 10;
  const int num_kb = ep.num_kb1 + ep.num_kb2;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms need 1024-byte alignment.
  uint8_t* sW = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));   // BRES: resident W slabs
  uint8_t* smem = sW + Cfg::RESIDENT_BYTES;                                                                       // the ring
  uint8_t* store_stage = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(store_stage + Cfg::STORE_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;     // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint64_t* st_full = acc_empty + 2;           // [2][ST_NS]: the four warps of an epilogue half have filled an output slab
  uint64_t* st_free = st_full + 2 * ST_NS;     // [2][ST_NS]: the bulk store that read the slab has finished reading it
  uint64_t* w_full = st_free + 2 * ST_NS;      // BRES: the resident W slice has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_n = (ep.N + BN - 1) / BN;
  const int num_tiles = ep.num_m * num_n * (CONV ? ep.splits : 1);
  // tile walk: c, c + grid, ... (n fastest); BRES: CTA c owns column block c % num_n and walks M tiles c / num_n, + grid / num_n, ...
  // (expressed in the same linear tile index: tile = m * num_n + n)
  const int n_fixed = BRES ? (int)(blockIdx.x % num_n) : 0;
  const int t_first = BRES ? (int)(blockIdx.x / num_n) * num_n + n_fixed : (int)blockIdx.x;
  const int t_step = BRES ? (int)(gridDim.x / num_n) * num_n : (int)gridDim.x;
  const int t_end = num_tiles;

  if (warp == kTma && lane == 0) {
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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 256);
    }
    for (int i = 0; i < 2 * ST_NS; ++i) {
      mbar_init(&st_full[i], 4);
      mbar_init(&st_free[i], 1);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  } else if (warp == kMma) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // this kernel touches global memory only after its predecessor has completed
  AF_PDL_TRIGGER_EARLY();

  if (warp == kTma) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t it = 0;
      if constexpr (BRES) {
        mbar_arrive_expect_tx(w_full, (uint32_t)(num_kb * Cfg::B_STAGE_BYTES));
        for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(sW + kb * Cfg::B_STAGE_BYTES, &tmB, w_full, kb * GEMM_BK, n_fixed * BN);
        for (int tile = t_first; tile < t_end; tile += t_step) {
          const int m0 = (tile / num_n) * GEMM_BM;
          for (int kb = 0; kb < num_kb; ++kb, ++it) {
            const int s = it % STAGES;
            const bool tr = ep.trace && blockIdx.x == 0 && it < 128;
            if (tr) ep.trace[it * 2] = clock64();
            mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
            if (tr) ep.trace[it * 2 + 1] = clock64();
            mbar_arrive_expect_tx(&full_bar[s], A_STAGE_BYTES);
            tma_load_2d(smem + s * Cfg::STAGE_BYTES, &tmA, &full_bar[s], kb * GEMM_BK, m0);
          }
        }
      }
      for (int tile = BRES ? t_end : t_first; tile < t_end; tile += t_step) {
        int tile_mn = tile, kb_lo = 0, kb_hi = num_kb;
        if constexpr (CONV) {
          if (ep.splits > 1) {
            const int sp = tile / (ep.num_m * num_n);
            tile_mn = tile - sp * (ep.num_m * num_n);
            kb_lo = sp * ep.kb_per_split;
            kb_hi = min(num_kb, kb_lo + ep.kb_per_split);
          }
        }
        const int m_blk = tile_mn / num_n, n0 = (tile_mn % num_n) * BN;
        int m0 = m_blk * GEMM_BM, img0 = 0, y0 = 0;
        if constexpr (CONV) {
          if (ep.cv_imgs > 1) {
            img0 = m_blk * ep.cv_imgs;
            m0 = img0 * ep.cv_HW;
          } else {
            img0 = m_blk / ep.cv_tpi;
            const int yc = m_blk - img0 * ep.cv_tpi;
            y0 = yc * (ep.cv_img_rows / ep.cv_W);
            m0 = img0 * ep.cv_HW + yc * ep.cv_img_rows;
          }
        }
        // (tap, channel chunk) of the K block, advanced incrementally: the producer thread was the convolution's bottleneck -- 500 clk
        // per K block against 320 clk of MMA at BN = 160 (traced), a good part of it two integer divisions by run-time values
        int cv_ch = 0, cv_kx = 0, cv_ky = 0;
        uint32_t cv_tx = 0;
        if constexpr (CONV) {
          const int tap0 = kb_lo / ep.cv_kc;
          cv_ch = kb_lo - tap0 * ep.cv_kc;
          cv_ky = tap0 / 3;
          cv_kx = tap0 - cv_ky * 3;
          cv_tx = (uint32_t)(ep.cv_img_rows * ep.cv_imgs * 128 + Cfg::B_STAGE_BYTES);   // the A box holds cv_img_rows * cv_imgs pixels x 128 bytes (zero-filled parts included)
        }
        const int cv_kc = CONV ? ep.cv_kc : 1, cv_stride = CONV ? ep.cv_stride : 1, kb1 = ep.num_kb1;
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          const int s = it % STAGES;
          const bool tr = ep.trace && blockIdx.x == 0 && it < 128;
          if (tr) ep.trace[it * 2] = clock64();
          mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          if (tr) ep.trace[it * 2 + 1] = clock64();
          uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if (CONV && kb < kb1) {
            if constexpr (CONV) {
              mbar_arrive_expect_tx(&full_bar[s], cv_tx);
              if (cv_stride == 1) {
                tma_load_4d(sa, &cmaps.m[0], &full_bar[s], cv_ch * GEMM_BK, cv_kx - 1, y0 + cv_ky - 1, img0);
              } else {
                // input (2 y + ky - 1, 2 x + kx - 1): parity 1 / index -1 for k = 0, parity 0 / index 0 for k = 1,
                // parity 1 / index 0 for k = 2
                const int py = cv_ky != 1, px = cv_kx != 1;
                tma_load_4d(sa, &cmaps.m[py * 2 + px], &full_bar[s], cv_ch * GEMM_BK, cv_kx == 0 ? -1 : 0, y0 + (cv_ky == 0 ? -1 : 0), img0);
              }
              tma_load_2d(sb, &tmB, &full_bar[s], kb * GEMM_BK, n0);
              if (++cv_ch == cv_kc) {
                cv_ch = 0;
                if (++cv_kx == 3) {
                  cv_kx = 0;
                  ++cv_ky;
                }
              }
            }
            continue;
          }
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
    }
  } else if (warp == kMma) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(GEMM_BM, BN);
      uint32_t it = 0, t = 0;
      if constexpr (BRES) {
        mbar_wait(w_full, 0);
        tc_fence_after();
      }
      for (int tile = t_first; tile < t_end; tile += t_step, ++t) {
        const uint32_t buf = t & 1;
        if (ep.trace && blockIdx.x == 0 && t < 32) ep.trace[1024 + t * 2] = clock64();
        mbar_wait(&acc_empty[buf], ((t >> 1) & 1) ^ 1);      // epilogue drained this accumulator (2 tiles ago)
        if (ep.trace && blockIdx.x == 0 && t < 32) ep.trace[1024 + t * 2 + 1] = clock64();
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        int kb_lo = 0, kb_hi = num_kb;
        if constexpr (CONV) {
          if (ep.splits > 1) {
            kb_lo = (tile / (ep.num_m * num_n)) * ep.kb_per_split;
            kb_hi = min(num_kb, kb_lo + ep.kb_per_split);
          }
        }
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          const int s = it % STAGES;
          const bool tr = ep.trace && blockIdx.x == 0 && it < 128;
          if (tr) ep.trace[256 + it * 3] = clock64();
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          if (tr) ep.trace[256 + it * 3 + 1] = clock64();
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint32_t sb = BRES ? smem_u32(sW + kb * Cfg::B_STAGE_BYTES) : sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advancing 16 K-elements inside the swizzle span = +32 bytes on the start address
            umma_bf16(d_tmem, make_smem_desc_sw128(sa + k * 32), make_smem_desc_sw128(sb + k * 32), idesc,
                      (kb > kb_lo || k != 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);   // smem stage reusable once these MMAs have read it
          if (tr) ep.trace[256 + it * 3 + 2] = clock64();
        }
        umma_commit(&acc_full[buf]);    // accumulator complete
      }
    }
  } else if (warp == kStore) {
    // ------------------------------------------------------------------ store warp: one thread turns filled slabs into TMA stores
    if (ep.tma_store && !ep.dbg && elect_one()) {
      const bool geglu = ep.act == ADAFACE_ACT_GEGLU;
      uint32_t seq[2] = {0, 0};
      int prev = -1;
      for (int tile = t_first; tile < t_end; tile += t_step) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;      // (tma_store implies splits == 1)
        int row0 = m_blk * GEMM_BM;
        if constexpr (CONV) {
          if (ep.cv_imgs > 1) {
            row0 = m_blk * ep.cv_imgs * ep.cv_HW;
          } else {
            const int img = m_blk / ep.cv_tpi, yc = m_blk - img * ep.cv_tpi;
            row0 = img * ep.cv_HW + yc * ep.cv_img_rows;
          }
        }
        const int cols_here = min(BN, ep.N - n_blk * BN);
        const int n_chunks = geglu ? BN / 64 : (cols_here + 31) / 32;
        for (int ci = 0; ci < n_chunks; ++ci) {
          if (!geglu && n_blk * BN + ci * 32 + 32 > ep.N) continue;      // ragged chunk: stored by its warps directly
          const int h = ci & 1, slot = (int)(seq[h] % ST_NS), id = h * ST_NS + slot;
          const int bx = (int)(seq[0] + seq[1]);
          const bool str_ = ep.trace && blockIdx.x == 0 && bx < 48;
          if (str_) ep.trace[1600 + bx * 4] = clock64();
          mbar_wait(&st_full[id], (seq[h] / ST_NS) & 1);
          if (str_) ep.trace[1600 + bx * 4 + 1] = clock64();
          ++seq[h];
          tma_store_2d(&tmY, store_stage + id * 8192, (geglu ? n_blk * (BN / 2) : n_blk * BN) + ci * 32, row0);
          tma_store_commit();
          if (str_) ep.trace[1600 + bx * 4 + 2] = clock64();
          tma_store_wait_read<1>();                  // every store but the one just issued has read its slab
          if (str_) ep.trace[1600 + bx * 4 + 3] = clock64();
          if (prev >= 0) mbar_arrive(&st_free[prev]);
          prev = id;
        }
      }
      tma_store_wait_all();                          // the slabs must outlive the bulk stores that read them
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0..7)
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = warp >> 2;         // which alternate 32-column chunks this warp owns
    const bool geglu = ep.act == ADAFACE_ACT_GEGLU;
    uint32_t t = 0, st_cnt = 0;
    for (int tile = t_first; tile < t_end; tile += t_step, ++t) {
      const uint32_t buf = t & 1;
      int tile_mn = tile, split = 0;
      if constexpr (CONV) {
        if (ep.splits > 1) {
          split = tile / (ep.num_m * num_n);
          tile_mn = tile - split * (ep.num_m * num_n);
        }
      }
      const int m_blk = tile_mn / num_n, n_blk = tile_mn % num_n;
      int row = m_blk * GEMM_BM + q * 32 + lane;
      int row_end = ep.M;
      const float* rowbias = nullptr;
      if constexpr (CONV) {
        int base;
        if (ep.cv_imgs > 1) {
          base = m_blk * ep.cv_imgs * ep.cv_HW;
          row_end = min(ep.M, base + ep.cv_imgs * ep.cv_HW);
        } else {
          const int img = m_blk / ep.cv_tpi, yc = m_blk - img * ep.cv_tpi;
          base = img * ep.cv_HW + yc * ep.cv_img_rows;
          row_end = min((img + 1) * ep.cv_HW, base + ep.cv_img_rows);
        }
        row = base + q * 32 + lane;
        if (ep.rowbias) rowbias = ep.rowbias + (long long)(min(row, ep.M - 1) / ep.cv_HW) * ep.N;
      }
      const bool row_ok = row < row_end;
      const bool etr = ep.trace && blockIdx.x == 0 && t < 32 && (warp == 0 || warp == 7) && lane == 0;
      if (etr) ep.trace[1200 + (warp ? 100 : 0) + t * 3] = clock64();
      mbar_wait(&acc_full[buf], (t >> 1) & 1);
      if (etr) ep.trace[1200 + (warp ? 100 : 0) + t * 3 + 1] = clock64();
      tc_fence_after();
      const uint32_t t_acc = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
      const int cols_here = min(BN, ep.N - n_blk * BN);                       // real columns of this tile
      const int n_chunks = geglu ? BN / 64 : (cols_here + 31) / 32;
      bool arrived = false;
      if (half >= n_chunks) {
        tc_fence_before();
        mbar_arrive(&acc_empty[buf]);
        arrived = true;
      }
      if (ep.dbg & 2) {
        if (!arrived) {
          tc_fence_before();
          mbar_arrive(&acc_empty[buf]);
        }
        continue;
      }
      if constexpr (LEAN) {
        // LEAN instantiation (chosen per launch): Y = bf16(X W^T [+ bias]), every chunk full, output through the store warp.
        // The general epilogue below carries every option as a warp-uniform branch (~120 issued instructions per chunk at
        // ~9.5 clk per instruction with two epilogue warps per scheduler: 1150 clk per chunk, 3500 per 128 x 192 tile against
        // 1920 clk of MMA at K = 320); this one is the ~50 instructions the plain projection needs.
        uint8_t* slab_half = store_stage + half * (ST_NS * 8192) + q * 32 * 64 + lane * 64;
        uint64_t* stf = st_full + half * ST_NS;
        uint64_t* stfree = st_free + half * ST_NS;
        const int swz = (lane >> 1) & 3;
        auto lean_chunk = [&](uint32_t (&v)[32], int ci) {
          if (ep.bias) {
            const float4* b4 = reinterpret_cast<const float4*>(ep.bias + n_blk * BN + ci * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(b4 + j);
              v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + b.x);
              v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + b.y);
              v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + b.z);
              v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + b.w);
            }
          }
          const uint32_t seq = st_cnt++;
          const int slot = (int)(seq % ST_NS);
          if (seq >= ST_NS) mbar_wait(&stfree[slot], ((seq / ST_NS) - 1) & 1);
          uint8_t* sl = slab_half + slot * 8192;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(sl + ((j ^ swz) << 4)) =
                make_uint4(pack_bf16(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1])), pack_bf16(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3])),
                           pack_bf16(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5])), pack_bf16(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7])));
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&stf[slot]);
        };
        uint32_t va[32], vb[32];
        int ci = half;
        if (ci < n_chunks) tmem_ld_32x32b_x32_nowait(t_acc + (uint32_t)(ci * 32), va);
#pragma unroll 1
        while (ci < n_chunks) {
          tmem_ld_wait_x32(va);
          if (ci + 2 < n_chunks) tmem_ld_32x32b_x32_nowait(t_acc + (uint32_t)((ci + 2) * 32), vb);
          else if (!arrived) {
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
            arrived = true;
          }
          lean_chunk(va, ci);
          ci += 2;
          if (ci >= n_chunks) break;
          tmem_ld_wait_x32(vb);
          if (ci + 2 < n_chunks) tmem_ld_32x32b_x32_nowait(t_acc + (uint32_t)((ci + 2) * 32), va);
          else if (!arrived) {
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
            arrived = true;
          }
          lean_chunk(vb, ci);
          ci += 2;
        }
        if (etr) ep.trace[1200 + (warp ? 100 : 0) + t * 3 + 2] = clock64();
        continue;
      }
      // General epilogue: one chunk = 32 accumulator columns of this warp's 32 rows.
      auto process = [&](const uint32_t (&v)[32], const uint32_t (&g)[32], int ci) {
        const int c0 = ci * 32;
        if constexpr (CONV) {
          if (ep.splits > 1) {      // raw fp32 partial tile: the epilogue terms are applied by the reduce kernel
            if (row_ok) {
              float* dst = ep.splitk_ws + ((long long)split * ep.M + row) * ep.N + n_blk * BN + c0;
              const int nv = min(32, ep.N - (n_blk * BN + c0));
              if (nv == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < nv) dst[j] = __uint_as_float(v[j]);
              }
            }
            return;
          }
        }
        const bool full = geglu || (n_blk * BN + c0 + 32 <= ep.N);
        if (full) epilogue_chunk32<BN, true, ST_NS>(ep, v, g, row, row_ok, row_end, rowbias, n_blk, c0, store_stage + warp * 4096, lane, store_stage + half * (ST_NS * 8192), st_full + half * ST_NS, st_free + half * ST_NS, q * 32, st_cnt);
        else epilogue_chunk32<BN, false, ST_NS>(ep, v, g, row, row_ok, row_end, rowbias, n_blk, c0, store_stage + warp * 4096, lane, store_stage + half * (ST_NS * 8192), st_full + half * ST_NS, st_free + half * ST_NS, q * 32, st_cnt);
      };
      auto release_acc = [&]() {
        // last TMEM read of this warp for this tile has landed: hand the accumulator back before the stores
        tc_fence_before();
        mbar_arrive(&acc_empty[buf]);
        arrived = true;
      };
      if (geglu) {
#pragma unroll 1
        for (int ci = half; ci < n_chunks; ci += 2) {
          uint32_t v[32], g[32];
          tmem_ld_32x32b_x32_nowait(t_acc + (uint32_t)(ci * 32), v);
          tmem_ld_32x32b_x32_wait(t_acc + (uint32_t)(ci * 32 + BN / 2), g);
          tmem_ld_wait_x32(v);
          if (ci + 2 >= n_chunks && !arrived) release_acc();
          process(v, g, ci);
        }
      } else {
#pragma unroll 1
        for (int ci = half; ci < n_chunks; ci += 2) {
          uint32_t v[32];
          tmem_ld_32x32b_x32_wait(t_acc + (uint32_t)(ci * 32), v);
          if (ci + 2 >= n_chunks && !arrived) release_acc();
          process(v, v, ci);
        }
      }
      if (etr) ep.trace[1200 + (warp ? 100 : 0) + t * 3 + 2] = clock64();
    }
  }
  AF_PDL_TRIGGER_LATE();
  tc_fence_before();
  __syncthreads();
  if (warp == kMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
extern long long g_launch_count;

template <int BN, bool CONV = false>
static int launch_gemm(const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tA2, const CUtensorMap& tB2,
                       const CUtensorMap& tY, const GemmEpilogue& ep, int n_tiles, cudaStream_t stream,
                       const typename ConvArg<CONV>::type& cmaps = typename ConvArg<CONV>::type()) {
  using Cfg = GemmCfg<BN>;
  constexpr bool HAS_BRES = !CONV && (BN == 128 || BN == 160 || BN == 192);
  using CfgR = GemmCfg<BN, HAS_BRES>;
  static DevOnce configured;
  const int cfg_dev = af_device();
  if (!configured.done(cfg_dev)) {
    AF_CUDA(cudaFuncSetAttribute((gemm_tn_tcgen05_kernel<BN, CONV, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    AF_CUDA(cudaFuncSetAttribute((gemm_tn_tcgen05_kernel<BN, CONV, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    if constexpr (HAS_BRES) {
      AF_CUDA(cudaFuncSetAttribute((gemm_tn_tcgen05_kernel<BN, CONV, false, HAS_BRES>), cudaFuncAttributeMaxDynamicSharedMemorySize, CfgR::SMEM_BYTES));
      AF_CUDA(cudaFuncSetAttribute((gemm_tn_tcgen05_kernel<BN, CONV, true, HAS_BRES>), cudaFuncAttributeMaxDynamicSharedMemorySize, CfgR::SMEM_BYTES));
    }
    configured.set(cfg_dev);
  }
  const int num_sms = af_num_sms();
  const int tiles = n_tiles * ep.num_m * (CONV ? ep.splits : 1);
  dim3 grid(tiles < num_sms ? tiles : num_sms);
  // LEAN epilogue: plain bf16 projection (optional bias) whose every 32-column chunk is full and leaves through the store warp
  static int lean_on = -1, bres_on = -1;
  if (lean_on < 0) {
    const char* e = getenv("ADAFACE_GEMM_LEAN");      // A/B switches (default on)
    lean_on = (e && e[0] == '0') ? 0 : 1;
    e = getenv("ADAFACE_GEMM_BRES");
    bres_on = (e && e[0] == '0') ? 0 : 1;
  }
  const bool lean = lean_on && ep.tma_store && !ep.dbg && !ep.colscale && !ep.rowbias && !ep.residual && ep.act == ADAFACE_ACT_NONE && !ep.y_f32 &&
                    ep.hs_d == 0 && ep.splits == 1 && ep.N % 32 == 0 && (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0;
  if constexpr (HAS_BRES) {
    // W-stationary schedule: short K, no LoRA tail, whole column blocks, and every CTA gets >= 4 M tiles
    const int per_col = num_sms / n_tiles;
    if (bres_on && ep.num_kb2 == 0 && ep.num_kb1 <= BRES_KB && n_tiles <= num_sms && per_col >= 1 && ep.num_m >= 4 * per_col) {
      const dim3 g2(per_col * n_tiles);
      if (lean)
        AF_CUDA(launch_pdl(1, gemm_tn_tcgen05_kernel<BN, CONV, true, HAS_BRES>, g2, dim3(GEMM_THREADS), CfgR::SMEM_BYTES, stream, tA, tB, tA2, tB2, cmaps, tY, ep));
      else
        AF_CUDA(launch_pdl(1, gemm_tn_tcgen05_kernel<BN, CONV, false, HAS_BRES>, g2, dim3(GEMM_THREADS), CfgR::SMEM_BYTES, stream, tA, tB, tA2, tB2, cmaps, tY, ep));
      AF_CUDA(cudaGetLastError());
      ++g_launch_count;
      return 0;
    }
  }
  if (lean)
    AF_CUDA(launch_pdl(1, gemm_tn_tcgen05_kernel<BN, CONV, true>, grid, dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tA, tB, tA2, tB2, cmaps, tY, ep));
  else
    AF_CUDA(launch_pdl(1, gemm_tn_tcgen05_kernel<BN, CONV, false>, grid, dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tA, tB, tA2, tB2, cmaps, tY, ep));
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

// Diagnosis (env ADAFACE_GEMM_TRACE): CTA 0 stamps clock64 at its producer / MMA / epilogue / store hand-offs; dumped to stderr.
static long long* trace_begin() {
  static long long* trace_buf = nullptr;
  if (!getenv("ADAFACE_GEMM_TRACE")) return nullptr;
  if (!trace_buf) cudaMalloc(&trace_buf, 2048 * 8);
  cudaMemset(trace_buf, 0, 2048 * 8);
  return trace_buf;
}
static void trace_dump(const long long* dev, const char* what, long long M, long long N, long long K, int BN, int num_kb) {
  static long long h[2048];
  cudaDeviceSynchronize();
  cudaMemcpy(h, dev, sizeof(h), cudaMemcpyDeviceToHost);
  const long long t0 = h[0];
  fprintf(stderr, "%s trace M=%lld N=%lld K=%lld BN=%d kb=%d\n", what, M, N, K, BN, num_kb);
  for (int i = 0; i < 64; ++i)
    fprintf(stderr, "it %3d  tma: wait %6lld..%6lld | mma: wait %6lld..%6lld issued %6lld\n", i, h[i * 2] - t0, h[i * 2 + 1] - t0, h[256 + i * 3] - t0,
            h[256 + i * 3 + 1] - t0, h[256 + i * 3 + 2] - t0);
  for (int t = 0; t < 12; ++t)
    fprintf(stderr, "tile %2d  mma acc_empty wait %6lld..%6lld | epi w0: acc_full wait %6lld..%6lld done %6lld | epi w7: %6lld..%6lld done %6lld\n", t,
            h[1024 + t * 2] - t0, h[1024 + t * 2 + 1] - t0, h[1200 + t * 3] - t0, h[1200 + t * 3 + 1] - t0, h[1200 + t * 3 + 2] - t0, h[1300 + t * 3] - t0,
            h[1300 + t * 3 + 1] - t0, h[1300 + t * 3 + 2] - t0);
  for (int b = 0; b < 12; ++b)
    fprintf(stderr, "staging %2d (warp 0): enter %6lld  slab free %6lld  written %6lld  fenced %6lld  arrived %6lld\n", b, h[1800 + b * 5] - t0, h[1800 + b * 5 + 1] - t0,
            h[1800 + b * 5 + 2] - t0, h[1800 + b * 5 + 3] - t0, h[1800 + b * 5 + 4] - t0);
  for (int b = 0; b < 30; ++b)
    fprintf(stderr, "store box %2d: wait full %6lld..%6lld  issued %6lld  prev read done %6lld\n", b, h[1600 + b * 4] - t0, h[1600 + b * 4 + 1] - t0, h[1600 + b * 4 + 2] - t0,
            h[1600 + b * 4 + 3] - t0);
}

static bool tma_store_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("ADAFACE_GEMM_TMA_STORE");      // A/B switch (default on)
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

// Tile width: the widest of {256, 192, 160, 128, 64} that wastes no column and still gives every SM a tile.
static int pick_tile_width(long long N, long long m_tiles, int act) {
  int BN;
  if (act == ADAFACE_ACT_GEGLU) {
    BN = 128;
  } else {
    BN = 0;
    // cost model: rounds over the 148 SMs x (tile width + fixed per-tile overhead); exact divisors of N only
    const int cand[5] = {256, 192, 160, 128, 64};
    long long best = -1;
    for (int i = 0; i < 5; ++i) {
      if (N % cand[i]) continue;
      const long long tiles = m_tiles * (N / cand[i]);
      const long long cost = ((tiles + 147) / 148) * (cand[i] + 64);
      if (best < 0 || cost < best) { best = cost; BN = cand[i]; }
    }
    if (!BN) BN = (N <= 64 || N % 128 <= 64) && N < 256 ? 64 : 128;   // ragged N: masked last tile
  }
  {
    static int force_bn = -1;
    if (force_bn < 0) {
      const char* e = getenv("ADAFACE_GEMM_BN");     // tuning knob: force the tile width (64 / 128 / 160)
      force_bn = e ? atoi(e) : 0;
    }
    if (force_bn && act != ADAFACE_ACT_GEGLU) BN = force_bn;
  }
  return BN;
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

  const long long m_tiles = (M + GEMM_BM - 1) / GEMM_BM;
  if (act == ADAFACE_ACT_GEGLU)
    AF_CHECK(N % 128 == 0, "proj_lora_fwd: GEGLU needs N %% 128 == 0 (packed [a|g] tiles), got %lld", (long long)N);
  const int BN = pick_tile_width(N, m_tiles, act);
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
  ep.cv_kc = ep.cv_W = ep.cv_HW = ep.cv_img_rows = ep.cv_tpi = ep.cv_imgs = ep.cv_stride = 0;
  ep.num_m = (int)m_tiles;
  ep.rowbias = nullptr;
  ep.splits = 1;
  ep.kb_per_split = 0;
  ep.splitk_ws = nullptr;
  {
    const char* e = getenv("ADAFACE_GEMM_DBG");
    ep.dbg = e ? atoi(e) : 0;
  }
  const int n_tiles = (int)((N + BN - 1) / BN);
  ep.trace = trace_begin();
  CUtensorMap tY = tA;
  ep.tma_store = 0;
  if (tma_store_enabled() && hs_d == 0 && y_dtype == ADAFACE_BF16 && ldy % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    if (make_tmap_bf16_store32(&tY, y, (uint64_t)M, (uint64_t)(act == ADAFACE_ACT_GEGLU ? N / 2 : N), (uint64_t)ldy)) return 3;
    ep.tma_store = 1;
  }
  int rc = -1;
  switch (BN) {
    case 64: rc = launch_gemm<64>(tA, tB, tA2, tB2, tY, ep, n_tiles, stream); break;
    case 128: rc = launch_gemm<128>(tA, tB, tA2, tB2, tY, ep, n_tiles, stream); break;
    case 160: rc = launch_gemm<160>(tA, tB, tA2, tB2, tY, ep, n_tiles, stream); break;
    case 192: rc = launch_gemm<192>(tA, tB, tA2, tB2, tY, ep, n_tiles, stream); break;
    case 256: rc = launch_gemm<256>(tA, tB, tA2, tB2, tY, ep, n_tiles, stream); break;
  }
  if (rc < 0) {
    set_error("proj_lora_fwd: unreachable tile width %d", BN);
    return 1;
  }
  if (ep.trace && rc == 0) trace_dump(ep.trace, "GEMM", (long long)M, (long long)N, (long long)K, BN, ep.num_kb1 + ep.num_kb2);
  return rc;
}

// y[row, c] = act(colscale[c] * sum_s ws[s][row][c] + bias[c] + rowbias[row / HW][c]) + residual[row, c]: the epilogue of a
// split-K convolution.  One thread per 4 columns (N % 4 == 0), partials summed in split order (deterministic).
__global__ void __launch_bounds__(256) conv_splitk_reduce_kernel(const GemmEpilogue ep) {
  const int n4 = ep.N >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)ep.M * n4) return;
  const int row = (int)(idx / n4), c = (int)(idx - (long long)row * n4) << 2;
  const long long plane = (long long)ep.M * ep.N;
  const float* p = ep.splitk_ws + (long long)row * ep.N + c;
  float4 acc = *reinterpret_cast<const float4*>(p);
  for (int s = 1; s < ep.splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(p + s * plane);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float f[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (ep.colscale) f[j] *= ep.colscale[c + j];
    if (ep.bias) f[j] += ep.bias[c + j];
    if (ep.rowbias) f[j] += ep.rowbias[(long long)(row / ep.cv_HW) * ep.N + c + j];
    if (ep.act == ADAFACE_ACT_QUICK_GELU) f[j] = f[j] / (1.f + __expf(-1.702f * f[j]));
    if (ep.residual) {
      f[j] += ep.res_f32 ? reinterpret_cast<const float*>(ep.residual)[(long long)row * ep.ldr + c + j]
                         : __bfloat162float(reinterpret_cast<const bf16*>(ep.residual)[(long long)row * ep.ldr + c + j]);
    }
  }
  if (ep.y_f32) {
    float* y = reinterpret_cast<float*>(ep.y) + (long long)row * ep.ldy + c;
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = f[j];
  } else {
    bf16* y = reinterpret_cast<bf16*>(ep.y) + (long long)row * ep.ldy + c;
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = __float2bfloat16(f[j]);
  }
}

// Split-K workspace: one buffer per device, grown on demand (stream-ordered allocation is not needed: the buffer is only
// ever touched by kernels of the calling stream, and a larger request replaces it after a device synchronisation).
// Single-stream use per device: two streams running split-K convolutions concurrently would share it.
// The 32 MB floor is never outgrown by conv3x3_fwd: it splits only when tiles <= 74 and takes splits <= 148 / tiles, so
// splits * M * Cout <= 148 tiles * 128 rows * 256 columns = 4.85 M floats < 8 M.  The regrow branch therefore never runs after
// the first call, and pointers baked into captured CUDA graphs stay valid.
static float* splitk_workspace(size_t floats) {
  constexpr int MAX_DEV = 64;
  static float* ws[MAX_DEV] = {nullptr};
  static size_t cap[MAX_DEV] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return nullptr;
  if (floats > cap[dev]) {
    if (ws[dev]) {
      if (cudaDeviceSynchronize() != cudaSuccess) return nullptr;
      cudaFree(ws[dev]);
      ws[dev] = nullptr;
      cap[dev] = 0;
    }
    const size_t want = floats < (size_t(8) << 20) ? (size_t(8) << 20) : floats;      // >= 32 MB: covers every SD-1.5 shape
    if (cudaMalloc(&ws[dev], want * sizeof(float)) != cudaSuccess) return nullptr;
    cap[dev] = want;
  }
  return ws[dev];
}

// 3x3 convolution (padding 1, stride 1 | 2) over an NHWC activation: see the CONV notes at the kernel.
int conv3x3_fwd(const void* x, int64_t B, int64_t H, int64_t W, int64_t Cin, const void* w, const void* t, int64_t ldt,
                const void* bs, int64_t R, const float* colscale, const float* bias, const float* rowbias, const void* residual,
                int64_t ldr, int residual_dtype, void* y, int64_t ldy, int y_dtype, int64_t Cout, int stride, int act,
                cudaStream_t stream) {
  AF_CHECK(x && w && y, "conv3x3_fwd: null x/w/y");
  AF_CHECK(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3x3_fwd: empty problem");
  AF_CHECK(stride == 1 || stride == 2, "conv3x3_fwd: stride %d (1 | 2)", stride);
  AF_CHECK(stride == 1 || (H % 2 == 0 && W % 2 == 0), "conv3x3_fwd: stride 2 needs even H, W (got %lld x %lld)", (long long)H, (long long)W);
  AF_CHECK(Cin % 8 == 0, "conv3x3_fwd: Cin (%lld) must be a multiple of 8 (16-byte TMA strides)", (long long)Cin);
  AF_CHECK((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0, "conv3x3_fwd: x / w must be 16-byte aligned");
  AF_CHECK(act == ADAFACE_ACT_NONE || act == ADAFACE_ACT_QUICK_GELU, "conv3x3_fwd: bad act %d", act);
  const bool lora = (t != nullptr) || (bs != nullptr) || R > 0;
  if (lora) {
    AF_CHECK(t && bs && R > 0 && R % 8 == 0 && ldt % 8 == 0, "conv3x3_fwd: t, bs, R (multiple of 8) must be given together");
    AF_CHECK((reinterpret_cast<uintptr_t>(t) & 15) == 0 && (reinterpret_cast<uintptr_t>(bs) & 15) == 0, "conv3x3_fwd: t / bs must be 16-byte aligned");
  }
  const int64_t Ho = H / stride, Wo = W / stride;
  AF_CHECK(Wo <= 128, "conv3x3_fwd: output width %lld > 128 is not tiled yet", (long long)Wo);
  const int64_t M = B * Ho * Wo;
  AF_CHECK(M < (1ll << 31) && Cout < (1ll << 31), "conv3x3_fwd: problem too large");
  int64_t rh = 128 / Wo;                       // image rows per tile
  if (rh > Ho) rh = Ho;
  int64_t imgs = 1, tpi;
  if (rh == Ho) {
    imgs = 128 / (Wo * Ho);
    if (imgs < 1) imgs = 1;
    if (imgs > B) imgs = B;
    tpi = 1;
  } else {
    tpi = (Ho + rh - 1) / rh;
  }
  const int64_t kc = (Cin + GEMM_BK - 1) / GEMM_BK;

  GemmEpilogue ep;
  ep.colscale = colscale;
  ep.bias = bias;
  ep.residual = residual;
  ep.y = y;
  ep.ldr = ldr;
  ep.ldy = ldy;
  ep.M = (int)M;
  ep.N = (int)Cout;
  ep.num_kb1 = (int)(9 * kc);
  ep.num_kb2 = lora ? (int)((R + GEMM_BK - 1) / GEMM_BK) : 0;
  ep.act = act;
  ep.y_f32 = y_dtype == ADAFACE_F32;
  ep.res_f32 = residual_dtype == ADAFACE_F32;
  ep.hs_d = ep.hs_dpad = ep.hs_H = ep.hs_C = ep.hs_rows = ep.hs_B = 0;
  ep.cv_kc = (int)kc;
  ep.cv_W = (int)Wo;
  ep.cv_HW = (int)(Ho * Wo);
  ep.cv_img_rows = (int)(Wo * rh);
  ep.cv_tpi = (int)tpi;
  ep.cv_imgs = (int)imgs;
  ep.cv_stride = stride;
  ep.num_m = (int)(imgs > 1 ? (B + imgs - 1) / imgs : B * tpi);
  ep.rowbias = rowbias;
  ep.splits = 1;
  ep.kb_per_split = 0;
  ep.splitk_ws = nullptr;
  ep.dbg = 0;
  ep.trace = trace_begin();

  int BN = pick_tile_width(Cout, ep.num_m, act);
  // The tile-width cost model is the projection GEMM's (short K, per-tile overhead matters).  With K = 9 Cin the main loop
  // dominates and the widest tile that divides Cout wins on arithmetic intensity: measured (profiles/r01_conv_bn_sweep.txt)
  // 160 beats 64 / 128 / 256 at every SD-1.5 shape, also where it leaves SMs idle (B = 2, 32 x 32, 1280 -> 640: 44 vs 60 us).
  if (Cout % 160 == 0 && getenv("ADAFACE_GEMM_BN") == nullptr) BN = 160;
  {
    // split-K when the output tiles cannot fill half the SMs and the K loop is long (levels C / D of the U-Net)
    static int splitk = -1;
    if (splitk < 0) {
      const char* e = getenv("ADAFACE_CONV_SPLITK");       // A/B switch: 0 = never split
      splitk = e ? atoi(e) : 1;
    }
    const int num_kb = ep.num_kb1 + ep.num_kb2;
    if (splitk && Cout % 4 == 0 && num_kb >= 48) {
      int bn = BN;
      if (Cout % 128 == 0 && (long long)ep.num_m * (Cout / 128) <= 74) bn = 128;      // wider tiles: each A tile is re-read by fewer CTAs
      const long long tiles = (long long)ep.num_m * ((Cout + bn - 1) / bn);
      if (tiles <= 74) {
        int sp = (int)(148 / tiles);
        if (sp > 8) sp = 8;
        if (sp > num_kb / 16) sp = num_kb / 16;
        if (sp >= 2) {
          const int per = (num_kb + sp - 1) / sp;
          sp = (num_kb + per - 1) / per;                   // no empty split
          float* ws = splitk_workspace((size_t)sp * (size_t)M * (size_t)Cout);
          AF_CHECK(ws != nullptr, "conv3x3_fwd: cannot allocate the split-K workspace");
          ep.splits = sp;
          ep.kb_per_split = per;
          ep.splitk_ws = ws;
          BN = bn;
        }
      }
    }
  }
  ConvMaps cm;
  const bf16* xb = reinterpret_cast<const bf16*>(x);
  if (stride == 1) {
    if (make_tmap_bf16_nhwc(&cm.m[0], xb, (uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B, (uint64_t)Cin, (uint64_t)(W * Cin),
                            (uint64_t)(H * W * Cin), (uint32_t)Wo, (uint32_t)rh, (uint32_t)imgs))
      return 3;
    cm.m[1] = cm.m[2] = cm.m[3] = cm.m[0];
  } else {
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px)
        if (make_tmap_bf16_nhwc(&cm.m[py * 2 + px], xb + (py * W + px) * Cin, (uint64_t)Cin, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B,
                                (uint64_t)(2 * Cin), (uint64_t)(2 * W * Cin), (uint64_t)(H * W * Cin), (uint32_t)Wo, (uint32_t)rh,
                                (uint32_t)imgs))
          return 3;
  }
  CUtensorMap tB, tA2, tB2;
  if (make_tmap_bf16_2d(&tB, w, (uint64_t)Cout, (uint64_t)(9 * kc * GEMM_BK), (uint64_t)(9 * kc * GEMM_BK), (uint32_t)BN)) return 3;
  if (lora) {
    if (make_tmap_bf16_2d(&tA2, t, (uint64_t)M, (uint64_t)R, (uint64_t)ldt, GEMM_BM)) return 3;
    if (make_tmap_bf16_2d(&tB2, bs, (uint64_t)Cout, (uint64_t)R, (uint64_t)R, (uint32_t)BN)) return 3;
  } else {
    tA2 = tB;
    tB2 = tB;
  }
  const int n_tiles = (int)((Cout + BN - 1) / BN);
  // TMA-store epilogue: only when every tile is 128 real output rows (no rows of the tile fall outside its image)
  CUtensorMap tY = tB;
  ep.tma_store = 0;
  const bool full_tiles = imgs > 1 ? (imgs * Ho * Wo == GEMM_BM) : (Wo * rh == GEMM_BM && Ho % rh == 0);
  if (tma_store_enabled() && ep.splits == 1 && full_tiles && y_dtype == ADAFACE_BF16 && ldy % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    if (make_tmap_bf16_store32(&tY, y, (uint64_t)M, (uint64_t)Cout, (uint64_t)ldy)) return 3;
    ep.tma_store = 1;
  }
  int rc;
  switch (BN) {
    case 64: rc = launch_gemm<64, true>(tB, tB, tA2, tB2, tY, ep, n_tiles, stream, cm); break;
    case 128: rc = launch_gemm<128, true>(tB, tB, tA2, tB2, tY, ep, n_tiles, stream, cm); break;
    case 160: rc = launch_gemm<160, true>(tB, tB, tA2, tB2, tY, ep, n_tiles, stream, cm); break;
    case 192: rc = launch_gemm<192, true>(tB, tB, tA2, tB2, tY, ep, n_tiles, stream, cm); break;
    case 256: rc = launch_gemm<256, true>(tB, tB, tA2, tB2, tY, ep, n_tiles, stream, cm); break;
    default:
      set_error("conv3x3_fwd: unreachable tile width %d", BN);
      return 1;
  }
  if (ep.trace && rc == 0) trace_dump(ep.trace, "conv3x3", (long long)M, (long long)Cout, (long long)(9 * kc * GEMM_BK), BN, ep.num_kb1 + ep.num_kb2);
  if (rc || ep.splits == 1) return rc;
  const long long n_thr = (long long)M * (Cout / 4);
  conv_splitk_reduce_kernel<<<(unsigned)((n_thr + 255) / 256), 256, 0, stream>>>(ep);
  AF_CUDA(cudaGetLastError());
  ++g_launch_count;
  return 0;
}

}  // namespace adaface

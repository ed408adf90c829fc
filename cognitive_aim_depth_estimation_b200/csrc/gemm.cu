// Persistent, warp-specialised tcgen05 GEMM for sm_100a:  C = epilogue(A * W^T), bf16 operands, fp32
// accumulation in TMEM.
//
//   warp 0      : TMA producer   (one lane) — A/W tiles -> 128B-swizzled smem ring, mbarrier complete_tx
//   warp 1      : MMA issuer     (one lane) — tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16 per instruction,
//                                  tcgen05.commit releases smem stages and publishes accumulators;
//                                  the whole warp allocates / frees TMEM
//   warps 2..9  : epilogue       — tcgen05.ld from the warp's TMEM lane quadrant (warp_id % 4), two warps per
//                                  quadrant split the BN columns; fused bias / GELU / LayerScale+residual /
//                                  pos-embed / softmax-statistics epilogues straight to global memory, and the
//                                  LayerNorm-folding pair (gemm.cuh EPI_LN_* / EPI_RESID_LN_F32)
//
// Two TMEM accumulator stages (2 x BN columns) let the epilogue of tile i overlap the MMAs of tile i+1.
// Tiles are scheduled statically (tile = blockIdx.x + i * gridDim.x) with the N index fastest so the CTAs
// working at the same time share A tiles through L2.
//
// Replaces the ATen/cuBLAS call sites listed in SURVEY.md §2.1 (HF modeling_dinov2.py:148,199-201,246,324-328;
// reference src/model.py:192-197).
#include "common.cuh"
#include "gemm.cuh"

#include <cuda_fp16.h>
#include "host.h"

#include <stdlib.h>

#ifndef CA_GEMM_RELEASE_ARRIVE
#define CA_GEMM_RELEASE_ARRIVE 0  // 1 = the former cluster-scope-release accumulator hand-back (A/B only)
#endif

namespace ca {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // one 128-byte swizzle atom of bf16
constexpr int kNumEpiWarps = 8;
constexpr int kGemmThreads = (2 + kNumEpiWarps) * 32;
// Epilogue staging: one 32-row x 128-byte output chunk per warp (32 fp32 or 64 bf16 columns).  The 16-byte piece j of
// row r lives at r*128 + ((j ^ (r & 7)) << 4): the thread-per-row writes (fixed j, 8 consecutive rows per quarter
// warp) and the row-contiguous reads (fixed row, 8 pieces per quarter warp) are both bank-conflict free, and every
// global access of the read side covers whole 128-byte lines (4 rows x 128 B per warp instruction).
constexpr int kEpiStageBytes = 32 * 128;
__device__ __forceinline__ uint32_t epi_off(int r, int piece) { return r * 128 + ((piece ^ (r & 7)) << 4); }

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // 512 or 256: power of two
  static constexpr int kStagingBytes = kNumEpiWarps * kEpiStageBytes;  // epilogue transposition buffers
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmKernelArgs {
  int M, N, K;
  int m_tiles, n_tiles, total_tiles;
  int w_batched;
  void* out;
  int ldo;
  long long out_batch_stride;
  const float* bias;
  const float* ls;
  const float* pos;
  int patches_per_img;
  float scale_log2;
  float* part_a;
  float* part_b;
  const float* col_max;
  const float* col_rinv;
  int partials;
  float2* ln_stats;        // EPI_LN_* (read) / EPI_RESID_LN_F32 (written): [M, ln_slots] (sum, M2) per 128-column span
  int ln_slots;
  float ln_eps;
  __nv_bfloat16* shadow;   // EPI_RESID_LN_F32: bf16 copy of the updated rows
  int ld_shadow;
  int dbg;  // CA_GEMM_DEBUG experiments (bit 0: skip B loads, bit 1: skip A loads) — results are then garbage
};

// erf-GELU of two values, x * Phi(x), without the MUFU: Phi(x) - 1/2 is an odd degree-17 minimax polynomial on
// |x| <= 4.2 whose leading coefficient is positive, so one saturating FMA (FFMA.SAT clamps to [0, 1]) both adds the 1/2
// and supplies the exact limits beyond the fit range.  |Phi error| <= 1.3e-5 and |GELU error| <= 5.5e-5 for EVERY x
// (checked in fp32 Horner arithmetic over [-12, 12], tests/test_gemm_gpu.py), one to two orders below the bf16 rounding
// of the output.  8 FFMA2 + 2 FMUL2 per pair + 1 FFMA.SAT per value: 7 issue slots per value against ~20 for an
// exp/rcp formulation whose 2 MUFU per value alone would need 4096 clk of a 6144-clk 128 x 256 x 768 tile.
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  float s0, s1, p0, p1;
  fmul2(s0, s1, x0, x1, x0, x1);
  ffma2(p0, p1, s0, s1, 6.013480685629347e-11f, -5.644098521884189e-09f);
  ffma2v(p0, p1, p0, p1, s0, s1, 2.3467518417419342e-07f, 2.3467518417419342e-07f);
  ffma2v(p0, p1, p0, p1, s0, s1, -5.765378773503471e-06f, -5.765378773503471e-06f);
  ffma2v(p0, p1, p0, p1, s0, s1, 9.46131840464659e-05f, 9.46131840464659e-05f);
  ffma2v(p0, p1, p0, p1, s0, s1, -0.0011143309529870749f, -0.0011143309529870749f);
  ffma2v(p0, p1, p0, p1, s0, s1, 0.009830592200160027f, 0.009830592200160027f);
  ffma2v(p0, p1, p0, p1, s0, s1, -0.06636093556880951f, -0.06636093556880951f);
  ffma2v(p0, p1, p0, p1, s0, s1, 0.39890772104263306f, 0.39890772104263306f);
  const float phi0 = ffma_sat(p0, x0, 0.5f), phi1 = ffma_sat(p1, x1, 0.5f);
  fmul2(x0, x1, x0, x1, phi0, phi1);
}

// 1 / sqrt(var + eps) of a residual row from the per-128-column (sum, M2) pairs the residual epilogue (or ca_ln_shadow) left
// (Chan's combination); 1 for rows past M.
__device__ __forceinline__ float ln_row_rstd(const GemmKernelArgs& p, int row) {
  if (row >= p.M) return 1.f;
  const float2* st = p.ln_stats + static_cast<size_t>(row) * p.ln_slots;
  const float inv_dim = 1.f / (128.f * p.ln_slots);
  float tot = 0.f, m2 = 0.f;
  for (int i = 0; i < p.ln_slots; ++i) {
    const float2 t = __ldg(st + i);
    tot += t.x;
    m2 += t.y;
  }
  const float mean = tot * inv_dim;
  for (int i = 0; i < p.ln_slots; ++i) {
    const float d = __ldg(st + i).x * (1.f / 128.f) - mean;
    m2 = fmaf(128.f * d, d, m2);
  }
  return rsqrtf(m2 * inv_dim + p.ln_eps);
}

// One epilogue warp: rows [32*q, 32*q+32) of the tile (q = warp_id % 4), columns [col0, col0 + BN/2).
template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(const GemmKernelArgs& p, const CUtensorMap* tmap_x, uint32_t tmem_acc, int b,
                                              int mt, int nt, int quad, int half, uint8_t* stage, float rstd = 1.f) {
  constexpr int kSpan = BN / 2;
  const int lane = lane_id();
  const int row = mt * BM + quad * 32 + lane;           // row inside batch b
  const int col_base = nt * BN + half * kSpan;          // first column this warp owns
  const bool row_ok = row < p.M;
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(half * kSpan);

  if constexpr (EPI == EPI_ROWSTATS) {
    // Softmax statistics per 64-column span (the granularity ca_rowstats_merge / ca_colsum_e work in): this warp owns
    // kSpan / 64 of them (one with 128-wide tiles, two with the 256-wide tiles of the CTA-pair kernel).
    // pass 1: max ; pass 2: sum exp2(s - max).  TMEM re-read is cheaper than holding the span in registers.
    constexpr int kSub = kSpan / 64;
    const uint32_t stage_s = smem_u32(stage);
#pragma unroll 1
    for (int sub = 0; sub < kSub; ++sub) {
      const int col_sub = col_base + sub * 64;
      const uint32_t taddr_sub = taddr + static_cast<uint32_t>(sub * 64);
      // spans that lie wholly inside [0, N) — all but the last one or two of a row — skip the per-element bounds test
      // (a compare and a select per score in an epilogue that already carries one MUFU per score)
      const bool full = col_sub + 64 <= p.N;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr_sub + c, v);
        tmem_ld_wait();
        if (full) {
          float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]);
#pragma unroll
          for (int j = 2; j < 32; j += 2) {
            m0 = fmaxf(m0, __uint_as_float(v[j]));
            m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
          }
          mx = fmaxf(mx, fmaxf(m0, m1));  // raw scores: the (positive) scale is applied once, below
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col_sub + c + j < p.N) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
      }
      mx *= p.scale_log2;  // max_j (s_j * scale) = scale * max_j s_j  (scale > 0); fl(s * scale) is monotone in s
      float sum = 0.f;
      // optional: keep the span-relative exponentials E = exp2(s - span max) as fp16 so that the column sums of the
      // softmax are a bandwidth pass over E instead of a second Q K^T (ca_colsum_e)
      const bool keep_e = p.out != nullptr && col_sub < p.ldo;  // warp-uniform; ldo is a multiple of 64: whole spans
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr_sub + c, v);
        tmem_ld_wait();
        float ev[32];
        if (full) {
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            ev[j] = fast_exp2(__uint_as_float(v[j]) * p.scale_log2 - mx);
            ev[j + 1] = fast_exp2(__uint_as_float(v[j + 1]) * p.scale_log2 - mx);
            s0 += ev[j];
            s1 += ev[j + 1];
          }
          sum += s0 + s1;
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float s = __uint_as_float(v[j]) * p.scale_log2;
            ev[j] = (col_sub + c + j < p.N) ? fast_exp2(s - mx) : 0.f;
            sum += ev[j];
          }
        }
        if (keep_e) {  // thread-per-row -> the warp's swizzled staging buffer (16-byte piece 4 * (c / 32) + j / 8 of the row)
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 w;
            __half2 h0 = __floats2half2_rn(ev[j + 0], ev[j + 1]), h1 = __floats2half2_rn(ev[j + 2], ev[j + 3]);
            __half2 h2 = __floats2half2_rn(ev[j + 4], ev[j + 5]), h3 = __floats2half2_rn(ev[j + 6], ev[j + 7]);
            w.x = *reinterpret_cast<uint32_t*>(&h0);
            w.y = *reinterpret_cast<uint32_t*>(&h1);
            w.z = *reinterpret_cast<uint32_t*>(&h2);
            w.w = *reinterpret_cast<uint32_t*>(&h3);
            sts128(stage_s + epi_off(lane, (c >> 3) + (j >> 3)), w);
          }
        }
      }
      if (keep_e) {
        // ... and out of it with 8 lanes per 128-byte row segment, 4 rows per instruction: whole lines per store (16-byte
        // stores straight from the thread-per-row layout touch 32 different lines per instruction)
        __syncwarp();
        __half* ebase = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(b) * p.out_batch_stride + col_sub + (lane & 7) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = (lane >> 3) + 4 * i;
          const int grow = mt * BM + quad * 32 + rr;
          const uint4 w = lds128(stage_s + epi_off(rr, lane & 7));
          if (grow < p.M) *reinterpret_cast<uint4*>(ebase + static_cast<size_t>(grow) * p.ldo) = w;
        }
        __syncwarp();
      }
      if (row_ok) {
        // span-major [batch, P, M]: the 32 rows of this warp are 128 contiguous bytes (row-major [batch, M, P] made every
        // lane's 4-byte store its own sector, 2.1 M store requests per focal iteration)
        const size_t o = (static_cast<size_t>(b) * p.partials + nt * (BN / 64) + half * kSub + sub) * p.M + row;
        p.part_a[o] = mx;
        p.part_b[o] = sum;
      }
    }
    return;
  } else if constexpr (EPI == EPI_COLSUM) {
    const float* cmax = p.col_max + static_cast<size_t>(b) * p.N;
    const float* cinv = p.col_rinv + static_cast<size_t>(b) * p.N;
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < kSpan; c += 32) {
      uint32_t v[32];
      tmem_ld32(taddr + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = col_base + c + j;
        if (col < p.N) {
          const float s = __uint_as_float(v[j]) * p.scale_log2;
          sum += fast_exp2(s - __ldg(cmax + col)) * __ldg(cinv + col);
        }
      }
    }
    if (row_ok) {
      const size_t o = (static_cast<size_t>(b) * p.partials + nt * 2 + half) * p.M + row;
      p.part_a[o] = sum;
    }
    return;
  } else {
    // Dense epilogues.  TMEM hands each thread one accumulator ROW; global memory wants each warp instruction to
    // cover whole 32-byte sectors of a few rows.  So every 64-byte-per-row output chunk (32 bf16 or 16 fp32 columns)
    // is finished per row in registers (bias, GELU), transposed through a padded smem buffer, and then read /
    // modified / written with lanes laid out as 8 rows x 4 x 16 B: every global access moves full sectors.
    const int lrow = lane >> 3;  // row within a group of 4 (8 passes cover the warp's 32 rows)
    const int seg = lane & 7;    // 16-byte piece of the 128-byte chunk row
    const uint32_t stage_s = smem_u32(stage);  // explicit .shared accesses (LDS / STS instead of generic LD.E / ST.E)
    if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_GELU_BF16 || EPI == EPI_LN_BIAS_BF16 || EPI == EPI_LN_GELU_BF16) {
      constexpr bool kLn = (EPI == EPI_LN_BIAS_BF16 || EPI == EPI_LN_GELU_BF16);
      constexpr bool kGelu = (EPI == EPI_GELU_BF16 || EPI == EPI_LN_GELU_BF16);
      // LayerNorm folded into this GEMM: A holds the raw residual rows, W is gamma-scaled with every row CENTRED
      // (sum_k W'[n,k] = 0), so the accumulator already is sum_k (x_k - mean) W'[n,k]; what is left for the epilogue is
      // the row's 1/std (`rstd`, ln_row_rstd: fetched while the accumulator was still being computed).
#pragma unroll 1
      for (int c = 0; c < kSpan; c += 64) {
        const int col = col_base + c;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          tmem_ld32(taddr + c + 32 * hh, v);
          tmem_ld_wait();
          if (col < p.N) {  // warp-uniform (N % 64 == 0 is checked on the host)
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col + 32 * hh);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float a[8];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const float4 t = __ldg(b4 + 2 * j + h);
                if constexpr (kLn) {
                  a[4 * h + 0] = fmaf(rstd, __uint_as_float(v[8 * j + 4 * h + 0]), t.x);
                  a[4 * h + 1] = fmaf(rstd, __uint_as_float(v[8 * j + 4 * h + 1]), t.y);
                  a[4 * h + 2] = fmaf(rstd, __uint_as_float(v[8 * j + 4 * h + 2]), t.z);
                  a[4 * h + 3] = fmaf(rstd, __uint_as_float(v[8 * j + 4 * h + 3]), t.w);
                } else {
                  a[4 * h + 0] = __uint_as_float(v[8 * j + 4 * h + 0]) + t.x;
                  a[4 * h + 1] = __uint_as_float(v[8 * j + 4 * h + 1]) + t.y;
                  a[4 * h + 2] = __uint_as_float(v[8 * j + 4 * h + 2]) + t.z;
                  a[4 * h + 3] = __uint_as_float(v[8 * j + 4 * h + 3]) + t.w;
                }
              }
              if constexpr (kGelu) {
#pragma unroll
                for (int i = 0; i < 8; i += 2) gelu_erf2(a[i], a[i + 1]);
              }
              uint4 w;
              w.x = pack_bf16x2(a[0], a[1]);
              w.y = pack_bf16x2(a[2], a[3]);
              w.z = pack_bf16x2(a[4], a[5]);
              w.w = pack_bf16x2(a[6], a[7]);
              sts128(stage_s + epi_off(lane, 4 * hh + j), w);
            }
          }
        }
        if (col < p.N) {
          __syncwarp();
          __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(b) * p.out_batch_stride +
                                 col + seg * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = lrow + 4 * i;
            const int grow = mt * BM + quad * 32 + rr;
            const uint4 w = lds128(stage_s + epi_off(rr, seg));
            if (grow < p.M) *reinterpret_cast<uint4*>(obase + static_cast<size_t>(grow) * p.ldo) = w;
          }
          __syncwarp();
        }
      }
    } else if constexpr (EPI == EPI_RESID_F32) {
      // x += ls * (acc + bias), in place, WITHOUT reading x into the SM: each 32-row x 32-column piece is finished in
      // registers, laid out in the warp's 128B-swizzled staging buffer and handed to the TMA as a reduce-add
      // (cp.reduce.async.bulk.tensor .add, fp32): the L2 performs the read-modify-write.  x does not fit L2 between
      // layers, so a load-modify-store epilogue is bound by the latency of its residual reads (8 warps x 4 KB in
      // flight per SM); the reduce form only ever writes.  Rows past M are clipped by the tensor map.
      constexpr int kChunks = kSpan / 32;
#pragma unroll 1
      for (int g = 0; g < kChunks; ++g) {
        const int c = g * 32;
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
        tmem_ld_wait();
        const int col = col_base + c;
        if (lane == 0) bulk_wait_group_read<0>();  // the previous reduce has finished reading the staging buffer
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + col) + j);
          const float4 l4 = __ldg(reinterpret_cast<const float4*>(p.ls + col) + j);
          sts128f(stage_s + epi_off(lane, j),
              make_float4(l4.x * (__uint_as_float(v[4 * j + 0]) + t.x), l4.y * (__uint_as_float(v[4 * j + 1]) + t.y),
                          l4.z * (__uint_as_float(v[4 * j + 2]) + t.z), l4.w * (__uint_as_float(v[4 * j + 3]) + t.w)));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(tmap_x, stage, col, mt * BM + quad * 32);
          bulk_commit_group();
        }
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < kSpan; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
        tmem_ld_wait();
        const int col = col_base + c;
        if (col < p.N) {  // warp-uniform (N % 32 == 0)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 a = make_float4(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1]),
                                   __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            if constexpr (EPI != EPI_F32) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + col) + j);
              a.x += t.x;
              a.y += t.y;
              a.z += t.z;
              a.w += t.w;
            }
            sts128f(stage_s + epi_off(lane, j), a);
          }
          __syncwarp();
          float* obase = reinterpret_cast<float*>(p.out) + static_cast<size_t>(b) * p.out_batch_stride + col + seg * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = lrow + 4 * i;
            const int grow = mt * BM + quad * 32 + rr;
            const float4 a = lds128f(stage_s + epi_off(rr, seg));
            if (grow < p.M) {
              if constexpr (EPI == EPI_PATCH_F32) {
                const int img = grow / p.patches_per_img;
                const int pidx = grow - img * p.patches_per_img;
                const size_t orow = static_cast<size_t>(img) * (p.patches_per_img + 1) + 1 + pidx;
                const float4 q = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(1 + pidx) * p.N + col) + seg);
                *reinterpret_cast<float4*>(obase + orow * p.ldo) = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
              } else {  // EPI_F32
                *reinterpret_cast<float4*>(obase + static_cast<size_t>(grow) * p.ldo) = a;
              }
            }
          }
          __syncwarp();
        }
      }
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)  // 10 warps -> 3 warps on two SMSPs -> 168 registers per thread
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_x, const GemmKernelArgs p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms are 1024 B; align the ring explicitly (dynamic smem is only 16 B aligned by contract).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::kStages * Cfg::kABytes;
  uint8_t* smem_stage = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;                        // [kStages] TMA -> MMA
  uint64_t* empty_bar = bars + Cfg::kStages;        // [kStages] MMA -> TMA
  uint64_t* acc_full = bars + 2 * Cfg::kStages;     // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;               // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = warp_id();
  const int lane = lane_id();
  const int kblocks = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kNumEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_sync();  // PDL: barriers, TMEM and descriptors are set up; from here on the kernel touches its operands
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        const int rest = tile / p.n_tiles;
        const int mt = rest % p.m_tiles;
        const int b = rest / p.m_tiles;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          const bool la = !(p.dbg & 2) || kb == 0, lb = !(p.dbg & 1) || kb == 0;
          mbar_arrive_expect_tx(&full_bar[stage], (la ? Cfg::kABytes : 0) + (lb ? Cfg::kBBytes : 0));
          if (la) tma_load_3d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * BK, mt * BM, b);
          if (lb)
            tma_load_3d(smem_b + stage * Cfg::kBBytes, &tmap_w, &full_bar[stage], kb * BK, nt * BN,
                        p.w_batched ? b : 0);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_smem_desc_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes per 16-element K step inside the swizzle atom  (descriptor address unit = 16 B)
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&acc_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    const int e = warp - 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int half = e >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int rest = tile / p.n_tiles;
      const int mt = rest % p.m_tiles;
      const int b = rest / p.m_tiles;
      float rstd = 1.f;
      if constexpr (EPI == EPI_LN_BIAS_BF16 || EPI == EPI_LN_GELU_BF16) rstd = ln_row_rstd(p, mt * BM + quad * 32 + lane);
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<BN, EPI>(p, &tmap_x, tmem_base + static_cast<uint32_t>(acc * BN), b, mt, nt, quad, half,
                             smem_stage + e * kEpiStageBytes, rstd);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if constexpr (EPI == EPI_RESID_F32) {
      if (lane == 0) bulk_wait_group<0>();  // every reduce of this warp has been performed
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2) for the dense epilogues: a cluster of two CTAs on one TPC computes a
// 256 x 256 tile; each CTA loads its own 128 rows of A and HALF of the 256 weight rows, the leader issues
// 256 x 256 x 16 MMAs that read both shared memories, and each CTA's TMEM receives its 128 accumulator rows.
// Per CTA and 64-wide K block the shared memory sees 32 KB of TMA fill + 32 KB of operand reads instead of 48 + 48:
// the single-CTA kernel is shared-memory-bandwidth bound (ablation: 1250 TF/s with loads, 1700 without).
// ---------------------------------------------------------------------------------------------------------------
struct Gemm2Cfg {
  static constexpr int BN = 256;
  static constexpr int kStages = 6;
  static constexpr int kABytes = BM * BK * 2;         // 16 KB: this CTA's 128 rows of A
  static constexpr int kBBytes = (BN / 2) * BK * 2;   // 16 KB: this CTA's 128 of the 256 weight rows
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kStagingBytes = kNumEpiWarps * kEpiStageBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 + 256;
};

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                     const __grid_constant__ CUtensorMap tmap_x, const GemmKernelArgs p) {
  using Cfg = Gemm2Cfg;
  constexpr int BN = Cfg::BN;
  constexpr bool kResidLn = (EPI == EPI_RESID_LN_F32);
  constexpr int n_stages = Cfg::kStages;
  constexpr int staging_bytes = Cfg::kStagingBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + n_stages * Cfg::kABytes;
  uint8_t* smem_stage = smem + n_stages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + staging_bytes);
  uint64_t* full_bar = bars;                        // [kStages] TMA (both CTAs) -> leader's MMA warp; leader's copy used
  uint64_t* empty_bar = bars + Cfg::kStages;        // [kStages] leader's MMA (multicast commit) -> each CTA's producer
  uint64_t* acc_full = bars + 2 * Cfg::kStages;     // [2] leader's MMA (multicast commit) -> each CTA's epilogue
  uint64_t* acc_empty = acc_full + 2;               // [2] epilogue warps of BOTH CTAs -> leader's MMA; leader's copy used
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = warp_id();
  const int lane = lane_id();
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int kblocks = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 2 * kNumEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  griddep_sync();  // PDL: barriers, TMEM and descriptors are set up; from here on the kernel touches its operands
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters) {
        const int nt = tile % p.n_tiles;
        const int rest = tile / p.n_tiles;
        const int mt2 = rest % p.m_tiles;
        const int b = rest / p.m_tiles;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          const uint32_t leader_full = mapa_cluster(smem_u32(&full_bar[stage]), 0);
          // CA_GEMM_DEBUG experiments (results are garbage): 1 = no B loads after the first K block, 2 = no A loads,
          // 16 = every cluster loads tile (0, 0) (same L2 lines for everyone)
          const bool la = !(p.dbg & 2) || kb == 0, lb = !(p.dbg & 1) || kb == 0;
          const int mt_l = (p.dbg & 16) ? 0 : mt2, nt_l = (p.dbg & 16) ? 0 : nt;
          if (rank == 0)  // both CTAs' bytes land on the leader's barrier
            mbar_arrive_expect_tx(&full_bar[stage], 2 * ((la ? Cfg::kABytes : 0) + (lb ? Cfg::kBBytes : 0)));
          if (la) tma_load_3d_2sm(smem_a + stage * Cfg::kABytes, &tmap_a, leader_full, kb * BK, mt_l * 2 * BM + rank * BM, b);
          if (lb)
            tma_load_3d_2sm(smem_b + stage * Cfg::kBBytes, &tmap_w, leader_full, kb * BK, nt_l * BN + rank * (BN / 2),
                            p.w_batched ? b : 0);
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_smem_desc_sw128(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_bf16_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage], 3);  // frees the stage in BOTH CTAs
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_2sm(&acc_full[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    const int e = warp - 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int half = e >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (kResidLn) {
      // x += ls * (acc + bias) with the OLD rows read into the SM, because the LayerNorm that follows needs the new
      // values.  Plain loads and stores, no TMA and so no proxy fence (a fence waits for the warp's earlier global stores
      // to be acknowledged — measured: that serialised every chunk behind the previous chunk's stores): each warp walks
      // its 32-row x 32-column chunks (4 per tile) as ONE stream across tiles.  Per chunk the update ls * (acc + bias)
      // goes from the accumulator's thread-per-row layout through the swizzled staging buffer into the layout global
      // memory wants (8 lanes per 128-byte row segment, 4 rows per instruction); there the old values — loaded one
      // chunk ahead into the registers the previous chunk has just released — are added, and the fp32 rows, their
      // bf16 copy (the next GEMM's A operand) and the running (sum, sum of squares) of each row leave from registers.
      const int lrow = lane >> 3, seg = lane & 7;
      const uint32_t stage_s = smem_u32(smem_stage + e * kEpiStageBytes);
      const int my_tiles = cluster_id < p.total_tiles ? (p.total_tiles - cluster_id + n_clusters - 1) / n_clusters : 0;
      const int total_chunks = my_tiles * 4;
      // (first column of this warp's 128-column span, first row of its 32 rows) of the i-th tile of this cluster
      auto tile_origin = [&](int i, int& col0, int& row0, int& nt) {
        const int tile = cluster_id + i * n_clusters;
        nt = tile % p.n_tiles;
        const int mt2 = (tile / p.n_tiles) % p.m_tiles;
        col0 = nt * BN + half * 128;
        row0 = (mt2 * 2 + static_cast<int>(rank)) * BM + quad * 32;
      };
      float* const xbase = reinterpret_cast<float*>(p.out);
      float4 old[8];  // piece i: row lrow + 4 i of the chunk, columns seg * 4 .. + 3
      int n_col0 = 0, n_row0 = 0, n_nt = 0;  // origin of the tile the load stream is in
      auto load_piece = [&](int s, int i) {  // chunk s, piece i -> old[i]
        const int row = n_row0 + lrow + 4 * i;
        old[i] = row < p.M ? *reinterpret_cast<const float4*>(xbase + static_cast<size_t>(row) * p.ldo + n_col0 + (s & 3) * 32 + seg * 4)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      if (total_chunks > 0) {
        tile_origin(0, n_col0, n_row0, n_nt);
#pragma unroll
        for (int i = 0; i < 8; ++i) load_piece(0, i);
      }
      float rs[8], rq[8];
      int col0 = 0, row0 = 0, nt = 0;
      int pf_col0 = 0, pf_row0 = 0;
      for (int s = 0; s < total_chunks; ++s) {
        const int g = s & 3;
        if (g == 0) {
          tile_origin(s >> 2, col0, row0, nt);
          int pf_nt;
          tile_origin((s >> 2) + 1, pf_col0, pf_row0, pf_nt);
#pragma unroll
          for (int i = 0; i < 8; ++i) rs[i] = rq[i] = 0.f;
          mbar_wait(&acc_full[acc], acc_phase);
          tc_fence_after();
        }
        const int col = col0 + g * 32;
        // the registers hold ONE chunk of old values in flight; the rows a tile further on start their way from HBM into
        // the L2 now, so that load finds them there
        if (s + 4 < total_chunks && pf_row0 + lane < p.M)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(xbase + static_cast<size_t>(pf_row0 + lane) * p.ldo + pf_col0 + g * 32));
        uint32_t v[32];
        tmem_ld32(tmem_base + static_cast<uint32_t>(acc * BN) + (static_cast<uint32_t>(quad * 32) << 16) +
                      static_cast<uint32_t>(half * 128 + g * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + col) + j);
          const float4 l4 = __ldg(reinterpret_cast<const float4*>(p.ls + col) + j);
          float4 d;
          fadd2v(d.x, d.y, __uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1]), t.x, t.y);
          fadd2v(d.z, d.w, __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]), t.z, t.w);
          fmul2(d.x, d.y, d.x, d.y, l4.x, l4.y);
          fmul2(d.z, d.w, d.z, d.w, l4.z, l4.w);
          sts128f(stage_s + epi_off(lane, j), d);
        }
        if (g == 3) {  // the accumulator has been read: hand it back before the global-memory half of the chunk
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(mapa_cluster(smem_u32(&acc_empty[acc]), 0));
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        } else {
          __syncwarp();
        }
        // origin of the chunk after this one (its old values replace this chunk's, piece by piece)
        const bool more = s + 1 < total_chunks;
        if (more && ((s + 1) & 3) == 0) tile_origin((s + 1) >> 2, n_col0, n_row0, n_nt);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = lrow + 4 * i;
          const int row = row0 + rr;
          const float4 d = lds128f(stage_s + epi_off(rr, seg));
          float4 n;
          fadd2v(n.x, n.y, old[i].x, old[i].y, d.x, d.y);
          fadd2v(n.z, n.w, old[i].z, old[i].w, d.z, d.w);
          if (more) load_piece(s + 1, i);
          if (row < p.M) {
            *reinterpret_cast<float4*>(xbase + static_cast<size_t>(row) * p.ldo + col + seg * 4) = n;
            *reinterpret_cast<uint2*>(p.shadow + static_cast<size_t>(row) * p.ld_shadow + col + seg * 4) =
                make_uint2(pack_bf16x2(n.x, n.y), pack_bf16x2(n.z, n.w));
          }
          rs[i] += (n.x + n.y) + (n.z + n.w);
          rq[i] = fmaf(n.x, n.x, fmaf(n.y, n.y, fmaf(n.z, n.z, fmaf(n.w, n.w, rq[i]))));
        }
        __syncwarp();  // the staging buffer is free for the next chunk
        if (g == 3) {
          // the 8 lanes of a row segment hold the pieces of one row: (sum, sum of squares) of its 128 columns
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], o);
              rq[i] += __shfl_xor_sync(0xffffffffu, rq[i], o);
            }
          }
          if (seg == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = row0 + lrow + 4 * i;
              // M2 about the span mean; the squares are taken about 0, which costs a relative 1e-7 (1 + mean^2 / var)
              if (row < p.M)
                p.ln_stats[static_cast<size_t>(row) * p.ln_slots + nt * 2 + half] =
                    make_float2(rs[i], fmaxf(rq[i] - rs[i] * rs[i] * (1.f / 128.f), 0.f));
            }
          }
        }
      }
    } else
    for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters) {
      const int nt = tile % p.n_tiles;
      const int rest = tile / p.n_tiles;
      const int mt = (rest % p.m_tiles) * 2 + static_cast<int>(rank);  // this CTA's 128-row tile
      const int b = rest / p.m_tiles;
      float rstd = 1.f;
      if constexpr (EPI == EPI_LN_BIAS_BF16 || EPI == EPI_LN_GELU_BF16) rstd = ln_row_rstd(p, mt * BM + quad * 32 + lane);
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<BN, EPI>(p, &tmap_x, tmem_base + static_cast<uint32_t>(acc * BN), b, mt, nt, quad, half,
                             smem_stage + e * kEpiStageBytes, rstd);
      tc_fence_before();
      __syncwarp();
#if CA_GEMM_RELEASE_ARRIVE
      if (lane == 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&acc_empty[acc]), 0));
#else
      // "this warp has read its part of the accumulator out of TMEM": no memory is published to the other CTA through
      // this barrier, so the arrive carries CTA-scope release only (the cluster-scope release made every epilogue warp
      // wait for its result stores to be acknowledged: 18 % of the fc1 kernel's warp samples sat on the resulting
      // MEMBAR.ALL.GPU / ERRBAR)
      if (lane == 0) mbar_arrive_remote(mapa_cluster(smem_u32(&acc_empty[acc]), 0));
#endif
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if constexpr (EPI == EPI_RESID_F32) {
      if (lane == 0) bulk_wait_group<0>();  // every reduce of this warp has been performed
    }
  }

  // the peer's shared memory is read by the leader's MMAs and its barriers are signalled remotely: leave together
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

template <int EPI>
int launch_inst2(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tx, const GemmKernelArgs& ka,
                 cudaStream_t stream) {
  auto kern = gemm2_tcgen05_kernel<EPI>;
  const int smem = Gemm2Cfg::kSmemBytes;
  static PerDeviceOnce configured;  // per instantiation and per device (the opt-in is per context)
  CA_TRY(configured.run([&]() -> int {
    CA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg::kSmemBytes));
    return 0;
  }));
  const int max_clusters = sm_count() / 2;
  const int clusters = ka.total_tiles < max_clusters ? ka.total_tiles : max_clusters;
  CA_TRY(launch_kernel(kern, dim3(2 * clusters), dim3(kGemmThreads), smem, stream, ta, tw, tx, ka));
  CA_CUDA(cudaGetLastError());
  return 0;
}

template <int BN, int EPI>
int launch_inst(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& tx, const GemmKernelArgs& ka,
                cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_tcgen05_kernel<BN, EPI>;
  static PerDeviceOnce configured;  // per instantiation and per device
  CA_TRY(configured.run([&]() -> int {
    CA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    return 0;
  }));
  int grid = ka.total_tiles < sm_count() ? ka.total_tiles : sm_count();
  CA_TRY(launch_kernel(kern, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream, ta, tw, tx, ka));
  CA_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int gemm_launch(const GemmArgs& a, cudaStream_t stream) {
  CA_REQUIRE(a.A && a.W, "gemm: null operand");
  CA_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0 && a.batch > 0, "gemm: non-positive dimension");
  CA_REQUIRE(a.lda % 8 == 0 && a.ldw % 8 == 0, "gemm: lda/ldw must be multiples of 8 elements (16 B) for TMA");
  CA_REQUIRE(a.K <= a.lda && a.K <= a.ldw, "gemm: K exceeds a leading dimension");
  const bool stats = (a.epilogue == EPI_ROWSTATS || a.epilogue == EPI_COLSUM);
  // dense epilogues and the row-statistics pass run on CTA pairs (cta_group::2, 256 x 256 pair tiles) unless
  // CA_GEMM_1CTA is set; the (test-only) column-sum recompute pass keeps the single-CTA 128 x 128 kernel, whose operand
  // traffic (256 B/clk of fill + reads against a 128 B/clk shared-memory port) caps it near half the tensor rate
  static const bool force_1cta = getenv("CA_GEMM_1CTA") != nullptr;
  const bool resid_ln = a.epilogue == EPI_RESID_LN_F32;
  const bool ln_in = a.epilogue == EPI_LN_BIAS_BF16 || a.epilogue == EPI_LN_GELU_BF16;
  const bool pair = (!force_1cta || resid_ln) && a.epilogue != EPI_COLSUM;
  const int bn = (stats && !pair) ? 128 : 256;
  if (!stats) {
    CA_REQUIRE(a.N % 32 == 0, "gemm: N must be a multiple of 32 for the dense epilogues");
    const bool bf16_out = a.epilogue == EPI_BIAS_BF16 || a.epilogue == EPI_GELU_BF16 || ln_in;
    CA_REQUIRE(!bf16_out || a.N % 64 == 0, "gemm: the bf16 epilogues need N % 64 == 0");
    CA_REQUIRE(a.out != nullptr, "gemm: null output");
    CA_REQUIRE(a.epilogue == EPI_F32 || a.bias != nullptr, "gemm: null bias");
    const int vec = bf16_out ? 8 : 4;
    CA_REQUIRE(a.ldo % vec == 0, "gemm: ldo must keep rows 16-byte aligned");
    const bool resid = a.epilogue == EPI_RESID_F32 || resid_ln;
    CA_REQUIRE(!resid || a.ls != nullptr, "gemm: null LayerScale");
    CA_REQUIRE(!resid || a.N % 256 == 0, "gemm: the residual epilogues need N % 256 == 0");
    CA_REQUIRE(!(ln_in || resid_ln) || (a.ln_stats != nullptr && a.ln_slots > 0 && a.batch == 1),
               "gemm: the LayerNorm-folding epilogues need a statistics buffer and are not batched");
    CA_REQUIRE(!ln_in || (a.K == 128 * a.ln_slots && a.ln_eps > 0.f), "gemm: EPI_LN_* need K == 128 * ln_slots and eps > 0");
    CA_REQUIRE(!resid_ln || (a.shadow != nullptr && a.N == 128 * a.ln_slots && a.ld_shadow % 8 == 0 &&
                             (reinterpret_cast<uintptr_t>(a.shadow) & 15) == 0),
               "gemm: EPI_RESID_LN_F32 needs a 16-byte aligned bf16 shadow and N == 128 * ln_slots");
    CA_REQUIRE(a.epilogue != EPI_PATCH_F32 || (a.pos != nullptr && a.patches_per_img > 0), "gemm: null pos-embed");
  } else {
    CA_REQUIRE(a.part_a != nullptr, "gemm: null partial buffer");
    CA_REQUIRE(a.epilogue != EPI_ROWSTATS || a.part_b != nullptr, "gemm: null partial-sum buffer");
    CA_REQUIRE(a.epilogue != EPI_COLSUM || (a.col_max && a.col_rinv), "gemm: null column statistics");
    CA_REQUIRE(a.epilogue != EPI_ROWSTATS || a.out == nullptr || (a.ldo % 64 == 0 && a.ldo >= a.N),
               "gemm: the exponential matrix needs a leading dimension that is a multiple of 64 and >= N");
  }

  CUtensorMap ta, tw;
  const long long abs = a.batch > 1 ? a.a_batch_stride : static_cast<long long>(a.M) * a.lda;
  CA_REQUIRE(abs % 8 == 0, "gemm: A batch stride must be a multiple of 8 elements");
  CA_TRY(make_tmap_3d(&ta, a.A, a.batch, a.M, a.K, a.lda, abs, BM));
  const bool wb = a.batch > 1 && a.w_batch_stride != 0;
  const long long wbs = wb ? a.w_batch_stride : static_cast<long long>(a.N) * a.ldw;
  CA_REQUIRE(wbs % 8 == 0, "gemm: W batch stride must be a multiple of 8 elements");
  CA_TRY(make_tmap_3d(&tw, a.W, wb ? a.batch : 1, a.N, a.K, a.ldw, wbs, pair ? bn / 2 : bn));

  CUtensorMap tx = ta;  // only the residual epilogue reads it
  if (a.epilogue == EPI_RESID_F32 || resid_ln) {
    CA_REQUIRE(a.batch == 1, "gemm: the residual epilogue is not batched");
    CA_REQUIRE(a.ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0, "gemm: residual rows must be 16-byte aligned");
    CA_TRY(make_tmap_f32_2d(&tx, a.out, a.M, a.N, a.ldo, 32));
  }

  GemmKernelArgs ka;
  ka.M = a.M;
  ka.N = a.N;
  ka.K = a.K;
  ka.m_tiles = pair ? (a.M + 2 * BM - 1) / (2 * BM) : (a.M + BM - 1) / BM;  // pair kernel: 256-row tiles
  ka.n_tiles = (a.N + bn - 1) / bn;
  if (stats && bn == 128) ka.n_tiles = (ka.n_tiles + 1) & ~1;  // every one of the P = 4 ceil(N / 256) slots gets written
  ka.total_tiles = ka.m_tiles * ka.n_tiles * a.batch;
  ka.w_batched = wb ? 1 : 0;
  ka.out = a.out;
  ka.ldo = a.ldo;
  ka.out_batch_stride = a.out_batch_stride;
  ka.bias = a.bias;
  ka.ls = a.ls;
  ka.pos = a.pos;
  ka.patches_per_img = a.patches_per_img;
  ka.scale_log2 = a.scale_log2;
  ka.part_a = a.part_a;
  ka.part_b = a.part_b;
  ka.col_max = a.col_max;
  ka.col_rinv = a.col_rinv;
  ka.partials = gemm_stats_partials(a.N);
  ka.ln_stats = reinterpret_cast<float2*>(a.ln_stats);
  ka.ln_slots = a.ln_slots;
  ka.ln_eps = a.ln_eps;
  ka.shadow = a.shadow;
  ka.ld_shadow = a.ld_shadow;
  static const int dbg = getenv("CA_GEMM_DEBUG") ? atoi(getenv("CA_GEMM_DEBUG")) : 0;
  ka.dbg = dbg;

  if (pair) {
    switch (a.epilogue) {
      case EPI_BIAS_BF16: return launch_inst2<EPI_BIAS_BF16>(ta, tw, tx, ka, stream);
      case EPI_GELU_BF16: return launch_inst2<EPI_GELU_BF16>(ta, tw, tx, ka, stream);
      case EPI_RESID_F32: return launch_inst2<EPI_RESID_F32>(ta, tw, tx, ka, stream);
      case EPI_PATCH_F32: return launch_inst2<EPI_PATCH_F32>(ta, tw, tx, ka, stream);
      case EPI_F32: return launch_inst2<EPI_F32>(ta, tw, tx, ka, stream);
      case EPI_ROWSTATS: return launch_inst2<EPI_ROWSTATS>(ta, tw, tx, ka, stream);
      case EPI_LN_BIAS_BF16: return launch_inst2<EPI_LN_BIAS_BF16>(ta, tw, tx, ka, stream);
      case EPI_LN_GELU_BF16: return launch_inst2<EPI_LN_GELU_BF16>(ta, tw, tx, ka, stream);
      case EPI_RESID_LN_F32: return launch_inst2<EPI_RESID_LN_F32>(ta, tw, tx, ka, stream);
      default: return invalid("gemm: unknown epilogue");
    }
  }
  switch (a.epilogue) {
    case EPI_BIAS_BF16: return launch_inst<256, EPI_BIAS_BF16>(ta, tw, tx, ka, stream);
    case EPI_GELU_BF16: return launch_inst<256, EPI_GELU_BF16>(ta, tw, tx, ka, stream);
    case EPI_RESID_F32: return launch_inst<256, EPI_RESID_F32>(ta, tw, tx, ka, stream);
    case EPI_PATCH_F32: return launch_inst<256, EPI_PATCH_F32>(ta, tw, tx, ka, stream);
    case EPI_F32: return launch_inst<256, EPI_F32>(ta, tw, tx, ka, stream);
    case EPI_LN_BIAS_BF16: return launch_inst<256, EPI_LN_BIAS_BF16>(ta, tw, tx, ka, stream);
    case EPI_LN_GELU_BF16: return launch_inst<256, EPI_LN_GELU_BF16>(ta, tw, tx, ka, stream);
    case EPI_ROWSTATS: return launch_inst<128, EPI_ROWSTATS>(ta, tw, tx, ka, stream);
    case EPI_COLSUM: return launch_inst<128, EPI_COLSUM>(ta, tw, tx, ka, stream);
    default: return invalid("gemm: unknown epilogue");
  }
}

}  // namespace ca

// Flash-style multi-head self-attention forward for the DINOv2 backbone on sm_100a (head_dim 64, no mask).
// Replaces `F.scaled_dot_product_attention` at HF modeling_dinov2.py:215-229 (scale 64^-0.5, softmax over keys).
//
// One CTA = one (image, head) x one tile of 128 queries; two CTAs co-reside per SM (112 KB smem, 256 TMEM
// columns each) so that one CTA's softmax (MUFU-bound) overlaps the other's tensor-core work.
//
//   warp 0     : TMA producer — Q once, then K/V tiles of 128 keys into a 2-stage ring (128B swizzle)
//   warp 1     : MMA issuer   — S = Q K^T   (tcgen05.mma 128x128x16 x4, both operands K-major)
//                               O += P V    (tcgen05.mma 128x64x16  x8, P K-major from smem, V MN-major as loaded)
//                               owns the TMEM allocation (S: 128 fp32 columns, O: 64 fp32 columns)
//   warps 2..5 : softmax      — one query row per thread: tcgen05.ld S, online max / sum in fp32 (base-2 domain),
//                               P -> bf16 -> swizzled smem, rescale O in TMEM when the running max moved,
//                               final O / l -> bf16 -> global (token-major, head h at columns [64h, 64h+64))
//
// Input is the fused QKV activation [B*T, 3*H*64] written by the QKV GEMM (Q | K | V column blocks), read in place
// through one 3-D tensor map (col, token, image): no head-major reshuffle pass exists.
#include "attention.cuh"
#include "common.cuh"
#include "host.h"

namespace ca {
namespace {

constexpr int kHeadDim = 64;
constexpr int kTileQ = 128;
constexpr int kTileK = 128;
constexpr int kKVStages = 2;
constexpr int kAttnThreads = 6 * 32;
constexpr int kTileBytes = 128 * kHeadDim * 2;                 // 16 KB: 128 rows x 128 B
constexpr int kSmemQ = 0;
constexpr int kSmemK = kSmemQ + kTileBytes;
constexpr int kSmemV = kSmemK + kKVStages * kTileBytes;
constexpr int kSmemP = kSmemV + kKVStages * kTileBytes;          // 2 chunks of 64 keys
constexpr int kSmemBar = kSmemP + 2 * kTileBytes;
constexpr int kAttnSmemBytes = kSmemBar + 128;
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kTmemS = 0;
constexpr uint32_t kTmemO = 128;

struct AttnArgs {
  int T;          // tokens per image
  int H;          // heads
  int n_kv;       // key tiles
  float scale_log2;
  __nv_bfloat16* out;  // [B*T, H*64]
  int ldo;
};

// One softmax step for one query row (thread) over a 128-key S tile held in TMEM:
//   pass 1: row max of the raw accumulators (4 independent chains, no per-element scaling)
//   pass 2: P = exp2(s*scale - max) -> bf16 -> 128B-swizzled smem (A operand of the PV MMA), fp32 row sum
// TMEM loads are software-pipelined (chunk c+1 is in flight while chunk c is processed).  kMasked is only
// instantiated for the ragged last tile, so the steady-state path carries no predicates.
template <bool kMasked>
__device__ __forceinline__ float softmax_tile(uint32_t t_s, uint8_t* p_row, int r, int valid, float scale_log2,
                                              float m_run, float& sum_out) {
  uint32_t v[2][32];
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  tmem_ld32(t_s, v[0]);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tmem_ld_wait();
    if (c < 3) tmem_ld32(t_s + (c + 1) * 32, v[(c + 1) & 1]);
    else tmem_ld32(t_s, v[0]);  // chunk 0 again for pass 2
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float x0 = __uint_as_float(v[c & 1][i + 0]), x1 = __uint_as_float(v[c & 1][i + 1]);
      float x2 = __uint_as_float(v[c & 1][i + 2]), x3 = __uint_as_float(v[c & 1][i + 3]);
      if (kMasked) {
        x0 = (c * 32 + i + 0 < valid) ? x0 : -INFINITY;
        x1 = (c * 32 + i + 1 < valid) ? x1 : -INFINITY;
        x2 = (c * 32 + i + 2 < valid) ? x2 : -INFINITY;
        x3 = (c * 32 + i + 3 < valid) ? x3 : -INFINITY;
      }
      m0 = fmaxf(m0, x0);
      m1 = fmaxf(m1, x1);
      m2 = fmaxf(m2, x2);
      m3 = fmaxf(m3, x3);
    }
  }
  const float mx = fmaxf(m_run, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * scale_log2);  // scale_log2 > 0
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tmem_ld_wait();
    if (c < 3) tmem_ld32(t_s + (c + 1) * 32, v[(c + 1) & 1]);
    uint8_t* chunk_base = p_row + (c >> 1) * kTileBytes;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        e[i] = fast_exp2(fmaf(__uint_as_float(v[c & 1][8 * t + i]), scale_log2, -mx));
        if (kMasked) e[i] = (c * 32 + 8 * t + i < valid) ? e[i] : 0.f;
      }
      s0 += e[0] + e[4];
      s1 += e[1] + e[5];
      s2 += e[2] + e[6];
      s3 += e[3] + e[7];
      uint4 w;
      w.x = pack_bf16x2(e[0], e[1]);
      w.y = pack_bf16x2(e[2], e[3]);
      w.z = pack_bf16x2(e[4], e[5]);
      w.w = pack_bf16x2(e[6], e[7]);
      *reinterpret_cast<uint4*>(chunk_base + sw128_offset(r, (c & 1) * 4 + t)) = w;
    }
  }
  sum_out = (s0 + s1) + (s2 + s3);
  return mx;
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = warp_id();
  const int lane = lane_id();
  const int qt = blockIdx.x;
  const int bh = blockIdx.y;
  const int b = bh / p.H;
  const int h = bh - b * p.H;
  const int q0 = qt * kTileQ;
  const int col_q = h * kHeadDim;
  const int col_k = (p.H + h) * kHeadDim;
  const int col_v = (2 * p.H + h) * kHeadDim;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[cogaim] attention: dynamic smem base not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kTileBytes);
      tma_load_3d(smem + kSmemQ, &tmap_qkv, q_full, col_q, q0, b);
      for (int j = 0; j < p.n_kv; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1u);
        mbar_arrive_expect_tx(&kv_full[s], 2 * kTileBytes);
        tma_load_3d(smem + kSmemK + s * kTileBytes, &tmap_qkv, &kv_full[s], col_k, j * kTileK, b);
        tma_load_3d(smem + kSmemV + s * kTileBytes, &tmap_qkv, &kv_full[s], col_v, j * kTileK, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);  // B (=V) is MN-major
      const uint32_t q_addr = smem_u32(smem + kSmemQ);
      const uint32_t p_addr = smem_u32(smem + kSmemP);
      mbar_wait(q_full, 0);
      for (int j = 0; j < p.n_kv; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint64_t qd = umma_smem_desc_sw128(q_addr);
        const uint64_t kd = umma_smem_desc_sw128(smem_u32(smem + kSmemK + s * kTileBytes));
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) umma_bf16(tmem_base + kTmemS, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
        // P (bf16, smem) and the rescaled O are published by the softmax warps
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(smem + kSmemV + s * kTileBytes);
#pragma unroll
        for (int k = 0; k < kTileK / 16; ++k) {
          const uint64_t pd = umma_smem_desc_sw128(p_addr + (k >> 2) * kTileBytes + (k & 3) * 32);
          const uint64_t vd = umma_smem_desc_sw128(v_addr + k * 2048, 1024, 1024);  // 16 keys = 2 groups of 8 rows
          umma_bf16(tmem_base + kTmemO, pd, vd, idesc_o, (j | k) != 0);
        }
        umma_commit(&kv_empty[s]);
        if (j == p.n_kv - 1) umma_commit(o_full);
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + kTmemS;
    const uint32_t t_o = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + kTmemO;
    uint8_t* p_row = smem + kSmemP;
    float m_run = -INFINITY;
    float l_run = 0.f;
    for (int j = 0; j < p.n_kv; ++j) {
      const int valid = p.T - j * kTileK;  // >= 128 on every tile but the last
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float mx, sum;
      if (valid >= kTileK) {
        mx = softmax_tile<false>(t_s, p_row, r, kTileK, p.scale_log2, m_run, sum);
      } else {
        mx = softmax_tile<true>(t_s, p_row, r, valid, p.scale_log2, m_run, sum);
      }
      const float alpha = fast_exp2(m_run - mx);  // 0 on the first tile (m_run = -inf)
      l_run = l_run * alpha + sum;
      m_run = mx;
      // ---- rescale the O accumulator when any row of this warp moved its max (warp-uniform branch) ----
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(t_o + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st32(t_o + c * 32, v);
        }
        tmem_st_wait();
      }
      fence_proxy_async_smem();  // P visible to the tensor-core (async) proxy
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 -> global ----
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const int q = q0 + r;
    __nv_bfloat16* orow = p.out + (static_cast<size_t>(b) * p.T + q) * p.ldo + h * kHeadDim;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld32(t_o + c * 32, v);
      tmem_ld_wait();
      if (q < p.T) {
        uint4* o4 = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(v[8 * t + 0]) * inv, __uint_as_float(v[8 * t + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(v[8 * t + 2]) * inv, __uint_as_float(v[8 * t + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(v[8 * t + 4]) * inv, __uint_as_float(v[8 * t + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(v[8 * t + 6]) * inv, __uint_as_float(v[8 * t + 7]) * inv);
          o4[t] = w;
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

int attention_launch(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, cudaStream_t stream) {
  CA_REQUIRE(qkv && out, "attention: null pointer");
  CA_REQUIRE(B > 0 && T > 0 && H > 0, "attention: non-positive dimension");
  const int ld = 3 * H * kHeadDim;
  CUtensorMap tm;
  CA_TRY(make_tmap_3d(&tm, qkv, B, T, ld, ld, static_cast<uint64_t>(T) * ld, 128));
  static bool configured = false;
  if (!configured) {
    CA_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
    configured = true;
  }
  AttnArgs a;
  a.T = T;
  a.H = H;
  a.n_kv = (T + kTileK - 1) / kTileK;
  a.scale_log2 = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)
  a.out = out;
  a.ldo = H * kHeadDim;
  dim3 grid((T + kTileQ - 1) / kTileQ, B * H);
  attention_fwd_kernel<<<grid, kAttnThreads, kAttnSmemBytes, stream>>>(tm, a);
  CA_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ca

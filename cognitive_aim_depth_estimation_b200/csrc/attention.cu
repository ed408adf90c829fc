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

#include <stdlib.h>

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
  int dbg;  // CA_ATTN_DEBUG timing experiments (results are garbage when non-zero)
  long long* trace;  // CA_ATTN_TRACE=1: per-tile clock64 stamps of CTA (5, 40) [n_kv][10]
};
#define CA_TRACE(slot)                                                                                  \
  if (p.trace && blockIdx.x == 5 && blockIdx.y == 40) p.trace[j * 10 + (slot)] = clock64();


// One softmax step for one query row (thread) over a 128-key S tile held in TMEM:
//   pass 1: row max of the raw accumulators (4 independent chains, no per-element scaling)
//   pass 2: P = exp2(s*scale - max) -> bf16 -> 128B-swizzled smem (A operand of the PV MMA), fp32 row sum
// TMEM loads are software-pipelined (chunk c+1 is in flight while chunk c is processed).  kMasked is only
// instantiated for the ragged last tile, so the steady-state path carries no predicates.
template <bool kMasked>
__device__ __forceinline__ float softmax_tile(uint32_t t_s, uint8_t* p_row, int r, int valid, float scale_log2,
                                              float m_run, float& sum_out, int dbg, uint64_t* s_free,
                                              uint64_t* pv_done, int j) {
  uint32_t v[2][32];
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  tmem_ld32(t_s, v[0]);
  if (dbg & 1) { m0 = 0.f; }
  else
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
  if (dbg & 1) { tmem_ld_wait(); tmem_ld32(t_s, v[0]); }
  const float mx = fmaxf(m_run, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * scale_log2);  // scale_log2 > 0
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tmem_ld_wait();
    if (c < 3) {
      tmem_ld32(t_s + (c + 1) * 32, v[(c + 1) & 1]);
    } else {  // S_j is now entirely in registers: let the MMA warp overwrite it with S_{j+1}
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(s_free);
    }
    if (c == 0 && j > 0) mbar_wait(pv_done, (j - 1) & 1);  // P_{j-1} V_{j-1} finished: P buffer and O are ours again
    uint8_t* chunk_base = p_row + (c >> 1) * kTileBytes;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        e[i] = (dbg & 2) ? fmaf(__uint_as_float(v[c & 1][8 * t + i]), scale_log2, -mx)
                         : fast_exp2(fmaf(__uint_as_float(v[c & 1][8 * t + i]), scale_log2, -mx));
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
      if (!(dbg & 4)) *reinterpret_cast<uint4*>(chunk_base + sw128_offset(r, (c & 1) * 4 + t)) = w;
    }
  }
  sum_out = (s0 + s1) + (s2 + s3);
  return mx;
}

// Steady-state softmax step: ONE pass over the S tile (TMEM read bandwidth, 64 B/clk/SM, is as scarce here as the
// MUFU).  The exponent reference `m_use` is the running max of the PREVIOUS tiles, so exp2 needs no max pass; the
// max of this tile is gathered on the side for the next tile.  Returns false — without having signalled s_free —
// when some row of the warp exceeds the reference by more than 2^kLagLimit; the caller then redoes the tile with
// the exact two-pass path (always the case for the first tile, whose reference is -inf).
constexpr float kLagLimit = 16.0f;
template <bool kMasked>
__device__ __forceinline__ bool softmax_tile_fast(uint32_t t_s, uint8_t* p_row, int r, int valid, float scale_log2,
                                                  float m_use, float& sum_out, float& tile_max, uint64_t* s_free,
                                                  uint64_t* pv_done, int j) {
  uint32_t v[2][32];
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  bool ok = true;
  tmem_ld32(t_s, v[0]);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tmem_ld_wait();
    if (c < 3) tmem_ld32(t_s + (c + 1) * 32, v[(c + 1) & 1]);
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
    if (c == 3) {  // whole tile seen: decide, and hand S back to the tensor core as early as possible
      tile_max = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * scale_log2;
      ok = !__any_sync(0xffffffffu, tile_max > m_use + kLagLimit);
      if (!ok) return false;
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(s_free);
    }
    if (c == 0) mbar_wait(pv_done, (j - 1) & 1);  // P_{j-1} V_{j-1} finished: the P buffer is ours again (j > 0 here)
    uint8_t* chunk_base = p_row + (c >> 1) * kTileBytes;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        e[i] = fast_exp2(fmaf(__uint_as_float(v[c & 1][8 * t + i]), scale_log2, -m_use));
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
  return true;
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;   // MMA -> softmax : S_j is in TMEM
  uint64_t* p_full = bars + 6;   // softmax -> MMA : P_j is in smem and O is rescaled
  uint64_t* o_full = bars + 7;   // MMA -> softmax : last P V finished
  uint64_t* s_free = bars + 8;   // softmax -> MMA : S_j has been read out of TMEM (S_{j+1} may overwrite it)
  uint64_t* pv_done = bars + 9;  // MMA -> softmax : P_j V_j finished (P buffer and O may be touched again)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

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
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    mbar_init(s_free, 4);
    mbar_init(pv_done, 1);
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
    // All 32 lanes walk the protocol (waits are cheap); the tensor-core instructions are issued by one elected lane.
    // Order on the (in-order) tensor pipe:  S_0 | S_1 PV_0 | S_2 PV_1 | ...   S_{j+1} is issued as soon as the softmax
    // warps have pulled S_j out of TMEM, so it runs under their exp / store phase and is off the critical path.
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);  // B (=V) is MN-major
    const uint64_t qd = umma_smem_desc_sw128(smem_u32(smem + kSmemQ));
    const uint64_t kd0 = umma_smem_desc_sw128(smem_u32(smem + kSmemK));
    const uint64_t pd0 = umma_smem_desc_sw128(smem_u32(smem + kSmemP));
    const uint64_t vd0 = umma_smem_desc_sw128(smem_u32(smem + kSmemV), 1024, 1024);
    constexpr uint64_t kStageStep = kTileBytes >> 4;  // descriptor address units are 16 bytes
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    if (elect_one_sync()) {
#pragma unroll
      for (int k = 0; k < kHeadDim / 16; ++k) umma_bf16(tmem_base + kTmemS, qd + 2 * k, kd0 + 2 * k, idesc_s, k != 0);
      umma_commit(s_full);
    }
    __syncwarp();
    for (int j = 0; j < p.n_kv; ++j) {
      const int s = j & 1;
      if (j + 1 < p.n_kv) {
        mbar_wait(&kv_full[s ^ 1], ((j + 1) >> 1) & 1);
        mbar_wait(s_free, j & 1);
        tc_fence_after();
        CA_TRACE(0)
        if (elect_one_sync()) {
          const uint64_t kd = kd0 + (s ^ 1) * kStageStep;
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k) umma_bf16(tmem_base + kTmemS, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
          umma_commit(s_full);
        }
        __syncwarp();
        CA_TRACE(1)
      }
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      CA_TRACE(2)
      if (elect_one_sync()) {
        const uint64_t vd = vd0 + s * kStageStep;
#pragma unroll
        for (int k = 0; k < kTileK / 16; ++k) {
          // P: 64-key chunk (k >> 2), +32 B per 16 keys inside the swizzle atom; V: 16 keys = 2048 B further down
          umma_bf16(tmem_base + kTmemO, pd0 + (k >> 2) * kStageStep + 2 * (k & 3), vd + k * (2048 >> 4), idesc_o,
                    (j | k) != 0);
        }
        umma_commit(&kv_empty[s]);
        umma_commit(pv_done);
        if (j == p.n_kv - 1) umma_commit(o_full);
      }
      __syncwarp();
      CA_TRACE(3)
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + kTmemS;
    const uint32_t t_o = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + kTmemO;
    uint8_t* p_row = smem + kSmemP;
    float m_hint = -INFINITY;  // running row max over the tiles seen so far
    float m_acc = -INFINITY;   // reference the O accumulator and l_run are currently expressed in
    float l_run = 0.f;
    for (int j = 0; j < p.n_kv; ++j) {
      const int valid = p.T - j * kTileK;  // >= 128 on every tile but the last
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (threadIdx.x == 64) { CA_TRACE(4) }
      // Exponent reference of this tile: the running max of the previous tiles (single-pass fast path), or the exact
      // max including this tile (two-pass path: first tile, or a row jumped by more than 2^kLagLimit).
      float m_use = m_hint, sum, tile_max = -INFINITY;
      bool done = false;
      if (j > 0 && !(p.dbg & 16)) {
        done = (valid >= kTileK)
                   ? softmax_tile_fast<false>(t_s, p_row, r, kTileK, p.scale_log2, m_use, sum, tile_max, s_free, pv_done, j)
                   : softmax_tile_fast<true>(t_s, p_row, r, valid, p.scale_log2, m_use, sum, tile_max, s_free, pv_done, j);
      }
      if (!done) {
        m_use = (valid >= kTileK)
                    ? softmax_tile<false>(t_s, p_row, r, kTileK, p.scale_log2, m_hint, sum, p.dbg, s_free, pv_done, j)
                    : softmax_tile<true>(t_s, p_row, r, valid, p.scale_log2, m_hint, sum, p.dbg, s_free, pv_done, j);
        tile_max = m_use;
      }
      const float alpha = fast_exp2(m_acc - m_use);  // O and l are expressed relative to m_acc; 0 on the first tile
      if (threadIdx.x == 64) { CA_TRACE(5) }
      l_run = l_run * alpha + sum;
      m_acc = m_use;
      m_hint = fmaxf(m_hint, tile_max);
      // ---- rescale the O accumulator when any row of this warp moved its max (warp-uniform branch) ----
      if (j > 0 && !(p.dbg & 8) && __any_sync(0xffffffffu, alpha != 1.0f)) {
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
      if (threadIdx.x == 64) { CA_TRACE(6) }
      fence_proxy_async_smem();  // P visible to the tensor-core (async) proxy
      if (threadIdx.x == 64) { CA_TRACE(7) }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (threadIdx.x == 64) { CA_TRACE(8) }
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
  static const int dbg = getenv("CA_ATTN_DEBUG") ? atoi(getenv("CA_ATTN_DEBUG")) : 0;
  a.dbg = dbg;
  a.trace = nullptr;
  static const bool want_trace = getenv("CA_ATTN_TRACE") != nullptr;
  static long long* d_trace = nullptr;
  if (want_trace) {
    if (!d_trace) CA_CUDA(cudaMalloc(&d_trace, 64 * 10 * sizeof(long long)));
    CA_CUDA(cudaMemsetAsync(d_trace, 0, 64 * 10 * sizeof(long long), stream));
    a.trace = d_trace;
  }
  dim3 grid((T + kTileQ - 1) / kTileQ, B * H);
  attention_fwd_kernel<<<grid, kAttnThreads, kAttnSmemBytes, stream>>>(tm, a);
  CA_CUDA(cudaGetLastError());
  if (want_trace) {
    static int dumps = 0;
    long long h[64 * 10];
    CA_CUDA(cudaStreamSynchronize(stream));
    CA_CUDA(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
    if (dumps++ == 4) {
      const long long t0 = h[0];
      for (int j = 0; j < a.n_kv && j < 64; ++j) {
        fprintf(stderr, "tile %2d:", j);
        for (int k = 0; k < 9; ++k) fprintf(stderr, " %7lld", h[j * 10 + k] - t0);
        fprintf(stderr, "\n");
      }
    }
  }
  return 0;
}

}  // namespace ca

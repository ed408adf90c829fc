// Flash-style multi-head self-attention forward for the DINOv2 backbone on sm_100a (head_dim 64, no mask).
// Replaces `F.scaled_dot_product_attention` at HF modeling_dinov2.py:215-229 (scale 64^-0.5, softmax over keys).
//
// Two budgets shape the kernel.  (1) One exp2 per score: the MUFU (16 /clk/SM) needs 1024 clk per 128x128 score tile,
// the tensor core 512.  (2) Shared-memory bandwidth (128 B/clk/SM): a 128x64x16 tcgen05.mma with both operands in smem
// reads 6 KB in 32 clk = 192 B/clk, so an "SS" kernel is smem-bound before it is MUFU-bound.  Hence BOTH A operands
// live in TMEM: Q (stored once per CTA) and P (written by the softmax warps over the S columns they just read);
// shared memory only carries the K / V stream (TMA in, one tensor-core read each).
//
// One CTA = one (image, head) x one tile of 128 queries; two CTAs co-reside per SM (64 KB smem, 256 TMEM columns
// each).  The key axis is walked in steps of 64 keys, S double-buffered in TMEM:
//
//   warp 0     : TMA producer — K and V steps (64 keys = 8 KB each) into two independent 4-slot rings
//                (K runs two steps ahead of V, the order the tensor core consumes them in)
//   warp 1     : MMA issuer   — S_i = Q K_i^T  (tcgen05.mma 128x64x16 x4, A = Q in TMEM) into S buffer i&1
//                               O += P_i V_i   (tcgen05.mma 128x64x16 x4, A = P_i in TMEM, V MN-major as loaded)
//                               issue order S_0 S_1 | PV_0 S_2 | PV_1 S_3 | ...  (S_{i+2} overwrites the buffer that held
//                               S_i / P_i; the tensor pipe executes in issue order, so it follows PV_i's operand reads)
//   warps 2..5 : softmax      — one query row per thread: Q row global -> TMEM once; per step the 64-wide S_i row is
//                               pulled into registers (two tcgen05.ld), exact row max in registers, the accumulator
//                               reference only moves when a row grew by more than 2^8 (lazy rescale),
//                               P = exp2(s*scale - ref) -> bf16x2 -> tcgen05.st over the first 32 columns of the buffer.
//
// Input is the fused QKV activation [B*T, 3*H*64] written by the QKV GEMM (Q | K | V column blocks), read in place
// (K, V through a 3-D tensor map (col, token, image)): no head-major reshuffle pass exists.
#include "attention.cuh"
#include "common.cuh"
#include "host.h"

#include <stdlib.h>

namespace ca {
namespace {

constexpr int kHeadDim = 64;
constexpr int kTileQ = 128;
constexpr int kSubK = 64;                          // keys per pipeline step
constexpr int kRing = 4;                           // K ring slots == V ring slots
constexpr int kAttnThreads = 6 * 32;
constexpr int kSubBytes = kSubK * kHeadDim * 2;    // 8 KB: 64 rows x 128 B
constexpr int kSmemK = 0;
constexpr int kSmemV = kSmemK + kRing * kSubBytes;
constexpr int kSmemBar = kSmemV + kRing * kSubBytes;
constexpr int kAttnSmemBytes = kSmemBar + 256;
static_assert(2 * (kAttnSmemBytes + 1024) <= 227 * 1024, "two CTAs must fit one SM");
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kTmemS = 0;      // two S buffers of 64 fp32 columns; P_i (bf16x2) overwrites the first 32 of buffer i&1
constexpr uint32_t kTmemO = 128;    // 64 fp32 columns
constexpr uint32_t kTmemQ = 192;    // 32 columns: Q tile as bf16x2 (A operand of every S MMA)
constexpr float kLazyLimit = 8.0f;  // the accumulator reference moves only when a row max grew by > 2^8

struct AttnArgs {
  int T;          // tokens per image
  int H;          // heads
  int n_sub;      // 64-key steps
  float scale_log2;
  const __nv_bfloat16* qkv;  // [B*T, 3*H*64]
  int ld;
  __nv_bfloat16* out;  // [B*T, H*64]
  int ldo;
};

// d = a * s + c on two packed fp32 lanes (FFMA2)
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float s, float c) {
  asm("{\n\t.reg .b64 ra, rs, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rs, {%4, %4};\n\tmov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rs, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(s), "f"(c));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1) {
  asm("{\n\t.reg .b64 ra, rd;\n\t"
      "mov.b64 rd, {%0, %1};\n\tmov.b64 ra, {%2, %3};\n\t"
      "add.rn.f32x2 rd, rd, ra;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1));
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_kv, const AttnArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
  uint64_t* k_full = bars + 0;     // [4] TMA -> MMA
  uint64_t* k_empty = bars + 4;    // [4] MMA (commit) -> TMA
  uint64_t* v_full = bars + 8;     // [4]
  uint64_t* v_empty = bars + 12;   // [4]
  uint64_t* q_ready = bars + 16;   // softmax -> MMA : Q is in TMEM
  uint64_t* s_full = bars + 17;    // [2] MMA -> softmax : S_i is in TMEM buffer i&1
  uint64_t* p_full = bars + 19;    // [2] softmax -> MMA : P_i is in TMEM (over S_i) and O is rescaled
  uint64_t* pv_done = bars + 21;   // [2] MMA -> softmax : P_i V_i finished (O quiescent up to step i)
  uint64_t* o_full = bars + 23;    // MMA -> softmax : last P V finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = warp_id();
  const int lane = lane_id();
  const int qt = blockIdx.x;
  const int bh = blockIdx.y;
  const int b = bh / p.H;
  const int h = bh - b * p.H;
  const int q0 = qt * kTileQ;
  const int col_k = (p.H + h) * kHeadDim;
  const int col_v = (2 * p.H + h) * kHeadDim;
  const int n_sub = p.n_sub;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[cogaim] attention: dynamic smem base not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmap_kv);
    for (int s = 0; s < kRing; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(q_ready, 4);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);
      mbar_init(&pv_done[s], 1);
    }
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
      auto load_k = [&](int i) {
        const int s = i & (kRing - 1);
        mbar_wait(&k_empty[s], ((i / kRing) & 1) ^ 1u);
        mbar_arrive_expect_tx(&k_full[s], kSubBytes);
        tma_load_3d(smem + kSmemK + s * kSubBytes, &tmap_kv, &k_full[s], col_k, i * kSubK, b);
      };
      load_k(0);
      if (n_sub > 1) load_k(1);
      for (int i = 0; i < n_sub; ++i) {
        const int s = i & (kRing - 1);
        mbar_wait(&v_empty[s], ((i / kRing) & 1) ^ 1u);
        mbar_arrive_expect_tx(&v_full[s], kSubBytes);
        tma_load_3d(smem + kSmemV + s * kSubBytes, &tmap_kv, &v_full[s], col_v, i * kSubK, b);
        if (i + 2 < n_sub) load_k(i + 2);
      }
    }
  } else if (warp == 1) {
    // All 32 lanes walk the protocol (waits are cheap); the tensor-core instructions are issued by one elected lane.
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, kSubK, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);  // B (=V) is MN-major
    const uint64_t kd0 = umma_smem_desc_sw128(smem_u32(smem + kSmemK));
    const uint64_t vd0 = umma_smem_desc_sw128(smem_u32(smem + kSmemV), 1024, 1024);
    constexpr uint64_t kSubStep = kSubBytes >> 4;  // descriptor address units are 16 bytes
    const uint32_t t_q = tmem_base + kTmemQ;
    auto issue_s = [&](int i) {
      const int s = i & (kRing - 1);
      mbar_wait(&k_full[s], (i / kRing) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t kd = kd0 + s * kSubStep;
        const uint32_t d = tmem_base + kTmemS + (i & 1) * kSubK;
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) umma_bf16_ts(d, t_q + 8 * k, kd + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[i & 1]);
        umma_commit(&k_empty[s]);
      }
      __syncwarp();
    };
    mbar_wait(q_ready, 0);
    issue_s(0);
    if (n_sub > 1) issue_s(1);
    for (int i = 0; i < n_sub; ++i) {
      const int bb = i & 1;
      const int s = i & (kRing - 1);
      mbar_wait(&v_full[s], (i / kRing) & 1);
      mbar_wait(&p_full[bb], (i >> 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t vd = vd0 + s * kSubStep;
        const uint32_t t_p = tmem_base + kTmemS + bb * kSubK;
#pragma unroll
        for (int k = 0; k < kSubK / 16; ++k) {
          // P: 16 keys = 8 packed columns; V: 16 keys = 16 rows of 128 B = 2048 B further down
          umma_bf16_ts(tmem_base + kTmemO, t_p + 8 * k, vd + k * (2048 >> 4), idesc_o, (i | k) != 0);
        }
        umma_commit(&pv_done[bb]);
        umma_commit(&v_empty[s]);
        if (i == n_sub - 1) umma_commit(o_full);
      }
      __syncwarp();
      if (i + 2 < n_sub) issue_s(i + 2);  // overwrites S_i / P_i: ordered behind PV_i by the in-order tensor pipe
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // query row inside the tile == TMEM lane
    const int q = q0 + r;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t t_o = t_lane + kTmemO;
    const float scale = p.scale_log2;
    // ---- Q row: global -> registers -> TMEM (bf16x2 per column); rows past the end of the image are zero ----
    {
      uint32_t qv[32];
      if (q < p.T) {
        const uint4* src = reinterpret_cast<const uint4*>(p.qkv + (static_cast<size_t>(b) * p.T + q) * p.ld + h * kHeadDim);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint4 w = __ldg(src + t);
          qv[4 * t + 0] = w.x;
          qv[4 * t + 1] = w.y;
          qv[4 * t + 2] = w.z;
          qv[4 * t + 3] = w.w;
        }
      } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) qv[t] = 0u;
      }
      tmem_st32(t_lane + kTmemQ, qv);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_ready);
    }
    float m_acc = -INFINITY;  // reference (log2 domain) the O accumulator and l_run are expressed in
    float l_run = 0.f;
    for (int i = 0; i < n_sub; ++i) {
      const int bb = i & 1;
      const int valid = p.T - i * kSubK;  // >= 64 on every step but the last
      const uint32_t t_s = t_lane + kTmemS + bb * kSubK;
      uint32_t v[64];
      mbar_wait(&s_full[bb], (i >> 1) & 1);
      tc_fence_after();
      {
        uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
        uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
        tmem_ld32(t_s, v0);
        tmem_ld32(t_s + 32, v1);
        tmem_ld_wait();
      }
      if (valid < kSubK) {
#pragma unroll
        for (int c = 0; c < kSubK; ++c)
          if (c >= valid) v[c] = 0xff800000u;  // -inf: exp2 -> 0, ignored by the max
      }
      // ---- exact row max, four independent chains ----
      float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]),
            m3 = __uint_as_float(v[3]);
#pragma unroll
      for (int c = 4; c < kSubK; c += 8) {
        m0 = fmaxf(m0, fmaxf(__uint_as_float(v[c + 0]), __uint_as_float(v[c + 1])));
        m1 = fmaxf(m1, fmaxf(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])));
        if (c + 4 < kSubK) {
          m2 = fmaxf(m2, fmaxf(__uint_as_float(v[c + 4]), __uint_as_float(v[c + 5])));
          m3 = fmaxf(m3, fmaxf(__uint_as_float(v[c + 6]), __uint_as_float(v[c + 7])));
        }
      }
      const float tile_max = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * scale;  // scale > 0
      // ---- lazy reference update (warp-uniform decision; always taken on the first step) ----
      if (__any_sync(0xffffffffu, tile_max > m_acc + kLazyLimit)) {
        const float m_new = fmaxf(m_acc, tile_max);
        const float alpha = fast_exp2(m_acc - m_new);  // 0 on the first step (m_acc = -inf)
        l_run *= alpha;
        m_acc = m_new;
        if (i > 0) {
          mbar_wait(&pv_done[bb ^ 1], ((i - 1) >> 1) & 1);  // P_{i-1} V_{i-1} (hence every earlier one) has finished
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[16];
            tmem_ld16(t_o + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
            tmem_st16(t_o + c * 16, o);
          }
        }
      }
      // ---- P = exp2(s * scale - m_acc) -> bf16x2 -> TMEM (over the S columns just read) ; fp32 row sum ----
      const float neg_m = -m_acc;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int t = 0; t < kSubK / 8; ++t) {
        float e[8];
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          ffma2(e[k], e[k + 1], __uint_as_float(v[8 * t + k]), __uint_as_float(v[8 * t + k + 1]), scale, neg_m);
          e[k] = fast_exp2(e[k]);
          e[k + 1] = fast_exp2(e[k + 1]);
        }
        fadd2(s0, s1, e[0], e[1]);
        fadd2(s2, s3, e[2], e[3]);
        fadd2(s0, s1, e[4], e[5]);
        fadd2(s2, s3, e[6], e[7]);
        pk[4 * t + 0] = pack_bf16x2(e[0], e[1]);
        pk[4 * t + 1] = pack_bf16x2(e[2], e[3]);
        pk[4 * t + 2] = pack_bf16x2(e[4], e[5]);
        pk[4 * t + 3] = pack_bf16x2(e[6], e[7]);
      }
      tmem_st32(t_s, pk);
      l_run += (s0 + s1) + (s2 + s3);
      tmem_st_wait();     // P (and a rescaled O) are in TMEM
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[bb]);
    }
    // ---- epilogue: O / l -> bf16 -> global ----
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    __nv_bfloat16* orow = p.out + (static_cast<size_t>(b) * p.T + q) * p.ldo + h * kHeadDim;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(t_o + c * 32, o);
      tmem_ld_wait();
      if (q < p.T) {
        uint4* o4 = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[8 * t + 0]) * inv, __uint_as_float(o[8 * t + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[8 * t + 2]) * inv, __uint_as_float(o[8 * t + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[8 * t + 4]) * inv, __uint_as_float(o[8 * t + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[8 * t + 6]) * inv, __uint_as_float(o[8 * t + 7]) * inv);
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
  CUtensorMap tm_kv;
  CA_TRY(make_tmap_3d(&tm_kv, qkv, B, T, ld, ld, static_cast<uint64_t>(T) * ld, kSubK));
  static bool configured = false;
  if (!configured) {
    CA_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
    configured = true;
  }
  AttnArgs a;
  a.T = T;
  a.H = H;
  a.n_sub = (T + kSubK - 1) / kSubK;
  a.scale_log2 = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)
  a.qkv = qkv;
  a.ld = ld;
  a.out = out;
  a.ldo = H * kHeadDim;
  dim3 grid((T + kTileQ - 1) / kTileQ, B * H);
  attention_fwd_kernel<<<grid, kAttnThreads, kAttnSmemBytes, stream>>>(tm_kv, a);
  CA_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ca

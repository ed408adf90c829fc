// Flash-style multi-head self-attention forward for the DINOv2 backbone on sm_100a (head_dim 64, no mask).
// Replaces `F.scaled_dot_product_attention` at HF modeling_dinov2.py:215-229 (scale 64^-0.5, softmax over keys).
//
// The kernel is MUFU-bound (one exp2 per score: 16 /clk/SM against 8192 tensor flops/clk/SM), so the design goal
// is to keep FOUR softmax warps resident per SM sub-partition, each running a short serial chain, instead of
// making one warp fast.  One CTA = one (image, head) x one tile of 128 queries; two CTAs co-reside per SM
// (112 KB smem, 256 TMEM columns each).  Inside a CTA the key axis is cut into 64-key steps which are dealt
// alternately to two independent STREAMS (even steps / odd steps).  A stream owns an S buffer and an O accumulator
// in TMEM, a P buffer in smem, its own softmax reference (m, l), its own MMA-issuing warp and softmax warpgroup;
// the two partial results are merged in the epilogue (flash-decoding style split over keys, inside the CTA).
//
//   warp 0      : TMA producer — Q once; K and V steps (64 keys = 8 KB each) into two 4-slot rings
//                 (K runs two steps ahead of V, the order the tensor core consumes them in)
//   warp 1+s    : MMA issuer of stream s — S_i = Q K_i^T (tcgen05.mma 128x64x16 x4) as soon as the softmax warps have
//                 pulled S_{i-2} out of TMEM; O_s += P_i V_i (128x64x16 x4, P K-major from smem, V MN-major as loaded)
//   warp 3      : idle (keeps the control warps one aligned warpgroup for setmaxnreg)
//   warps 4+4s..: softmax of stream s — one query row per thread: the 64-wide S_i row is pulled into registers and
//                 the TMEM buffer handed back at once; exact row max in registers; the accumulator reference moves
//                 only when a row grew by more than 2^8 (lazy rescale); P = exp2(s*scale - ref) -> bf16 -> swizzled
//                 smem.  Registers are re-budgeted with setmaxnreg (control warps 32, softmax warps 104).
//
// Input is the fused QKV activation [B*T, 3*H*64] written by the QKV GEMM (Q | K | V column blocks), read in place
// through 3-D tensor maps (col, token, image): no head-major reshuffle pass exists.
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
constexpr int kAttnThreads = 12 * 32;
constexpr int kQBytes = kTileQ * kHeadDim * 2;     // 16 KB: 128 rows x 128 B
constexpr int kSubBytes = kSubK * kHeadDim * 2;    // 8 KB: 64 rows x 128 B
constexpr int kPBytes = kTileQ * kSubK * 2;        // 16 KB: 128 rows x 128 B (64 keys)
constexpr int kSmemQ = 0;
constexpr int kSmemK = kSmemQ + kQBytes;
constexpr int kSmemV = kSmemK + kRing * kSubBytes;
constexpr int kSmemP = kSmemV + kRing * kSubBytes;   // one P buffer per stream
constexpr int kSmemStats = kSmemQ;                   // epilogue only (Q is dead by then): float2 (m, l) [2][128]
constexpr int kSmemBar = kSmemP + 2 * kPBytes;
constexpr int kAttnSmemBytes = kSmemBar + 256;
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kTmemS = 0;      // S buffer of stream s: 64 fp32 columns at 64 s
constexpr uint32_t kTmemO = 128;    // O accumulator of stream s: 64 fp32 columns at 128 + 64 s
constexpr float kLazyLimit = 8.0f;  // the accumulator reference moves only when a row max grew by > 2^8
static_assert(2 * (kAttnSmemBytes + 1024) <= 227 * 1024, "two CTAs must fit one SM");
constexpr int kCtrlRegs = 32;
constexpr int kSoftmaxRegs = 104;

struct AttnArgs {
  int T;          // tokens per image
  int H;          // heads
  int n_sub;      // 64-key steps
  float scale_log2;
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
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     const AttnArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;     // [4] TMA -> MMA
  uint64_t* k_empty = bars + 5;    // [4] MMA (commit) -> TMA
  uint64_t* v_full = bars + 9;     // [4]
  uint64_t* v_empty = bars + 13;   // [4]
  uint64_t* s_full = bars + 17;    // [stream] MMA -> softmax : S of the stream's current step is in TMEM
  uint64_t* s_free = bars + 19;    // [stream] softmax -> MMA : that S is in registers, the buffer may be overwritten
  uint64_t* p_full = bars + 21;    // [stream] softmax -> MMA : P is in smem (and O_s is rescaled)
  uint64_t* p_free = bars + 23;    // [stream] MMA -> softmax : P V finished (P buffer reusable, O_s quiescent)
  uint64_t* o_full = bars + 25;    // [stream] MMA -> softmax : the stream's last P V finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 27);

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
  const int n_sub = p.n_sub;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[cogaim] attention: dynamic smem base not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    mbar_init(q_full, 1);
    for (int s = 0; s < kRing; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_free[s], 4);
      mbar_init(&p_full[s], 4);
      mbar_init(&p_free[s], 1);
      mbar_init(&o_full[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    setmaxnreg_dec<kCtrlRegs>();
    if (warp == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, kQBytes);
        tma_load_3d(smem + kSmemQ, &tmap_q, q_full, col_q, q0, b);
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
    } else if (warp <= 2) {
      // MMA issuer of stream st.  All 32 lanes walk the protocol; the tensor-core instructions are issued by one
      // elected lane.  Issue order per stream:  S_0 | S_1 PV_0 | S_2 PV_1 | ...  (steps counted inside the stream).
      const int st = warp - 1;
      const int n_st = (n_sub - st + 1) >> 1;  // steps of this stream
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kSubK, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);  // B (=V) is MN-major
      const uint64_t qd = umma_smem_desc_sw128(smem_u32(smem + kSmemQ));
      const uint64_t kd0 = umma_smem_desc_sw128(smem_u32(smem + kSmemK));
      const uint64_t pd = umma_smem_desc_sw128(smem_u32(smem + kSmemP + st * kPBytes));
      const uint64_t vd0 = umma_smem_desc_sw128(smem_u32(smem + kSmemV), 1024, 1024);
      constexpr uint64_t kSubStep = kSubBytes >> 4;  // descriptor address units are 16 bytes
      const uint32_t d_s = tmem_base + kTmemS + st * kSubK;
      const uint32_t d_o = tmem_base + kTmemO + st * kHeadDim;
      auto issue_s = [&](int i) {
        const int s = i & (kRing - 1);
        mbar_wait(&k_full[s], (i / kRing) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t kd = kd0 + s * kSubStep;
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k) umma_bf16(d_s, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
          umma_commit(&s_full[st]);
          umma_commit(&k_empty[s]);
        }
        __syncwarp();
      };
      if (n_st > 0) {
        mbar_wait(q_full, 0);
        issue_s(st);
      }
      for (int n = 0; n < n_st; ++n) {
        const int i = 2 * n + st;
        const int s = i & (kRing - 1);
        if (n + 1 < n_st) {
          mbar_wait(&s_free[st], n & 1);  // S_i has been read out of TMEM
          issue_s(i + 2);
        }
        mbar_wait(&v_full[s], (i / kRing) & 1);
        mbar_wait(&p_full[st], n & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t vd = vd0 + s * kSubStep;
#pragma unroll
          for (int k = 0; k < kSubK / 16; ++k) {
            // P: +32 B per 16 keys inside the swizzle atom; V: 16 keys = 16 rows of 128 B = 2048 B further down
            umma_bf16(d_o, pd + 2 * k, vd + k * (2048 >> 4), idesc_o, (n | k) != 0);
          }
          umma_commit(&p_free[st]);
          umma_commit(&v_empty[s]);
          if (n == n_st - 1) umma_commit(&o_full[st]);
        }
        __syncwarp();
      }
    }
  } else {
    setmaxnreg_inc<kSoftmaxRegs>();
    const int st = (warp - 4) >> 2;
    const int n_st = (n_sub - st + 1) >> 1;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t t_s = t_lane + kTmemS + st * kSubK;
    const uint32_t t_o = t_lane + kTmemO + st * kHeadDim;
    uint8_t* p_buf = smem + kSmemP + st * kPBytes;
    const float scale = p.scale_log2;
    float m_acc = -INFINITY;  // reference (log2 domain) O_s and l_run are expressed in
    float l_run = 0.f;
    for (int n = 0; n < n_st; ++n) {
      const int valid = p.T - (2 * n + st) * kSubK;  // >= 64 on every step but the last
      uint32_t v[64];
      mbar_wait(&s_full[st], n & 1);
      tc_fence_after();
      {
        uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
        uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
        tmem_ld32(t_s, v0);
        tmem_ld32(t_s + 32, v1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[st]);  // S lives in registers now: the stream's next S may overwrite it
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
      if (n > 0) mbar_wait(&p_free[st], (n - 1) & 1);  // previous P V of the stream finished: P buffer ours, O_s quiet
      // ---- lazy reference update (warp-uniform decision; always taken on the first step) ----
      if (__any_sync(0xffffffffu, tile_max > m_acc + kLazyLimit)) {
        const float m_new = fmaxf(m_acc, tile_max);
        const float alpha = fast_exp2(m_acc - m_new);  // 0 on the first step (m_acc = -inf)
        l_run *= alpha;
        m_acc = m_new;
        if (n > 0) {
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
          tmem_st_wait();
        }
      }
      // ---- P = exp2(s * scale - m_acc) -> bf16 -> swizzled smem ; fp32 row sum ----
      const float neg_m = -m_acc;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
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
        uint4 w;
        w.x = pack_bf16x2(e[0], e[1]);
        w.y = pack_bf16x2(e[2], e[3]);
        w.z = pack_bf16x2(e[4], e[5]);
        w.w = pack_bf16x2(e[6], e[7]);
        *reinterpret_cast<uint4*>(p_buf + sw128_offset(r, t)) = w;
      }
      l_run += (s0 + s1) + (s2 + s3);
      fence_proxy_async_smem();  // P visible to the tensor-core (async) proxy
      tc_fence_before();         // orders the O rescale (tcgen05.st) before the PV issue
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[st]);
    }
    // ---- epilogue: merge the two streams, O / l -> bf16 -> global.  Stream st writes head dims [32 st, 32 st + 32). ----
    float2* stats = reinterpret_cast<float2*>(smem + kSmemStats);
    const int n_other = (n_sub - (st ^ 1) + 1) >> 1;
    if (n_st > 0) mbar_wait(&o_full[st], 0);  // every tensor-core op of the stream has finished (Q is dead, O_s final)
    if (n_other > 0) mbar_wait(&o_full[st ^ 1], 0);
    tc_fence_after();
    stats[st * kTileQ + r] = make_float2(m_acc, l_run);
    named_bar_sync(1, 8 * 32);
    const float2 other = stats[(st ^ 1) * kTileQ + r];
    const float m_all = fmaxf(m_acc, other.x);
    const float a_self = fast_exp2(m_acc - m_all), a_other = fast_exp2(other.x - m_all);
    const float inv = 1.0f / (l_run * a_self + other.y * a_other);
    const float w_self = a_self * inv, w_other = a_other * inv;
    const int q = q0 + r;
    uint32_t os[32], oo[32];
    tmem_ld32(t_lane + kTmemO + st * kHeadDim + st * 32, os);
    if (n_other > 0) tmem_ld32(t_lane + kTmemO + (st ^ 1) * kHeadDim + st * 32, oo);
    tmem_ld_wait();
    if (n_other == 0 || n_st == 0) {  // an empty stream's accumulator was never written
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        if (n_other == 0) oo[k] = 0u;
        if (n_st == 0) os[k] = 0u;
      }
    }
    if (q < p.T) {
      __nv_bfloat16* orow = p.out + (static_cast<size_t>(b) * p.T + q) * p.ldo + h * kHeadDim + st * 32;
      uint4* o4 = reinterpret_cast<uint4*>(orow);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float f[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          f[k] = __uint_as_float(os[8 * t + k]) * w_self + __uint_as_float(oo[8 * t + k]) * w_other;
        uint4 w;
        w.x = pack_bf16x2(f[0], f[1]);
        w.y = pack_bf16x2(f[2], f[3]);
        w.z = pack_bf16x2(f[4], f[5]);
        w.w = pack_bf16x2(f[6], f[7]);
        o4[t] = w;
      }
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
  CUtensorMap tm_q, tm_kv;
  CA_TRY(make_tmap_3d(&tm_q, qkv, B, T, ld, ld, static_cast<uint64_t>(T) * ld, kTileQ));
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
  a.out = out;
  a.ldo = H * kHeadDim;
  dim3 grid((T + kTileQ - 1) / kTileQ, B * H);
  attention_fwd_kernel<<<grid, kAttnThreads, kAttnSmemBytes, stream>>>(tm_q, tm_kv, a);
  CA_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ca

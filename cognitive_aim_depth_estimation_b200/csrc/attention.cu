// Flash-style multi-head self-attention forward for the DINOv2 backbone on sm_100a (head_dim 64, no mask).
// Replaces `F.scaled_dot_product_attention` at HF modeling_dinov2.py:215-229 (scale 64^-0.5, softmax over keys).
//
// Two budgets shape the kernel.  (1) One exp2 per score: the MUFU (16 /clk/SM) needs 1024 clk per 128x128 score tile,
// the tensor core 512.  (2) Shared-memory bandwidth (128 B/clk/SM): a 128x64x16 tcgen05.mma with both operands in smem
// reads 6 KB in 32 clk = 192 B/clk, so an "SS" kernel is smem-bound before it is MUFU-bound.  Hence BOTH A operands
// live in TMEM: Q (stored once per CTA) and P (written by the softmax warps over the S columns they just read);
// shared memory only carries the K / V stream (TMA in, one tensor-core read each).
//
// Work item = one (image, head) x one tile of 128 queries.  The kernel is PERSISTENT: two CTAs co-reside per SM
// (64 KB smem, 256 TMEM columns each) and each walks items blockIdx.x, blockIdx.x + gridDim.x, ... as ONE flattened
// sequence of 64-key steps, so the K/V stream, the S MMAs and the Q tile of the next item are already in flight
// while the current item finishes (a per-item launch loses ~20 % to TMEM allocation, first-load latency and the
// drain of the last P V at T = 1370).  S is double-buffered in TMEM, Q double-buffered per item:
//
//   warp 0     : TMA producer — K and V steps (64 keys = 8 KB each) into two independent 4-slot rings
//                (K runs two steps ahead of V, the order the tensor core consumes them in)
//   warp 1     : MMA issuer   — S_i = Q K_i^T  (tcgen05.mma 128x64x16 x4, A = Q in TMEM) into S buffer i&1
//                               O += P_i V_i   (tcgen05.mma 128x64x16 x4, A = P_i in TMEM, V MN-major as loaded)
//                               issue order S_0 S_1 | PV_0 S_2 | PV_1 S_3 | ...  (S_{i+2} overwrites the buffer that held
//                               S_i / P_i; the tensor pipe executes in issue order, so it follows PV_i's operand reads)
//   warps 2..5 : softmax      — one query row per thread: Q row global -> TMEM (for the NEXT item, one item ahead); the
//                               epilogue of item k (O / l -> global) runs inside step 0 of item k+1, after that step's
//                               exponentials, so the drain of the last P V is hidden; per step the 64-wide S_i row is
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

#include <atomic>
#include <mutex>

#include <type_traits>

namespace ca {
namespace {

constexpr int kHeadDim = 64;
constexpr int kTileQ = 128;
constexpr int kSubK = 64;                          // keys per pipeline step
constexpr int kRing = 6;                           // K ring slots == V ring slots
constexpr int kAttnThreads = 6 * 32;
constexpr int kSubBytes = kSubK * kHeadDim * 2;    // 8 KB: 64 rows x 128 B
constexpr int kSmemK = 0;
constexpr int kSmemV = kSmemK + kRing * kSubBytes;
constexpr int kSmemOut = kSmemV + kRing * kSubBytes;  // 4 x 2 KB: per softmax warp, 32 rows x 64 B of the output tile
constexpr int kOutStageBytes = 32 * 64;
constexpr int kSmemBar = kSmemOut + 4 * kOutStageBytes;
constexpr int kAttnSmemBytes = kSmemBar + 512;
static_assert(2 * (kAttnSmemBytes + 1024) <= 227 * 1024, "two CTAs must fit one SM");
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kTmemS = 0;      // two S buffers of 64 fp32 columns; P_i (bf16x2) overwrites the first 32 of buffer i&1
constexpr uint32_t kTmemO = 128;    // 64 fp32 columns
constexpr uint32_t kTmemQ = 192;    // 2 x 32 columns: Q tile of item k&1 as bf16x2 (A operand of every S MMA)
constexpr int kItemQ = 8;            // depth of the per-CTA item queue (power of two)
// Split of the exp2 work between the MUFU and the FMA-pipe polynomial: 2 bits per group of 8 scores (4 pairs) = how many
// of its pairs take the polynomial path; 8 groups per 64-key step.  0x5555 = one pair in every group = 25 %.
#ifndef CA_ATTN_POLY_MASK
#define CA_ATTN_POLY_MASK 0x5555
#endif
constexpr unsigned kPolyMask = CA_ATTN_POLY_MASK;
// Waits of the control warps (TMA producer on scheduler 0, MMA issuer on scheduler 1, which they share with softmax
// warps 4 and 5 of both resident CTAs): parked by the hardware instead of polled.
// CA_ATTN_ABLATE (compile time, timing experiments only — results are wrong): 1 = exponentials do not wait for the
// step's row maximum, 2 = no exponentials, 4 = no K / V traffic after the first ring fill.  What they showed
// (B=32, T=1370): 0.263 ms -> 0.251 (1) / 0.234 (2) / 0.228 (3) / 0.262 (4) / 0.221 (7): neither the MUFU nor the L2
// stream bounds the kernel; the per-step protocol does (profiles/README.md, "attention timeline").
#ifndef CA_ATTN_ABLATE
#define CA_ATTN_ABLATE 0
#endif
// CA_ATTN_PREFETCH=1: pull the next step's S row into registers before this step's store wait / hand-over.  Measured
// same-box: SLOWER (0.265 -> 0.303 ms; 168 registers and a few spills instead of 152), so it is off.
#ifndef CA_ATTN_PREFETCH
#define CA_ATTN_PREFETCH 0
#endif
// CA_ATTN_TRACE=1 (compile time) + CA_ATTN_DEBUG (run time): CTA 0 records clock64() at the phase boundaries of its first
// kTraceSteps steps (softmax warp 2 lane 0: 4 slots; MMA warp: 4 slots); the launcher prints the timeline.
#ifndef CA_ATTN_TRACE
#define CA_ATTN_TRACE 0
#endif
constexpr int kTraceSteps = 96;
constexpr int kTraceBase = 2 * 1024;  // offset into the debug buffer (roles: 0 softmax, 1 MMA, 2 boundary detail)
#if CA_ATTN_TRACE
#define TRACE(role, step, slot)                                                                              \
  do {                                                                                                       \
    if (p.dbg && blockIdx.x == 0 && lane == 0 && (step) < kTraceSteps)                                       \
      p.dbg[kTraceBase + ((role) * kTraceSteps + (step)) * 4 + (slot)] = clock64();                          \
  } while (0)
#else
#define TRACE(role, step, slot) do {} while (0)
#endif
#ifndef CA_ATTN_PARKED_WAITS
#define CA_ATTN_PARKED_WAITS 1
#endif
#if CA_ATTN_PARKED_WAITS
#define CTRL_WAIT(bar, parity) mbar_wait_parked(bar, parity)
#else
#define CTRL_WAIT(bar, parity) mbar_wait(bar, parity)
#endif
// The softmax reference of a row is set kRefBias (log2 units) ABOVE the largest score seen when it is set, and only moves
// again when a later score exceeds that maximum by more than 2^kRefBias (see the step function).
constexpr float kRefBias = 64.0f;
// CA_ATTN_EXACT_MAX=1 (compile time): every step computes its exact row maximum (the A/B control of the max-free step)
#ifndef CA_ATTN_EXACT_MAX
#define CA_ATTN_EXACT_MAX 0
#endif

struct AttnArgs {
  int T;          // tokens per image
  int H;          // heads
  int n_sub;      // 64-key steps per item
  int n_qt;       // query tiles per (image, head)
  int n_items;    // n_qt * B * H
  float scale_log2;
  const __nv_bfloat16* qkv;  // [B*T, 3*H*64]
  int ld;
  __nv_bfloat16* out;  // [B*T, H*64]
  int ldo;
  int stagger;     // cycles the second CTA of each SM waits before its first step
  int* counter;    // [0] work counter, [1] CTAs that have finished — THIS launch's own slot of a ring of such pairs, zero when
                   // the launch starts; the last CTA to leave zeroes it again for the launch that reuses the slot
  long long* dbg;  // CA_ATTN_DEBUG=1: per-CTA {cycles, smid}
};

// 2^t for two lanes on the FMA / ALU pipes instead of the MUFU (t <= ~100): round-to-nearest split t = n + f with the
// 1.5 * 2^23 trick, cubic minimax of 2^f on [-0.5, 0.5] (max relative error 7.6e-5, far below the bf16 rounding of P),
// then n is added into the exponent field.  The MUFU does 16 exp2 / clk / SM; this path takes the overflow.
__device__ __forceinline__ void exp2_poly2(float& t0, float& t1) {
  constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23
  t0 = fmaxf(t0, -125.0f);
  t1 = fmaxf(t1, -125.0f);
  float r0, r1, n0, n1, f0, f1, p0, p1;
  ffma2v(r0, r1, t0, t1, 1.0f, 1.0f, kMagic, kMagic);
  ffma2v(n0, n1, r0, r1, 1.0f, 1.0f, -kMagic, -kMagic);
  ffma2v(f0, f1, n0, n1, -1.0f, -1.0f, t0, t1);
  ffma2v(p0, p1, f0, f1, 0.05520550534129143f, 0.05520550534129143f, 0.24261397123336792f, 0.24261397123336792f);
  ffma2v(p0, p1, p0, p1, f0, f1, 0.6932547688484192f, 0.6932547688484192f);
  ffma2v(p0, p1, p0, p1, f0, f1, 0.9999276995658875f, 0.9999276995658875f);
  t0 = __int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23));
  t1 = __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23));
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_kv, const AttnArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemBar);
  uint64_t* k_full = bars + 0;               // [kRing] TMA -> MMA
  uint64_t* k_empty = bars + kRing;          // [kRing] MMA (commit) -> TMA
  uint64_t* v_full = bars + 2 * kRing;       // [kRing]
  uint64_t* v_empty = bars + 3 * kRing;      // [kRing]
  uint64_t* q_ready = bars + 4 * kRing;      // [2] softmax -> MMA : Q of item k is in TMEM buffer k&1
  uint64_t* s_full = q_ready + 2;    // [2] MMA -> softmax : S of global step g is in TMEM buffer g&1
  uint64_t* p_full = q_ready + 4;    // [2] softmax -> MMA : P of step g is in TMEM (over S), O rescaled / handed over
  uint64_t* pv_done = q_ready + 6;   // [2] MMA -> softmax : P V of step g finished (O quiescent up to step g)
  uint64_t* o_full = q_ready + 8;    // MMA -> softmax : last P V of an item finished
  uint64_t* item_full = q_ready + 9;  // [8] scheduler -> everyone : item_q[k & 7] holds the k-th item of this CTA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(item_full + kItemQ);
  volatile int* item_q = reinterpret_cast<volatile int*>(item_full + kItemQ + 1);  // [8] item index or -1 (no more work)
  static_assert((4 * kRing + 9 + kItemQ + 1) * 8 + kItemQ * 4 <= 512, "barrier block overflows its 512 bytes");

  const int warp = warp_id();
  const int lane = lane_id();
  const long long t_start = clock64();
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
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_ready[s], 4);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);
      mbar_init(&pv_done[s], 1);
    }
    mbar_init(o_full, 1);
    for (int s = 0; s < kItemQ; ++s) mbar_init(&item_full[s], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_sync();  // PDL: everything above is CTA-local; the QKV activation (and the work counter) are touched below
  const uint32_t tmem_base = *tmem_slot;

  // k-th item of this CTA (blocks until the scheduler has published it); -1 = the CTA has run out of work
  auto read_item = [&](int k) -> int {
    mbar_wait(&item_full[k & (kItemQ - 1)], (k / kItemQ) & 1);
    return item_q[k & (kItemQ - 1)];
  };

  if (warp == 0) {
    if (lane == 0) {
      // Scheduler + TMA producer.  Items are claimed from a global counter (the two CTAs of an SM do not progress at
      // the same rate, and neither do all SMs: a static split leaves the slowest CTA running alone at the end).
      // Item k+2 is claimed when the K stream enters item k, so every consumer finds its item id long published.
      int n_claimed = 0;
      bool exhausted = false;
      auto claim = [&]() {
        int it = -1;
        if (!exhausted) {
          it = atomicAdd(p.counter, 1);
          if (it >= p.n_items) { it = -1; exhausted = true; }
        }
        item_q[n_claimed & (kItemQ - 1)] = it;
        mbar_arrive(&item_full[n_claimed & (kItemQ - 1)]);  // release: the id is visible to whoever passes the wait
        ++n_claimed;
      };
      claim();
      claim();
      int k_item = 0, k_step = 0, k_id = item_q[0], kg = 0;  // K stream cursor: item, step, item id, global step
      int v_item = 0, v_step = 0, v_id = k_id, vg = 0;       // V stream cursor
      int k_b = 0, k_col = 0, v_b = 0, v_col = 0;            // image and column of the cursor's item
      auto load_k = [&]() {  // one K step (no-op once the K stream has run out of items)
        if (k_id < 0) return;
        if (k_step == 0) {
          claim();  // entering item k_item: claim item k_item + 2
          const int bh = k_id / p.n_qt;
          k_b = bh / p.H;
          k_col = (p.H + bh - k_b * p.H) * kHeadDim;
        }
        const int s = kg % kRing;
        CTRL_WAIT(&k_empty[s], ((kg / kRing) & 1) ^ 1u);
#if CA_ATTN_ABLATE & 4   // timing experiment: no K / V traffic after the first ring fill
        if (kg >= kRing) { mbar_arrive(&k_full[s]); } else
#endif
        {
        mbar_arrive_expect_tx(&k_full[s], kSubBytes);
        tma_load_3d(smem + kSmemK + s * kSubBytes, &tmap_kv, &k_full[s], k_col, k_step * kSubK, k_b);
        }
        ++kg;
        if (++k_step == n_sub) {
          k_step = 0;
          ++k_item;
          k_id = item_q[k_item & (kItemQ - 1)];  // written by this thread
        }
      };
      load_k();
      load_k();
      while (v_id >= 0) {
        if (v_step == 0) {
          const int bh = v_id / p.n_qt;
          v_b = bh / p.H;
          v_col = (2 * p.H + bh - v_b * p.H) * kHeadDim;
        }
        const int s = vg % kRing;
        CTRL_WAIT(&v_empty[s], ((vg / kRing) & 1) ^ 1u);
#if CA_ATTN_ABLATE & 4
        if (vg >= kRing) { mbar_arrive(&v_full[s]); } else
#endif
        {
        mbar_arrive_expect_tx(&v_full[s], kSubBytes);
        tma_load_3d(smem + kSmemV + s * kSubBytes, &tmap_kv, &v_full[s], v_col, v_step * kSubK, v_b);
        }
        ++vg;
        if (++v_step == n_sub) {
          v_step = 0;
          ++v_item;
          v_id = item_q[v_item & (kItemQ - 1)];
        }
        load_k();
      }
    }
  } else if (warp == 1) {
    // All 32 lanes walk the protocol (waits are cheap); the tensor-core instructions are issued by one elected lane.
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, kSubK, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);  // B (=V) is MN-major
    const uint64_t kd0 = umma_smem_desc_sw128(smem_u32(smem + kSmemK));
    const uint64_t vd0 = umma_smem_desc_sw128(smem_u32(smem + kSmemV), 1024, 1024);
    constexpr uint64_t kSubStep = kSubBytes >> 4;  // descriptor address units are 16 bytes
    int s_item = 0, s_step = 0, sg = 0;  // (item, step, global step) of the next S to issue
    bool s_more = true;
    auto issue_s = [&]() {
      if (!s_more) return;
      if (s_step == 0) {  // first S of an item: the item must exist and its Q tile must be in TMEM
        if (read_item(s_item) < 0) { s_more = false; return; }
        CTRL_WAIT(&q_ready[s_item & 1], (s_item >> 1) & 1);
      }
      const int s = sg % kRing;
      CTRL_WAIT(&k_full[s], (sg / kRing) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t kd = kd0 + s * kSubStep;
        const uint32_t d = tmem_base + kTmemS + (sg & 1) * kSubK;
        const uint32_t t_q = tmem_base + kTmemQ + (s_item & 1) * 32;
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) umma_bf16_ts(d, t_q + 8 * k, kd + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[sg & 1]);
        umma_commit(&k_empty[s]);
      }
      __syncwarp();
      ++sg;
      if (++s_step == n_sub) { s_step = 0; ++s_item; }
    };
    issue_s();
    issue_s();
    int g = 0;
    for (int k = 0; read_item(k) >= 0; ++k) {
      for (int i = 0; i < n_sub; ++i, ++g) {
        const int bb = g & 1;
        const int s = g % kRing;
        TRACE(1, g, 0);
        CTRL_WAIT(&v_full[s], (g / kRing) & 1);
        TRACE(1, g, 1);
        CTRL_WAIT(&p_full[bb], (g >> 1) & 1);  // on step 0 of an item this also hands O over (previous item read out)
        TRACE(1, g, 2);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t vd = vd0 + s * kSubStep;
          const uint32_t t_p = tmem_base + kTmemS + bb * kSubK;
#pragma unroll
          for (int kk = 0; kk < kSubK / 16; ++kk) {
            // P: 16 keys = 8 packed columns; V: 16 keys = 16 rows of 128 B = 2048 B further down
            umma_bf16_ts(tmem_base + kTmemO, t_p + 8 * kk, vd + kk * (2048 >> 4), idesc_o, (i | kk) != 0);
          }
          umma_commit(&pv_done[bb]);
          umma_commit(&v_empty[s]);
          if (i == n_sub - 1) umma_commit(o_full);
        }
        __syncwarp();
        issue_s();  // S of step g+2: overwrites S_g / P_g, ordered behind PV_g by the in-order tensor pipe
        TRACE(1, g, 3);
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t t_o = t_lane + kTmemO;
    const float scale = p.scale_log2;
    int g = 0;  // global 64-key step of this CTA
    // Q tile of an item: global -> registers with 4 lanes per row (8 rows x 64 contiguous bytes per load instruction; a
    // thread-per-row load touches 32 lines per instruction), rows past the end of the image zero.  q_rows() turns the
    // registers into the thread-per-row layout tcgen05.st wants, through the warp's 2 KB staging buffer.
    const uint32_t stg = smem_u32(smem + kSmemOut + quad * kOutStageBytes);
    const int orr = lane >> 2, opc = lane & 3;
    auto q_load = [&](int it, uint32_t (&qv)[32]) {
      const int bh = it / p.n_qt;
      const int q0 = (it - bh * p.n_qt) * kTileQ + quad * 32;
      const int b = bh / p.H;
      const int h = bh - b * p.H;
      const __nv_bfloat16* src = p.qkv + static_cast<size_t>(b) * p.T * p.ld + h * kHeadDim + opc * 8;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = q0 + orr + 8 * i;
          const uint4 w = row < p.T ? __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(row) * p.ld + c * 32))
                                    : make_uint4(0u, 0u, 0u, 0u);
          qv[16 * c + 4 * i + 0] = w.x;
          qv[16 * c + 4 * i + 1] = w.y;
          qv[16 * c + 4 * i + 2] = w.z;
          qv[16 * c + 4 * i + 3] = w.w;
        }
      }
    };
    auto q_rows = [&](uint32_t (&qv)[32]) {  // in place: 4-lanes-per-row pieces -> this thread's row (64 bf16)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = orr + 8 * i;
          sts128(stg + rr * 64 + ((opc ^ ((rr >> 1) & 3)) << 4),
                 make_uint4(qv[16 * c + 4 * i + 0], qv[16 * c + 4 * i + 1], qv[16 * c + 4 * i + 2], qv[16 * c + 4 * i + 3]));
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint4 w = lds128(stg + lane * 64 + ((t ^ ((lane >> 1) & 3)) << 4));
          qv[16 * c + 4 * t + 0] = w.x;
          qv[16 * c + 4 * t + 1] = w.y;
          qv[16 * c + 4 * t + 2] = w.z;
          qv[16 * c + 4 * t + 3] = w.w;
        }
        __syncwarp();
      }
    };
    // epilogue of the k-th item (global item index `it`): O / l -> bf16 -> global (token-major, head h at [64h, 64h+64))
    auto epilogue = [&](int k, int it, float l) {
      const int bh = it / p.n_qt;
      const int q = (it - bh * p.n_qt) * kTileQ + r;
      const int b = bh / p.H;
      const int h = bh - b * p.H;
      mbar_wait(o_full, k & 1);
      if (warp == 2) TRACE(2, g, 0);
      tc_fence_after();
      const float inv = 1.0f / l;
      // TMEM hands every thread one output ROW; written straight from there each 16-byte store of a warp touches 32
      // different lines (8 such instructions per item and warp).  Each 32-column half goes through the warp's 2 KB
      // staging buffer instead (16-byte pieces XOR-swizzled with bits 1-2 of the row: both sides conflict-free) and
      // leaves with 4 lanes per row: 8 rows x 64 contiguous bytes per store instruction.
      const int q0 = q - lane;  // first row of this warp
      __nv_bfloat16* obase = p.out + static_cast<size_t>(b) * p.T * p.ldo + h * kHeadDim + opc * 8;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(t_o + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[8 * t + 0]) * inv, __uint_as_float(o[8 * t + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(o[8 * t + 2]) * inv, __uint_as_float(o[8 * t + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(o[8 * t + 4]) * inv, __uint_as_float(o[8 * t + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(o[8 * t + 6]) * inv, __uint_as_float(o[8 * t + 7]) * inv);
          sts128(stg + lane * 64 + ((t ^ ((lane >> 1) & 3)) << 4), w);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = orr + 8 * i;
          const uint4 w = lds128(stg + rr * 64 + ((opc ^ ((rr >> 1) & 3)) << 4));
          if (q0 + rr < p.T) *reinterpret_cast<uint4*>(obase + static_cast<size_t>(q0 + rr) * p.ldo + c * 32) = w;
        }
        __syncwarp();
      }
    };
    if (p.stagger > 0 && blockIdx.x >= gridDim.x / 2) {  // de-phase the two CTAs of an SM
      const long long t0 = clock64();
      while (clock64() - t0 < p.stagger) {}
    }
    int it_cur = read_item(0), it_prev = -1, it_next = -1;
    if (it_cur >= 0) {  // Q of the first item
      uint32_t qv[32];
      q_load(it_cur, qv);
      q_rows(qv);
      tmem_st32(t_lane + kTmemQ, qv);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_ready[0]);
    }
    float l_prev = 1.f;
    int k = 0;
    float m_acc = -INFINITY, l_run = 0.f;  // reference (log2 domain) O and l_run are expressed in; running row sum
    uint32_t v[64];        // the S row of the current step (fp32 bits)
    // Row statistics the exact way: row maximum of the 64 scores in `v`, reference raised to it (plus the bias below) when
    // it grew, running sum and — on a step that is not the first of its item — the O accumulator rescaled to the new
    // reference.  Runs on the first step of every item and on the (rare) steps the fast path below hands back.
    auto exact_reference = [&](const uint32_t (&v)[64], bool first_step, int bb) {
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
      const float m_new = fmaxf(m_acc, tile_max + kRefBias);
      const float alpha = fast_exp2(m_acc - m_new);  // 0 on the first step (m_acc = -inf), 1 for rows that did not grow
      l_run *= alpha;
      m_acc = m_new;
      if (!first_step) {
        mbar_wait(&pv_done[bb ^ 1], ((g - 1) >> 1) & 1);  // P V of step g-1 (hence every earlier one) has finished
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t o[16];
          tmem_ld16(t_o + c * 16, o);
          tmem_ld_wait();
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) o[kk] = __float_as_uint(__uint_as_float(o[kk]) * alpha);
          tmem_st16(t_o + c * 16, o);
        }
      }
    };
    // P = exp2(s * scale - m_acc) -> bf16x2 in pk; returns the fp32 row sum of the step
    auto exponentials = [&](const uint32_t (&v)[64], uint32_t (&pk)[32]) -> float {
      const float neg_m = -m_acc;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int t = 0; t < kSubK / 8; ++t) {
        float e[8];
#pragma unroll
        for (int kk = 0; kk < 8; kk += 2) {
          ffma2(e[kk], e[kk + 1], __uint_as_float(v[8 * t + kk]), __uint_as_float(v[8 * t + kk + 1]), scale, neg_m);
          if (kk >= 8 - 2 * static_cast<int>((kPolyMask >> (2 * t)) & 3u)) {  // compile-time split MUFU / polynomial
            exp2_poly2(e[kk], e[kk + 1]);
          } else {
            e[kk] = fast_exp2(e[kk]);
            e[kk + 1] = fast_exp2(e[kk + 1]);
          }
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
      return (s0 + s1) + (s2 + s3);
    };
    // One 64-key step.  kBoundary = step 0 of an item, which also carries the previous item's epilogue and the next
    // item's Q tile; it is a separate instantiation so that the steady-state step stays lean.
    //
    // The steady-state step computes NO row maximum: the exponentials use the reference the row already has, which sits
    // kRefBias = 64 (log2 units) above the largest score seen when it was set, so P = 2^(s - ref) stays below 1 as long as
    // no score exceeds that maximum by more than 2^64 — and floating point does not care where in its range P, the row
    // sum and O live (the final O / l cancels the reference).  A score that does break out shows up as a P >= 2 (exponent
    // MSB of the packed bf16) or as a wrapped / negative polynomial result (sign bit): one OR-reduction of the 32 packed
    // words finds either, and the warp then redoes the step the exact way (re-reading S, which P has not overwritten
    // yet).  This takes the 31 max instructions, the vote on the maximum and — more important — the dependency of every
    // exponential on the row maximum off the step.
    auto step = [&](auto boundary_tag, const int i) {
      constexpr bool boundary = decltype(boundary_tag)::value;
      const int bb = g & 1;
      const int valid = p.T - i * kSubK;  // >= 64 on every step but the last
      const uint32_t t_s = t_lane + kTmemS + bb * kSubK;
      uint32_t qv[boundary ? 32 : 1];
      if constexpr (boundary) {
        it_next = read_item(k + 1);
        if (it_next >= 0) q_load(it_next, qv);  // in flight under this step's exponentials
      }
      uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
      uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
      auto load_row = [&]() {
        tmem_ld32(t_s, v0);
        tmem_ld32(t_s + 32, v1);
        tmem_ld_wait();
        if (valid < kSubK) {
#pragma unroll
          for (int c = 0; c < kSubK; ++c)
            if (c >= valid) v[c] = 0xff800000u;  // -inf: exp2 -> 0, ignored by the max
        }
      };
      if (warp == 2) TRACE(0, g, 0);
      mbar_wait(&s_full[bb], (g >> 1) & 1);
      tc_fence_after();
      if (warp == 2) TRACE(0, g, 1);
      load_row();
      uint32_t pk[32];
      float step_sum;
      if constexpr (boundary) {
        exact_reference(v, true, bb);
        step_sum = exponentials(v, pk);
      } else {
#if CA_ATTN_EXACT_MAX
        exact_reference(v, false, bb);
        step_sum = exponentials(v, pk);
#else
        step_sum = exponentials(v, pk);
        // Stale reference?  A P >= 2 out of the MUFU (up to +inf, or NaN) makes the row sum >= 2 — sums of a valid step are
        // below 64 * 2^-64 — so ONE compare covers the MUFU elements; the polynomial elements can also wrap into a
        // negative or denormal pattern for inputs beyond their range, so their packed words (8 of the 32 at the shipped
        // 25 % share) are still OR-ed and tested for the sign / exponent-MSB bits.
        uint32_t any = 0;
#pragma unroll
        for (int t = 0; t < kSubK / 8; ++t) {
#pragma unroll
          for (int w = 0; w < 4; ++w)
            if (w >= 4 - static_cast<int>((kPolyMask >> (2 * t)) & 3u)) any |= pk[4 * t + w];
        }
        if (__any_sync(0xffffffffu, (any & 0xC000C000u) != 0u || !(step_sum < 2.0f))) {  // the reference is stale
          load_row();
          exact_reference(v, false, bb);
          step_sum = exponentials(v, pk);
        }
#endif
      }
      if (warp == 2) TRACE(0, g, 2);
      tmem_st32(t_s, pk);
      l_run += step_sum;
      if constexpr (boundary) {
        // the previous item's last P V has long finished: read its O out before this step's P V may overwrite it
        if (k > 0) epilogue(k - 1, it_prev, l_prev);
        if (warp == 2) TRACE(2, g, 1);
        if (it_next >= 0) {  // Q of the next item (its buffer was last read by item k-1, which is complete)
          q_rows(qv);
          tmem_st32(t_lane + kTmemQ + ((k + 1) & 1) * 32, qv);
          tmem_st_wait();
          if (warp == 2) TRACE(2, g, 2);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&q_ready[(k + 1) & 1]);
        }
      }
      tmem_st_wait();     // P (and a rescaled O) are in TMEM
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[bb]);
      if (warp == 2) TRACE(0, g, 3);
      ++g;
    };
    while (it_cur >= 0) {
      m_acc = -INFINITY;
      l_run = 0.f;
      step(std::true_type{}, 0);
      for (int i = 1; i < n_sub; ++i) step(std::false_type{}, i);
      l_prev = l_run;
      it_prev = it_cur;
      it_cur = it_next;
      ++k;
    }
    if (k > 0) epilogue(k - 1, it_prev, l_prev);
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    // re-arm this launch's counter slot (every claim of the launch has returned by now: a CTA only gets here after its
    // scheduler drew the out-of-work sentinel).  No memset node between kernels: it would break the PDL chain.
    __threadfence();
    if (atomicAdd(p.counter + 1, 1) == static_cast<int>(gridDim.x) - 1) {
      p.counter[0] = 0;
      p.counter[1] = 0;
      __threadfence();
    }
  }
  if (p.dbg && threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.dbg[2 * blockIdx.x] = clock64() - t_start;
    p.dbg[2 * blockIdx.x + 1] = smid;
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

constexpr int kMaxDevices = 64;
constexpr unsigned kCounterRing = 1024;     // slots recycled among eager launches
constexpr unsigned kCapturedSlots = 65536;  // slots handed out ONCE to launches recorded into a CUDA graph
struct DeviceState {
  int* ring = nullptr;  // kCounterRing + kCapturedSlots (work counter, exit counter) pairs, zero when idle
  int sms = 0;
  std::atomic<unsigned> next{0};
  std::atomic<unsigned> next_captured{0};
};
DeviceState g_dev[kMaxDevices];
std::mutex g_mu;

}  // namespace

int attention_launch(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, cudaStream_t stream, int ldo) {
  CA_REQUIRE(qkv && out, "attention: null pointer");
  CA_REQUIRE(B > 0 && T > 0 && H > 0, "attention: non-positive dimension");
  const int ld = 3 * H * kHeadDim;
  CUtensorMap tm_kv;
  CA_TRY(make_tmap_3d(&tm_kv, qkv, B, T, ld, ld, static_cast<uint64_t>(T) * ld, kSubK));
  int dev = 0;
  CA_CUDA(cudaGetDevice(&dev));
  CA_REQUIRE(dev >= 0 && dev < kMaxDevices, "attention: device index out of range");
  DeviceState& st = g_dev[dev];
  {
    // once per device: the > 48 KB shared-memory opt-in is per context, and so is the ring of work counters
    std::lock_guard<std::mutex> lock(g_mu);
    if (!st.ring) {
      CA_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
      CA_CUDA(cudaDeviceGetAttribute(&st.sms, cudaDevAttrMultiProcessorCount, dev));
      CA_CUDA(cudaMalloc(&st.ring, 2 * (kCounterRing + kCapturedSlots) * sizeof(int)));
      CA_CUDA(cudaMemset(st.ring, 0, 2 * (kCounterRing + kCapturedSlots) * sizeof(int)));
      CA_CUDA(cudaDeviceSynchronize());  // (first use only: the zeroes must be in place whatever stream launches first)
    }
  }
  AttnArgs a;
  a.T = T;
  a.H = H;
  a.n_sub = (T + kSubK - 1) / kSubK;
  a.n_qt = (T + kTileQ - 1) / kTileQ;
  a.n_items = a.n_qt * B * H;
  a.scale_log2 = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)
  a.qkv = qkv;
  a.ld = ld;
  a.out = out;
  a.ldo = ldo > 0 ? ldo : H * kHeadDim;
  CA_REQUIRE(a.ldo >= H * kHeadDim && a.ldo % 8 == 0, "attention: output leading dimension must be >= H*64 and a multiple of 8");
  const int grid = a.n_items < 2 * st.sms ? a.n_items : 2 * st.sms;  // persistent: two CTAs per SM
  // Every launch owns one slot of a per-device ring of 1024 (work counter, exit counter) pairs: launches that overlap on
  // different streams never share a counter.  A slot is zero when its launch starts (zeroed at allocation, re-zeroed by
  // the last CTA of the launch that used it 1024 launches earlier) — no memset node sits between the kernels of a
  // captured forward, so the programmatic dependency chain from the QKV GEMM to this kernel stays intact.  (A launch
  // that faults leaves its slot dirty; a faulted context is unusable anyway.)  A launch that is being CAPTURED keeps its
  // slot for as long as the graph may be replayed, so it gets one that is never handed out again: an eager launch or
  // another graph running on a different stream can then never meet it on the same counter.
  cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
  if (stream != nullptr && stream != cudaStreamLegacy)  // the legacy stream cannot capture, and asking it would
    CA_CUDA(cudaStreamIsCapturing(stream, &capturing)); // invalidate a capture in progress on another stream
  if (capturing == cudaStreamCaptureStatusActive) {
    const unsigned slot = st.next_captured.fetch_add(1u);
    CA_REQUIRE(slot < kCapturedSlots, "attention: more than 65536 attention launches have been captured into CUDA graphs in this "
                                      "process (each keeps a work-counter slot for the life of the process)");
    a.counter = st.ring + 2 * (kCounterRing + slot);
  } else {
    a.counter = st.ring + 2 * (st.next.fetch_add(1u) % kCounterRing);
  }
  static const int stagger = getenv("CA_ATTN_STAGGER") ? atoi(getenv("CA_ATTN_STAGGER")) : 0;
  a.stagger = stagger;
  static const bool want_dbg = getenv("CA_ATTN_DEBUG") != nullptr;
  static long long* d_dbg = nullptr;
  a.dbg = nullptr;
  if (want_dbg) {
    if (!d_dbg) {
      CA_CUDA(cudaMalloc(&d_dbg, (kTraceBase + 3 * kTraceSteps * 4) * sizeof(long long)));
      CA_CUDA(cudaMemset(d_dbg, 0, (kTraceBase + 3 * kTraceSteps * 4) * sizeof(long long)));
    }
    a.dbg = d_dbg;
  }
  CA_TRY(launch_kernel(attention_fwd_kernel, dim3(grid), dim3(kAttnThreads), kAttnSmemBytes, stream, tm_kv, a));
  CA_CUDA(cudaGetLastError());
  if (want_dbg) {
    static int calls = 0;
    if (++calls == 5) {
      static long long h[2 * 1024];
      CA_CUDA(cudaStreamSynchronize(stream));
      CA_CUDA(cudaMemcpy(h, d_dbg, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost));
      long long mn = h[0], mx = h[0];
      double sum = 0;
      for (int i = 0; i < grid; ++i) {
        mn = h[2 * i] < mn ? h[2 * i] : mn;
        mx = h[2 * i] > mx ? h[2 * i] : mx;
        sum += h[2 * i];
      }
      fprintf(stderr, "[attn dbg] grid %d cycles/CTA min %lld mean %.0f max %lld\n", grid, mn, sum / grid, mx);
#if CA_ATTN_TRACE
      {
        static long long tr[3 * kTraceSteps * 4];
        CA_CUDA(cudaMemcpy(tr, d_dbg + kTraceBase, sizeof(tr), cudaMemcpyDeviceToHost));
        const long long t0 = tr[0];
        fprintf(stderr, "[attn trace] CTA 0, n_sub %d; per step: softmax {wait_s_begin, s_ready, exps_done, p_arrived} | "
                        "mma {begin, v_ready, p_ready, issued}  (cycles since the first event)\n", a.n_sub);
        for (int g2 = 0; g2 < kTraceSteps; ++g2) {
          const long long* sm = tr + g2 * 4;
          const long long* mm = tr + (kTraceSteps + g2) * 4;
          const long long* bd = tr + (2 * kTraceSteps + g2) * 4;
          fprintf(stderr, "  step %2d  sm %7lld %7lld %7lld %7lld | mma %7lld %7lld %7lld %7lld", g2, sm[0] - t0,
                  sm[1] - t0, sm[2] - t0, sm[3] - t0, mm[0] - t0, mm[1] - t0, mm[2] - t0, mm[3] - t0);
          if (bd[0]) fprintf(stderr, " | boundary {o_full, O stored, Q in TMEM} %7lld %7lld %7lld", bd[0] - t0, bd[1] - t0, bd[2] - t0);
          fprintf(stderr, "\n");
        }
      }
#endif
      for (int i = 0; i < grid; i += 16) fprintf(stderr, "  cta %3d sm %3lld cycles %lld\n", i, h[2 * i + 1], h[2 * i]);
    }
  }
  return 0;
}

}  // namespace ca

// Launcher-side view of the tcgen05 GEMM (csrc/gemm.cu).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ca {

// Epilogues fused into the GEMM (the accumulator never round-trips through HBM in fp32).
enum GemmEpilogue : int {
  EPI_BIAS_BF16 = 0,   // out_bf16[m,n] = acc + bias[n]                         (QKV, focal Q|K)
  EPI_GELU_BF16 = 1,   // out_bf16[m,n] = gelu_erf(acc + bias[n])               (MLP fc1)
  EPI_RESID_F32 = 2,   // x_f32[m,n]   += ls[n] * (acc + bias[n])   in place    (attn proj, MLP fc2 + LayerScale + residual)
  EPI_PATCH_F32 = 3,   // x_f32[b*T+1+p, n] = acc + bias[n] + pos[1+p, n]       (patch embedding + position embedding)
  EPI_ROWSTATS = 4,    // per (row, 64-col span): max and sum-exp2 of acc*scale (focal softmax pass A)
  EPI_COLSUM = 5,      // per (row, 64-col span): sum_i w_i * exp2(acc*scale - rmax[i]) * rinv[i]   (focal pass B, transposed)
  EPI_F32 = 6          // out_f32[m,n] = acc                                    (tests)
};

struct GemmArgs {
  // Problem: for each batch b: C[b] (M x N) = A[b] (M x K) * W[b] (N x K)^T ; bf16 operands, K contiguous.
  const __nv_bfloat16* A;
  const __nv_bfloat16* W;
  int M, N, K;
  int lda, ldw;                    // leading dimensions (elements)
  int batch;                       // >= 1
  long long a_batch_stride;        // elements
  long long w_batch_stride;        // elements (0 = shared weight)
  int epilogue;                    // GemmEpilogue
  // Epilogue operands (unused ones may be null)
  void* out;                       // bf16 or f32 depending on epilogue
  int ldo;                         // elements
  long long out_batch_stride;      // elements
  const float* bias;               // [N]
  const float* ls;                 // [N] LayerScale (EPI_RESID_F32)
  const float* pos;                // [T, N] position embedding (EPI_PATCH_F32)
  int patches_per_img;             // EPI_PATCH_F32: Np ; output row = (m / Np) * (Np + 1) + 1 + m % Np
  float scale_log2;                // EPI_ROWSTATS / EPI_COLSUM: acc * scale_log2 is the base-2 logit
  float* part_a;                   // ROWSTATS: partial max  [batch, M, P] ; COLSUM: partial sums [batch, M, P]
  float* part_b;                   // ROWSTATS: partial sums [batch, M, P]
  const float* col_max;            // COLSUM: [batch, N] row-max (base-2 domain) of the softmax row that column i is
  const float* col_rinv;           // COLSUM: [batch, N] weight_i / sumexp_i
};

// P (partials per row) for the stats epilogues: one per 64-column span of the 256-wide pair tiles = 4 * ceil(N / 256).
inline int gemm_stats_partials(int N) { return 4 * ((N + 255) / 256); }

int gemm_launch(const GemmArgs& args, cudaStream_t stream);

}  // namespace ca

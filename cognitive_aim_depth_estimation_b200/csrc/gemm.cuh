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
  EPI_F32 = 6,         // out_f32[m,n] = acc                                    (tests)
  // LayerNorm folded into the GEMMs either side of it (HF modeling_dinov2.py:354,359 without a LayerNorm pass):
  //   LN(x) W^T + b = rstd * (x W'^T) + (b + W beta),   W' = W diag(gamma) with every ROW CENTRED: W'[n,:] -= mean_k W'[n,k]
  // (x W'^T = (x - mean(x)) W'^T when the rows of W' sum to zero).  The consumer multiplies the RAW bf16 residual rows
  // with W' and scales by the row's 1/std in its epilogue; the producer (residual epilogue) emits those raw bf16 rows
  // and the row statistics.
  EPI_LN_BIAS_BF16 = 7,  // out_bf16[m,n] = rstd[m] * acc + bias[n]                                      (QKV after norm1)
  EPI_LN_GELU_BF16 = 8,  // out_bf16[m,n] = gelu_erf(the same)                                           (fc1 after norm2)
  EPI_RESID_LN_F32 = 9   // x_f32[m,n] += ls[n] * (acc + bias[n]); shadow_bf16[m,n] = bf16(x); per (row, 128-column span)
                         // (sum, M2) of the new x into ln_stats                                          (proj, fc2)
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
  // LayerNorm folding (EPI_LN_* read, EPI_RESID_LN_F32 writes); null / 0 otherwise
  float* ln_stats = nullptr;         // [M, ln_slots, 2]: (sum, sum of squared deviations from the span mean) of each
                                     // 128-column span of the row of x the LayerNorm normalises
  int ln_slots = 0;                  // spans per row: the LayerNorm width / 128
  float ln_eps = 0.f;
  __nv_bfloat16* shadow = nullptr;   // EPI_RESID_LN_F32: bf16 copy of the updated rows [M, N]
  int ld_shadow = 0;
};

// P (partials per row) for the stats epilogues: one per 64-column span of the 256-wide pair tiles = 4 * ceil(N / 256).
inline int gemm_stats_partials(int N) { return 4 * ((N + 255) / 256); }

int gemm_launch(const GemmArgs& args, cudaStream_t stream);

}  // namespace ca

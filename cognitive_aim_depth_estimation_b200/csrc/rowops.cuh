// Launchers of the bandwidth-bound row kernels (csrc/rowops.cu).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ca {

constexpr int kPatchRowStride = 592;  // 3*14*14 = 588 padded to a 16-byte multiple (TMA row pitch)

int patchify_f32_launch(const float* images, __nv_bfloat16* patches, int B, int S, cudaStream_t stream);
int preprocess_u8_launch(const uint8_t* images, __nv_bfloat16* patches, int B, int S, const float* mean3,
                         const float* std3, cudaStream_t stream);
int fetch_pinned_launch(float* dst, const float* src_pinned_host, size_t n, cudaStream_t stream);
int cls_rows_launch(float* x, const float* cls, const float* pos, int B, int T, int D, cudaStream_t stream);
int layernorm_launch(const float* x, const float* gamma, const float* beta, void* out, int out_is_bf16, int rows, int D,
                     float eps, cudaStream_t stream, int ld_out = 0);  // ld_out: output row pitch in elements (0 = D)
// raw rows as bf16 + per (row, 128-column span) (sum, M2): the entry of the LayerNorm-folded GEMM chain (gemm.cuh)
int ln_shadow_launch(const float* x, __nv_bfloat16* shadow, int ld_shadow, float* stats, int rows, int D,
                     cudaStream_t stream);
int focal_input_launch(const float* tokens, const float* pe, const float* rowscale, __nv_bfloat16* xin, int B, int N,
                       int D, cudaStream_t stream);

}  // namespace ca

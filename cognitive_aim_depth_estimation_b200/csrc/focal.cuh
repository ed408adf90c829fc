// Launchers of the vector-sized focal / guidance kernels (csrc/focal.cu).
#pragma once

#include <cuda_runtime.h>

namespace ca {

int rowstats_merge_launch(const float* pm, const float* ps, const float* weight, float* rmax, float* rinv, float* wtab,
                          int rows, int rows_per_image, int P, cudaStream_t stream);
int colsum_e_launch(const void* E, int lde, long long e_batch_stride, const float* wtab, float* pc, int B, int N, int P,
                    cudaStream_t stream);
int focal_finalize_launch(const float* pc, const float* cbias, float* attn, const float* rs_in, float* rs_out, int B,
                          int N, int P, float focus_strength, int mode, const float* cur_weight, float adaptive_weight,
                          cudaStream_t stream);
int guided_softmax_launch(const float* base, const float* mask, long long mask_batch_stride, float* heat, int* argmax,
                          int B, int N, float alpha, float temperature, cudaStream_t stream);
int weighted_pool_launch(const float* src, long long src_batch_stride, int row_offset, const float* w, const float* w2,
                         float* partial, int B, int N, int D, int splits, cudaStream_t stream);

}  // namespace ca

// Launcher of the tcgen05 flash-attention forward (csrc/attention.cu).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace ca {

// qkv: [B*T, 3*H*64] bf16 (Q | K | V column blocks, head h at columns 64h within a block)
// out: [B*T, H*64] bf16.  softmax(Q K^T / 8) V per (image, head); no mask.
int attention_launch(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, cudaStream_t stream,
                     int ldo = 0);  // ldo: output row pitch in elements (0 = H * 64)

}  // namespace ca

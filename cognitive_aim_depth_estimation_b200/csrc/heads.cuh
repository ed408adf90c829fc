// Launchers of the fused head kernels (csrc/heads.cu).  The argument structs are the public C structs.
#pragma once

#include <cuda_runtime.h>

#include "../../include/cogaim_b200.h"

namespace ca {

using HeadsWeights = ca_heads_weights;
using HeadsInputs = ca_heads_inputs;
using FocalValueArgs = ca_focal_value_args;

int heads_launch(const HeadsWeights& w, const HeadsInputs& in, float* depth, float* conf, float* fused_out, int B,
                 cudaStream_t stream);
int focal_value_launch(const FocalValueArgs& a, int B, cudaStream_t stream);
int focal_fusion_launch(const float* feats, int n_iters, const float* w0, const float* b0, const float* w1,
                        const float* b1, float* out, int B, cudaStream_t stream);

}  // namespace ca

// Launcher of the heat-map visualisation post-processing (csrc/visual.cu).
#pragma once

#include <cuda_runtime.h>

namespace ca {

int focus_map_launch(const float* heat, int B, int g, int out_h, int out_w, float* norm, float* out, cudaStream_t stream);

}  // namespace ca

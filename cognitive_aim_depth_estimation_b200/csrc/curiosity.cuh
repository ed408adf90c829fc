// Launcher of the CuriosityModule kernel (csrc/curiosity.cu).
#pragma once

#include <cuda_runtime.h>

#include "../../include/cogaim_b200.h"

namespace ca {

using CuriosityWeights = ca_curiosity_weights;
using CuriosityModWeights = ca_curiosity_mod_weights;

int curiosity_launch(const CuriosityWeights& w, const float* tokens, int tokens_per_img, const float* eps,
                     const float* noise, float* reward_raw, float* reward, float* history, int history_len,
                     long long* pointer, int B, cudaStream_t stream);

int curiosity_modulation_launch(const CuriosityModWeights& w, const float* reward, float lo, float hi, float* cur_weight,
                                int B, int n_iters, int mod_hidden, cudaStream_t stream);

}  // namespace ca

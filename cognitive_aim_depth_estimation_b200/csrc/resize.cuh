// Launcher of the exact-Pillow bilinear resize (csrc/resize.cu).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ca {

size_t resize_tmp_bytes(int B, int H0, int W0, int out_h, int out_w);
int resize_u8_launch(const uint8_t* src, int B, int H0, int W0, int out_h, int out_w, uint8_t* tmp, uint8_t* out,
                     cudaStream_t stream);

}  // namespace ca

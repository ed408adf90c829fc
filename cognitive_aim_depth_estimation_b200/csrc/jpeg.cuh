// JPEG byte stream (host) -> uint8 RGB HWC (device) through nvJPEG (csrc/jpeg.cu).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ca {

int jpeg_info(const uint8_t* h_data, size_t len, int* width, int* height, int* components);
int jpeg_decode(const uint8_t* h_data, size_t len, uint8_t* out_rgb, int width, int height, cudaStream_t stream);
int jpeg_decode_batch(const uint8_t* const* h_data, const size_t* lens, int n, uint8_t* const* outs, const int* widths,
                      const int* heights, cudaStream_t stream);

}  // namespace ca

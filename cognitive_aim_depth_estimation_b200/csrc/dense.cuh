// Small fp32 dense layers evaluated by ONE CTA on vectors held in shared memory (heads, EXIF prior, focal value path,
// curiosity module): shared by csrc/heads.cu and csrc/curiosity.cu.
#pragma once

#include "common.cuh"

namespace ca {
namespace {

constexpr int kHeadsThreads = 1024;

// out[o] = act(bias[o] + sum_i W[o, i] * in[i]) for o in [0, n_out); W row-major [n_out, n_in]; in/out in smem.
// The layers are tiny (<= 768 x 768) and run with one CTA per image, so what matters is how many independent L2 reads
// a CTA keeps in flight: each warp computes FOUR outputs per pass with 16-byte weight loads (4 rows x up to 6 float4
// per lane outstanding), 32 warps per CTA.
__device__ __forceinline__ void dense(const float* __restrict__ W, const float* __restrict__ bias, const float* in,
                                      float* out, int n_in, int n_out, bool relu) {
  const int w = warp_id(), l = lane_id(), nw = blockDim.x >> 5;
  if ((n_in & 3) == 0) {
    const int n4 = n_in >> 2;
    const float4* in4 = reinterpret_cast<const float4*>(in);
    for (int o0 = w * 4; o0 < n_out; o0 += nw * 4) {
      const float4* r0 = reinterpret_cast<const float4*>(W + static_cast<size_t>(o0) * n_in);
      // rows past n_out alias the last valid row (results discarded): keeps the inner loop branch-free
      const float4* r1 = r0 + (o0 + 1 < n_out ? 1 : 0) * n4;
      const float4* r2 = r0 + (o0 + 2 < n_out ? 2 : 0) * n4;
      const float4* r3 = r0 + (o0 + 3 < n_out ? 3 : 0) * n4;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 2
      for (int i = l; i < n4; i += 32) {
        const float4 x = in4[i];
        const float4 w0 = __ldg(r0 + i), w1 = __ldg(r1 + i), w2 = __ldg(r2 + i), w3 = __ldg(r3 + i);
        a0 = fmaf(w0.x, x.x, fmaf(w0.y, x.y, fmaf(w0.z, x.z, fmaf(w0.w, x.w, a0))));
        a1 = fmaf(w1.x, x.x, fmaf(w1.y, x.y, fmaf(w1.z, x.z, fmaf(w1.w, x.w, a1))));
        a2 = fmaf(w2.x, x.x, fmaf(w2.y, x.y, fmaf(w2.z, x.z, fmaf(w2.w, x.w, a2))));
        a3 = fmaf(w3.x, x.x, fmaf(w3.y, x.y, fmaf(w3.z, x.z, fmaf(w3.w, x.w, a3))));
      }
      a0 = warp_sum(a0);
      a1 = warp_sum(a1);
      a2 = warp_sum(a2);
      a3 = warp_sum(a3);
      if (l < 4 && o0 + l < n_out) {
        float acc = l == 0 ? a0 : l == 1 ? a1 : l == 2 ? a2 : a3;
        acc += bias ? bias[o0 + l] : 0.f;
        out[o0 + l] = relu ? fmaxf(acc, 0.f) : acc;
      }
    }
  } else {
    for (int o = w; o < n_out; o += nw) {
      const float* row = W + static_cast<size_t>(o) * n_in;
      float acc = 0.f;
      for (int i = l; i < n_in; i += 32) acc = fmaf(__ldg(row + i), in[i], acc);
      acc = warp_sum(acc);
      if (l == 0) {
        acc += bias ? bias[o] : 0.f;
        out[o] = relu ? fmaxf(acc, 0.f) : acc;
      }
    }
  }
  __syncthreads();
}

}  // namespace
}  // namespace ca

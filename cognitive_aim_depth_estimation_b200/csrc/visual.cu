// Heat-map post-processing for visualisation, the step right after the hot path:
// reference demo.py:530-563 does it per image on the host with numpy + scipy after a D2H copy of the attention vector:
//   cube -> 70th-percentile threshold (values at or below it are scaled by 0.3) -> min-max normalise -> g x g grid ->
//   scipy.ndimage.zoom(order=1) to the image size.
// Here: one CTA per image for the statistics (the percentile needs two order statistics: exact rank counting in shared
// memory, N <= 16384), then a bandwidth-bound bilinear zoom.  The float32 arithmetic follows numpy's (np.percentile
// `linear` method with a float32 virtual index and its two-sided lerp); the zoom follows scipy's (double coordinates
// out * (in-1)/(out-1), order-1 spline weights).
#include "common.cuh"
#include "host.h"
#include "visual.cuh"

namespace ca {
namespace {

constexpr int kVisThreads = 1024;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : fminf(v, t);
  }
  const int w = warp_id(), l = lane_id(), nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < nw) ? red[l] : red[0];
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, t, o);
    t = is_max ? fmaxf(t, u) : fminf(t, u);
  }
  return t;
}

__global__ void __launch_bounds__(kVisThreads) focus_normalize_kernel(const float* __restrict__ heat,
                                                                       float* __restrict__ norm, int N) {
  griddep_sync();  // PDL: nothing before this line reads or writes global memory
  extern __shared__ float a[];
  __shared__ float red[32];
  __shared__ float sel[2];
  const int b = blockIdx.x;
  const float* hb = heat + static_cast<size_t>(b) * N;
  for (int i = threadIdx.x; i < N; i += kVisThreads) {
    const float x = hb[i];
    a[i] = __fmul_rn(__fmul_rn(x, x), x);  // np.power(x, 3) in float32
  }
  __syncthreads();
  // np.percentile(a, 70): float32 virtual index (n-1) * (70/100), neighbours by exact rank (ties broken by index)
  const float q = __fdiv_rn(70.0f, 100.0f);
  const float vi = __fmul_rn(static_cast<float>(N - 1), q);
  const int k = static_cast<int>(floorf(vi));
  const float gamma = __fsub_rn(vi, static_cast<float>(k));
  const int k1 = min(k + 1, N - 1);
  for (int i = threadIdx.x; i < N; i += kVisThreads) {
    const float v = a[i];
    int rank = 0;
    for (int j = 0; j < N; ++j) {
      const float u = a[j];  // broadcast read
      rank += (u < v) || (u == v && j < i);
    }
    if (rank == k) sel[0] = v;
    if (rank == k1) sel[1] = v;
  }
  __syncthreads();
  const float lo = sel[0], hi = sel[1];
  const float diff = __fsub_rn(hi, lo);
  const float thr = gamma >= 0.5f ? __fsub_rn(hi, __fmul_rn(diff, __fsub_rn(1.0f, gamma)))
                                  : __fadd_rn(lo, __fmul_rn(diff, gamma));
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < N; i += kVisThreads) {
    float v = a[i];
    v = v > thr ? v : __fmul_rn(v, 0.3f);
    a[i] = v;
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  mn = block_reduce(mn, red, false);
  mx = block_reduce(mx, red, true);
  const float den = __fadd_rn(__fsub_rn(mx, mn), 1e-8f);
  float* nb = norm + static_cast<size_t>(b) * N;
  for (int i = threadIdx.x; i < N; i += kVisThreads) nb[i] = __fdiv_rn(__fsub_rn(a[i], mn), den);
}

// out[b, y, x] = order-1 zoom of norm[b] (g x g) to (H, W), scipy.ndimage.zoom semantics (grid_mode=False)
__global__ void focus_zoom_kernel(const float* __restrict__ norm, float* __restrict__ out, int g, int H, int W) {
  griddep_sync();  // PDL: nothing before this line reads or writes global memory
  const int b = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const double zy = H > 1 ? static_cast<double>(g - 1) / static_cast<double>(H - 1) : 0.0;
  const double zx = W > 1 ? static_cast<double>(g - 1) / static_cast<double>(W - 1) : 0.0;
  const double cy = zy * y, cx = zx * x;
  int y0 = static_cast<int>(floor(cy)), x0 = static_cast<int>(floor(cx));
  y0 = min(y0, g - 1);
  x0 = min(x0, g - 1);
  const double fy = cy - y0, fx = cx - x0;
  const int y1 = min(y0 + 1, g - 1), x1 = min(x0 + 1, g - 1);
  const float* m = norm + static_cast<size_t>(b) * g * g;
  const double v00 = m[y0 * g + x0], v01 = m[y0 * g + x1], v10 = m[y1 * g + x0], v11 = m[y1 * g + x1];
  // separable order-1 spline: rows first, then columns (the order scipy's geometric transform accumulates in)
  const double top = v00 * (1.0 - fx) + v01 * fx;
  const double bot = v10 * (1.0 - fx) + v11 * fx;
  out[(static_cast<size_t>(b) * H + y) * W + x] = static_cast<float>(top * (1.0 - fy) + bot * fy);
}

}  // namespace

int focus_map_launch(const float* heat, int B, int g, int out_h, int out_w, float* norm, float* out,
                     cudaStream_t stream) {
  CA_REQUIRE(heat && norm, "focus_map: null pointer");
  CA_REQUIRE(B > 0 && g > 0, "focus_map: non-positive dimension");
  const int N = g * g;
  CA_REQUIRE(N <= 16384, "focus_map: grid larger than 128 x 128 is not supported");
  static PerDeviceOnce configured;
  CA_TRY(configured.run([&]() -> int {
    CA_CUDA(cudaFuncSetAttribute(focus_normalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4));
    return 0;
  }));
  CA_TRY(launch_kernel(focus_normalize_kernel, dim3(B), dim3(kVisThreads), N * sizeof(float), stream, heat, norm, N));
  CA_CUDA(cudaGetLastError());
  if (out != nullptr) {
    CA_REQUIRE(out_h > 0 && out_w > 0, "focus_map: non-positive output size");
    dim3 block(32, 8);
    dim3 grid((out_w + 31) / 32, (out_h + 7) / 8, B);
    CA_TRY(launch_kernel(focus_zoom_kernel, dim3(grid), dim3(block), 0, stream, norm, out, g, out_h, out_w));
    CA_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace ca

// Exact-Pillow antialiased bilinear resize on the GPU: the `Resize((S, S))` of demo.py:162-163 (torchvision on a PIL
// image = PIL.Image.resize(..., BILINEAR) = Pillow's ImagingResample with the triangle filter), batched.
//
// Pillow's algorithm is integer after the coefficient tables: per output column a window [xmin, xmin + n) of source
// columns with 22-bit fixed-point weights (double-precision triangle filter, normalised, rounded half away from zero),
// accumulate from 1 << 21, shift right by 22, clamp to 8 bits; horizontal pass first into a uint8 intermediate (only
// the source rows the vertical pass will read), then the vertical pass.  The tables are built on the host in double
// exactly as Resample.c builds them and cached per (source size, target size); the two passes are bit-exact against
// PIL (tests/test_rowops_gpu.py::test_resize_matches_pillow).
#include "common.cuh"
#include "host.h"
#include "resize.cuh"

#include <math.h>

#include <map>
#include <mutex>
#include <vector>

namespace ca {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

struct AxisTable {
  int ksize = 0;
  int first = 0;  // first source index any output reads
  int last = 0;   // one past the last source index any output reads
  int* d_xmin = nullptr;
  int* d_cnt = nullptr;
  int* d_kk = nullptr;
};

// Resample.c precompute_coeffs + normalize_coeffs_8bpc, triangle filter (support 1.0), box = the whole axis.
int build_axis_table(int in_size, int out_size, AxisTable* t) {
  const double scale = static_cast<double>(in_size) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  const int ksize = static_cast<int>(ceil(support)) * 2 + 1;
  std::vector<int> xmin(out_size), cnt(out_size), kk(static_cast<size_t>(out_size) * ksize, 0);
  std::vector<double> w(ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int lo = static_cast<int>(center - support + 0.5);
    if (lo < 0) lo = 0;
    int hi = static_cast<int>(center + support + 0.5);
    if (hi > in_size) hi = in_size;
    const int n = hi - lo;
    double ww = 0.0;
    for (int x = 0; x < n; ++x) {
      double a = (x + lo - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      w[x] = a < 1.0 ? 1.0 - a : 0.0;
      ww += w[x];
    }
    for (int x = 0; x < n; ++x) {
      if (ww != 0.0) w[x] /= ww;
      kk[static_cast<size_t>(xx) * ksize + x] = w[x] < 0 ? static_cast<int>(-0.5 + w[x] * (1 << kPrecisionBits))
                                                          : static_cast<int>(0.5 + w[x] * (1 << kPrecisionBits));
    }
    xmin[xx] = lo;
    cnt[xx] = n;
  }
  t->ksize = ksize;
  t->first = xmin[0];
  t->last = xmin[out_size - 1] + cnt[out_size - 1];
  CA_CUDA(cudaMalloc(&t->d_xmin, out_size * sizeof(int)));
  CA_CUDA(cudaMalloc(&t->d_cnt, out_size * sizeof(int)));
  CA_CUDA(cudaMalloc(&t->d_kk, kk.size() * sizeof(int)));
  CA_CUDA(cudaMemcpy(t->d_xmin, xmin.data(), out_size * sizeof(int), cudaMemcpyHostToDevice));
  CA_CUDA(cudaMemcpy(t->d_cnt, cnt.data(), out_size * sizeof(int), cudaMemcpyHostToDevice));
  CA_CUDA(cudaMemcpy(t->d_kk, kk.data(), kk.size() * sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

int axis_table(int in_size, int out_size, const AxisTable** out) {
  static std::mutex mu;
  static std::map<std::tuple<int, int, int>, AxisTable> cache;  // (device, in, out)
  int dev = 0;
  CA_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_tuple(dev, in_size, out_size);
  auto it = cache.find(key);
  if (it == cache.end()) {
    AxisTable t;
    CA_TRY(build_axis_table(in_size, out_size, &t));
    it = cache.emplace(key, t).first;
  }
  *out = &it->second;
  return 0;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// dst[b, y, xx, :] = sum_j src[b, row0 + y, xmin[xx] + j, :] * kk[xx, j]     (3 interleaved channels)
__global__ void resize_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H0, int W0, int row0,
                                int rows, int out_w, const int* __restrict__ xmin, const int* __restrict__ cnt,
                                const int* __restrict__ kk, int ksize) {
  griddep_sync();  // PDL: nothing before this line reads or writes global memory
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (xx >= out_w) return;
  const uint8_t* s = src + ((static_cast<size_t>(b) * H0 + row0 + y) * W0 + xmin[xx]) * 3;
  const int* k = kk + static_cast<size_t>(xx) * ksize;
  const int n = cnt[xx];
  int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  for (int j = 0; j < n; ++j) {
    const int c = k[j];
    a0 += s[3 * j + 0] * c;
    a1 += s[3 * j + 1] * c;
    a2 += s[3 * j + 2] * c;
  }
  uint8_t* d = dst + ((static_cast<size_t>(b) * rows + y) * out_w + xx) * 3;
  d[0] = clip8(a0);
  d[1] = clip8(a1);
  d[2] = clip8(a2);
}

// dst[b, yy, e] = sum_j src[b, ymin[yy] - row0 + j, e] * kk[yy, j]   over e in [0, 3 * width) (channels interleaved)
__global__ void resize_v_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int rows, int row0,
                                int row_elems, const int* __restrict__ ymin, const int* __restrict__ cnt,
                                const int* __restrict__ kk, int ksize, int out_h) {
  griddep_sync();  // PDL: nothing before this line reads or writes global memory
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int yy = blockIdx.y;
  const int b = blockIdx.z;
  if (e >= row_elems) return;
  const uint8_t* s = src + (static_cast<size_t>(b) * rows + (ymin[yy] - row0)) * row_elems + e;
  const int* k = kk + static_cast<size_t>(yy) * ksize;
  const int n = cnt[yy];
  int a = 1 << (kPrecisionBits - 1);
  for (int j = 0; j < n; ++j) a += s[static_cast<size_t>(j) * row_elems] * k[j];
  dst[(static_cast<size_t>(b) * out_h + yy) * row_elems + e] = clip8(a);
}

}  // namespace

size_t resize_tmp_bytes(int B, int H0, int W0, int out_h, int out_w) {
  if (H0 == out_h || W0 == out_w) return 0;
  return static_cast<size_t>(B) * H0 * out_w * 3;  // upper bound (the horizontal pass may skip unread rows)
}

int resize_u8_launch(const uint8_t* src, int B, int H0, int W0, int out_h, int out_w, uint8_t* tmp, uint8_t* out,
                     cudaStream_t stream) {
  CA_REQUIRE(src && out, "resize: null pointer");
  CA_REQUIRE(B > 0 && H0 > 0 && W0 > 0 && out_h > 0 && out_w > 0, "resize: non-positive dimension");
  const bool need_h = W0 != out_w, need_v = H0 != out_h;
  if (!need_h && !need_v) {  // Pillow returns a copy
    CA_CUDA(cudaMemcpyAsync(out, src, static_cast<size_t>(B) * H0 * W0 * 3, cudaMemcpyDeviceToDevice, stream));
    return 0;
  }
  const AxisTable *th = nullptr, *tv = nullptr;
  if (need_h) CA_TRY(axis_table(W0, out_w, &th));
  if (need_v) CA_TRY(axis_table(H0, out_h, &tv));
  const int row0 = (need_h && need_v) ? tv->first : 0;
  const int rows = (need_h && need_v) ? tv->last - tv->first : H0;
  const uint8_t* vsrc = src;
  if (need_h) {
    uint8_t* hdst = need_v ? tmp : out;
    CA_REQUIRE(hdst != nullptr, "resize: a two-pass resize needs the temporary buffer");
    dim3 grid((out_w + 127) / 128, rows, B);
    CA_TRY(launch_kernel(resize_h_kernel, dim3(grid), dim3(128), 0, stream, src, hdst, H0, W0, row0, rows, out_w, th->d_xmin, th->d_cnt, th->d_kk,
                                              th->ksize));
    CA_CUDA(cudaGetLastError());
    vsrc = hdst;
  }
  if (need_v) {
    const int row_elems = out_w * 3;
    dim3 grid((row_elems + 255) / 256, out_h, B);
    CA_TRY(launch_kernel(resize_v_kernel, dim3(grid), dim3(256), 0, stream, vsrc, out, rows, row0, row_elems, tv->d_xmin, tv->d_cnt, tv->d_kk,
                                              tv->ksize, out_h));
    CA_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace ca

// Bandwidth-bound row kernels of the backbone: preprocessing / patchify, CLS rows, LayerNorm, focal-stream input.
// All are one-pass, 128-bit vectorised, coalesced along the contiguous dimension; their roofline is HBM.
#include "common.cuh"
#include "host.h"
#include "rowops.cuh"

namespace ca {
namespace {

constexpr int kPatch = 14;
constexpr int kPatchK = 3 * kPatch * kPatch;  // 588
constexpr int kPatchKPad = 592;               // row stride of the patch matrix: 16-byte multiple for TMA

// ---------------------------------------------------------------------------------------------
// K1b: fp32 CHW image (already normalised; what `forward` receives) -> bf16 patch rows
//      row (b, py, px), column k = c*196 + ky*14 + kx   (Conv2d weight.flatten(1) order, HF modeling_dinov2.py:139-148)
// One warp per (patch row py, channel c, ky): it reads a contiguous image row segment of g*14 floats (coalesced) and
// scatters 14-element pieces to the g patch rows.
// ---------------------------------------------------------------------------------------------
__global__ void patchify_f32_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ patches, int B, int S,
                                    int g) {
  const int line = blockIdx.x * (blockDim.x >> 5) + warp_id();  // over B*3*g*14 image lines that matter
  const int lines = B * 3 * g * kPatch;
  if (line >= lines) return;
  const int ky = line % kPatch;
  int t = line / kPatch;
  const int py = t % g;
  t /= g;
  const int c = t % 3;
  const int b = t / 3;
  const float* src = img + ((static_cast<size_t>(b) * 3 + c) * S + (py * kPatch + ky)) * S;
  __nv_bfloat16* dst = patches + (static_cast<size_t>(b) * g * g + static_cast<size_t>(py) * g) * kPatchKPad +
                       c * (kPatch * kPatch) + ky * kPatch;
  const int w = g * kPatch;
  for (int x = lane_id(); x < w; x += 32) {
    const int px = x / kPatch;
    const int kx = x - px * kPatch;
    dst[static_cast<size_t>(px) * kPatchKPad + kx] = __float2bfloat16_rn(__ldg(src + x));
  }
}

// zero the 4 pad columns (588..591) of every patch row once per call
__global__ void patch_pad_kernel(__nv_bfloat16* __restrict__ patches, int rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) {
    uint2 z = make_uint2(0u, 0u);
    *reinterpret_cast<uint2*>(patches + static_cast<size_t>(r) * kPatchKPad + kPatchK) = z;
  }
}

// ---------------------------------------------------------------------------------------------
// K1a: uint8 HWC image at model resolution -> ToTensor (/255) -> Normalize(mean,std) -> bf16 patch rows
//      (demo.py:162-166 with a source already S x S, so Resize is the identity)
// One warp per (b, image row y): reads 3*S contiguous bytes, writes pieces of 14 to each patch row.
// ---------------------------------------------------------------------------------------------
__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ img, __nv_bfloat16* __restrict__ patches, int B, int S,
                                     int g, float3 mean, float3 inv_std) {
  const int line = blockIdx.x * (blockDim.x >> 5) + warp_id();
  const int lines = B * g * kPatch;
  if (line >= lines) return;
  const int y = line % (g * kPatch);
  const int b = line / (g * kPatch);
  const int py = y / kPatch;
  const int ky = y - py * kPatch;
  const uint8_t* src = img + (static_cast<size_t>(b) * S + y) * S * 3;
  __nv_bfloat16* dst = patches + (static_cast<size_t>(b) * g * g + static_cast<size_t>(py) * g) * kPatchKPad + ky * kPatch;
  const int w = g * kPatch * 3;
  for (int i = lane_id(); i < w; i += 32) {
    const int x = i / 3;
    const int c = i - x * 3;
    const int px = x / kPatch;
    const int kx = x - px * kPatch;
    const float m = c == 0 ? mean.x : (c == 1 ? mean.y : mean.z);
    const float s = c == 0 ? inv_std.x : (c == 1 ? inv_std.y : inv_std.z);
    const float v = (static_cast<float>(src[i]) / 255.0f - m) * s;
    dst[static_cast<size_t>(px) * kPatchKPad + c * (kPatch * kPatch) + kx] = __float2bfloat16_rn(v);
  }
}

// x[b, 0, :] = cls_token + pos[0, :]   (HF modeling_dinov2.py:108-112)
__global__ void cls_rows_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos, int B,
                                int T, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D;
  const int d = i - b * D;
  x[static_cast<size_t>(b) * T * D + d] = cls[d] + pos[d];
}

// ---------------------------------------------------------------------------------------------
// K3: LayerNorm over D = 768 (eps 1e-6), fp32 in, bf16 or fp32 out.  One warp per row, the row stays in registers
// (24 floats per lane), two-pass mean / variance like ATen.  HF modeling_dinov2.py:354,359,449.
// ---------------------------------------------------------------------------------------------
template <int D, typename OutT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, OutT* __restrict__ out, int rows,
                                                         float eps) {
  constexpr int kVec = D / 128;  // float4 per lane
  const int row = blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (row >= rows) return;
  const int lane = lane_id();
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
  float4 v[kVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i);
    const float4 be = __ldg(b4 + lane + 32 * i);
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + be.x;
    y.y = (v[i].y - mean) * rstd * g.y + be.y;
    y.z = (v[i].z - mean) * rstd * g.z + be.z;
    y.w = (v[i].w - mean) * rstd * g.w + be.w;
    if constexpr (sizeof(OutT) == 2) {
      uint2 w = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
      reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * D)[lane + 32 * i] = w;
    } else {
      reinterpret_cast<float4*>(out + static_cast<size_t>(row) * D)[lane + 32 * i] = y;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Focal-stream GEMM operand: xin[b, n, :] = bf16( tokens[b, 1+n, :] * rowscale[b, n] + PE[n, :] )
//   rowscale = prod_k (1 + focus_strength * a_k[b, n]) is the re-focus of reference src/model.py:426 folded into a
//   per-row scalar; PE is the 2-D sinusoidal table of src/model.py:140-184 (precomputed once per grid).
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) focal_input_kernel(const float* __restrict__ tokens, const float* __restrict__ pe,
                                                           const float* __restrict__ rowscale,
                                                           __nv_bfloat16* __restrict__ xin, int B, int N) {
  constexpr int kVec = D / 128;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (row >= B * N) return;
  const int b = row / N;
  const int n = row - b * N;
  const int lane = lane_id();
  const float sc = rowscale ? rowscale[row] : 1.0f;
  const float4* t4 = reinterpret_cast<const float4*>(tokens + (static_cast<size_t>(b) * (N + 1) + 1 + n) * D);
  const float4* p4 = reinterpret_cast<const float4*>(pe + static_cast<size_t>(n) * D);
  uint2* o = reinterpret_cast<uint2*>(xin + static_cast<size_t>(row) * D);
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const float4 t = t4[lane + 32 * i];
    const float4 p = __ldg(p4 + lane + 32 * i);
    o[lane + 32 * i] = make_uint2(pack_bf16x2(fmaf(t.x, sc, p.x), fmaf(t.y, sc, p.y)),
                                  pack_bf16x2(fmaf(t.z, sc, p.z), fmaf(t.w, sc, p.w)));
  }
}

}  // namespace

int patchify_f32_launch(const float* images, __nv_bfloat16* patches, int B, int S, cudaStream_t stream) {
  CA_REQUIRE(images && patches, "patchify: null pointer");
  CA_REQUIRE(B > 0 && S >= kPatch, "patchify: bad shape");
  const int g = S / kPatch;
  const int rows = B * g * g;
  patch_pad_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(patches, rows);
  const int lines = B * 3 * g * kPatch;
  patchify_f32_kernel<<<(lines + 7) / 8, 256, 0, stream>>>(images, patches, B, S, g);
  CA_CUDA(cudaGetLastError());
  return 0;
}

int preprocess_u8_launch(const uint8_t* images, __nv_bfloat16* patches, int B, int S, const float* mean3,
                         const float* std3, cudaStream_t stream) {
  CA_REQUIRE(images && patches && mean3 && std3, "preprocess: null pointer");
  CA_REQUIRE(B > 0 && S >= kPatch, "preprocess: bad shape");
  const int g = S / kPatch;
  const int rows = B * g * g;
  patch_pad_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(patches, rows);
  const int lines = B * g * kPatch;
  preprocess_u8_kernel<<<(lines + 7) / 8, 256, 0, stream>>>(
      images, patches, B, S, g, make_float3(mean3[0], mean3[1], mean3[2]),
      make_float3(1.0f / std3[0], 1.0f / std3[1], 1.0f / std3[2]));
  CA_CUDA(cudaGetLastError());
  return 0;
}

// dst[i] = src[i], src in PINNED HOST memory read by the SMs over PCIe (UVA).  For the few hundred KB of per-call inputs
// (per-call projection, curiosity draws): a cudaMemcpyAsync would queue on the H2D copy engine behind whatever bulk
// upload the application has in flight there (the next image batch, ~2 ms) and stall the compute stream for that long.
__global__ void fetch_pinned_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = src[i];
}

int fetch_pinned_launch(float* dst, const float* src_pinned_host, size_t n, cudaStream_t stream) {
  CA_REQUIRE(dst && src_pinned_host, "fetch_pinned: null pointer");
  if (n == 0) return 0;
  cudaPointerAttributes attr;
  CA_CUDA(cudaPointerGetAttributes(&attr, src_pinned_host));
  CA_REQUIRE(attr.type == cudaMemoryTypeHost, "fetch_pinned: the source must be page-locked (pinned) host memory");
  const int grid = static_cast<int>((n + 255) / 256 < 64 ? (n + 255) / 256 : 64);
  fetch_pinned_kernel<<<grid, 256, 0, stream>>>(dst, static_cast<const float*>(attr.devicePointer), n);
  CA_CUDA(cudaGetLastError());
  return 0;
}

int cls_rows_launch(float* x, const float* cls, const float* pos, int B, int T, int D, cudaStream_t stream) {
  CA_REQUIRE(x && cls && pos, "cls_rows: null pointer");
  cls_rows_kernel<<<(B * D + 255) / 256, 256, 0, stream>>>(x, cls, pos, B, T, D);
  CA_CUDA(cudaGetLastError());
  return 0;
}

int layernorm_launch(const float* x, const float* gamma, const float* beta, void* out, int out_is_bf16, int rows, int D,
                     float eps, cudaStream_t stream) {
  CA_REQUIRE(x && gamma && beta && out, "layernorm: null pointer");
  CA_REQUIRE(D == 768, "layernorm: only D = 768 (ViT-B) is instantiated");
  CA_REQUIRE(rows > 0, "layernorm: no rows");
  const int grid = (rows + 7) / 8;
  if (out_is_bf16)
    layernorm_kernel<768, __nv_bfloat16><<<grid, 256, 0, stream>>>(x, gamma, beta, static_cast<__nv_bfloat16*>(out),
                                                                   rows, eps);
  else
    layernorm_kernel<768, float><<<grid, 256, 0, stream>>>(x, gamma, beta, static_cast<float*>(out), rows, eps);
  CA_CUDA(cudaGetLastError());
  return 0;
}

int focal_input_launch(const float* tokens, const float* pe, const float* rowscale, __nv_bfloat16* xin, int B, int N,
                       int D, cudaStream_t stream) {
  CA_REQUIRE(tokens && pe && xin, "focal_input: null pointer");
  CA_REQUIRE(D == 768, "focal_input: only D = 768 is instantiated");
  const int rows = B * N;
  focal_input_kernel<768><<<(rows + 7) / 8, 256, 0, stream>>>(tokens, pe, rowscale, xin, B, N);
  CA_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ca

// Bandwidth-bound row kernels of the backbone: preprocessing / patchify, CLS rows, LayerNorm, focal-stream input.
// All are one-pass, 128-bit vectorised, coalesced along the contiguous dimension; their roofline is HBM.
#include "common.cuh"
#include "host.h"

#include <stdlib.h>
#include "rowops.cuh"

namespace ca {
namespace {

constexpr int kPatch = 14;
constexpr int kPatchK = 3 * kPatch * kPatch;  // 588
constexpr int kPatchKPad = 592;               // row stride of the patch matrix: 16-byte multiple for TMA

// ---------------------------------------------------------------------------------------------
// K1: image -> bf16 patch rows (im2col of the 14 x 14 / stride 14 patch embedding, HF modeling_dinov2.py:139-148)
//      row (b, py, px), column k = c*196 + ky*14 + kx  (Conv2d weight.flatten(1) order), columns 588..591 zero
//   kU8 = false: fp32 CHW, already normalised (what `forward` receives)
//   kU8 = true : uint8 HWC at model resolution -> ToTensor (/255) -> Normalize(mean, std)  (demo.py:162-166)
// One CTA per (image, patch row py).  Phase 1 stages the 14 image lines of that patch row in shared memory with fully
// coalesced reads (for HWC uint8 they are ONE contiguous region of 14*S*3 bytes, read 16 bytes per thread); phase 2
// writes the g patch rows whole — 1184 contiguous bytes each, 4-byte stores, consecutive lanes on consecutive words.
// (The first version scattered 28-byte pieces from one warp per image line: 0.16 / 0.49 of the HBM rate,
// profiles/README.md round 2.)
// ---------------------------------------------------------------------------------------------
template <bool kU8>
__global__ void __launch_bounds__(320) patch_rows_kernel(const void* __restrict__ img_v,
                                                          __nv_bfloat16* __restrict__ patches, int S, int g, float3 mean,
                                                          float3 inv_std) {
  griddep_sync();  // PDL: nothing before this line reads or writes global memory
  extern __shared__ __align__(16) uint8_t sm_all[];
  // uint8 path: the first 3 KB hold lut[c][v] = (v / 255 - mean_c) / std_c, the value ToTensor + Normalize give byte v of
  // channel c — one shared-memory read instead of an int->float conversion, an IEEE division and an FMA per element
  float* lut = reinterpret_cast<float*>(sm_all);
  uint8_t* sm = sm_all + (kU8 ? 3 * 256 * sizeof(float) : 0);
  const int py = blockIdx.x;
  const int b = blockIdx.y;
  const int w = g * kPatch;
  int head = 0;
  if constexpr (kU8) {
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
      const int c = i >> 8;
      const float m = c == 0 ? mean.x : (c == 1 ? mean.y : mean.z);
      const float sd = c == 0 ? inv_std.x : (c == 1 ? inv_std.y : inv_std.z);
      lut[i] = (static_cast<float>(i & 255) / 255.0f - m) * sd;
    }
    const uint8_t* src = static_cast<const uint8_t*>(img_v) + (static_cast<size_t>(b) * S + py * kPatch) * S * 3;
    const int len = kPatch * S * 3;
    head = static_cast<int>(reinterpret_cast<uintptr_t>(src) & 15u);
    const uint4* s4 = reinterpret_cast<const uint4*>(src - head);  // aligned down: stays inside the allocation
    const int n_full = (head + len) / 16;                          // 16-byte words that end inside the region
    for (int i = threadIdx.x; i < n_full; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = __ldg(s4 + i);
    for (int i = n_full * 16 + threadIdx.x; i < head + len; i += blockDim.x) sm[i] = __ldg(src - head + i);
  } else {
    const float* img = static_cast<const float*>(img_v);
    __nv_bfloat16* t = reinterpret_cast<__nv_bfloat16*>(sm);  // [3 * 14][w]
    for (int line = warp_id(); line < 3 * kPatch; line += (blockDim.x >> 5)) {
      const int c = line / kPatch;
      const int ky = line - c * kPatch;
      const float* src = img + ((static_cast<size_t>(b) * 3 + c) * S + (py * kPatch + ky)) * S;
      for (int x = lane_id(); x < w; x += 32) t[line * w + x] = __float2bfloat16_rn(__ldg(src + x));
    }
  }
  __syncthreads();
  // Phase 2: thread kp owns the bf16x2 word kp (columns 2kp, 2kp+1) of EVERY patch row of this CTA: the column -> (c, ky, kx)
  // decode happens once per thread, the loop over the g patches only steps the source by one patch width, and the
  // stores of a warp are 32 consecutive words of one patch row.
  constexpr int kPairs = kPatchKPad / 2;  // 296 bf16x2 words per patch row
  const int kp = threadIdx.x;
  if (kp >= kPairs) return;
  int off[2];
  bool ok[2];
  int ch[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int k = 2 * kp + e;
    ok[e] = k < kPatchK;
    const int c = ok[e] ? k / (kPatch * kPatch) : 0;
    const int r = ok[e] ? k - c * (kPatch * kPatch) : 0;
    const int ky = r / kPatch;
    const int kx = r - ky * kPatch;
    ch[e] = c;
    off[e] = kU8 ? head + (ky * S + kx) * 3 + c : (c * kPatch + ky) * w + kx;
  }
  uint32_t* out = reinterpret_cast<uint32_t*>(patches + (static_cast<size_t>(b) * g * g + static_cast<size_t>(py) * g) * kPatchKPad) + kp;
  const int step = kU8 ? kPatch * 3 : kPatch;
  if constexpr (kU8) {
    const float* l0 = lut + ch[0] * 256;
    const float* l1 = lut + ch[1] * 256;
#pragma unroll 4
    for (int px = 0; px < g; ++px) {
      const float a = ok[0] ? l0[sm[off[0] + px * step]] : 0.f;
      const float c = ok[1] ? l1[sm[off[1] + px * step]] : 0.f;
      out[static_cast<size_t>(px) * kPairs] = pack_bf16x2(a, c);
    }
  } else {
    const uint16_t* t = reinterpret_cast<const uint16_t*>(sm);
#pragma unroll 4
    for (int px = 0; px < g; ++px) {
      const uint32_t a = ok[0] ? t[off[0] + px * step] : 0u;
      const uint32_t c = ok[1] ? t[off[1] + px * step] : 0u;
      out[static_cast<size_t>(px) * kPairs] = a | (c << 16);
    }
  }
}

// x[b, 0, :] = cls_token + pos[0, :]   (HF modeling_dinov2.py:108-112)
__global__ void cls_rows_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos, int B,
                                int T, int D) {
  griddep_sync();  // PDL: nothing before this line reads or writes global memory
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
                                                         float eps, int ld_out, int reverse) {
  griddep_sync();  // PDL: nothing before this line reads or writes global memory
  constexpr int kVec = D / 128;  // float4 per lane
  int row = blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (row >= rows) return;
  // the GEMM before this kernel wrote the rows in ascending order, so the LAST rows are the ones still in the L2:
  // start there (the block scheduler hands out blockIdx in ascending order)
  if (reverse) row = rows - 1 - row;
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
      reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * ld_out)[lane + 32 * i] = w;
    } else {
      reinterpret_cast<float4*>(out + static_cast<size_t>(row) * ld_out)[lane + 32 * i] = y;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Entry of the LayerNorm-folded layer chain (gemm.cuh, EPI_LN_* / EPI_RESID_LN_F32): what the residual epilogue leaves
// behind for every later layer, produced here for the rows the embedding wrote — the raw rows as bf16 and, per 128-column
// span (the 4 floats x 32 lanes of one register slot), the sum and the sum of squared deviations from the span mean.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) ln_shadow_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ shadow,
                                                         int ld_shadow, float2* __restrict__ stats, int rows) {
  griddep_sync();  // PDL: nothing before this line reads or writes global memory
  constexpr int kVec = D / 128;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (row >= rows) return;
  const int lane = lane_id();
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
  float4 v[kVec];
#pragma unroll
  for (int i = 0; i < kVec; ++i) v[i] = xr[lane + 32 * i];
  float mine_sum = 0.f, mine_m2 = 0.f;  // lane i keeps span i
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const float sum = warp_sum((v[i].x + v[i].y) + (v[i].z + v[i].w));
    const float m = sum * (1.0f / 128.f);
    const float a = v[i].x - m, b = v[i].y - m, c = v[i].z - m, d = v[i].w - m;
    const float m2 = warp_sum((a * a + b * b) + (c * c + d * d));
    if (lane == i) {
      mine_sum = sum;
      mine_m2 = m2;
    }
    reinterpret_cast<uint2*>(shadow + static_cast<size_t>(row) * ld_shadow)[lane + 32 * i] =
        make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
  }
  if (lane < kVec) stats[static_cast<size_t>(row) * kVec + lane] = make_float2(mine_sum, mine_m2);
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
  griddep_sync();  // PDL: nothing before this line reads or writes global memory
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

template <bool kU8>
static int patch_rows_launch(const void* images, __nv_bfloat16* patches, int B, int S, float3 mean, float3 inv_std,
                             cudaStream_t stream) {
  const int g = S / kPatch;
  const size_t smem = kU8 ? static_cast<size_t>(kPatch) * S * 3 + 32 + 3 * 256 * sizeof(float) : static_cast<size_t>(3) * kPatch * g * kPatch * 2;
  CA_REQUIRE(smem <= 200 * 1024, "patchify: image side too large for the shared-memory staging (max ~2400)");
  static PerDeviceOnce configured;  // per instantiation and per device
  CA_TRY(configured.run([&]() -> int {
    CA_CUDA(cudaFuncSetAttribute(patch_rows_kernel<kU8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return 0;
  }));
  CA_TRY(launch_kernel(patch_rows_kernel<kU8>, dim3(dim3(g, B)), dim3(320), smem, stream, images, patches, S, g, mean, inv_std));
  CA_CUDA(cudaGetLastError());
  return 0;
}

int patchify_f32_launch(const float* images, __nv_bfloat16* patches, int B, int S, cudaStream_t stream) {
  CA_REQUIRE(images && patches, "patchify: null pointer");
  CA_REQUIRE(B > 0 && S >= kPatch, "patchify: bad shape");
  return patch_rows_launch<false>(images, patches, B, S, make_float3(0.f, 0.f, 0.f), make_float3(1.f, 1.f, 1.f), stream);
}

int preprocess_u8_launch(const uint8_t* images, __nv_bfloat16* patches, int B, int S, const float* mean3,
                         const float* std3, cudaStream_t stream) {
  CA_REQUIRE(images && patches && mean3 && std3, "preprocess: null pointer");
  CA_REQUIRE(B > 0 && S >= kPatch, "preprocess: bad shape");
  return patch_rows_launch<true>(images, patches, B, S, make_float3(mean3[0], mean3[1], mean3[2]),
                                 make_float3(1.0f / std3[0], 1.0f / std3[1], 1.0f / std3[2]), stream);
}

// dst[i] = src[i], src in PINNED HOST memory read by the SMs over PCIe (UVA).  For the few hundred KB of per-call inputs
// (per-call projection, curiosity draws): a cudaMemcpyAsync would queue on the H2D copy engine behind whatever bulk
// upload the application has in flight there (the next image batch, ~2 ms) and stall the compute stream for that long.
__global__ void fetch_pinned_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
  griddep_sync();  // PDL: nothing before this line reads or writes global memory
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
  CA_TRY(launch_kernel(fetch_pinned_kernel, dim3(grid), dim3(256), 0, stream, dst, static_cast<const float*>(attr.devicePointer), n));
  CA_CUDA(cudaGetLastError());
  return 0;
}

int cls_rows_launch(float* x, const float* cls, const float* pos, int B, int T, int D, cudaStream_t stream) {
  CA_REQUIRE(x && cls && pos, "cls_rows: null pointer");
  CA_TRY(launch_kernel(cls_rows_kernel, dim3((B * D + 255) / 256), dim3(256), 0, stream, x, cls, pos, B, T, D));
  CA_CUDA(cudaGetLastError());
  return 0;
}

int layernorm_launch(const float* x, const float* gamma, const float* beta, void* out, int out_is_bf16, int rows, int D,
                     float eps, cudaStream_t stream, int ld_out) {
  CA_REQUIRE(x && gamma && beta && out, "layernorm: null pointer");
  if (ld_out <= 0) ld_out = D;
  CA_REQUIRE(ld_out >= D && ld_out % (out_is_bf16 ? 4 : 4) == 0, "layernorm: output leading dimension must be >= D and keep rows 8/16-byte aligned");
  CA_REQUIRE(D == 768, "layernorm: only D = 768 (ViT-B) is instantiated");
  CA_REQUIRE(rows > 0, "layernorm: no rows");
  const int grid = (rows + 7) / 8;
  static const int reverse = getenv("CA_LN_FORWARD") ? 0 : 1;
  if (out_is_bf16)
    CA_TRY(launch_kernel(layernorm_kernel<768, __nv_bfloat16>, dim3(grid), dim3(256), 0, stream, x, gamma, beta, static_cast<__nv_bfloat16*>(out),
                                                                   rows, eps, ld_out, reverse));
  else
    CA_TRY(launch_kernel(layernorm_kernel<768, float>, dim3(grid), dim3(256), 0, stream, x, gamma, beta, static_cast<float*>(out), rows, eps, ld_out, reverse));
  CA_CUDA(cudaGetLastError());
  return 0;
}

int ln_shadow_launch(const float* x, __nv_bfloat16* shadow, int ld_shadow, float* stats, int rows, int D,
                     cudaStream_t stream) {
  CA_REQUIRE(x && shadow && stats, "ln_shadow: null pointer");
  CA_REQUIRE(D == 768, "ln_shadow: only D = 768 (ViT-B) is instantiated");
  CA_REQUIRE(rows > 0, "ln_shadow: no rows");
  CA_REQUIRE(ld_shadow >= D && ld_shadow % 4 == 0, "ln_shadow: the shadow rows must be >= D wide and 8-byte aligned");
  CA_TRY(launch_kernel(ln_shadow_kernel<768>, dim3((rows + 7) / 8), dim3(256), 0, stream, x, shadow, ld_shadow,
                       reinterpret_cast<float2*>(stats), rows));
  CA_CUDA(cudaGetLastError());
  return 0;
}

int focal_input_launch(const float* tokens, const float* pe, const float* rowscale, __nv_bfloat16* xin, int B, int N,
                       int D, cudaStream_t stream) {
  CA_REQUIRE(tokens && pe && xin, "focal_input: null pointer");
  CA_REQUIRE(D == 768, "focal_input: only D = 768 is instantiated");
  const int rows = B * N;
  CA_TRY(launch_kernel(focal_input_kernel<768>, dim3((rows + 7) / 8), dim3(256), 0, stream, tokens, pe, rowscale, xin, B, N));
  CA_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace ca

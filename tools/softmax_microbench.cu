// Micro-benchmark of the attention kernel's per-step softmax block (one query row per thread, 64 scores per step): how
// many clocks does ONE warp need for it alone on its SM sub-partition, and two warps sharing one?  Variants of the
// instruction mix / source order are compared here before they go into csrc/attention.cu.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cognitive_aim_depth_estimation_b200/csrc
//             tools/softmax_microbench.cu -o build/softmax_microbench
#include "common.cuh"
#include <stdio.h>
using namespace ca;

constexpr float kScale = 0.18033688f;

__device__ __forceinline__ void exp2_poly2_ref(float& t0, float& t1) {  // the shipped polynomial (degree 3, clamped)
  constexpr float kMagic = 12582912.0f;
  t0 = fmaxf(t0, -125.0f);
  t1 = fmaxf(t1, -125.0f);
  float r0, r1, n0, n1, f0, f1, p0, p1;
  ffma2v(r0, r1, t0, t1, 1.0f, 1.0f, kMagic, kMagic);
  ffma2v(n0, n1, r0, r1, 1.0f, 1.0f, -kMagic, -kMagic);
  ffma2v(f0, f1, n0, n1, -1.0f, -1.0f, t0, t1);
  ffma2v(p0, p1, f0, f1, 0.05520550534129143f, 0.05520550534129143f, 0.24261397123336792f, 0.24261397123336792f);
  ffma2v(p0, p1, p0, p1, f0, f1, 0.6932547688484192f, 0.6932547688484192f);
  ffma2v(p0, p1, p0, p1, f0, f1, 0.9999276995658875f, 0.9999276995658875f);
  t0 = __int_as_float(__float_as_int(p0) + (__float_as_int(r0) << 23));
  t1 = __int_as_float(__float_as_int(p1) + (__float_as_int(r1) << 23));
}

// VARIANT bits: 1 = no row sum (ones-column trick), 2 = MUFU issued first (breadth-first source order),
//               4 = degree-2 polynomial, 8 = no clamp
template <int kVariant, unsigned kPolyMask>
__device__ __forceinline__ float softmax_block(const uint32_t (&v)[64], uint32_t (&pk)[32], float& m_acc, float& l_run) {
  // row max
  float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
  for (int c = 4; c < 64; c += 8) {
    m0 = fmaxf(m0, fmaxf(__uint_as_float(v[c + 0]), __uint_as_float(v[c + 1])));
    m1 = fmaxf(m1, fmaxf(__uint_as_float(v[c + 2]), __uint_as_float(v[c + 3])));
    if (c + 4 < 64) {
      m2 = fmaxf(m2, fmaxf(__uint_as_float(v[c + 4]), __uint_as_float(v[c + 5])));
      m3 = fmaxf(m3, fmaxf(__uint_as_float(v[c + 6]), __uint_as_float(v[c + 7])));
    }
  }
  const float tile_max = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * kScale;
  if (__any_sync(0xffffffffu, tile_max > m_acc + 8.0f)) {
    const float m_new = fmaxf(m_acc, tile_max);
    l_run *= fast_exp2(m_acc - m_new);
    m_acc = m_new;
  }
  const float neg_m = -m_acc;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if constexpr ((kVariant & 2) == 0) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      float e[8];
#pragma unroll
      for (int kk = 0; kk < 8; kk += 2) {
        ffma2(e[kk], e[kk + 1], __uint_as_float(v[8 * t + kk]), __uint_as_float(v[8 * t + kk + 1]), kScale, neg_m);
        if (kk >= 8 - 2 * static_cast<int>((kPolyMask >> (2 * t)) & 3u)) {
          exp2_poly2_ref(e[kk], e[kk + 1]);
        } else {
          e[kk] = fast_exp2(e[kk]);
          e[kk + 1] = fast_exp2(e[kk + 1]);
        }
      }
      if constexpr ((kVariant & 1) == 0) {
        fadd2(s0, s1, e[0], e[1]);
        fadd2(s2, s3, e[2], e[3]);
        fadd2(s0, s1, e[4], e[5]);
        fadd2(s2, s3, e[6], e[7]);
      }
      pk[4 * t + 0] = pack_bf16x2(e[0], e[1]);
      pk[4 * t + 1] = pack_bf16x2(e[2], e[3]);
      pk[4 * t + 2] = pack_bf16x2(e[4], e[5]);
      pk[4 * t + 3] = pack_bf16x2(e[6], e[7]);
    }
  } else {
    // breadth-first: all scale FMAs, then the MUFU elements in program order with the polynomial stages spread between them
    float e[64];
#pragma unroll
    for (int c = 0; c < 64; c += 2) ffma2(e[c], e[c + 1], __uint_as_float(v[c]), __uint_as_float(v[c + 1]), kScale, neg_m);
    constexpr float kMagic = 12582912.0f;
    float r[64], f[64], p[64];
    // which pairs take the polynomial: the last `cnt` pairs of each group of 4 pairs
    auto is_poly = [](int pair) {
      const int t = pair >> 2, kk = (pair & 3) * 2;
      return kk >= 8 - 2 * static_cast<int>((kPolyMask >> (2 * t)) & 3u);
    };
#pragma unroll
    for (int stage = 0; stage < 8; ++stage) {
#pragma unroll
      for (int pair = 0; pair < 32; ++pair) {
        const int c = 2 * pair;
        if (is_poly(pair)) {
          if (stage == 0) {
            if constexpr ((kVariant & 8) == 0) {
              e[c] = fmaxf(e[c], -125.0f);
              e[c + 1] = fmaxf(e[c + 1], -125.0f);
            }
            ffma2v(r[c], r[c + 1], e[c], e[c + 1], 1.0f, 1.0f, kMagic, kMagic);
          } else if (stage == 1) {
            ffma2v(f[c], f[c + 1], r[c], r[c + 1], 1.0f, 1.0f, -kMagic, -kMagic);
          } else if (stage == 2) {
            ffma2v(f[c], f[c + 1], f[c], f[c + 1], -1.0f, -1.0f, e[c], e[c + 1]);
          } else if (stage == 3) {
            if constexpr (kVariant & 4)
              ffma2v(p[c], p[c + 1], f[c], f[c + 1], 0.2402265f, 0.2402265f, 0.6931472f, 0.6931472f);
            else
              ffma2v(p[c], p[c + 1], f[c], f[c + 1], 0.05520550534129143f, 0.05520550534129143f, 0.24261397123336792f, 0.24261397123336792f);
          } else if (stage == 4) {
            if constexpr (kVariant & 4)
              ffma2v(p[c], p[c + 1], p[c], p[c + 1], f[c], f[c + 1], 1.0f, 1.0f);
            else
              ffma2v(p[c], p[c + 1], p[c], p[c + 1], f[c], f[c + 1], 0.6932547688484192f, 0.6932547688484192f);
          } else if (stage == 5) {
            if constexpr ((kVariant & 4) == 0)
              ffma2v(p[c], p[c + 1], p[c], p[c + 1], f[c], f[c + 1], 0.9999276995658875f, 0.9999276995658875f);
          } else if (stage == 6) {
            e[c] = __int_as_float(__float_as_int(p[c]) + (__float_as_int(r[c]) << 23));
            e[c + 1] = __int_as_float(__float_as_int(p[c + 1]) + (__float_as_int(r[c + 1]) << 23));
          }
        } else {
          // 24 MUFU pairs spread over the first 6 stages: 4 pairs per stage
          int rank = 0;
          for (int q = 0; q < pair; ++q) rank += is_poly(q) ? 0 : 1;
          if (rank / 4 == stage) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[c]));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e[c + 1]));
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if constexpr ((kVariant & 1) == 0) {
        fadd2(s0, s1, e[8 * t + 0], e[8 * t + 1]);
        fadd2(s2, s3, e[8 * t + 2], e[8 * t + 3]);
        fadd2(s0, s1, e[8 * t + 4], e[8 * t + 5]);
        fadd2(s2, s3, e[8 * t + 6], e[8 * t + 7]);
      }
      pk[4 * t + 0] = pack_bf16x2(e[8 * t + 0], e[8 * t + 1]);
      pk[4 * t + 1] = pack_bf16x2(e[8 * t + 2], e[8 * t + 3]);
      pk[4 * t + 2] = pack_bf16x2(e[8 * t + 4], e[8 * t + 5]);
      pk[4 * t + 3] = pack_bf16x2(e[8 * t + 6], e[8 * t + 7]);
    }
  }
  l_run += (s0 + s1) + (s2 + s3);
  return l_run;
}

template <int kVariant, unsigned kPolyMask>
__global__ void __launch_bounds__(512) bench(int iters, long long* cycles, float* sink) {
  __shared__ uint4 sm_in[512 * 16 / 4];   // 64 floats per thread would be 64 KB: use a shared 16-float pattern per thread
  __shared__ uint4 sm_out[512 * 2];
  for (int i = threadIdx.x; i < 512 * 4; i += blockDim.x) {
    float a = -3.f + 0.01f * (i % 97), b = 1.f - 0.02f * (i % 53), c = 0.5f * (i % 7), d = -0.3f * (i % 11);
    sm_in[i] = make_uint4(__float_as_uint(a), __float_as_uint(b), __float_as_uint(c), __float_as_uint(d));
  }
  __syncthreads();
  float m_acc = -INFINITY, l_run = 0.f;
  uint32_t pk[32];
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[64];
#pragma unroll
    for (int q = 0; q < 16; ++q) {  // 16 LDS.128 stand in for the two tcgen05.ld
      const uint4 w = lds128(smem_u32(&sm_in[(threadIdx.x * 4 + ((q + it) & 3))]));  // volatile asm, like tcgen05.ld: no re-loads
      v[4 * q + 0] = w.x + q;
      v[4 * q + 1] = w.y ^ (q << 3);
      v[4 * q + 2] = w.z + (it & 1);
      v[4 * q + 3] = w.w;
    }
    softmax_block<kVariant, kPolyMask>(v, pk, m_acc, l_run);
#pragma unroll
    for (int q = 0; q < 2; ++q)  // stand-in for the tcgen05.st (keeps every packed value alive)
      sm_out[threadIdx.x * 2 + q] = make_uint4(pk[16 * q] ^ pk[16 * q + 4] ^ pk[16 * q + 8] ^ pk[16 * q + 12],
                                               pk[16 * q + 1] ^ pk[16 * q + 5] ^ pk[16 * q + 9] ^ pk[16 * q + 13],
                                               pk[16 * q + 2] ^ pk[16 * q + 6] ^ pk[16 * q + 10] ^ pk[16 * q + 14],
                                               pk[16 * q + 3] ^ pk[16 * q + 7] ^ pk[16 * q + 11] ^ pk[16 * q + 15]);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = l_run + __uint_as_float(sm_out[threadIdx.x].x);
}

template <int kVariant, unsigned kPolyMask>
void run(const char* name, long long* d, float* sink) {
  long long h[148];
  const int iters = 2000;
  for (int threads : {128, 256, 384, 512}) {
    bench<kVariant, kPolyMask><<<148, threads>>>(iters, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-44s warps/sub-partition %d: %s  %.0f clk per 64-column block per warp\n", name, threads / 128,
           cudaGetErrorString(e), double(h[0]) / iters);
  }
}

int main() {
  long long* d;
  float* sink;
  cudaMalloc(&d, 148 * 8);
  cudaMalloc(&sink, 148 * 512 * 4);
  run<0, 0x5555>("shipped (25% poly, sum)", d, sink);
  run<0, 0x0000>("all MUFU, sum", d, sink);
  run<0, 0x1111>("12.5% poly, sum", d, sink);
  run<1, 0x5555>("25% poly, no sum", d, sink);
  run<1, 0xAAAA>("50% poly, no sum", d, sink);
  run<2, 0x5555>("breadth-first 25% poly, sum", d, sink);
  run<3, 0x5555>("breadth-first 25% poly, no sum", d, sink);
  run<3, 0xAAAA>("breadth-first 50% poly, no sum", d, sink);
  run<7, 0xAAAA>("breadth-first 50% poly deg2, no sum", d, sink);
  run<15, 0xAAAA>("breadth-first 50% poly deg2 noclamp, no sum", d, sink);
  run<15, 0xFFFF>("breadth-first 100% poly deg2 noclamp, no sum", d, sink);
  run<3, 0x6666>("breadth-first 37.5% poly, no sum", d, sink);
  run<2, 0x6666>("breadth-first 37.5% poly, sum", d, sink);
  return 0;
}

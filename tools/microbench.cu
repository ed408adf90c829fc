// Micro-benchmarks that size the attention softmax loop on B200: TMEM read bandwidth (tcgen05.ld), MUFU.EX2 rate,
// FFMA2 rate.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I cognitive_aim_depth_estimation_b200/csrc
//                     tools/microbench.cu -o build/microbench ; run on the GPU box.
#include "common.cuh"
#include <stdio.h>
using namespace ca;

__global__ void tmem_ld_bw(int iters, long long* cycles, int shape) {
  __shared__ uint32_t slot;
  if (warp_id() == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((warp_id() & 3) * 32u << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (shape == 0) {
      uint32_t v[32];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(base + ((warp_id() >> 2) * 128 + c * 32) % 512, v);
        tmem_ld_wait();
        acc ^= v[0] ^ v[31];
      }
    } else {  // two loads in flight before the wait
      uint32_t v[32], w[32];
#pragma unroll
      for (int c = 0; c < 4; c += 2) {
        tmem_ld32(base + ((warp_id() >> 2) * 128 + c * 32) % 512, v);
        tmem_ld32(base + ((warp_id() >> 2) * 128 + c * 32 + 32) % 512, w);
        tmem_ld_wait();
        acc ^= v[0] ^ w[31];
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) printf("x");
  __syncthreads();
  if (warp_id() == 0) tmem_dealloc(slot, 512);
}

__global__ void mufu_rate(int iters, long long* cycles, float* sink, int mode) {
  float x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) x[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = fast_exp2(x[i]) - 1.0f;  // MUFU + FADD
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        asm volatile("{\n\t.reg .b64 ra, rd;\n\tmov.b64 ra, {%0, %1};\n\tfma.rn.f32x2 rd, ra, ra, ra;\n\tmov.b64 {%0, %1}, rd;\n\t}"
                     : "+f"(x[i]), "+f"(x[i + 1]));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += x[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  long long* d;
  float* sink;
  cudaMalloc(&d, 148 * 8);
  cudaMalloc(&sink, 148 * 1024 * 4);
  long long h[148];
  const int iters = 2000;
  for (int shape = 0; shape < 2; ++shape)
    for (int threads : {128, 256, 512}) {
      tmem_ld_bw<<<148, threads>>>(iters, d, shape);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double bytes = double(iters) * 4 * 32 * 32 * 4 * (threads / 32);
      printf("tmem_ld x32 shape%d threads %d: %s  %.1f B/clk/SM (cycles %lld)\n", shape, threads, cudaGetErrorString(e),
             bytes / h[0], h[0]);
    }
  for (int mode = 0; mode < 2; ++mode)
    for (int threads : {128, 256, 512}) {
      mufu_rate<<<148, threads>>>(iters, d, sink, mode);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double ops = double(iters) * 32 * threads * (mode ? 1 : 1);
      printf("%s threads %d: %s  %.2f lane-ops/clk/SM\n", mode ? "ffma2(pairs=2 flop-lanes)" : "mufu.ex2+fadd", threads,
             cudaGetErrorString(e), ops / h[0]);
    }
  return 0;
}

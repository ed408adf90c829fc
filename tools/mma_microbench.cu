// Micro-benchmarks that size the attention kernel's tensor-core protocol on B200 (round 2):
//   1. issue/execute rate of tcgen05.mma from ONE thread for the shapes the attention kernel uses
//      (A in TMEM or smem, N = 64 / 128 / 256, B K-major or MN-major), with a tcgen05.commit every `per_commit` MMAs
//   2. round trip "issue one MMA group + commit -> the same thread sees the mbarrier flip"
//   3. the minimal softmax <-> MMA ping-pong (tcgen05.ld, tcgen05.st, fence, arrive | wait, mma, commit) with no arithmetic
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I cognitive_aim_depth_estimation_b200/csrc
//             tools/mma_microbench.cu -o build/mma_microbench ; run on the GPU box.
#include "common.cuh"
#include <stdio.h>
#include <stdlib.h>
using namespace ca;

struct Cfg {
  int ts;          // 1: A from TMEM, 0: A from smem
  int n;           // MMA N
  int b_mn;        // 1: B MN-major (the P V form), 0: K-major (the Q K^T form)
  int per_commit;  // MMAs between commits (0 = commit only at the end)
  int iters;       // groups
  int group;       // MMAs per group
};

__global__ void __launch_bounds__(128) mma_rate(Cfg c, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp_id() == 0) tmem_alloc(&slot, 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp_id() == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, c.n, 0, c.b_mn);
    const uint64_t ad = umma_smem_desc_sw128(smem_u32(smem));
    const uint64_t bd = c.b_mn ? umma_smem_desc_sw128(smem_u32(smem + 16384), 1024, 1024)
                               : umma_smem_desc_sw128(smem_u32(smem + 16384));
    uint32_t n_commit = 0;
    const long long t0 = clock64();
    for (int it = 0; it < c.iters; ++it) {
      if (elect_one_sync()) {
        for (int j = 0; j < c.group; ++j) {
          if (c.ts)
            umma_bf16_ts(tm, tm + 224 + 8 * (j & 3), bd + 2 * (j & 3), idesc, 1);
          else
            umma_bf16(tm, ad + 2 * (j & 3), bd + 2 * (j & 3), idesc, 1);
          if (c.per_commit && ((j + 1) % c.per_commit) == 0) { umma_commit(&bar); ++n_commit; }
        }
      }
      __syncwarp();
    }
    const long long t_issue = clock64();
    // drain: one final commit, then wait for every arrival (phase flips once per arrival, count 1)
    n_commit = __shfl_sync(0xffffffffu, __reduce_max_sync(0xffffffffu, n_commit), 0);
    if (elect_one_sync()) umma_commit(&bar);
    __syncwarp();
    // the barrier completed n_commit + 1 phases; wait for the last one
    mbar_wait(&bar, n_commit & 1);
    const long long t1 = clock64();
    if (lane_id() == 0) {
      cycles[2 * blockIdx.x] = t1 - t0;
      cycles[2 * blockIdx.x + 1] = t_issue - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp_id() == 0) {
    tc_fence_after();
    tmem_dealloc(tm, 256);
  }
}

// one group of `group` MMAs + commit, then the issuing warp waits for the barrier: issue -> complete -> observed
__global__ void __launch_bounds__(128) mma_roundtrip(Cfg c, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp_id() == 0) tmem_alloc(&slot, 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp_id() == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, c.n, 0, c.b_mn);
    const uint64_t bd = c.b_mn ? umma_smem_desc_sw128(smem_u32(smem + 16384), 1024, 1024)
                               : umma_smem_desc_sw128(smem_u32(smem + 16384));
    const long long t0 = clock64();
    for (int it = 0; it < c.iters; ++it) {
      if (elect_one_sync()) {
        for (int j = 0; j < c.group; ++j) umma_bf16_ts(tm, tm + 224 + 8 * (j & 3), bd + 2 * (j & 3), idesc, 1);
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, it & 1);
      tc_fence_after();
    }
    const long long t1 = clock64();
    if (lane_id() == 0) cycles[2 * blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp_id() == 0) {
    tc_fence_after();
    tmem_dealloc(tm, 256);
  }
}

// minimal ping-pong: warps 4..7 ("softmax", one row per thread): wait s_full, tcgen05.ld `cols` columns, tcgen05.st cols/2,
// wait::st, fence, arrive p_full.  warp 0 ("MMA"): wait p_full, issue `group` MMAs, commit s_full.
__global__ void __launch_bounds__(256) pingpong(Cfg c, int cols, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t s_full, p_full;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&s_full, 1);
    mbar_init(&p_full, 4);
    fence_barrier_init();
  }
  if (warp_id() == 0) tmem_alloc(&slot, 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const long long t0 = clock64();
  if (warp_id() == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, c.n, 0, c.b_mn);
    const uint64_t bd = umma_smem_desc_sw128(smem_u32(smem + 16384));
    for (int it = 0; it < c.iters; ++it) {
      if (it > 0) {
        mbar_wait(&p_full, (it - 1) & 1);
        tc_fence_after();
      }
      if (elect_one_sync()) {
        for (int j = 0; j < c.group; ++j) umma_bf16_ts(tm, tm + 224 + 8 * (j & 3), bd + 2 * (j & 3), idesc, 1);
        umma_commit(&s_full);
      }
      __syncwarp();
    }
  } else if (warp_id() >= 4) {
    const uint32_t t_lane = tm + (((warp_id() & 3) * 32u) << 16);
    uint32_t acc = 0;
    for (int it = 0; it < c.iters; ++it) {
      mbar_wait(&s_full, it & 1);
      tc_fence_after();
      uint32_t v[32];
      for (int cc = 0; cc < cols; cc += 32) {
        tmem_ld32(t_lane + cc, v);
        tmem_ld_wait();
        acc ^= v[3];
      }
      for (int cc = 0; cc < cols / 2; cc += 32) tmem_st32(t_lane + cc, v);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&p_full);
    }
    if (acc == 0x1234567u) printf("x");
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 || threadIdx.x == 128) cycles[2 * blockIdx.x + (threadIdx.x ? 1 : 0)] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp_id() == 0) {
    tc_fence_after();
    tmem_dealloc(tm, 256);
  }
}

int main(int argc, char** argv) {
  long long* d;
  cudaMalloc(&d, 2 * 296 * 8);
  long long h[2 * 296];
  const int smem = 96 * 1024;
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_roundtrip, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(pingpong, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  auto report = [&](const char* what, const Cfg& c, int grid, cudaError_t e) {
    cudaMemcpy(h, d, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
    const double n_mma = double(c.iters) * c.group;
    printf("%-10s ts=%d n=%3d b_mn=%d per_commit=%d group=%d grid=%3d: %s  %.1f clk/MMA total (%.1f issue-only), "
           "ideal %d\n", what, c.ts, c.n, c.b_mn, c.per_commit, c.group, grid, cudaGetErrorString(e), h[0] / n_mma,
           h[1] / n_mma, c.n / 2);
  };
  for (int grid : {148, 296}) {
    for (int ts : {1, 0})
      for (int n : {64, 128, 256})
        for (int b_mn : {0, 1}) {
          if (b_mn && n != 64) continue;
          for (int per_commit : {0, 4, 1}) {
            Cfg c{ts, n, b_mn, per_commit, 2000, 8};
            mma_rate<<<grid, 128, smem>>>(c, d);
            report("rate", c, grid, cudaDeviceSynchronize());
          }
        }
  }
  for (int n : {64, 128})
    for (int group : {1, 4, 8}) {
      Cfg c{1, n, 0, 0, 2000, group};
      mma_roundtrip<<<148, 128, smem>>>(c, d);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, sizeof(long long) * 2 * 148, cudaMemcpyDeviceToHost);
      printf("roundtrip  n=%3d group=%d: %s  %.1f clk/iteration (tensor work %d)\n", n, group, cudaGetErrorString(e),
             double(h[0]) / c.iters, group * n / 2);
    }
  for (int grid : {148, 296})
    for (int n : {64, 128})
      for (int cols : {64, 128}) {
        Cfg c{1, n, 0, 0, 2000, 8};
        pingpong<<<grid, 256, smem>>>(c, cols, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
        printf("pingpong   grid=%d n=%3d group=8 cols=%d: %s  %.1f clk/iteration (tensor work %d)\n", grid, n, cols,
               cudaGetErrorString(e), double(h[0]) / c.iters, 8 * n / 2);
      }
  return 0;
}

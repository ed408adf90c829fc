// Host-side plumbing shared by the launchers: error capture and TMA tensor-map encoding.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>

namespace ca {

void set_error(const std::string& msg);  // records the message returned by ca_last_error()
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int invalid(const char* what);

#define CA_CUDA(call)                                                        \
  do {                                                                       \
    cudaError_t _e = (call);                                                 \
    if (_e != cudaSuccess) return ca::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define CA_REQUIRE(cond, msg)              \
  do {                                     \
    if (!(cond)) return ca::invalid(msg);  \
  } while (0)

#define CA_TRY(expr)            \
  do {                          \
    int _s = (expr);            \
    if (_s != 0) return _s;     \
  } while (0)

// bf16 tensor map, 128B swizzle, inner box = 64 elements (128 bytes).
// dims/strides are given innermost-first; strides in BYTES for dims 1..rank-1.
int make_tmap_bf16_sw128(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box);

// Convenience: row-major [rows, cols] bf16 matrix with leading dimension ld (elements),
// optionally batched (batch stride in elements); box = [box_rows, 64].
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t batch, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint64_t batch_stride, uint32_t box_rows);

// Row-major [rows, cols] fp32 matrix, 128B swizzle, box = [box_rows, 32 columns (128 bytes)] — the destination of the
// residual epilogue's TMA reduce-add.
int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

// Kernel launch, optionally with the programmatic-dependent-launch attribute (CA_PDL=1 in the environment): the kernel
// may then start while its predecessor on the stream drains; it synchronises with it through griddep_sync() (common.cuh).
// OFF by default: measured on B200 (round 2, same box, alternating runs) the guided forward at 32 x 518^2 takes 12.70 ms
// per step with it against 12.22 ms without, and a single image 1.39 ms either way — the early-resident CTAs of the
// next kernel cost more than the launch latency they hide.
bool pdl_enabled();
template <class... KArgs, class... Args>
int launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  CA_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  return 0;
}

int sm_count();         // SMs of the CURRENT device (cached per device)
int current_device();   // cudaGetDevice, -1 on failure

// One-time setup that is per CUDA context, not per process (cudaFuncSetAttribute opt-ins above 48 KB of shared memory):
// `run(f)` calls f() the first time it is reached with a given device current, under a lock, and returns f's status
// (0 = ok; a failure is retried on the next call).
class PerDeviceOnce {
 public:
  template <class F>
  int run(F&& f) {
    const int dev = current_device();
    if (dev < 0 || dev >= 64) return f();
    const unsigned long long bit = 1ull << dev;
    std::lock_guard<std::mutex> lock(mu_);
    if (done_ & bit) return 0;
    const int status = f();
    if (status == 0) done_ |= bit;
    return status;
  }

 private:
  std::mutex mu_;
  unsigned long long done_ = 0;
};

}  // namespace ca

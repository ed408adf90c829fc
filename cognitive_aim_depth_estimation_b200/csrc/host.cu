// Host-side plumbing: thread-local last-error string, TMA tensor-map encoder (driver entry point
// resolved at run time so the library links against cudart only), device queries.
#include "host.h"

#include <stdio.h>
#include <stdlib.h>

#include <mutex>

namespace ca {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
const char* last_error_cstr() { return g_last_error.c_str(); }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d in `%s`", static_cast<int>(e), cudaGetErrorString(e), file,
           line, what);
  set_error(buf);
  return 2;  // CA_ERR_CUDA
}

int invalid(const char* what) {
  set_error(std::string("invalid argument: ") + what);
  return 1;  // CA_ERR_INVALID
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int make_tmap_sw128(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box,
                           CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B);

int make_tmap_bf16_sw128(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box) {
  return make_tmap_sw128(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}

int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows};
  uint64_t strides[1] = {ld * 4};
  uint32_t box[2] = {32, box_rows};
  return make_tmap_sw128(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, 2, dims, strides, box);
}

static int make_tmap_sw128(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver / GPU?)");
    return 2;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 0xf) != 0) return invalid("TMA base address must be 16-byte aligned");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      if (strides_bytes[i - 1] % 16 != 0) return invalid("TMA global strides must be multiples of 16 bytes");
      gstr[i - 1] = strides_bytes[i - 1];
    }
  }
  CUresult r = fn(out, dtype, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                  gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu)",
             static_cast<int>(r), rank, static_cast<unsigned long long>(dims[0]),
             static_cast<unsigned long long>(rank > 1 ? dims[1] : 1));
    set_error(buf);
    return 2;
  }
  return 0;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows};
  uint64_t strides[1] = {ld * 2};
  uint32_t box[2] = {64, box_rows};
  return make_tmap_bf16_sw128(out, base, 2, dims, strides, box);
}

int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t batch, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint64_t batch_stride, uint32_t box_rows) {
  uint64_t dims[3] = {cols, rows, batch};
  uint64_t strides[2] = {ld * 2, batch_stride * 2};
  uint32_t box[3] = {64, box_rows, 1};
  return make_tmap_bf16_sw128(out, base, 3, dims, strides, box);
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("CA_PDL");
    return e != nullptr && atoi(e) != 0;
  }();
  return on;
}

int current_device() {
  int dev = -1;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

int sm_count() {
  static std::mutex mu;
  static int cached[64] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lock(mu);
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace ca

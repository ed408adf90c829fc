// JPEG ingest for the demo path: reference demo.py:312 `Image.open(path).convert('RGB')` (libjpeg-turbo inside Pillow)
// -> uint8 RGB HWC on the DEVICE, so that decode -> ca_resize_u8 -> ca_preprocess_u8 -> backbone never touches the host
// with pixels.  The decode itself is nvJPEG (CUDA toolkit library code, like cuBLAS would be for a plain GEMM): Huffman
// on the host, IDCT / upsampling / colour conversion on the GPU, on the caller's stream.  The library is loaded lazily
// with dlopen so that libcogaim_b200.so has no link-time dependency on it; when it is absent the entry points fail with
// CA_STATUS_UNSUPPORTED and a message — nothing falls back to a CPU decoder.
#include "jpeg.cuh"

#include <dlfcn.h>
#include <nvjpeg.h>

#include <mutex>
#include <vector>

#include "host.h"

namespace ca {
namespace {

struct NvJpeg {
  void* lib = nullptr;
  decltype(&nvjpegCreateEx) create = nullptr;
  decltype(&nvjpegJpegStateCreate) state_create = nullptr;
  decltype(&nvjpegGetImageInfo) info = nullptr;
  decltype(&nvjpegDecode) decode = nullptr;
  decltype(&nvjpegDecodeBatchedInitialize) batched_init = nullptr;
  decltype(&nvjpegDecodeBatched) batched = nullptr;
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state[64] = {};  // one decoder state per device (calls on one device are serialised by the caller)
  nvjpegJpegState_t bstate[64] = {};  // ... and one state for the batched decoder, initialised for bstate_n[dev] images
  int bstate_n[64] = {};
  std::mutex mu;
  bool ok = false;
  std::string why;
};

NvJpeg& nvjpeg() {
  static NvJpeg nj;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
      nj.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (nj.lib) break;
    }
    if (!nj.lib) {
      nj.why = "nvJPEG (libnvjpeg.so.12) not found: JPEG ingest is unavailable (there is no CPU decode fallback)";
      return;
    }
    nj.create = reinterpret_cast<decltype(nj.create)>(dlsym(nj.lib, "nvjpegCreateEx"));
    nj.state_create = reinterpret_cast<decltype(nj.state_create)>(dlsym(nj.lib, "nvjpegJpegStateCreate"));
    nj.info = reinterpret_cast<decltype(nj.info)>(dlsym(nj.lib, "nvjpegGetImageInfo"));
    nj.decode = reinterpret_cast<decltype(nj.decode)>(dlsym(nj.lib, "nvjpegDecode"));
    nj.batched_init = reinterpret_cast<decltype(nj.batched_init)>(dlsym(nj.lib, "nvjpegDecodeBatchedInitialize"));
    nj.batched = reinterpret_cast<decltype(nj.batched)>(dlsym(nj.lib, "nvjpegDecodeBatched"));
    if (!nj.create || !nj.state_create || !nj.info || !nj.decode) {
      nj.why = "nvJPEG symbols missing";
      return;
    }
    // Chroma planes of 4:2:0 / 4:2:2 streams are up-sampled WITH interpolation, which is what libjpeg-turbo's default
    // "fancy up-sampling" (the decoder inside Pillow) does; nvJPEG's default replicates chroma samples instead.
    const nvjpegStatus_t st = nj.create(NVJPEG_BACKEND_DEFAULT, nullptr, nullptr, NVJPEG_FLAGS_UPSAMPLING_WITH_INTERPOLATION,
                                        &nj.handle);
    if (st != NVJPEG_STATUS_SUCCESS) {
      nj.why = "nvjpegCreateEx failed with status " + std::to_string(static_cast<int>(st));
      return;
    }
    nj.ok = true;
  });
  return nj;
}

int unsupported(const std::string& why) {
  set_error(why);
  return 3;  // CA_STATUS_UNSUPPORTED
}

}  // namespace

int jpeg_info(const uint8_t* h_data, size_t len, int* width, int* height, int* components) {
  CA_REQUIRE(h_data && len > 0 && width && height, "jpeg_info: null argument");
  NvJpeg& nj = nvjpeg();
  if (!nj.ok) return unsupported(nj.why);
  int n = 0, ws[NVJPEG_MAX_COMPONENT] = {}, hs[NVJPEG_MAX_COMPONENT] = {};
  nvjpegChromaSubsampling_t sub;
  const nvjpegStatus_t st = nj.info(nj.handle, h_data, len, &n, &sub, ws, hs);
  if (st != NVJPEG_STATUS_SUCCESS) {
    set_error("jpeg_info: not a decodable JPEG stream (nvjpeg status " + std::to_string(static_cast<int>(st)) + ")");
    return 1;
  }
  *width = ws[0];
  *height = hs[0];
  if (components) *components = n;
  return 0;
}

int jpeg_decode(const uint8_t* h_data, size_t len, uint8_t* out_rgb, int width, int height, cudaStream_t stream) {
  CA_REQUIRE(h_data && len > 0 && out_rgb, "jpeg_decode: null argument");
  int w = 0, h = 0;
  CA_TRY(jpeg_info(h_data, len, &w, &h, nullptr));
  CA_REQUIRE(w == width && h == height, "jpeg_decode: output size does not match the stream (call ca_jpeg_info first)");
  NvJpeg& nj = nvjpeg();
  int dev = 0;
  CA_CUDA(cudaGetDevice(&dev));
  CA_REQUIRE(dev >= 0 && dev < 64, "jpeg_decode: device index out of range");
  if (!nj.state[dev]) {
    const nvjpegStatus_t st = nj.state_create(nj.handle, &nj.state[dev]);
    if (st != NVJPEG_STATUS_SUCCESS) return unsupported("nvjpegJpegStateCreate failed");
  }
  nvjpegImage_t dst = {};
  dst.channel[0] = out_rgb;  // NVJPEG_OUTPUT_RGBI: interleaved RGB in channel 0 (grayscale streams are expanded,
  dst.pitch[0] = static_cast<size_t>(width) * 3;  // which is what PIL's .convert('RGB') does)
  const nvjpegStatus_t st = nj.decode(nj.handle, nj.state[dev], h_data, len, NVJPEG_OUTPUT_RGBI, &dst, stream);
  if (st != NVJPEG_STATUS_SUCCESS) {
    set_error("jpeg_decode: nvjpegDecode failed with status " + std::to_string(static_cast<int>(st)));
    return 2;
  }
  return 0;
}

// demo.py:406-432 (`predict_batch`) opens its files one by one; here n JPEG streams are handed to nvJPEG's batched
// decoder in ONE call (host-side Huffman decoding of the whole batch on a thread pool, one set of GPU kernels for all
// images) instead of n nvjpegDecode calls from a Python loop.
int jpeg_decode_batch(const uint8_t* const* h_data, const size_t* lens, int n, uint8_t* const* outs, const int* widths,
                      const int* heights, cudaStream_t stream) {
  CA_REQUIRE(h_data && lens && outs && widths && heights && n > 0, "jpeg_decode_batch: null argument");
  NvJpeg& nj = nvjpeg();
  if (!nj.ok) return unsupported(nj.why);
  if (!nj.batched_init || !nj.batched) return unsupported("nvJPEG batched decode entry points missing");
  for (int i = 0; i < n; ++i) {
    int w = 0, h = 0;
    CA_REQUIRE(h_data[i] && lens[i] > 0 && outs[i], "jpeg_decode_batch: null image");
    CA_TRY(jpeg_info(h_data[i], lens[i], &w, &h, nullptr));
    CA_REQUIRE(w == widths[i] && h == heights[i], "jpeg_decode_batch: output size does not match the stream");
  }
  int dev = 0;
  CA_CUDA(cudaGetDevice(&dev));
  CA_REQUIRE(dev >= 0 && dev < 64, "jpeg_decode_batch: device index out of range");
  std::lock_guard<std::mutex> lock(nj.mu);
  if (!nj.bstate[dev]) {
    if (nj.state_create(nj.handle, &nj.bstate[dev]) != NVJPEG_STATUS_SUCCESS) return unsupported("nvjpegJpegStateCreate failed");
  }
  if (nj.bstate_n[dev] != n) {
    const nvjpegStatus_t st = nj.batched_init(nj.handle, nj.bstate[dev], n, 8, NVJPEG_OUTPUT_RGBI);
    if (st != NVJPEG_STATUS_SUCCESS) {
      set_error("jpeg_decode_batch: nvjpegDecodeBatchedInitialize failed with status " + std::to_string(static_cast<int>(st)));
      return 2;
    }
    nj.bstate_n[dev] = n;
  }
  std::vector<nvjpegImage_t> dst(n);
  for (int i = 0; i < n; ++i) {
    dst[i] = nvjpegImage_t{};
    dst[i].channel[0] = outs[i];
    dst[i].pitch[0] = static_cast<size_t>(widths[i]) * 3;
  }
  const nvjpegStatus_t st = nj.batched(nj.handle, nj.bstate[dev], h_data, lens, dst.data(), stream);
  if (st != NVJPEG_STATUS_SUCCESS) {
    set_error("jpeg_decode_batch: nvjpegDecodeBatched failed with status " + std::to_string(static_cast<int>(st)));
    return 2;
  }
  return 0;
}

}  // namespace ca

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

#include "host.h"

namespace ca {
namespace {

struct NvJpeg {
  void* lib = nullptr;
  decltype(&nvjpegCreateEx) create = nullptr;
  decltype(&nvjpegJpegStateCreate) state_create = nullptr;
  decltype(&nvjpegGetImageInfo) info = nullptr;
  decltype(&nvjpegDecode) decode = nullptr;
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state[64] = {};  // one decoder state per device (calls on one device are serialised by the caller)
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

}  // namespace ca

// CuriosityModule.forward (reference src/model.py:586-688, called with the CLS token and exif_data=None at
// :1185, :1104, :1138): variational encoder / decoder, uncertainty head, local-sensitivity probe, weighted reward, and
// the exploration ring buffer (:760-773).  Output-dead under the effective configuration (curiosity_guided=False), but
// it is observable state: `curiosity_module.exploration_history` / `history_pointer` are state_dict buffers, and with
// curiosity_guided=True the reward modulates the focal attention.  The two Gaussian draws (:609 eps[B,192], :744
// noise[B,768]) come from the caller, who takes them from the CPU generator exactly where the reference does.
// One CTA per image, fp32; a second single-thread kernel replays the reference's sequential ring-buffer writes
// (the reference does B host round-trips with .item() here).
// curiosity_modulation: the curiosity-guided attention weights of IterativeFocalStream / FocalStream
// (src/model.py:333-339 amplifier, :73-79 modulator, :406-417, :266-269), one thread per image.
#include "common.cuh"
#include "curiosity.cuh"
#include "dense.cuh"
#include "host.h"

namespace ca {
namespace {

__device__ __forceinline__ float block_sum_1024(float v, float* red) {
  v = warp_sum(v);
  const int w = warp_id(), l = lane_id(), nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < nw) ? red[l] : 0.f;
  return warp_sum(t);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(kHeadsThreads) curiosity_kernel(const CuriosityWeights wt, const float* __restrict__ tokens,
                                                                  int tokens_per_img, const float* __restrict__ eps,
                                                                  const float* __restrict__ noise,
                                                                  float* __restrict__ reward_raw,
                                                                  float* __restrict__ reward) {
  griddep_sync();  // PDL: nothing before this line reads or writes global memory
  __shared__ __align__(16) float s_cls[768];
  __shared__ __align__(16) float s_h[768];
  __shared__ __align__(16) float s_mu[192];
  __shared__ __align__(16) float s_lv[192];
  __shared__ __align__(16) float s_z[192];
  __shared__ __align__(16) float s_o[192];
  __shared__ float red[32];
  __shared__ float s_scalar[4];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* cls = tokens + static_cast<size_t>(b) * tokens_per_img * 768;
  for (int i = tid; i < 768; i += kHeadsThreads) s_cls[i] = cls[i];
  __syncthreads();
  // 1-2. variational encoder, reparameterisation with the caller's eps
  dense(wt.em_w0, wt.em_b0, s_cls, s_h, 768, 384, true);
  dense(wt.em_w1, wt.em_b1, s_h, s_mu, 384, 192, false);
  dense(wt.el_w0, wt.el_b0, s_cls, s_h, 768, 384, true);
  dense(wt.el_w1, wt.el_b1, s_h, s_lv, 384, 192, false);
  float kl_part = 0.f;
  for (int i = tid; i < 192; i += kHeadsThreads) {
    const float mu = s_mu[i], lv = s_lv[i];
    s_z[i] = mu + eps[static_cast<size_t>(b) * 192 + i] * expf(0.5f * lv);
    kl_part += 1.0f + lv - mu * mu - expf(lv);
  }
  const float kl = -0.5f * block_sum_1024(kl_part, red);
  __syncthreads();
  // 3-4. decoder and robust reconstruction error against the first 192 feature dimensions
  dense(wt.dec_w0, wt.dec_b0, s_z, s_h, 192, 384, true);
  dense(wt.dec_w1, wt.dec_b1, s_h, s_o, 384, 192, false);
  float err_part = 0.f;
  for (int i = tid; i < 192; i += kHeadsThreads) {
    const float d = s_o[i] - s_cls[i];
    err_part += d * d;
  }
  float err = sqrtf(block_sum_1024(err_part, red) + 1e-8f);
  err = err / (1.0f + err);
  // 6. auxiliary uncertainty head (Softplus, beta 1, threshold 20)
  dense(wt.unc_w0, wt.unc_b0, s_cls, s_h, 768, 192, true);
  dense(wt.unc_w1, wt.unc_b1, s_h, s_o, 192, 1, false);
  if (tid == 0) {
    const float zz = s_o[0];
    s_scalar[0] = zz > 20.0f ? zz : log1pf(expf(zz));
  }
  __syncthreads();
  const float unc = s_scalar[0];
  const float basic = fmaxf(err, 0.f) + 0.1f * fmaxf(kl, 0.f) + 0.1f * fminf(fmaxf(unc, 0.f), 10.0f);
  float out = basic;
  if (wt.loc_w0 != nullptr) {  // hierarchical curiosity (always on under the effective configuration)
    // local exploration: sensitivity of the local head to the caller's 0.01-scaled Gaussian perturbation
    dense(wt.loc_w0, wt.loc_b0, s_cls, s_h, 768, 128, true);
    dense(wt.loc_w1, wt.loc_b1, s_h, s_o, 128, 1, false);
    if (tid == 0) s_scalar[1] = sigmoidf_(s_o[0]);
    __syncthreads();
    for (int i = tid; i < 768; i += kHeadsThreads) s_cls[i] += noise[static_cast<size_t>(b) * 768 + i] * 0.01f;
    __syncthreads();
    dense(wt.loc_w0, wt.loc_b0, s_cls, s_h, 768, 128, true);
    dense(wt.loc_w1, wt.loc_b1, s_h, s_o, 128, 1, false);
    if (tid == 0) {
      const float base = s_scalar[1], nz = sigmoidf_(s_o[0]);
      const float local = fminf(fmaxf(base + fabsf(base - nz) * 0.2f, 0.f), 1.0f);
      const float geo = 0.5f;  // exif_data is None at every call site (src/model.py:697-700)
      const float c0 = wt.cur_w[0], c1 = wt.cur_w[1], c2 = wt.cur_w[2];
      const float m = fmaxf(c0, fmaxf(c1, c2));
      const float e0 = expf(c0 - m), e1 = expf(c1 - m), e2 = expf(c2 - m);
      const float inv = 1.0f / (e0 + e1 + e2);
      s_scalar[2] = (e0 * geo + e1 * local + e2 * basic) * inv;
    }
    __syncthreads();
    out = s_scalar[2];
  }
  if (tid == 0) {
    reward_raw[b] = out;                            // what the ring buffer records (:685)
    reward[b] = fminf(fmaxf(out, 0.f), 100.0f);     // what the caller gets (:691)
  }
}

// The reference walks the batch in order: history[ptr] = reward; ptr = (ptr + 1) % len   (src/model.py:770-773)
__global__ void history_update_kernel(const float* __restrict__ reward_raw, int B, float* __restrict__ history, int len,
                                      long long* __restrict__ pointer) {
  griddep_sync();  // PDL: nothing before this line reads or writes global memory
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  long long p = *pointer;
  for (int b = 0; b < B; ++b) {
    const int idx = static_cast<int>(((p % len) + len) % len);
    history[idx] = reward_raw[b];
    p = (p + 1) % len;
  }
  *pointer = p;
}

// One thread per image: iteration weights = softmax(amplifier(score)), then per iteration the head-mean of
// sigmoid(modulator_i(score * weight_i)).  Amplifier hidden width is 32 (src/model.py:335), 8 modulator outputs (:77).
__global__ void curiosity_modulation_kernel(const CuriosityModWeights wt, const float* __restrict__ reward, float lo,
                                            float hi, float* __restrict__ cur_weight, int B, int n_iters,
                                            int mod_hidden) {
  griddep_sync();  // PDL: nothing before this line reads or writes global memory
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = reward[b];
  if (lo <= hi) s = fminf(fmaxf(s, lo), hi);
  float logit[8];
  for (int i = 0; i < n_iters; ++i) logit[i] = wt.amp_b1[i];
  for (int h = 0; h < 32; ++h) {
    const float a = fmaxf(fmaf(wt.amp_w0[h], s, wt.amp_b0[h]), 0.f);
    for (int i = 0; i < n_iters; ++i) logit[i] = fmaf(wt.amp_w1[i * 32 + h], a, logit[i]);
  }
  float m = logit[0];
  for (int i = 1; i < n_iters; ++i) m = fmaxf(m, logit[i]);
  float den = 0.f;
  for (int i = 0; i < n_iters; ++i) {
    logit[i] = expf(logit[i] - m);
    den += logit[i];
  }
  for (int i = 0; i < n_iters; ++i) {
    const float si = s * (logit[i] / den);  // :412 iter_curiosity = curiosity_score * iteration_weights[:, i]
    float o[8];
    for (int k = 0; k < 8; ++k) o[k] = wt.mod_b1[i][k];
    for (int h = 0; h < mod_hidden; ++h) {
      const float a = fmaxf(fmaf(wt.mod_w0[i][h], si, wt.mod_b0[i][h]), 0.f);
      for (int k = 0; k < 8; ++k) o[k] = fmaf(wt.mod_w1[i][k * mod_hidden + h], a, o[k]);
    }
    float mean = 0.f;
    for (int k = 0; k < 8; ++k) mean += sigmoidf_(o[k]);
    cur_weight[static_cast<size_t>(i) * B + b] = mean * 0.125f;  // :269 mean over the 8 "heads"
  }
}

}  // namespace

int curiosity_modulation_launch(const CuriosityModWeights& w, const float* reward, float lo, float hi, float* cur_weight,
                                int B, int n_iters, int mod_hidden, cudaStream_t stream) {
  CA_REQUIRE(reward && cur_weight, "curiosity_modulation: null pointer");
  CA_REQUIRE(n_iters >= 1 && n_iters <= 8, "curiosity_modulation: 1..8 iterations");
  CA_REQUIRE(mod_hidden >= 1 && B > 0, "curiosity_modulation: bad sizes");
  CA_REQUIRE(w.amp_w0 && w.amp_b0 && w.amp_w1 && w.amp_b1, "curiosity_modulation: null amplifier weights");
  for (int i = 0; i < n_iters; ++i)
    CA_REQUIRE(w.mod_w0[i] && w.mod_b0[i] && w.mod_w1[i] && w.mod_b1[i], "curiosity_modulation: null modulator weights");
  CA_TRY(launch_kernel(curiosity_modulation_kernel, dim3((B + 127) / 128), dim3(128), 0, stream, w, reward, lo, hi, cur_weight, B, n_iters, mod_hidden));
  CA_CUDA(cudaGetLastError());
  return 0;
}

int curiosity_launch(const CuriosityWeights& w, const float* tokens, int tokens_per_img, const float* eps,
                     const float* noise, float* reward_raw, float* reward, float* history, int history_len,
                     long long* pointer, int B, cudaStream_t stream) {
  CA_REQUIRE(tokens && eps && reward_raw && reward, "curiosity: null pointer");
  CA_REQUIRE(w.loc_w0 == nullptr || noise != nullptr, "curiosity: the hierarchical path needs the noise draw");
  CA_REQUIRE(B > 0, "curiosity: empty batch");
  CA_TRY(launch_kernel(curiosity_kernel, dim3(B), dim3(kHeadsThreads), 0, stream, w, tokens, tokens_per_img, eps, noise, reward_raw, reward));
  CA_CUDA(cudaGetLastError());
  if (history != nullptr && w.loc_w0 != nullptr) {  // the reference only records in the hierarchical branch (:683-685)
    CA_REQUIRE(pointer != nullptr && history_len > 0, "curiosity: history without pointer");
    CA_TRY(launch_kernel(history_update_kernel, dim3(1), dim3(32), 0, stream, reward_raw, B, history, history_len, pointer));
    CA_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace ca

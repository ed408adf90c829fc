/* cogaim_b200.h — C-ABI of the B200-native Cognitive-Aim forward path.
 *
 * The reference (yenjane-dot/cognitive-aim-depth-estimation) has no FFI: its boundary is the Python
 * nn.Module surface of `CognitiveAimModel` (reference src/model.py:1064 `forward`, :1157
 * `forward_with_guidance`, :1534 `create_model`).  The Python host module in
 * `cognitive_aim_depth_estimation_b200/model.py` mirrors that surface and calls the entry points below through
 * ctypes.  Every entry point cites the reference call site whose arithmetic it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless named `h_*`;
 *   - every function returns 0 on success, non-zero (ca_status) on failure and never throws;
 *     `ca_last_error()` returns a thread-local description of the last failure;
 *   - the caller supplies the CUDA stream (as a `void*` holding a `cudaStream_t`), all outputs and
 *     all workspaces; nothing is allocated behind the caller's back on the hot path;
 *   - bf16 buffers are raw 16-bit storage (`uint16_t`), row-major.
 *   - there is NO CPU fallback: without an sm_100 device every compute entry point fails.
 */
#ifndef COGAIM_B200_H_
#define COGAIM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ca_status {
  CA_STATUS_OK = 0,
  CA_STATUS_INVALID = 1,    /* bad shape / alignment / null pointer */
  CA_STATUS_CUDA = 2,       /* CUDA runtime or driver failure */
  CA_STATUS_UNSUPPORTED = 3 /* no sm_100 device */
} ca_status;

/* GEMM epilogues (see ca_gemm_bf16). */
typedef enum ca_epilogue {
  CA_EPI_BIAS_BF16 = 0, /* out_bf16 = acc + bias                      HF modeling_dinov2.py:199-201 (q/k/v Linear); src/model.py:192-193 */
  CA_EPI_GELU_BF16 = 1, /* out_bf16 = gelu_erf(acc + bias)            HF modeling_dinov2.py:324-326 (fc1 + GELU) */
  CA_EPI_RESID_F32 = 2, /* x_f32 += ls * (acc + bias), in place       HF modeling_dinov2.py:246,278,376-385 (dense + LayerScale + residual) */
  CA_EPI_PATCH_F32 = 3, /* x_f32[b*T+1+p] = acc + bias + pos[1+p]     HF modeling_dinov2.py:148,112 (patch conv + pos-embed) */
  CA_EPI_ROWSTATS = 4,  /* softmax row statistics of acc*scale        src/model.py:197-200 (pass A, no N x N materialisation) */
  CA_EPI_COLSUM = 5,    /* column sums of softmax(acc*scale)          src/model.py:234 (pass B, transposed) */
  CA_EPI_F32 = 6        /* out_f32 = acc (test hook) */
} ca_epilogue;

/* Library / device introspection -------------------------------------------------------------- */
const char* ca_last_error(void);
int ca_version(void);                       /* ABI version, currently 1 */
int ca_device_check(int device);            /* CA_STATUS_OK iff `device` is compute capability 10.x */

/* Dense contraction on tcgen05 tensor cores --------------------------------------------------- */
/* C[b] = epilogue(A[b] (M x K, lda) * W[b] (N x K, ldw)^T), bf16 operands, fp32 TMEM accumulators.
 * w_batch_stride == 0 shares W across the batch.  Unused epilogue operands may be NULL.
 * For CA_EPI_ROWSTATS/COLSUM the per-row partial buffers have P = 2*ceil(N/128) entries per row. */
int ca_gemm_bf16(const uint16_t* A, const uint16_t* W, int M, int N, int K, int lda, int ldw, int batch,
                 long long a_batch_stride, long long w_batch_stride, int epilogue, void* out, int ldo,
                 long long out_batch_stride, const float* bias, const float* ls, const float* pos,
                 int patches_per_img, float scale_log2, float* part_a, float* part_b, const float* col_max,
                 const float* col_rinv, void* stream);

/* Backbone attention --------------------------------------------------------------------------- */
/* out[B*T, H*64] = softmax(Q K^T / 8) V per (image, head); qkv is the fused [B*T, 3*H*64] activation
 * (Q | K | V column blocks).  Replaces F.scaled_dot_product_attention at HF modeling_dinov2.py:215-229. */
int ca_attention_bf16(const uint16_t* qkv, uint16_t* out, int B, int T, int H, void* stream);

/* Bandwidth-bound row kernels ------------------------------------------------------------------- */
/* fp32 CHW images [B,3,S,S] (the tensor `forward` receives) -> bf16 patch rows [B*(S/14)^2, 592]
 * (column k = c*196 + ky*14 + kx, columns 588..591 zero).  HF modeling_dinov2.py:139-148 (im2col of the conv). */
int ca_patchify_f32(const float* images, uint16_t* patches, int B, int S, void* stream);
/* uint8 HWC images [B,S,S,3] -> /255 -> (x-mean)/std -> bf16 patch rows.  demo.py:162-166 (ToTensor + Normalize;
 * the source is already S x S so Resize is the identity).  h_mean3 / h_std3 are HOST pointers to 3 floats. */
int ca_preprocess_u8(const uint8_t* images, uint16_t* patches, int B, int S, const float* h_mean3,
                     const float* h_std3, void* stream);
/* x[b, 0, :] = cls + pos[0]  (HF modeling_dinov2.py:108-112). x is the fp32 residual stream [B, T, D]. */
int ca_cls_rows(float* x, const float* cls, const float* pos, int B, int T, int D, void* stream);
/* LayerNorm(D=768) of fp32 rows -> bf16 (out_is_bf16=1) or fp32.  HF modeling_dinov2.py:354,359,449. */
int ca_layernorm(const float* x, const float* gamma, const float* beta, void* out, int out_is_bf16, int rows, int D,
                 float eps, void* stream);
/* xin[b,n,:] = bf16(tokens[b,1+n,:] * rowscale[b,n] + pe[n,:]); rowscale may be NULL (=1).
 * reference src/model.py:184 (PE add) and :426 (re-focus, folded into a per-row scale). */
int ca_focal_input(const float* tokens, const float* pe, const float* rowscale, uint16_t* xin, int B, int N, int D,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COGAIM_B200_H_ */

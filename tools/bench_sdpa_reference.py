"""Reference point for the attention kernel: what do the LIBRARY kernels in this image reach at the same shape
(B = 32, H = 12, T = 1370, head_dim 64, bf16, no mask)?  torch SDPA with the cuDNN / flash / mem-efficient backends and
flash_attn 2.8 — library code, not part of the product; printed next to csrc/attention.cu for context."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, '/root/repo')
from cognitive_aim_depth_estimation_b200 import ops

B, T, H, D = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (32, 1370, 12, 64)
dev = 'cuda'
flops = 4.0 * B * H * T * T * D


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


qkv = torch.randn(B * T, 3 * H * D, device=dev).bfloat16()
out = torch.empty(B * T, H * D, device=dev, dtype=torch.bfloat16)
ms = timed(lambda: ops.attention(qkv, out, B, T, H))
print(f"csrc/attention.cu (this repo)          : {ms:.4f} ms  {flops / ms / 1e9:7.1f} TF/s")
q, k, v = qkv.view(B, T, 3, H, D).permute(2, 0, 3, 1, 4).contiguous()  # [3][B,H,T,D]
from torch.nn.attention import SDPBackend, sdpa_kernel
for name, be in (("cuDNN", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION),
                 ("mem-efficient", SDPBackend.EFFICIENT_ATTENTION)):
    try:
        with sdpa_kernel(be):
            ms = timed(lambda: F.scaled_dot_product_attention(q, k, v))
        print(f"torch SDPA, {name:13s} backend        : {ms:.4f} ms  {flops / ms / 1e9:7.1f} TF/s")
    except Exception as e:  # noqa: BLE001
        print(f"torch SDPA, {name} backend: unavailable ({str(e)[:80]})")
try:
    from flash_attn import flash_attn_func
    qf, kf, vf = (t.transpose(1, 2).contiguous() for t in (q, k, v))  # [B,T,H,D]
    ms = timed(lambda: flash_attn_func(qf, kf, vf))
    print(f"flash_attn 2.8 flash_attn_func         : {ms:.4f} ms  {flops / ms / 1e9:7.1f} TF/s")
except Exception as e:  # noqa: BLE001
    print(f"flash_attn: unavailable ({str(e)[:80]})")
try:
    import flashinfer
    qi = q.permute(0, 2, 1, 3).reshape(B * T, H, D).contiguous()
    ki = k.permute(0, 2, 1, 3).reshape(B * T, H, D).contiguous()
    vi = v.permute(0, 2, 1, 3).reshape(B * T, H, D).contiguous()
    indptr = torch.arange(0, B + 1, device=dev, dtype=torch.int32) * T
    for backend in ("cutlass", "fa2"):
        try:
            wsb = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
            w = flashinfer.BatchPrefillWithRaggedKVCacheWrapper(wsb, "NHD", backend=backend)
            w.plan(indptr, indptr, H, H, D, causal=False, q_data_type=torch.bfloat16)
            ms = timed(lambda: w.run(qi, ki, vi))
            print(f"flashinfer ragged prefill, {backend:8s}    : {ms:.4f} ms  {flops / ms / 1e9:7.1f} TF/s")
        except Exception as e:  # noqa: BLE001
            print(f"flashinfer {backend}: unavailable ({str(e)[:100]})")
except Exception as e:  # noqa: BLE001
    print(f"flashinfer: unavailable ({str(e)[:80]})")

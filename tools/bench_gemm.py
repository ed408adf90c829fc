"""Device-time of the backbone GEMM shapes (B=32, 518x518) — tools/bench_gemm.py [CA_GEMM_DEBUG=n]"""
import sys
import torch
sys.path.insert(0, '/root/repo')
from cognitive_aim_depth_estimation_b200 import ops, _lib
import os, pathlib
if os.environ.get("CA_LIB_OVERRIDE"):
    _lib.LIB_PATH = pathlib.Path(os.environ["CA_LIB_OVERRIDE"])  # tooling only: same-box A/B of two builds
M = 32 * 1370
dev = 'cuda'
shapes = [("qkv", 2304, 768, ops.EPI_BIAS_BF16), ("proj", 768, 768, ops.EPI_RESID_F32),
          ("fc1", 3072, 768, ops.EPI_GELU_BF16), ("fc1-nogelu", 3072, 768, ops.EPI_BIAS_BF16),
          ("fc2", 768, 3072, ops.EPI_RESID_F32), ("focal-qk", 1536, 768, ops.EPI_BIAS_BF16)]
for name, N, K, epi in shapes:
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    bias = torch.zeros(N, device=dev)
    ls = torch.ones(N, device=dev)
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if epi == ops.EPI_RESID_F32 else torch.bfloat16)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    for _ in range(2):
        ops.gemm(A, W, epi, out, bias=bias, ls=ls)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(A, W, epi, out, bias=bias, ls=ls); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:10s} N={N:5d} K={K:5d}  {t*1e3:8.1f} us  {2*M*N*K/t/1e9:8.1f} TF/s")

"""Per-shape timing of the LayerNorm-folded GEMM epilogues against the plain ones (B = 32 x 518^2: M = 43840 rows).

    python tools/bench_gemm_ln.py            # CA_GEMM_LN_NBUF=2|3 selects the residual lookahead of EPI_RESID_LN_F32
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cognitive_aim_depth_estimation_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M = 32 * 1370
g = torch.Generator(device="cpu").manual_seed(0)
r = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)  # noqa: E731
x = r(M, 768)
h = r(M, 768).bfloat16()
att = r(M, 768).bfloat16()
mlp = r(M, 3072).bfloat16()
qkv = torch.empty(M, 2304, device=dev, dtype=torch.bfloat16)
stats = torch.zeros(M, 6, 2, device=dev)
wqkv, w1, wo, w2 = r(2304, 768, sc=.03).bfloat16(), r(3072, 768, sc=.03).bfloat16(), r(768, 768, sc=.03).bfloat16(), r(768, 3072, sc=.03).bfloat16()
b2304, b3072, b768, ls = r(2304), r(3072), r(768), r(768)
ops.ln_shadow(x, h, stats)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)

cases = {
    "qkv  bias": lambda: ops.gemm(h, wqkv, ops.EPI_BIAS_BF16, qkv, bias=b2304),
    "qkv  ln+bias": lambda: ops.gemm_ln(h, wqkv, ops.EPI_LN_BIAS_BF16, qkv, bias=b2304, stats=stats),
    "fc1  gelu": lambda: ops.gemm(h, w1, ops.EPI_GELU_BF16, mlp, bias=b3072),
    "fc1  ln+gelu": lambda: ops.gemm_ln(h, w1, ops.EPI_LN_GELU_BF16, mlp, bias=b3072, stats=stats),
    "proj resid (reduce-add)": lambda: ops.gemm(att, wo, ops.EPI_RESID_F32, x, bias=b768, ls=ls),
    "proj resid+ln": lambda: ops.gemm_ln(att, wo, ops.EPI_RESID_LN_F32, x, bias=b768, ls=ls, stats=stats, shadow=h),
    "fc2  resid (reduce-add)": lambda: ops.gemm(mlp, w2, ops.EPI_RESID_F32, x, bias=b768, ls=ls),
    "fc2  resid+ln": lambda: ops.gemm_ln(mlp, w2, ops.EPI_RESID_LN_F32, x, bias=b768, ls=ls, stats=stats, shadow=h),
    "layernorm": lambda: ops.layernorm(x, b768, b768, h),
}
print("CA_GEMM_LN_NBUF =", os.environ.get("CA_GEMM_LN_NBUF", "(default: 3 for K<=1024, else 2)"))
for name, fn in cases.items():
    ls.zero_()  # keep x bounded over the repetitions
    for _ in range(3):
        fn()
    ts = []
    for _ in range(12):
        flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"{name:28s} median {ts[len(ts) // 2]:7.1f} us   min {ts[0]:7.1f} us")

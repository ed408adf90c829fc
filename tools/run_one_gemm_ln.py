"""One LayerNorm-folded residual GEMM shape, a few launches: the target of an ncu capture.
    python tools/run_one_gemm_ln.py proj|fc2|qkv|fc1|fc1plain|qkvplain"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cognitive_aim_depth_estimation_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "proj"
dev = torch.device("cuda:0")
M = 32 * 1370
g = torch.Generator(device="cpu").manual_seed(0)
r = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).to(dev)  # noqa: E731
x, h = r(M, 768), r(M, 768).bfloat16()
stats = torch.zeros(M, 6, 2, device=dev)
ops.ln_shadow(x, h, stats)
ls = torch.zeros(768, device=dev)
if which == "proj":
    a, w, b = r(M, 768).bfloat16(), r(768, 768, sc=.03).bfloat16(), r(768)
    fn = lambda: ops.gemm_ln(a, w, ops.EPI_RESID_LN_F32, x, bias=b, ls=ls, stats=stats, shadow=h)  # noqa: E731
elif which == "fc2":
    a, w, b = r(M, 3072).bfloat16(), r(768, 3072, sc=.03).bfloat16(), r(768)
    fn = lambda: ops.gemm_ln(a, w, ops.EPI_RESID_LN_F32, x, bias=b, ls=ls, stats=stats, shadow=h)  # noqa: E731
elif which == "qkv":
    w, b, o = r(2304, 768, sc=.03).bfloat16(), r(2304), torch.empty(M, 2304, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm_ln(h, w, ops.EPI_LN_BIAS_BF16, o, bias=b, stats=stats)  # noqa: E731
elif which == "fc1plain":
    w, b, o = r(3072, 768, sc=.03).bfloat16(), r(3072), torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(h, w, ops.EPI_GELU_BF16, o, bias=b)  # noqa: E731
elif which == "qkvplain":
    w, b, o = r(2304, 768, sc=.03).bfloat16(), r(2304), torch.empty(M, 2304, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(h, w, ops.EPI_BIAS_BF16, o, bias=b)  # noqa: E731
else:
    w, b, o = r(3072, 768, sc=.03).bfloat16(), r(3072), torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm_ln(h, w, ops.EPI_LN_GELU_BF16, o, bias=b, stats=stats)  # noqa: E731
for _ in range(4):
    fn()
torch.cuda.synchronize()
print("ok")

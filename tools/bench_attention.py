"""Device time of the attention kernel alone: python tools/bench_attention.py [B T H]"""
import sys
import torch
sys.path.insert(0, '/root/repo')
from cognitive_aim_depth_estimation_b200 import ops, _lib
import os, pathlib
if os.environ.get('CA_LIB_OVERRIDE'):
    _lib.LIB_PATH = pathlib.Path(os.environ['CA_LIB_OVERRIDE'])  # tooling only: same-box A/B of two builds
cfgs = [(32, 1370, 12), (8, 5477, 12), (64, 257, 12)] if len(sys.argv) < 4 else [tuple(int(a) for a in sys.argv[1:4])]
for B, T, H in cfgs:
    qkv = (torch.randn(B * T, 3 * H * 64, device='cuda')).bfloat16()
    out = torch.empty(B * T, H * 64, device='cuda', dtype=torch.bfloat16)
    for _ in range(3):
        ops.attention(qkv, out, B, T, H)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.attention(qkv, out, B, T, H)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f'attention B={B} T={T} H={H}: {ms:.4f} ms  {4 * B * H * T * T * 64 / ms / 1e9:.1f} TF/s')

"""Device time and achieved GB/s of the bandwidth / latency kernels either side of the backbone (SURVEY.md §8 a1, f1-f3 and
the focal vector stages) at the benchmark shape (32 images of 518 x 518), each launched alone with an L2 flush between
repetitions.  Algorithmic bytes = what the kernel must read + write once.  Used as the ncu target for
profiles/r02_small_kernels.md:   python tools/bench_small_kernels.py [--once]   (--once: one launch each, for ncu)"""
import sys
import torch
sys.path.insert(0, '/root/repo')
from cognitive_aim_depth_estimation_b200 import ops, tables

once = "--once" in sys.argv
dev = torch.device("cuda:0")
B, S = 32, 518
g = S // 14
N, T, D = g * g, g * g + 1, 768
PEAK = 6539.9
flush = torch.zeros(40 * 1024 * 1024, dtype=torch.int32, device=dev)   # 160 MB > the 126 MB L2


def timed(name, fn, bytes_alg, reps=10):
    fn()
    torch.cuda.synchronize()
    if once:
        print(f"{name}: launched")
        return
    tot = 0.0
    for _ in range(reps):
        flush.sum()   # READ 160 MB: evicts the kernel's inputs and leaves clean lines (a write would leave 126 MB of dirty
                      # lines whose write-back then competes with the timed kernel's reads)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / reps
    gbs = bytes_alg / ms / 1e6
    print(f"{name:34s} {ms * 1e3:8.1f} us  {bytes_alg / 1e6:8.1f} MB  {gbs:7.0f} GB/s  {gbs / PEAK:5.2f} of {PEAK:.0f}")


gen = torch.Generator().manual_seed(1)
# a1: Pillow-exact resize 480 x 640 -> 518 x 518 (two passes, uint8 intermediate), then normalise + patchify
src = torch.randint(0, 256, (B, 480, 640, 3), generator=gen, dtype=torch.uint8).to(dev)
timed("resize_u8 480x640 -> 518x518", lambda: ops.resize_u8(src, S, S),
      B * (480 * 640 * 3 + 2 * 480 * S * 3 + S * S * 3))
u8 = torch.randint(0, 256, (B, S, S, 3), generator=gen, dtype=torch.uint8).to(dev)
patches = torch.empty(B * N, ops.PATCH_ROW_STRIDE, device=dev, dtype=torch.bfloat16)
timed("preprocess_u8 (norm + patchify)", lambda: ops.preprocess_u8(u8, patches), u8.numel() + patches.numel() * 2)
xf = torch.randn(B, 3, S, S, device=dev)
timed("patchify_f32", lambda: ops.patchify_f32(xf, patches), xf.numel() * 4 + patches.numel() * 2)
# a3: LayerNorm fp32 -> bf16
x = torch.randn(B * T, D, device=dev)
h = torch.empty(B * T, D, device=dev, dtype=torch.bfloat16)
w, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
timed("layernorm fp32 -> bf16", lambda: ops.layernorm(x, w, b, h), x.numel() * 4 + h.numel() * 2)
# a8: focal input (tokens + PE, row scale -> bf16), column sums of the stored exponentials, weighted pooling
tokens = torch.randn(B, T, D, device=dev)
pe = tables.focal_position_encoding(N, D).to(dev)
xin = torch.empty(B * N, D, device=dev, dtype=torch.bfloat16)
rs = torch.rand(B, N, device=dev) + 0.5
timed("focal_input", lambda: ops.focal_input(tokens, pe, rs, xin, B, N, D), B * N * D * 4 + pe.numel() * 4 + xin.numel() * 2)
P = ops.stats_partials(N)
E = (torch.rand(B, N, 64 * ((N + 63) // 64), device=dev) * 0.9 + 0.05).half()
wtab = torch.rand(B, P, N, device=dev)
pc = torch.empty(B, P, N, device=dev)
timed("colsum_e", lambda: ops.colsum_e(E, wtab, pc, B, N), B * N * N * 2 + wtab.numel() * 4 + pc.numel() * 4)
heat = torch.softmax(torch.randn(B, N, device=dev), -1)
pool = torch.empty(B, 32, D, device=dev)
timed("weighted_pool (32 splits)", lambda: ops.weighted_pool(tokens, T * D, 1, heat, None, pool, B, N, D, 32),
      B * N * D * 4 + heat.numel() * 4 + pool.numel() * 4)
# f2: heat-map post-processing (cube, 70th percentile, min-max, order-1 zoom to the source size)
norm = torch.empty(B, N, device=dev)
zoom = torch.empty(B, 480, 640, device=dev)
timed("focus_map -> 480x640", lambda: ops.focus_map(heat, g, 480, 640, norm, zoom), heat.numel() * 4 + norm.numel() * 4 + zoom.numel() * 4)

"""Does pulling the next batch straight out of pinned host memory (patchify reads it over PCIe) disturb the running step
less than a DMA H2D copy + patchify?  Interleaved A/B/C so that clock drift cancels."""
import sys, time
import torch
sys.path.insert(0, '/root/repo')
from cognitive_aim_depth_estimation_b200 import ops
from cognitive_aim_depth_estimation_b200.model import create_model
from oracle import cogaim_oracle as orc
dev = torch.device('cuda:0')
B, S = 32, 518
CFG = {"model": {"cognitive_modules": ["ambient_stream", "iterative_focal_stream", "exif_prior_database"]}}
model = create_model(CFG, {"num_cameras": 71}, device=dev)
model.load_state_dict(orc.build_state_dict(0))
model.validate_inputs = False
x = orc.synthetic_images(B, S).to(dev)
ex = {k: v.to(dev) for k, v in orc.synthetic_exif(B).items()}
host = orc.synthetic_images(B, S, seed=5).pin_memory()
dst = torch.empty(B, 3, S, S, device=dev)
patches2 = torch.empty(B * 37 * 37, ops.PATCH_ROW_STRIDE, device=dev, dtype=torch.bfloat16)
cs = torch.cuda.Stream()


def step():
    torch.manual_seed(11)
    return model.forward_with_guidance(x, ex, "center", return_attention=True)


def dma():
    dst.copy_(host, non_blocking=True)
    ops.patchify_f32(dst, patches2)


class HostView:  # the pinned tensor seen as a device pointer (UVA): what ops.patchify_f32 needs
    def __init__(self, t):
        self.t = t
        self.shape, self.dtype, self.is_cuda = t.shape, t.dtype, True
    def data_ptr(self):
        return self.t.data_ptr()
    def numel(self):
        return self.t.numel()


def pull():
    ops.patchify_f32(HostView(host), patches2)


def run(side, n=10):
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        if side is not None:
            cs.wait_stream(torch.cuda.current_stream()) if False else None
            with torch.cuda.stream(cs):
                side()
        step()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


# correctness of the pull path
pull(); torch.cuda.synchronize(); a = patches2.clone()
dma(); torch.cuda.synchronize()
print("pull == dma+patchify:", torch.equal(a, patches2))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); pull(); e1.record(); torch.cuda.synchronize()
print(f"zero-copy patchify alone: {e0.elapsed_time(e1):.2f} ms")
for rnd in range(3):
    print(f"round {rnd}: alone {run(None):.3f}  dma+patchify {run(dma):.3f}  zero-copy patchify {run(pull):.3f} ms/step", flush=True)

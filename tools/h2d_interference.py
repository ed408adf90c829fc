"""How much does a concurrent copy slow the forward?  (same step, different traffic on a side stream)"""
import sys, time
import torch
sys.path.insert(0, '/root/repo')
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
host = torch.empty(B, 3, S, S).pin_memory()
host_q = torch.empty(B, 3, S, S // 4).pin_memory()
dst = torch.empty(B, 3, S, S, device=dev)
dst_q = torch.empty(B, 3, S, S // 4, device=dev)
src_d = torch.empty(B, 3, S, S, device=dev)
cs = torch.cuda.Stream()
def step():
    torch.manual_seed(11)
    return model.forward_with_guidance(x, ex, "center", return_attention=True)
def run(name, side):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 12
    for i in range(n):
        if side is not None:
            with torch.cuda.stream(cs):
                side()
        step()
    torch.cuda.synchronize()
    print(f"{name:34s} {(time.perf_counter() - t0) / n * 1e3:7.3f} ms/step")
chunks_h = host.view(16, -1)
chunks_d = dst.view(16, -1)
run("no side traffic", None)
run("H2D 103 MB", lambda: dst.copy_(host, non_blocking=True))
run("H2D 26 MB", lambda: dst_q.copy_(host_q, non_blocking=True))
run("H2D 103 MB in 16 chunks", lambda: [chunks_d[i].copy_(chunks_h[i], non_blocking=True) for i in range(16)])
run("D2H 103 MB", lambda: host.copy_(dst, non_blocking=True))
run("D2D 103 MB (side stream)", lambda: dst.copy_(src_d, non_blocking=True))
run("no side traffic (again)", None)

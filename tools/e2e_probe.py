"""Where does the end-to-end step spend its time?  H2D bandwidth, host time per step, overlap."""
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
host = [orc.synthetic_images(B, S, seed=1234 + i).pin_memory() for i in range(2)]
ex = {k: v.to(dev) for k, v in orc.synthetic_exif(B).items()}
stage = [torch.empty(B, 3, S, S, device=dev) for _ in range(2)]
# (a) H2D bandwidth
torch.cuda.synchronize()
for _ in range(2):
    stage[0].copy_(host[0], non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    stage[0].copy_(host[0], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"H2D 103 MB pinned: {dt*1e3:.2f} ms  {host[0].numel()*4/dt/1e9:.1f} GB/s")
def step(x):
    torch.manual_seed(11)
    return model.forward_with_guidance(x, ex, "center", return_attention=True)
for _ in range(4):
    step(stage[0])
torch.cuda.synchronize()
# (b) host time per step (no sync) vs device time
t0 = time.perf_counter()
for _ in range(10):
    step(stage[0])
t_host = (time.perf_counter() - t0) / 10
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 10
print(f"host enqueue per step {t_host*1e3:.2f} ms ; wall per step {t_all*1e3:.2f} ms")
# (c) with concurrent H2D on a copy stream
cs = torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    with torch.cuda.stream(cs):
        stage[1].copy_(host[1], non_blocking=True)
    step(stage[0])
torch.cuda.synchronize()
print(f"step + concurrent (unsynchronised) H2D: {(time.perf_counter()-t0)/10*1e3:.2f} ms per step")
# (d) host-side pieces
t0 = time.perf_counter()
for _ in range(20):
    torch.manual_seed(11); torch.randn(B, 192); torch.randn(B, 768); l = torch.nn.Linear(768, 64)
print(f"rng replay + Linear init: {(time.perf_counter()-t0)/20*1e3:.3f} ms")
t0 = time.perf_counter()
for _ in range(20):
    w = l.weight.detach().to(dev, non_blocking=True)
torch.cuda.synchronize()
print(f"pageable 196 KB H2D: {(time.perf_counter()-t0)/20*1e3:.3f} ms")
import cProfile, pstats
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step(stage[0])
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr).sort_stats("tottime"); st.print_callers("synchronize"); st.print_callers("format_frame_summary")

"""Device time of the whole guided step (B=32, 518x518) for same-box A/B of two library builds under the power cap:
python tools/ab.py tools/bench_step.py libA.so libB.so"""
import os, pathlib, sys
import torch
sys.path.insert(0, '/root/repo')
from cognitive_aim_depth_estimation_b200 import _lib
if os.environ.get("CA_LIB_OVERRIDE"):
    _lib.LIB_PATH = pathlib.Path(os.environ["CA_LIB_OVERRIDE"])  # tooling only
from cognitive_aim_depth_estimation_b200.model import create_model
from oracle import cogaim_oracle as orc
dev = torch.device('cuda:0')
B, S = 32, 518
CFG = {"model": {"cognitive_modules": ["ambient_stream", "iterative_focal_stream", "exif_prior_database"]}}
model = create_model(CFG, {"num_cameras": 71}, device=dev)
model.load_state_dict(orc.build_state_dict(0))
model.validate_inputs = False
xs = [orc.synthetic_images(B, S, seed=i).to(dev) for i in range(3)]
ex = {k: v.to(dev) for k, v in orc.synthetic_exif(B).items()}
for i in range(6):
    torch.manual_seed(11); model.forward_with_guidance(xs[i % 3], ex, "center")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 40
e0.record()
for i in range(n):
    torch.manual_seed(11); model.forward_with_guidance(xs[i % 3], ex, "center")
e1.record(); torch.cuda.synchronize()
print(f"step {e0.elapsed_time(e1) / n:.3f} ms  {B * n / e0.elapsed_time(e1) * 1e3:.0f} images/s")

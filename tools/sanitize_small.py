"""Small guided + un-guided forward for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import sys
import torch
sys.path.insert(0, '/root/repo')
from cognitive_aim_depth_estimation_b200.model import create_model
from oracle import cogaim_oracle as orc
cfg = {"model": {"cognitive_modules": ["ambient_stream", "iterative_focal_stream", "exif_prior_database"]}}
m = create_model(cfg, {"num_cameras": 71}, device="cuda:0")
m.load_state_dict(orc.build_state_dict(0))
m.use_cuda_graphs = False
S = int(sys.argv[1]) if len(sys.argv) > 1 else 126
x = orc.synthetic_images(2, S).cuda()
ex = {k: v.cuda() for k, v in orc.synthetic_exif(2).items()}
torch.manual_seed(11)
d, c, h = m.forward_with_guidance(x, ex, ["left", "center"], return_attention=True)
d2, c2, h2 = m(x, ex, return_attention=True)
fm = m.focus_map((64, 80))
u8 = torch.randint(0, 256, (2, 90, 120, 3), dtype=torch.uint8)
p = m.preprocess(u8, S)
torch.cuda.synchronize()
print("ok", d.flatten().tolist(), h.argmax(-1).tolist(), float(fm.max()), tuple(p.shape))

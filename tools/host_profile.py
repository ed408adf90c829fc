"""Where do the ~1.4 ms of host time per forward_with_guidance call go?  (single image: the call is host-bound)"""
import cProfile, pstats, sys, time
import torch
sys.path.insert(0, '/root/repo')
from cognitive_aim_depth_estimation_b200.model import create_model
from oracle import cogaim_oracle as orc
dev = torch.device('cuda:0')
B, S = 1, int(sys.argv[1]) if len(sys.argv) > 1 else 518
CFG = {"model": {"cognitive_modules": ["ambient_stream", "iterative_focal_stream", "exif_prior_database"]}}
model = create_model(CFG, {"num_cameras": 71}, device=dev)
model.load_state_dict(orc.build_state_dict(0))
model.validate_inputs = False
x = orc.synthetic_images(B, S).to(dev)
ex = {k: v.to(dev) for k, v in orc.synthetic_exif(B).items()}
for _ in range(5):
    model.forward_with_guidance(x, ex, "center", return_attention=True)
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n):
    model.forward_with_guidance(x, ex, "center", return_attention=True)
t_host = (time.perf_counter() - t0) / n
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / n
print(f"host enqueue {t_host * 1e3:.3f} ms / call, wall {t_all * 1e3:.3f} ms / call")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    model.forward_with_guidance(x, ex, "center", return_attention=True)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)

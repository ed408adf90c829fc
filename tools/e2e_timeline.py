"""Per-step device timeline of bench.py's end-to-end loop: where do the ~1.2 ms / step between `value` and `e2e` go?"""
import sys, time
import torch
sys.path.insert(0, '/root/repo')
from cognitive_aim_depth_estimation_b200.model import create_model
from oracle import cogaim_oracle as orc
dev = torch.device('cuda:0')
B, S, N = 32, 518, 18
CFG = {"model": {"cognitive_modules": ["ambient_stream", "iterative_focal_stream", "exif_prior_database"]}}
model = create_model(CFG, {"num_cameras": 71}, device=dev)
model.load_state_dict(orc.build_state_dict(0))
model.validate_inputs = False
host = [orc.synthetic_images(B, S, seed=1234 + i).pin_memory() for i in range(3)]
dev_imgs = [h.to(dev) for h in host]
ex = {k: v.to(dev) for k, v in orc.synthetic_exif(B).items()}
INS = ["center", "left", "right", "top", "bottom", "top-left", "top-right", "bottom-left", "bottom-right"]
copy_stream = torch.cuda.Stream(device=dev)
stage = [torch.empty_like(dev_imgs[0]) for _ in range(2)]
out_host = [torch.empty(B, 1).pin_memory(), torch.empty(B, 1).pin_memory(), torch.empty(B, 1369).pin_memory()]


def step(i, imgs):
    torch.manual_seed(11)
    return model.forward_with_guidance(imgs, ex, INS[i % 9], return_attention=True)


def e2e_loop(n, d2h=True, h2d=True, trace=None):
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    with torch.cuda.stream(copy_stream):
        if h2d:
            stage[0].copy_(host[0], non_blocking=True)
        ready[0].record()
    for i in range(n):
        cur = i % 2
        torch.cuda.current_stream().wait_event(ready[cur])
        if i + 1 < n:
            with torch.cuda.stream(copy_stream):
                if i >= 1:
                    copy_stream.wait_event(done[1 - cur])
                if h2d:
                    stage[1 - cur].copy_(host[(i + 1) % 3], non_blocking=True)
                ready[1 - cur].record()
        if trace is not None:
            e = torch.cuda.Event(enable_timing=True); e.record(); trace.append(e)
        d, c, h = step(i, stage[cur])
        done[cur].record()
        if d2h:
            out_host[0].copy_(d, non_blocking=True)
            out_host[1].copy_(c, non_blocking=True)
            out_host[2].copy_(h, non_blocking=True)
    if trace is not None:
        e = torch.cuda.Event(enable_timing=True); e.record(); trace.append(e)
    torch.cuda.synchronize()


def wall(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / N


for i in range(4):
    step(i, dev_imgs[i % 3])
e2e_loop(4)
t0 = time.perf_counter(); stage[0].copy_(host[0], non_blocking=True); torch.cuda.synchronize()
print(f"H2D 103 MB: {(time.perf_counter() - t0) * 1e3:.2f} ms")
for rnd in range(2):
    a = wall(lambda: [step(i, dev_imgs[i % 3]) for i in range(N)])
    b = wall(lambda: e2e_loop(N))
    c = wall(lambda: e2e_loop(N, d2h=False))
    d = wall(lambda: e2e_loop(N, h2d=False))
    e = wall(lambda: e2e_loop(N, d2h=False, h2d=False))
    print(f"round {rnd}: resident {a:.3f}  e2e {b:.3f}  e2e-no-d2h {c:.3f}  e2e-no-h2d {d:.3f}  loop-only {e:.3f} ms/step", flush=True)
tr = []
e2e_loop(N, trace=tr)
print("per-step device ms:", " ".join(f"{tr[i].elapsed_time(tr[i + 1]):.2f}" for i in range(N)))
t0 = time.perf_counter()
for i in range(N):
    step(i, dev_imgs[i % 3])
t_host = (time.perf_counter() - t0) * 1e3 / N
torch.cuda.synchronize()
print(f"host enqueue per step: {t_host:.3f} ms")

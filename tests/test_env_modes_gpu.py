"""Process-wide switches of libcogaim_b200.so that are read once per process, exercised in a child process:
CA_PDL=1 (programmatic dependent launch on every kernel, csrc/host.h launch_kernel) must give the same results as the
default launch path — through the eager first call, the CUDA-graph capture and a replay."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, sys, torch
sys.path.insert(0, %r)
from oracle import cogaim_oracle as orc
from cognitive_aim_depth_estimation_b200.model import create_model
cfg = {"model": {"cognitive_modules": ["ambient_stream", "iterative_focal_stream", "exif_prior_database"]}}
m = create_model(cfg, {"num_cameras": 71}, device="cuda:0")
m.load_state_dict(orc.build_state_dict(0))
x, ex = orc.synthetic_images(2, 224), orc.synthetic_exif(2)
ex = {k: v.cuda() for k, v in ex.items()}
out = []
for _ in range(3):   # eager, captured, replayed
    torch.manual_seed(11)
    d, c, h = m.forward_with_guidance(x.cuda(), ex, "top-left", return_attention=True)
    out.append([d.flatten().tolist(), c.flatten().tolist(), h.argmax(-1).tolist(), float(h.double().sum())])
print("RESULT " + json.dumps(out))
""" % ROOT


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    res = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_programmatic_dependent_launch_gives_identical_results(cuda_device):
    base = _run({"CA_PDL": "0"})
    pdl = _run({"CA_PDL": "1"})
    assert base[0] == base[1] == base[2]          # eager == captured == replayed
    assert pdl == base

"""Build the C-ABI library of another git revision (or the working tree) next to the current one, for same-box A/B
timing:  python tools/build_variant.py <rev|WORK> <name> [-DMACRO=v ...]  ->  build/variants/<name>/libcogaim_b200.so"""
import concurrent.futures as cf
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rev, name = sys.argv[1], sys.argv[2]
dst = ROOT / "build" / "variants" / name
(dst / "pkg" / "csrc").mkdir(parents=True, exist_ok=True)
(dst / "include").mkdir(exist_ok=True)
rel = "cognitive_aim_depth_estimation_b200/csrc"
if rev == "WORK":
    files = [f"{rel}/{f.name}" for f in (ROOT / rel).iterdir() if f.suffix in (".cu", ".cuh", ".h")] + ["include/cogaim_b200.h"]
    get = lambda f: (ROOT / f).read_bytes()  # noqa: E731
else:
    ls = subprocess.run(["git", "ls-tree", "--name-only", rev, f"{rel}/", "include/"], cwd=ROOT, capture_output=True, text=True).stdout.split()
    files = [f for f in ls if f.endswith((".cu", ".cuh", ".h"))]
    get = lambda f: subprocess.run(["git", "show", f"{rev}:{f}"], cwd=ROOT, capture_output=True).stdout  # noqa: E731
for f in files:
    out = dst / ("include" if f.startswith("include/") else "pkg/csrc") / os.path.basename(f)
    out.write_bytes(get(f))
srcs = sorted((dst / "pkg" / "csrc").glob("*.cu"))
flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", f"-I{dst / 'include'}", *sys.argv[3:]]


def cc(s):
    o = s.with_suffix(".o")
    subprocess.run(["nvcc", *flags, "-c", str(s), "-o", str(o)], check=True)
    return o


with cf.ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(cc, srcs))
lib = dst / "libcogaim_b200.so"
subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib), *map(str, objs)], check=True)
print(lib)

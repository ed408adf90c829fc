"""In-tree build of libcogaim_b200.so (the C-ABI library holding every sm_100a kernel).

`python -m cognitive_aim_depth_estimation_b200.build` or `__graft_entry__.build()` runs this.  nvcc cross-compiles for
sm_100a without a GPU.  Objects are cached in `csrc/_obj/` keyed on source mtimes so an edit to one kernel
recompiles one file.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OBJ = CSRC / "_obj"
LIB = PKG_DIR / "libcogaim_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build the sm_100a kernels")
    return nvcc


def _headers_mtime() -> float:
    hs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((PKG_DIR.parent / "include").glob("*.h"))
    return max(h.stat().st_mtime for h in hs)


def _compile_one(src: Path, verbose: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    newest = max(src.stat().st_mtime, _headers_mtime())
    if obj.exists() and obj.stat().st_mtime >= newest:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    """Compile every csrc/*.cu for sm_100a and link libcogaim_b200.so next to this file."""
    OBJ.mkdir(exist_ok=True)
    if force:
        for o in OBJ.glob("*.o"):
            o.unlink()
    srcs = sorted(CSRC.glob("*.cu"))
    if not srcs:
        raise RuntimeError("no CUDA sources found")
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, verbose), srcs))
    if LIB.exists() and all(LIB.stat().st_mtime >= o.stat().st_mtime for o in objs) and not force:
        return LIB
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(path)

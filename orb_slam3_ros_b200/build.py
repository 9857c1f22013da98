"""Build liborbb200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

The library is cross-compiled where there is no GPU (nvcc needs none) and travels to the GPU box with the repo
snapshot.  -fmad=false: the un-fused float32 result is the specification (orientation / descriptor / stereo
parabola must match the reference's float arithmetic bit for bit).
"""
import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "liborbb200.so"
SOURCES = ["orbb_extract.cu", "orbb_match.cu", "orbb_bow.cu", "orbb_nccl.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2", "--shared", "-cudart", "shared",
]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def is_stale():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.inc")) + [PKG.parent / "include" / "orbb200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [str(CSRC / s) for s in SOURCES] + ["-o", str(LIB), "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build_library(force=True, verbose="-v" in sys.argv))

"""In-tree build of the CUDA extension (sm_100a only): ``python -m genvarloader_b200._build``.

nvcc cross-compiles without a GPU; the resulting ``_lib/libgvl_b200.so`` is git-ignored but
travels with the source tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "_lib" / os.environ.get("GVL_LIB_NAME", "libgvl_b200.so")
SOURCES = ["gvl_ctx.cu", "gvl_hap.cu", "gvl_svar2.cu", "gvl_tracks.cu", "gvl_aux.cu", "gvl_host.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--fmad=false",  # f64 Lagrange arithmetic of Interpolate must match the reference's un-fused ops
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-cudart", "static",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found")
    return exe


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "gvl_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    srcs = [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
    cmd = [nvcc(), *NVCC_FLAGS, *os.environ.get("GVL_EXTRA_NVCC_FLAGS", "").split(), "-shared", "-o", str(LIB), *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)

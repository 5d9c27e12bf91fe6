"""In-tree build of the CUDA extension (sm_100a only): ``python -m genvarloader_b200._build``.

nvcc cross-compiles without a GPU; the resulting ``_lib/libgvl_b200.so`` is git-ignored but
travels with the source tree.  Every translation unit is compiled to its own object (in parallel, cached by
modification time under ``_lib/obj``), then linked.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "_lib" / os.environ.get("GVL_LIB_NAME", "libgvl_b200.so")
SOURCES = ["gvl_ctx.cu", "gvl_hap.cu", "gvl_svar2.cu", "gvl_tracks.cu", "gvl_aux.cu", "gvl_batch.cu", "gvl_variants.cu", "gvl_host.cu"]
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


def _deps() -> list[Path]:
    return list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "gvl_b200.h"]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(d.stat().st_mtime > t for d in _deps())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    extra = os.environ.get("GVL_EXTRA_NVCC_FLAGS", "").split()
    objdir = LIB.parent / ("obj" if LIB.name == "libgvl_b200.so" and not extra else "obj_" + LIB.stem)
    objdir.mkdir(exist_ok=True)
    headers_t = max(d.stat().st_mtime for d in _deps() if d.suffix != ".cu")
    srcs = [CSRC / s for s in SOURCES if (CSRC / s).exists()]

    def compile_one(src: Path) -> Path:
        obj = objdir / (src.stem + ".o")
        if not force and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, headers_t):
            return obj
        cmd = [nvcc(), *NVCC_FLAGS, *extra, "-c", "-o", str(obj), str(src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, srcs))
    subprocess.check_call([nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
                           "-o", str(LIB), *map(str, objs)])
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)

"""In-tree build of librbd_b200.so with nvcc for sm_100a (no torch extension machinery).

    python -m rbdreference_b200.build [--force]

The library is a plain C-ABI shared object (include/rbd_b200.h); it is git-ignored but travels
with the working tree to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "librbd_b200.so")
SOURCES = ["rbd_capi.cu"]
HEADERS = ["rbd_common.cuh", "rbd_pass_kernels.cuh", "rbd_fused_kernels.cuh", "rbd_grad_kernels.cuh",
           "rbd_minv_kernels.cuh", "rbd_coop_kernels.cuh", "rbd_coop_minv_kernels.cuh", "rbd_lane_minv_kernels.cuh", "rbd_fd_kernels.cuh", os.path.join("..", "..", "include", "rbd_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--expt-relaxed-constexpr",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = True, extra_flags=()) -> str:
    if not force and not needs_build():
        return LIB_PATH
    cmd = [find_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print("[rbdreference_b200.build]", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB_PATH)

"""In-tree build of librbd_b200.so with nvcc for sm_100a (no torch extension machinery).

    python -m rbdreference_b200.build [--force]

The library is a plain C-ABI shared object (include/rbd_b200.h); it is git-ignored but travels
with the working tree to the GPU box.  The translation units are compiled in parallel
(`nvcc -c`, one process each) and linked with `nvcc -shared`; the launcher files
are compiled once per precision (-DRBD_LAUNCH_T=double / float).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(CSRC, "_build")
LIB_PATH = os.path.join(PKG_DIR, "librbd_b200.so")

# (object name, source, extra defines)
UNITS = [("rbd_capi.o", "rbd_capi.cu", []), ("rbd_launch_ee.o", "rbd_launch_ee.cu", []),
         ("rbd_launch_fb.o", "rbd_launch_fb.cu", [])]
for _src in ("rbd_launch_rnea.cu", "rbd_launch_grad.cu", "rbd_launch_minv.cu", "rbd_launch_pass.cu"):
    for _t in ("double", "float"):
        UNITS.append(("%s_%s.o" % (_src[:-3], _t), _src, ["-DRBD_LAUNCH_T=%s" % _t]))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(PKG_DIR, "..", "include", "rbd_b200.h"))
    hs.append(os.path.abspath(__file__))
    return hs


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    hs = _headers()
    return any(_stale(os.path.join(OBJ_DIR, o), hs + [os.path.join(CSRC, s)]) for o, s, _ in UNITS) or \
        _stale(LIB_PATH, [os.path.join(OBJ_DIR, o) for o, _, _ in UNITS])


def build(force: bool = False, verbose: bool = True, extra_flags=()) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    hs = _headers()
    jobs = []
    for obj, src, defs in UNITS:
        o, s = os.path.join(OBJ_DIR, obj), os.path.join(CSRC, src)
        if force or _stale(o, hs + [s]):
            jobs.append([nvcc] + NVCC_FLAGS + list(extra_flags) + defs + ["-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print("[rbdreference_b200.build]", " ".join(cmd), flush=True)
        subprocess.run(cmd, check=True, cwd=CSRC)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(run, jobs))
    run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] +
        [os.path.join(OBJ_DIR, o) for o, _, _ in UNITS])
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB_PATH)

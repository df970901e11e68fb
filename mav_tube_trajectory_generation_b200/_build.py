"""In-tree build of libmtg_cuda.so (sm_100a only).

    python -m mav_tube_trajectory_generation_b200._build [--force]

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to
the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libmtg_cuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--fmad=true",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                  glob.glob(os.path.join(CSRC, "*.cpp")) + glob.glob(os.path.join(CSRC, "*.h")) +
                  [os.path.join(os.path.dirname(PKG), "include", "mtg_cuda.h")])


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def _run(cmd, verbose):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(cmd))


def _stale(obj, deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles every csrc/*.cu to an object (in parallel; only the stale ones) and links
    libmtg_cuda.so. Each .cu is a self-contained translation unit (no -rdc)."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(BUILD, exist_ok=True)
    headers = [s for s in _sources() if s.endswith((".cuh", ".h"))]
    jobs, objs = [], []
    tables_o = os.path.join(BUILD, "tables.o")
    objs.append(tables_o)
    if force or _stale(tables_o, [os.path.join(CSRC, "tables.cpp"), os.path.join(CSRC, "tables.h")]):
        # g++ from PATH: the image exports CXX=/opt/gcc/bin/g++, which is not what nvcc pairs with
        jobs.append(["g++", "-O2", "-fPIC", "-std=c++14", "-c", os.path.join(CSRC, "tables.cpp"), "-o", tables_o])
    for cu in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        obj = os.path.join(BUILD, os.path.basename(cu)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [cu] + headers):
            jobs.append([_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", cu, "-o", obj])
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(lambda c: _run(c, verbose), jobs))
    _run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-ldl"], verbose)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

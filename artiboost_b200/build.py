"""Builds the C-ABI CUDA library in-tree: artiboost_b200/csrc/*.cu -> artiboost_b200/libartiboost_b200.so.

sm_100a only (`-gencode arch=compute_100a,code=sm_100a`); nvcc cross-compiles without a GPU.  Objects are rebuilt
only when their source (or a header) is newer.  `python -m artiboost_b200.build [--force] [--verbose]`.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(HERE, "libartiboost_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr"]
# The rasteriser's rule set requires individually rounded fp32 operations (see oracle/raster.c header).
# augment.cu reproduces Pillow's fp32 / fp64 arithmetic operation by operation (oracle/augment.py): no FMA contraction.
PER_FILE = {"raster.cu": ["-fmad=false"], "augment.cu": ["-fmad=false"]}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: artiboost_b200 needs the CUDA toolkit to build its sm_100a library")


def _newest_header() -> float:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "artiboost_b200.h"))
    return max(os.path.getmtime(p) for p in deps if os.path.exists(p))


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdr_t = _newest_header()
    jobs = []
    objs = []
    for s in srcs:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ_DIR, s[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            cmd = [nvcc, *ARCH, *COMMON, *PER_FILE.get(s, []), *os.environ.get("AB_EXTRA_NVCC_FLAGS", "").split(), "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB_PATH):
        # -Xcompiler -fvisibility=hidden + AB_API keeps the export list equal to include/artiboost_b200.h
        run([nvcc, *ARCH, "-shared", "-o", LIB_PATH, *objs, "-lcuda"])
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

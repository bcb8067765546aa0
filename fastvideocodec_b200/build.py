"""Builds libfvc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m fastvideocodec_b200.build [--force]

The shared library has a plain C ABI (include/fvc_b200.h), links cudart statically and does not
depend on libcuda at load time (the driver entry point for TMA descriptors is resolved at run time).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libfvc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
SPLIT = os.environ.get("FVC_SPLIT", "fp16").lower()   # fp16 (default, 22-bit pairs) | bf16 (16-bit pairs)
FLAGS = ["-DFVC_SPLIT_FP16=%d" % (0 if SPLIT == "bf16" else 1), "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "-Xptxas", "-v"] + os.environ.get("FVC_NVCC_EXTRA", "").split()   # e.g. -DFVC_TC_ACCDBG (accumulator-warp counters)


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [NVCC] + ARCH + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ, src[:-3] + ".ptxas.log")
    with open(log, "w") as f:
        f.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout + r.stderr))
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError("nvcc not found at %s and %s is missing or stale" % (NVCC, LIB))
    with ThreadPoolExecutor(max_workers=min(8, len(sources()))) as ex:
        objs = list(ex.map(_compile, sources()))
    cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fvisibility=hidden"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)

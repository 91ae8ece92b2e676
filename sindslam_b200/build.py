"""Builds libsindyn_cuda.so in-tree with nvcc for sm_100a (B200) -- no other architecture.

    python -m sindslam_b200.build [--force] [--verbose]

The built library is git-ignored but travels to the GPU box with gpurun snapshots.
"""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libsindyn_cuda.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "-ccbin", "/usr/bin/g++"]
# developer builds only (e.g. SINDYN_NVCC_EXTRA=-DSINDYN_BROX_PHASE_CLOCKS for the phase-clock instrumentation of k_brox_sor)
COMMON += os.environ.get("SINDYN_NVCC_EXTRA", "").split()
# Bit-exact integer/label stages are compiled without FMA contraction so that float/double
# expressions evaluate like the reference's plain IEEE arithmetic; the flow solvers (tolerance
# parity) keep FMA.
FMAD_ON = {"brox.cu", "varref.cu"}


def _sig(path: str, flags) -> str:
    h = hashlib.sha1()
    h.update(" ".join(flags).encode())
    for p in [path] + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "sindyn.h")]:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _compile(src: str, force: bool, verbose: bool):
    name = os.path.basename(src)
    obj = os.path.join(OBJ, name + ".o")
    flags = ARCH + COMMON + ([] if name in FMAD_ON else ["-fmad=false"])
    if verbose:
        flags = flags + ["-Xptxas", "-v"]
    sig = _sig(src, flags)
    sigf = obj + ".sig"
    if not force and os.path.exists(obj) and os.path.exists(sigf) and open(sigf).read() == sig:
        return obj, ""
    cmd = [NVCC] + flags + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (name, r.stdout, r.stderr))
    with open(sigf, "w") as f:
        f.write(sig)
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    objs = [o for o, _ in res]
    logs = "".join(l for _, l in res)
    if verbose and logs:
        print(logs)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    build_examples(force)
    return LIB


EXAMPLE_BIN = os.path.join(HERE, "bin", "rgbd_tum_noros")


def build_examples(force: bool = False) -> str:
    """The reference-shaped C++ driver (examples/rgbd_tum_noros.cpp) on top of include/sindyn_classes.hpp."""
    src = os.path.join(HERE, "..", "examples", "rgbd_tum_noros.cpp")
    deps = [src, os.path.join(HERE, "..", "include", "sindyn_classes.hpp"), os.path.join(HERE, "..", "include", "sindyn.h"), LIB]
    os.makedirs(os.path.dirname(EXAMPLE_BIN), exist_ok=True)
    if not force and os.path.exists(EXAMPLE_BIN) and os.path.getmtime(EXAMPLE_BIN) >= max(os.path.getmtime(d) for d in deps):
        return EXAMPLE_BIN
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", src, "-o", EXAMPLE_BIN, "-L" + HERE, "-lsindyn_cuda", "-Wl,-rpath," + HERE,
           "-Wl,-rpath,$ORIGIN/.."]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed for the example driver:\n%s\n%s" % (r.stdout, r.stderr))
    return EXAMPLE_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

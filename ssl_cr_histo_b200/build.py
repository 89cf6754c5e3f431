"""In-tree nvcc build of libb2n.so (sm_100a only) and the stand-alone device test binary.

The shared library links the CUDA runtime statically and resolves the two driver entry points
it needs (cuTensorMapEncode*) at run time, so it also loads on a machine without a GPU driver
(the CPU-side symbol-export test relies on that).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb2n.so")
SOURCES = ["api.cu", "errors.cu", "tmap.cu", "conv_launch.cu", "bn.cu", "stem_pool.cu", "pack.cu",
           "linear.cu", "loss_lerp.cu", "optim.cu", "augment.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo"]


def _deps():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    files.append(os.path.join(HERE, "..", "include", "b2n.h"))
    return files


def _stale(target: str) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in _deps())


def build_lib(force: bool = False, verbose: bool = False) -> str:
    """Compile libb2n.so in-tree.  Several ranks may import the package at once on a box where the
    library is missing: an exclusive file lock lets one of them compile (into a temporary name,
    renamed atomically) while the others wait and then find it fresh."""
    import fcntl

    if not force and not _stale(LIB):
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale(LIB):       # another process built it while we waited
                return LIB
            tmp = "%s.%d.tmp" % (LIB, os.getpid())
            cmd = ["nvcc", *ARCH, *COMMON, "-shared", "-Xcompiler", "-fPIC", "-cudart", "static",
                   "-o", tmp, *[os.path.join(CSRC, s) for s in SOURCES]]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            subprocess.run(cmd, check=True)
            os.replace(tmp, LIB)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


def build_devtest(force: bool = False) -> str:
    out = os.path.normpath(os.path.join(HERE, "..", "build", "devtest"))
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not force and not _stale(out):
        return out
    srcs = ["devtest.cu", "conv_launch.cu", "tmap.cu", "errors.cu"]
    cmd = ["nvcc", *ARCH, *COMMON, "-cudart", "static", "-o", out,
           *[os.path.join(CSRC, s) for s in srcs]]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_devtest(force="--force" in sys.argv)
    print(LIB)

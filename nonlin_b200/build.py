"""Build the engine's shared library in-tree with nvcc for sm_100a (no JIT cache, no arch list).

    python nonlin_b200/build.py [--force] [--verbose]

The parity build passes -fmad=false: every multiply-add is a DMUL followed by a DADD, as in a
default gfortran build of the reference.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnonlin_b200.so")
SOURCES = ["nlb_api.cu", "coop_kernels.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",
    "-shared", "-Xcompiler", "-fPIC",
    "-diag-suppress", "177",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the engine has no prebuilt fallback")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "nonlin_batch.h")]
    return any(os.path.getmtime(d) > t for d in deps)


LIB_FAST = os.path.join(HERE, "libnonlin_b200_fast.so")


def build_fast(force=False):
    """The same sources WITH FMA contraction (nvcc's default): not bit-compatible with the reference's arithmetic, kept as
    a separate library that is only ever loaded explicitly (NLB_LIB=...; scripts/fast_build_stats.py measures how far its
    results move: x / f agreement and equality of the iteration counts against the CPU oracle)."""
    if not force and os.path.exists(LIB_FAST) and os.path.getmtime(LIB_FAST) >= max(
            os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)):
        return LIB_FAST
    flags = [f for f in NVCC_FLAGS if f != "-fmad=false"]
    cmd = [_nvcc()] + flags + ["-o", LIB_FAST] + [os.path.join(CSRC, s) for s in SOURCES]
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    subprocess.check_call(cmd)
    return LIB_FAST


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    # nvcc's host compiler must be the distribution g++ (an /opt wrapper g++ on PATH lacks libgomp specs)
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    if "--fast" in sys.argv:
        print(build_fast(force="--force" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

"""Builds granne_b200/libgranne_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "granne_b200.cu")
DEPS = sorted(os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))) + [
    os.path.join(ROOT, "include", "granne_b200.h")]
LIB = os.path.join(HERE, "libgranne_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    # no --use_fast_math, and contraction left to explicit __fmaf_rn/__fadd_rn intrinsics in the distance code
    "-fmad=false",
]
# `--split-compile 0` halves the build time (~55 kernel instantiations) but measured 2% slower code (80 vs 96 registers
# in the bench kernel); set GRANNE_B200_FAST_BUILD=1 to use it while iterating.
if os.environ.get("GRANNE_B200_FAST_BUILD"):
    NVCC_FLAGS += ["--split-compile", "0"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    subprocess.check_call(cmd)
    return LIB


def build_variant(tag, defines=(), extra_flags=()):
    """Experimental build next to the default library: libgranne_b200_<tag>.so with extra -D defines (e.g.
    GB_MIN_BLOCKS=28) — select it at run time with GRANNE_B200_LIB=<path> (tools/ab_bench.py compares variants)."""
    out = os.path.join(HERE, "libgranne_b200_%s.so" % tag)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D%s" % d for d in defines] + list(extra_flags) + ["-o", out, SRC]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

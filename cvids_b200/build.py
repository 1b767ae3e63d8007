"""Build the CUDA shared library in-tree: cvids_b200/libchisel_b200.so (sm_100a only).

    python -m cvids_b200.build [--force]

nvcc cross-compiles without a GPU. -fmad=false: nothing on the exact-arithmetic path may be contracted into
FMA; -lineinfo so that ncu's source page maps to the .cu files.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libchisel_b200.so")
SOURCES = ["capi.cu", "integrate.cu", "integrate_batch_half.cu", "integrate_batch_quarter.cu", "mesh.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-Wno-unused-function",
    "-shared", "-cudart", "shared", "-ldl",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(HERE, "..", "include", "chisel_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str = OUT, defines=()) -> str:
    """defines / out: differently tuned variants for A/B runs (tools/ab_build.py); the product build uses neither."""
    if out == OUT and not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + [os.path.join(SRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

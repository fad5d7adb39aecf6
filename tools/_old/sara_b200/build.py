"""Builds the C-ABI shared library (hand-written sm_100a CUDA kernels) in-tree.

    python -m sara_b200.build [--force]

nvcc cross-compiles without a GPU.  -fmad=false is part of the numerical
contract: the pyramid must be bit-identical to the reference's fp32 arithmetic
(separate multiply and add, no contraction).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsara_b200.so")
SOURCES = ["ctx.cu", "match.cu", "ingest.cu", "pyramid.cu", "pyramid_fused.cu", "pyramid_stage.cu", "pyramid_march.cu", "extrema.cu", "describe.cu"]
HEADERS = ["common.cuh", "match.cuh", "scan.cuh", "fp32x2_tma.cuh", os.path.join("..", "..", "include", "sara_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
    "--shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    if not all(os.path.exists(s) for s in srcs):
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("CUDA sources missing and no prebuilt libsara_b200.so")
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

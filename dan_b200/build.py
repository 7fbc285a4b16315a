"""Build the C-ABI shared library dan_b200/libdan_b200.so for sm_100a with nvcc.

    python -m dan_b200.build            # build if stale
    python -m dan_b200.build --force

The library is built IN-TREE so that it travels with the repository snapshot to the
GPU box; it is git-ignored (*.so)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdan_b200.so")
SOURCES = ["api.cu", "anchors.cu", "encode.cu", "postprocess.cu", "mining.cu", "routing.cu", "vote.cu", "handoff.cu"]
HEADERS = ["common.cuh", "heap_order.cuh", "sort.cuh", os.path.join(ROOT, "include", "dan_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # no FMA contraction: every fp32 op separately rounded
    "-Xcompiler", "-fPIC", "-shared",
    "-I" + os.path.join(ROOT, "include"),
    "-I" + CSRC,
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES]
    deps += [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB)

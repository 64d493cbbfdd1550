"""Build the in-tree CUDA library (nvcc, sm_100a only).  Stand-alone on purpose: it must be runnable
before the package can be imported (the package refuses to import without the library), e.g.
``python pointnet_refine_b200/build.py`` or ``__graft_entry__.build()``."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblrn_b200.so")
CSRC = os.path.join(HERE, "csrc")
HEADER = os.path.join(os.path.dirname(HERE), "include", "lrn_b200.h")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


TIMELINE_LIB_PATH = os.path.join(HERE, "liblrn_b200_timeline.so")   # tuning variant with clock64() stamps (tools/timeline.py)


def build(force: bool = False, verbose: bool = False, timeline: bool = False) -> str:
    out = TIMELINE_LIB_PATH if timeline else LIB_PATH
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + [HEADER]
    if not force and os.path.exists(out) and all(os.path.getmtime(s) <= os.path.getmtime(out) for s in srcs):
        return out
    cmd = ["nvcc", *NVCC_FLAGS, *(["-DLRN_TIMELINE"] if timeline else []), "-o", out, os.path.join(CSRC, "lrn_abi.cu")]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, timeline="--timeline" in sys.argv))

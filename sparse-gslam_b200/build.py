"""Build recipe for libsgb.so (CUDA kernels + C ABI), in-tree, sm_100a only.

`python -m sparse_gslam_b200.build` or `__graft_entry__.build()`; nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsgb.so")
SOURCES = ["sgb_backend.cu", "sgb_posegraph.cu", "sgb_frontend.cu", "sgb_structure.cpp", "sgb_partition.cpp", "sgb_g2o_io.cpp"]
HEADERS = ["sgb_kernels.cuh", "sgb_rows.h", "sgb_math.h", "sgb_types.h", "sgb_structure.h", "sgb_partition.h", "sgb_edits.h", "sgb_internal.h", "../../include/sgb_capi.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsgb.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [__file__]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    from ._buildlock import build_lock
    with build_lock(LIB) as tmp:
        if not force and not needs_build():  # another process built it while this one waited for the lock
            return LIB
        cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + \
              [os.path.join(CSRC, s) for s in SOURCES]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed building libsgb.so")
        if verbose:
            sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

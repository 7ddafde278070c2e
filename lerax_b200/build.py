"""Build the C-ABI shared library in-tree (lerax_b200/_C/liblerax_b200.so) for sm_100a.

`python -m lerax_b200.build` runs `make` in lerax_b200/csrc (nvcc cross-compiles without a
GPU).  The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""

from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "_C", "liblerax_b200.so")


def build(verbose: bool = False, jobs: int | None = None) -> str:
    jobs = jobs or min(8, os.cpu_count() or 1)
    proc = subprocess.run(["make", "-C", CSRC, f"-j{jobs}"], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("lerax_b200: building the CUDA library failed (see output above)")
    if not os.path.exists(LIB):
        raise RuntimeError(f"lerax_b200: make succeeded but {LIB} is missing")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))

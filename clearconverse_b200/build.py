"""Builds ``libresep_b200.so`` in-tree with nvcc for sm_100a (``python -m clearconverse_b200.build``)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")


def build(verbose: bool = False, clean: bool = False) -> str:
    """Compile every CUDA source with ``-gencode arch=compute_100a,code=sm_100a -lineinfo``
    (see csrc/Makefile) and return the path of the shared library."""
    if clean:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, capture_output=not verbose)
    r = subprocess.run(["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc build of libresep_b200.so failed (see output above)")
    out = os.path.join(HERE, "libresep_b200.so")
    if not os.path.exists(out):
        raise RuntimeError(f"build finished but {out} is missing")
    return out


if __name__ == "__main__":
    print(build(verbose=True, clean="--clean" in sys.argv))

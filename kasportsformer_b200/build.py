"""Build libkasf.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m kasportsformer_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libkasf.so")
STAMP = OUT + ".stamp"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr"]


def _digest() -> str:
    h = hashlib.sha256(" ".join(FLAGS).encode())
    files = sorted(glob.glob(os.path.join(CSRC, "*"))) + [os.path.join(HERE, "..", "include", "kasf.h")]
    for f in files:
        with open(f, "rb") as fh:
            h.update(f.encode() + fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    dg = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(STAMP) and open(STAMP).read() == dg:
        return OUT
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building libkasf.so")
    with open(STAMP, "w") as f:
        f.write(dg)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

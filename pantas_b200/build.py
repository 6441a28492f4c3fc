"""Build libpantas_aug.so in-tree with nvcc for sm_100a (no JIT cache: the .so
travels with the repository snapshot to the GPU box)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(PKG, "csrc", "pantas_aug.cu")
DEPS = [SRC, os.path.join(PKG, "csrc", "aug_kernels.cuh"), os.path.join(PKG, "csrc", "tables.cuh"), os.path.join(PKG, "csrc", "line_core.cuh"), os.path.join(PKG, "csrc", "team_tiles.cuh"), os.path.join(os.path.dirname(PKG), "include", "pantas_aug.h")]
LIB = os.path.join(PKG, "libpantas_aug.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + p.stdout + p.stderr)
    if verbose:
        print(p.stderr)
    return LIB


if __name__ == "__main__":
    import sys

    print(build_library(force=True, verbose="-v" in sys.argv))

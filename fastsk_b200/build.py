"""In-tree build of the CUDA extension: nvcc -> fastsk_b200/libfastsk_b200.so (sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfastsk_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "fsk_lib.cu")]
DEPS = SOURCES + [os.path.join(HERE, "csrc", "fsk_kernels.cuh"), os.path.join(HERE, "csrc", "fsk_dense.cuh"), os.path.join(HERE, "csrc", "fsk_bucket.cuh"), os.path.join(HERE, "csrc", "fsk_segment.cuh"),
                  os.path.join(os.path.dirname(HERE), "include", "fastsk_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # fp64 normalisation / Welford must not be contracted into FMAs (bit parity with x86-64)
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return cand


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in DEPS)


def build_extension(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_extension(force=True, verbose=True))

"""Build the CUDA library in-tree: ``gadfit_b200/libgadfit_b200.so`` (sm_100a).

nvcc cross-compiles without a GPU.  The ``.so`` is git-ignored but travels to
the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libgadfit_b200.so")
SOURCES = ["solver.cu", "microbench.cu"]
HEADERS = [
    os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")
] + [
    os.path.join(ROOT, "include", "gadfit_b200.h"),
] + [
    os.path.join(ROOT, "include", "gadfit_b200", f)
    for f in os.listdir(os.path.join(ROOT, "include", "gadfit_b200"))
]

NVCC_FLAGS = [
    "-std=c++20", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [__file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cmd = [_nvcc(), *NVCC_FLAGS,
           "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           *[os.path.join(CSRC, s) for s in SOURCES],
           "-o", OUT, "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    env = dict(os.environ)
    # nvcc must use the system g++, not a wrapper that lacks libgomp specs
    env.pop("CXX", None)
    env.pop("CC", None)
    subprocess.run(cmd, check=True, env=env)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)

"""Build the CUDA library in-tree: ``gadfit_b200/libgadfit_b200.so`` (sm_100a).

nvcc cross-compiles without a GPU.  The kernel instantiations of each group of
fitting functions are a separate translation unit (``model_group.cu`` compiled
with ``-DGADFIT_MODEL_GROUP=k``); the units build in parallel and are linked
into one shared library.  The ``.so`` is git-ignored but travels to the GPU
box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
OUT = os.path.join(HERE, "libgadfit_b200.so")
N_MODEL_GROUPS = 30  # models::n_model_groups in include/gadfit_b200/fit_models.h

NVCC_FLAGS = [
    "-std=c++20", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _headers() -> list[str]:
    inc = os.path.join(ROOT, "include", "gadfit_b200")
    return ([os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
            + [os.path.join(ROOT, "include", "gadfit_b200.h")]
            + [os.path.join(inc, f) for f in os.listdir(inc)])


def _units() -> list[tuple[str, str, list[str]]]:
    """(source, object, extra flags) of every translation unit."""
    units = [(os.path.join(CSRC, "solver.cu"), os.path.join(OBJ, "solver.o"), []),
             (os.path.join(CSRC, "microbench.cu"), os.path.join(OBJ, "microbench.o"), [])]
    for g in range(N_MODEL_GROUPS):
        units.append((os.path.join(CSRC, "model_group.cu"), os.path.join(OBJ, f"model_group_{g}.o"),
                      [f"-DGADFIT_MODEL_GROUP={g}"]))
    return units


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _env() -> dict:
    env = dict(os.environ)
    # nvcc must use the system g++, not a wrapper that lacks libgomp specs
    env.pop("CXX", None)
    env.pop("CC", None)
    return env


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    deps = _headers() + [__file__] + [u[0] for u in _units()]
    return _stale(OUT, deps)


def build(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> str:
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    common = [*NVCC_FLAGS, *(extra or []), "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        common.insert(0, "-Xptxas=-v")
    headers = _headers() + [__file__]

    def compile_unit(unit):
        src, obj, flags = unit
        if not force and not _stale(obj, headers + [src]):
            return ""
        cmd = [nvcc, *common, *flags, "-c", src, "-o", obj]
        r = subprocess.run(cmd, env=_env(), capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        logs = list(pool.map(compile_unit, _units()))
    if verbose:
        sys.stderr.write("".join(logs))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
            *[u[1] for u in _units()], "-o", OUT, "-ldl"]
    subprocess.run(link, check=True, env=_env())
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)

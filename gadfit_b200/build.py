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
N_MODEL_GROUPS = 3  # models::n_model_groups in include/gadfit_b200/fit_models.h
N_TEST_MODELS = 30  # models::n_test_models in include/gadfit_b200/test_models.h
TEST_MODELS_OUT = os.path.join(HERE, "libgadfit_b200_testmodels.so")

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


def _compile_and_link(units, out, force, verbose, extra, libs, lineinfo=True):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    common = [*NVCC_FLAGS, *(extra or []), "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if not lineinfo:
        common.remove("-lineinfo")
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
        logs = list(pool.map(compile_unit, units))
    if verbose:
        sys.stderr.write("".join(logs))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
            *[u[1] for u in units], "-o", out, *libs]
    subprocess.run(link, check=True, env=_env())
    return out


def build(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> str:
    """libgadfit_b200.so: solver, kernels of the built-in fitting functions."""
    if not force and not needs_build():
        return OUT
    return _compile_and_link(_units(), OUT, force, verbose, extra, ["-ldl"])


def build_variant(tag: str, extra: list[str]) -> str:
    """A/B builds for kernel tuning: the same library compiled with extra
    flags (-DGADFIT_...=...) as gadfit_b200/_variants/libgadfit_b200_<tag>.so;
    select it at run time with GADFIT_B200_LIB=<path>."""
    out_dir = os.path.join(HERE, "_variants")
    os.makedirs(out_dir, exist_ok=True)
    units = [(src, os.path.join(OBJ, f"{tag}_{os.path.basename(obj)}"), flags) for src, obj, flags in _units()]
    return _compile_and_link(units, os.path.join(out_dir, f"libgadfit_b200_{tag}.so"), True, False, extra,
                             ["-ldl"])


def build_plugin(sources: list[str], out: str, force: bool = True, verbose: bool = False,
                 extra: list[str] | None = None) -> str:
    """A model plug-in (include/gadfit_b200/model_plugin.h) from user .cu files."""
    tag = os.path.splitext(os.path.basename(out))[0]
    units = [(os.path.abspath(src), os.path.join(OBJ, f"{tag}_{i}.o"), [])
             for i, src in enumerate(sources)]
    return _compile_and_link(units, os.path.abspath(out), force, verbose, extra, [])


def _test_model_units():
    src = os.path.join(CSRC, "test_models_plugin.cu")
    return [(src, os.path.join(OBJ, f"test_model_{k}.o"), [f"-DGADFIT_TEST_MODEL={k}"])
            for k in range(N_TEST_MODELS)]


def build_test_models(force: bool = False, verbose: bool = False) -> str:
    """libgadfit_b200_testmodels.so: the plug-in with the functions of the
    reference's nested / direct double-integral tests (test_models.h).  Built
    without -lineinfo: these 27 template-heavy functions are parity fixtures,
    not profiling targets, and line tables triple the size of the library that
    every gpurun call has to ship (187 MB -> 60 MB)."""
    units = _test_model_units()
    deps = _headers() + [__file__, units[0][0]]
    if not force and not _stale(TEST_MODELS_OUT, deps):
        return TEST_MODELS_OUT
    return _compile_and_link(units, TEST_MODELS_OUT, force, verbose, None, [], lineinfo=False)


DEVICE_CHECKS_OUT = os.path.join(ROOT, "tests", "_build", "libdevice_dual_check.so")


def build_device_checks(force: bool = False) -> str:
    """tests/_build/libdevice_dual_check.so (test infrastructure): the
    reference's AD known-answer expressions evaluated in a kernel,
    tests/device_dual_check.cu."""
    src = os.path.join(ROOT, "tests", "device_dual_check.cu")
    deps = [src, os.path.join(ROOT, "tests", "dual_check_expr.h"),
            os.path.join(ROOT, "include", "gadfit_b200", "dual.h")]
    if not force and not _stale(DEVICE_CHECKS_OUT, deps):
        return DEVICE_CHECKS_OUT
    os.makedirs(os.path.dirname(DEVICE_CHECKS_OUT), exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "tests"), src, "-o", DEVICE_CHECKS_OUT]
    r = subprocess.run(cmd, env=_env(), capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return DEVICE_CHECKS_OUT


if __name__ == "__main__":
    if "--plugin" in sys.argv:
        # python -m gadfit_b200.build --plugin a.cu [b.cu ...] -o libmy_models.so
        args = sys.argv[sys.argv.index("--plugin") + 1:]
        out = args[args.index("-o") + 1]
        srcs = [a for a in args[:args.index("-o")] if a.endswith(".cu")]
        print(build_plugin(srcs, out, verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
        if "--test-models" in sys.argv:
            print(build_test_models(force="--force" in sys.argv, verbose="-v" in sys.argv))

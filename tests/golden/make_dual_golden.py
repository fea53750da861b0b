"""TEST INFRASTRUCTURE.  Generates tests/golden/dual_ops_ref.json: every
elemental operation of the reference's AD variable (reference
c++/gadfit/automatic_differentiation.cpp:59-680) evaluated one at a time in
forward mode by the UNMODIFIED reference, compiled where it lies.  Run in the
build container only (needs /root/reference):

    python tests/golden/make_dual_golden.py
"""
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GADFIT_REFERENCE", "/root/reference")


def main():
    env = dict(os.environ)
    env.pop("CXX", None)
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "make_dual_golden")
        subprocess.run(["g++", "-std=c++20", "-O2", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(REF, "c++"), "-I", os.path.join(ROOT, "oracle", "spdlog_stub"),
                        os.path.join(HERE, "make_dual_golden.cpp"),
                        os.path.join(REF, "c++", "gadfit", "automatic_differentiation.cpp"),
                        os.path.join(REF, "c++", "gadfit", "exceptions.cpp"),
                        "-o", exe], check=True, env=env)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    values = {ln.split()[0]: [float(v) for v in ln.split()[1:]] for ln in out.splitlines()}
    with open(os.path.join(HERE, "dual_ops_ref.json"), "w") as f:
        json.dump({"source": "gadfit::AdVar forward mode, reference v2.0.1, g++ -O2 (make_dual_golden.py)",
                   "values": values}, f, indent=1)
    print(len(values), "entries")


if __name__ == "__main__":
    main()

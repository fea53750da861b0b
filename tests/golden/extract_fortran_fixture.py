"""Extract the data arrays of the reference's Fortran tests
(reference fortran/tests/{1_gaussian,2_integral_single,3_integral_double,4_multiple_curves}_data.F90)
into tests/golden/fortran_fixture.json.

Run in the build container, where /root/reference is mounted:
    python tests/golden/extract_fortran_fixture.py
Only DATA (the test vectors) is extracted, no code.
"""
import json
import os
import re

SRC = "/root/reference/fortran/tests"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fortran_fixture.json")


def arrays(path: str) -> dict:
    src = open(path).read()
    out = {}
    for m in re.finditer(r"real\(kp\), parameter :: (\w+)\(\*\) = \[(.*?)\]", src, re.S):
        body = m.group(2).replace("&", " ").replace("\n", " ")
        out[m.group(1)] = [float(t.strip().replace("d", "e").replace("_kp", ""))
                           for t in body.split(",") if t.strip()]
    return out


if __name__ == "__main__":
    out = {}
    for stem, prefix in (("1_gaussian", "gaussian"), ("2_integral_single", "integral_single"),
                         ("3_integral_double", "integral_double"), ("4_multiple_curves", "multiple_curves")):
        for name, values in arrays(os.path.join(SRC, stem + "_data.F90")).items():
            out[f"{prefix}_{name}"] = values
    out["_source"] = "reference fortran/tests/*_data.F90"
    json.dump(out, open(OUT, "w"))
    print(OUT, {k: len(v) for k, v in out.items() if k != "_source"})

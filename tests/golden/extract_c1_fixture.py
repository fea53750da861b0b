"""Extract the data arrays of the reference's LM-solver tests
(reference c++/tests/fixtures.h:9-68) into tests/golden/c1_fixture.json.

Run in the build container, where /root/reference is mounted:
    python tests/golden/extract_c1_fixture.py
Only DATA (the test vectors) is extracted, no code.
"""
import json
import os
import re

SRC = "/root/reference/c++/tests/fixtures.h"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_fixture.json")


def array(src: str, name: str) -> list:
    m = re.search(r"(?:constexpr std::array|const std::vector<double>) %s\s*\{(.*?)\};" % name, src, re.S)
    return [float(t) for t in m.group(1).replace("\n", " ").split(",") if t.strip()]


if __name__ == "__main__":
    src = open(SRC).read()
    names = ["fix_d", "x_data_1", "y_data_1", "x_data_2", "y_data_2",
             "x_data_single", "y_data_single",  # fixtures.h:72-151: integration tests
             "x_data_double", "y_data_double", "weights_double",  # fixtures.h:153-237
             "y_data_double_direct"]  # fixtures.h:239-265
    out = {n: array(src, n) for n in names}
    out["_source"] = "reference c++/tests/fixtures.h:9-68"
    json.dump(out, open(OUT, "w"))
    print(OUT, {k: len(v) for k, v in out.items() if k != "_source"})

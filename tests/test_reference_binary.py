"""The UNMODIFIED reference program (oracle/_ref/ref_output = g++ -O3 of /root/reference/project/raytracer/main.cpp,
built where the sources lie) run end to end on the golden bunny: its output.ppm is the shipped golden frame.  Runs
wherever the binary travelled (build container and GPU box); it is the baseline (A) of SURVEY.md §8d."""
import hashlib
import os
import re
import subprocess

import pytest

import conftest as T

BIN = os.path.join(T.ROOT, "oracle", "_ref", "ref_output")


def test_unmodified_reference_binary_reproduces_output_ppm(tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/ref_output not built (needs /root/reference)")
    (tmp_path / "models").mkdir()
    with open(tmp_path / "models" / "bunny.obj", "w") as f:
        for x, y, z in T.bunny_vertices():
            f.write("v %.9g %.9g %.9g\n" % (x, y, z))
    r = subprocess.run([BIN], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0
    ppm = open(tmp_path / "output.ppm", "rb").read()
    assert hashlib.md5(ppm).hexdigest() == "c69c66375f2c6bda433f9f457a4b2b2e"
    assert "Total number of nodes: 71895" in r.stdout
    assert "Number of Sphere intersection tests: 323685 test" in r.stdout
    m = re.search(r"Total time spent: ([0-9.eE+-]+)s", r.stdout)
    assert m and float(m.group(1)) > 0

"""The C++ host mirror (include/biogarden.hpp) is compiled against the C ABI here; on the GPU box the
reference's doctests, written against that mirror, are executed."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "doctests.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "doctests")


def _build():
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), SRC, "-o", EXE,
           "-L" + os.path.join(ROOT, "biogarden_b200"), "-lbgalign", "-Wl,-rpath," + os.path.join(ROOT, "biogarden_b200"),
           "-L/usr/local/cuda/lib64", "-lcudart"]
    subprocess.check_call(cmd)


def test_cpp_mirror_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_runs_reference_doctests():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cpp doctests ok" in out.stdout

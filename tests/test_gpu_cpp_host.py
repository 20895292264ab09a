"""Builds the C++ host mirror (cortex_b200/host/vector_index.hpp) together with the
reference's unit tests restated in C++ (tests/cpp/test_vector_index.cpp) and runs them."""
import os
import subprocess

import pytest

from cortex_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compile(tmp_path):
    lib = build.build()
    exe = str(tmp_path / "test_vector_index")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_vector_index.cpp"),
           lib, f"-Wl,-rpath,{os.path.dirname(lib)}", "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_host_mirror_compiles(tmp_path):
    """CPU: the header and the C++ tests compile and link against the C ABI."""
    _compile(tmp_path)


@pytest.mark.gpu
def test_cpp_reference_unit_tests(tmp_path):
    exe = _compile(tmp_path)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ok (0 failed of 12)" in r.stdout

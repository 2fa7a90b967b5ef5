"""Runs the C++ operator mirror test (tests/cpp/test_operators.cpp): ColumnBatchIterator -> ColumnTableScan /
FilterExec / SnappyHashAggregateExec (C++ host side over the C ABI) -> CollectAggregateExec, against the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_operators")


def build_binary():
    from oracle import oracle
    oracle.build()
    src = os.path.join(ROOT, "tests", "cpp", "test_operators.cpp")
    lib_dir = os.path.join(ROOT, "snappydata_b200", "csrc")
    deps = [src, os.path.join(lib_dir, "sd_operators.hpp"), os.path.join(ROOT, "include", "snappy_gpu.h")]
    if not os.path.exists(BIN) or any(os.path.getmtime(d) > os.path.getmtime(BIN) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", BIN, src, "-L" + lib_dir, "-lsnappygpu",
                               "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                               "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-ldl", "-lpthread"])
    return BIN


def test_cpp_operator_mirror_builds():
    """CPU: the C++ host mirror compiles and links against the C ABI (no CUDA call)."""
    assert os.path.exists(build_binary())


@pytest.mark.gpu
def test_cpp_operator_mirror_matches_oracle():
    out = subprocess.run([build_binary(), "150001"], capture_output=True, text=True, timeout=600)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok ") == 2

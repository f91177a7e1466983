"""Builds and runs the C++ test of the header-only host API (bio_ik_b200/host/bioik_host_api.hpp)."""
import os
import subprocess

import pytest

import __graft_entry__ as ge
from bio_ik_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "cpp", "test_host_api")


def build():
    ge.build_cuda()
    libdir = os.path.dirname(_abi.LIB_PATH)
    cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O1", "-Wall", "-o", EXE, os.path.join(HERE, "cpp", "test_host_api.cpp"), "-L" + libdir, "-lbioik_b200", "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True)


def test_cpp_host_api_flattening_and_loud_failure_without_gpu():
    build()
    out = subprocess.run([EXE], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host api ok" in out.stdout


@pytest.mark.gpu
def test_cpp_host_api_round_trip_on_gpu():
    build()
    out = subprocess.run([EXE, "gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "gpu leg" in out.stdout

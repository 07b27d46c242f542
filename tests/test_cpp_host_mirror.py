"""The C++ host mirror (toyni_b200/host/toyni.hpp) over the C ABI: compiles everywhere, runs on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.bin")


def _build():
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "toyni_b200", "host"),
           "-I", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"),
           "-L", os.path.join(ROOT, "toyni_b200"), "-lntt_cuda", "-L", os.path.join(ROOT, "oracle"), "-loracle",
           "-L/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath," + os.path.join(ROOT, "toyni_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
           "-Wl,-rpath,/usr/local/cuda/lib64", "-o", BIN]
    subprocess.check_call(cmd)


def test_cpp_host_mirror_compiles_and_links():
    _build()
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_cpp_host_mirror_runs_the_reference_gpu_tests():
    _build()
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all C++ host-mirror tests passed" in out.stdout

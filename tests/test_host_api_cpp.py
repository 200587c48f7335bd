"""The C++ host layer (include/kmerutils_b200.hpp) that mirrors the reference's Rust API above the C ABI.

CPU: the test program compiles, links against the in-tree library and fails loudly without a device.
GPU: tests/cpp/test_host_api.cpp runs the reference's own unit tests (same inputs and assertions, cited there) and
compares every signature / count with the CPU oracle bit for bit.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp")
OUT = os.path.join(ROOT, "tests", "cpp", "_build", "test_host_api")


def build_program():
    import kmerutils_b200
    kmerutils_b200.load_library()  # builds / checks the CUDA library
    from oracle_lib import get_oracle
    get_oracle()
    deps = [SRC, os.path.join(ROOT, "include", "kmerutils_b200.hpp"), os.path.join(ROOT, "include", "kmerutils_b200.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call([
        "g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", SRC, "-o", OUT,
        "-L" + os.path.join(ROOT, "kmerutils_b200"), "-lkmerutils_b200", "-L" + os.path.join(ROOT, "oracle"), "-lkmer_oracle",
        "-Wl,-rpath,$ORIGIN/../../../kmerutils_b200", "-Wl,-rpath,$ORIGIN/../../../oracle"])
    return OUT


def test_host_api_compiles_and_refuses_cpu():
    import torch
    exe = build_program()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU path" in r.stderr, r.stderr


@pytest.mark.gpu
def test_host_api_reference_tests(tmp_path):
    exe = build_program()
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host api ok" in r.stdout

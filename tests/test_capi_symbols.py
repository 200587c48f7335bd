"""The C-ABI library loads without a GPU and exports every symbol include/kmerutils_b200.h declares.
No compute call is made here; without a device every compute entry point must fail loudly."""
import ctypes as C
import os
import re

import pytest

import kmerutils_b200 as kb
from kmerutils_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kmerutils_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kmu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = kb.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/kmerutils_b200.h but not exported"
    # and the Python binding covers every declared entry point
    missing = [n for n in names if n not in _lib.SIGNATURES]
    assert not missing, f"ctypes signatures missing for {missing}"


def test_version_string():
    lib = kb.load_library()
    assert b"sm_100a" in lib.kmu_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device behaviour is checked on the CPU box")
    lib = kb.load_library()
    ctx = C.c_void_p()
    rc = lib.kmu_ctx_create(0, C.byref(ctx))
    assert rc == _lib.KMU_ECUDA and not ctx.value
    assert b"no CPU path" in lib.kmu_last_error()
    with pytest.raises(kb.KmuError):
        kb.Engine(0)


def test_product_does_not_import_oracle():
    # the oracle is test infrastructure: nothing under kmerutils_b200/ may reference it
    pkg = os.path.join(ROOT, "kmerutils_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_lib" not in text and "kmer_oracle" not in text and "libkmer_oracle" not in text, f

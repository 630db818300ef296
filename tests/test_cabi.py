"""CPU tests of the boundary: libhlsd.so loads and exports exactly what include/hlsd.h declares;
argument errors are reported without a GPU; nothing under the product package touches oracle/."""
import ctypes
import os
import re

import numpy as np
import pytest

import hls_designs_b200 as hd
from conftest import ROOT
from hls_designs_b200 import _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hlsd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hlsd_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = hd.load()
    declared = _declared_symbols()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/hlsd.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "python binding and header disagree"
    assert b"sm_100a" in lib.hlsd_version()


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "hlsd.h")).read()
    assert "torch" not in text and "at::" not in text and "#include <cuda" not in text


def test_argument_errors_without_gpu():
    lib = hd.load()
    a = np.zeros((1, 4, 4), np.float32)
    assert lib.hlsd_cholesky_host(None, a.ctypes.data, 4, 1, 13, 0) == 1
    assert lib.hlsd_cholesky_host(a.ctypes.data, a.ctypes.data, 0, 1, 13, 0) == 1
    assert lib.hlsd_cholesky_host(a.ctypes.data, a.ctypes.data, 4, -1, 13, 0) == 1
    assert b"batch" in lib.hlsd_last_error()
    assert lib.hlsd_qr_host(a.ctypes.data, a.ctypes.data, a.ctypes.data, 3, 4, 1, 21, 0) == 1  # rows < cols
    assert lib.hlsd_cholesky_dev(1, 1, 4, 1, 99, 0, None) == 1  # unknown design, checked before any CUDA call
    assert b"design" in lib.hlsd_last_error()
    assert lib.hlsd_lu_dev(1, 1, 1, 1, 4, 1, 11, 0x1000, None) == 1  # unknown flag bits
    assert b"flag" in lib.hlsd_last_error()
    assert lib.hlsd_lu_dev(1, 1, 1, 1, 4, 1, 11, hd.FLAG_EXACT | hd.FLAG_FUSED, None) == 1  # exclusive
    assert lib.hlsd_cholesky_dev(1, 1, 256, 1, 13, hd.FLAG_EXACT | hd.PATHS["tensor"], None) == 1
    assert lib.hlsd_cholesky_host_multi(None, a.ctypes.data, 4, 1, 13, 0, None, 0) == 1
    devs = (ctypes.c_int * 1)(0)
    assert lib.hlsd_cholesky_host_multi(a.ctypes.data, a.ctypes.data, 4, 1, 13, 0, devs, 0) == 1  # a device list of length 0
    assert lib.hlsd_selftest_copy_host(None, a.ctypes.data, 64, 64, 1) == 1
    assert lib.hlsd_cholesky_host(a.ctypes.data, a.ctypes.data, 4, 0, 13, 0) == 0  # empty batch is a no-op


def test_python_mirror_validates():
    with pytest.raises(ValueError):
        hd.cholesky(np.zeros((2, 3, 4), np.float32))
    with pytest.raises(ValueError):
        hd.cholesky(np.zeros((4, 4), np.float32), design="cholesky_v9")
    with pytest.raises(ValueError):
        hd.lu(np.zeros((4, 4), np.float32), path="cpu")


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: no file of the product package may even name it."""
    pkg = os.path.join(ROOT, "hls_designs_b200")
    for dirpath, dirs, files in os.walk(pkg):
        dirs[:] = [d for d in dirs if d not in ("build", "__pycache__")]
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read().lower()
                for word in ("oracle", "restate", "librestate", "_ref"):
                    assert word not in text, f"{f} mentions {word!r}"

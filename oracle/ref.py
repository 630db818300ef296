"""TEST INFRASTRUCTURE ONLY (oracle/): ctypes access to the reference designs built by
build_ref.py into oracle/_ref/ (the reference's own `top()` compiled with g++ and the HLS header
stand-ins).  Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; the product package never imports this.
"""
import ctypes
import os

import numpy as np

_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_cache = {}

CHOL_DESIGNS = ("cholesky_v1.3", "cholesky_v2.2", "cholesky_v3.2", "cholesky_v4.0")
LU_DESIGNS = ("lu_v1.1", "lu_v1.2", "lu_v2.1")
QR_DESIGNS = ("qr_v1.1", "qr_v1.2", "qr_v2.1")


def lib_path(design, tag):
    return os.path.join(_REF_DIR, f"lib{design}_{tag}.so")


def available(design, tag):
    return os.path.exists(lib_path(design, tag))


def _load(design, tag):
    key = (design, str(tag))
    if key not in _cache:
        path = lib_path(design, tag)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `python oracle/build_ref.py` where /root/reference exists")
        _cache[key] = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    return _cache[key]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def ref_cholesky(A, design="cholesky_v1.3", tag=None, threads=1):
    """A: (batch, n, n) or (n, n) float32 -> L, via the reference's top(A, L)."""
    A = _f32(A)
    single = A.ndim == 2
    A3 = A.reshape((-1,) + A.shape[-2:])
    n = A3.shape[-1]
    lib = _load(design, tag or n)
    assert lib.ref_rows() == n, (lib.ref_rows(), n)
    L = np.empty_like(A3)
    lib.ref_chol.argtypes = [ctypes.POINTER(ctypes.c_float)] * 2 + [ctypes.c_long, ctypes.c_int]
    rc = lib.ref_chol(_fp(A3), _fp(L), A3.shape[0], threads)
    assert rc == 0
    return L[0] if single else L


def ref_lu(A, design="lu_v1.1", tag=None, threads=1):
    """A -> (L, U, P) via the reference's top(A, L, U, P)."""
    A = _f32(A)
    single = A.ndim == 2
    A3 = A.reshape((-1,) + A.shape[-2:])
    n = A3.shape[-1]
    lib = _load(design, tag or n)
    assert lib.ref_rows() == n
    L = np.empty_like(A3)
    U = np.empty_like(A3)
    P = np.empty((A3.shape[0], n), dtype=np.int32)
    lib.ref_lu.argtypes = [ctypes.POINTER(ctypes.c_float)] * 3 + [ctypes.POINTER(ctypes.c_int), ctypes.c_long, ctypes.c_int]
    rc = lib.ref_lu(_fp(A3), _fp(L), _fp(U), P.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), A3.shape[0], threads)
    assert rc == 0
    return (L[0], U[0], P[0]) if single else (L, U, P)


def ref_qr(A, design="qr_v2.1", tag=None, threads=1):
    """A: (batch, rows, cols) -> (Q, R) via the reference's top(A, Q, R)."""
    A = _f32(A)
    single = A.ndim == 2
    A3 = A.reshape((-1,) + A.shape[-2:])
    rows, cols = A3.shape[-2:]
    lib = _load(design, tag or f"{rows}x{cols}")
    assert (lib.ref_rows(), lib.ref_cols()) == (rows, cols)
    Q = np.empty((A3.shape[0], rows, rows), dtype=np.float32)
    R = np.empty_like(A3)
    lib.ref_qr.argtypes = [ctypes.POINTER(ctypes.c_float)] * 3 + [ctypes.c_long, ctypes.c_int]
    rc = lib.ref_qr(_fp(A3), _fp(Q), _fp(R), A3.shape[0], threads)
    assert rc == 0
    return (Q[0], R[0]) if single else (Q, R)

"""TEST INFRASTRUCTURE ONLY (oracle/): builds and wraps oracle/restate.c (the plain-C restatement
of the reference's arithmetic).  Importable only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; the product package never imports this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "restate.c")
_LIB = os.path.join(_HERE, "librestate.so")
_lib = None


def build(force=False):
    """gcc, baseline x86-64, no FMA contraction: each float op rounds once, like the csim."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fopenmp", _SRC, "-o", _LIB, "-lm"],
                       check=True)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
        _lib.oracle_cholesky.argtypes = [fp, fp, ctypes.c_int, ctypes.c_long, ctypes.c_int, ctypes.c_int]
        _lib.oracle_lu.argtypes = [fp, fp, fp, ip, ctypes.c_int, ctypes.c_long, ctypes.c_int, ctypes.c_int]
        _lib.oracle_qr.argtypes = [fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_int, ctypes.c_int]
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _batched(A):
    A = np.ascontiguousarray(A, dtype=np.float32)
    return A.ndim == 2, A.reshape((-1,) + A.shape[-2:])


CHOL_VARIANT = {"cholesky_v1.3": 0, "cholesky_v3.2": 0, "cholesky_v4.0": 0, "cholesky_v2.2": 1}
QR_VARIANT = {"qr_v2.1": 0, "qr_v1.1": 1, "qr_v1.2": 1}


def cholesky(A, design="cholesky_v1.3", threads=1):
    single, A3 = _batched(A)
    L = np.empty_like(A3)
    rc = _load().oracle_cholesky(_fp(A3), _fp(L), A3.shape[-1], A3.shape[0], CHOL_VARIANT[design], threads)
    assert rc == 0
    return L[0] if single else L


def lu(A, pivoting=True, threads=1):
    single, A3 = _batched(A)
    n = A3.shape[-1]
    L, U = np.empty_like(A3), np.empty_like(A3)
    P = np.empty((A3.shape[0], n), dtype=np.int32)
    rc = _load().oracle_lu(_fp(A3), _fp(L), _fp(U), P.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), n, A3.shape[0],
                           int(bool(pivoting)), threads)
    assert rc == 0
    return (L[0], U[0], P[0]) if single else (L, U, P)


def qr(A, design="qr_v2.1", threads=1):
    single, A3 = _batched(A)
    rows, cols = A3.shape[-2:]
    Q = np.empty((A3.shape[0], rows, rows), dtype=np.float32)
    R = np.empty_like(A3)
    rc = _load().oracle_qr(_fp(A3), _fp(Q), _fp(R), rows, cols, A3.shape[0], QR_VARIANT[design], threads)
    if rc != 0:
        raise ValueError("oracle_qr: rows must be >= cols >= 1")
    return (Q[0], R[0]) if single else (Q, R)

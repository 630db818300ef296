"""Host-side mirror of the reference's top-level functions, batched.

    L       = cholesky(A)            <->  top(A, L)        Cholesky/*/model4x4/chol.h:15-16
    L, U, P = lu(A)                  <->  top(A, L, U, P)  LU/*/model4x4/lu.h:19-22
    Q, R    = qr(A)                  <->  top(A, Q, R)     QR/*/model4x4/qr.h:22-24

`A` is either a numpy array on the host (goes through the hlsd_*_host C entry points: copy in,
factor on the GPU, copy out; with `devices=[...]` through hlsd_*_host_multi, which splits the
batch over several GPUs) or a torch CUDA tensor (hlsd_*_dev on torch's current stream, outputs
are torch tensors on the same device).  Shapes: (..., n, n) / (..., rows, cols); leading
dimensions are flattened into the batch.  Everything is float32, row-major, as in the reference.

`exact=True` (HLSD_FLAG_EXACT) keeps the reference design's arithmetic at every size (no tensor
path); `fused=True` (HLSD_FLAG_FUSED) lets the register-resident kernels contract multiply-subtract
pairs into FMAs (not bit-identical; see include/hlsd.h).
"""
import ctypes

import numpy as np

from . import _lib

CHOL_DESIGNS = {"cholesky_v1.3": 13, "cholesky_v2.2": 22, "cholesky_v3.2": 32, "cholesky_v4.0": 40}
LU_DESIGNS = {"lu_v1.1": 11, "lu_v1.2": 12, "lu_v2.1": 21, "nopivot": 1}
QR_DESIGNS = {"qr_v1.1": 11, "qr_v1.2": 12, "qr_v2.1": 21}
PATHS = {"auto": 0, "tpm": 1, "group": 2, "global": 3, "blocked": 4, "tensor": 5}
FLAG_EXACT, FLAG_FUSED = 0x100, 0x200


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _design(table, design):
    if design not in table:
        raise ValueError(f"unknown design {design!r}; one of {sorted(table)}")
    return table[design]


def _path(path, exact=False, fused=False):
    if path not in PATHS:
        raise ValueError(f"unknown path {path!r}; one of {sorted(PATHS)}")
    return PATHS[path] | (FLAG_EXACT if exact else 0) | (FLAG_FUSED if fused else 0)


def _devices(devices):
    """-> (ctypes int array or None, count) for the *_host_multi entry points."""
    if devices is None:
        return None, 0
    if isinstance(devices, int):  # the first `devices` GPUs
        return None, int(devices)
    arr = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
    return arr, len(devices)


def _np_in(A, square):
    A = np.ascontiguousarray(A, dtype=np.float32)
    if A.ndim < 2:
        raise ValueError("need at least a 2-D array")
    if square and A.shape[-1] != A.shape[-2]:
        raise ValueError("matrix must be square")
    batch = int(np.prod(A.shape[:-2])) if A.ndim > 2 else 1
    return A, batch


def _torch_in(A, square):
    import torch
    if not A.is_cuda:
        raise ValueError("torch input must live on a CUDA device (use a numpy array for host data)")
    if A.dtype != torch.float32:
        raise ValueError("float32 only (the reference's matrix_t is float)")
    if A.dim() < 2 or (square and A.shape[-1] != A.shape[-2]):
        raise ValueError("bad shape")
    A = A.contiguous()
    batch = 1
    for d in A.shape[:-2]:
        batch *= int(d)
    return A, batch


def _stream(t):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def cholesky(A, design="cholesky_v1.3", path="auto", exact=False, fused=False, devices=None):
    lib, d, p = _lib.load(), _design(CHOL_DESIGNS, design), _path(path, exact, fused)
    if _is_torch(A):
        import torch
        A, batch = _torch_in(A, True)
        L = torch.empty_like(A)
        with torch.cuda.device(A.device):
            _lib.check(lib.hlsd_cholesky_dev(A.data_ptr(), L.data_ptr(), A.shape[-1], batch, d, p, _stream(A)))
        return L
    A, batch = _np_in(A, True)
    L = np.empty_like(A)
    if devices is not None:
        devs, nd = _devices(devices)
        _lib.check(lib.hlsd_cholesky_host_multi(A.ctypes.data, L.ctypes.data, A.shape[-1], batch, d, p, devs, nd))
    else:
        _lib.check(lib.hlsd_cholesky_host(A.ctypes.data, L.ctypes.data, A.shape[-1], batch, d, p))
    return L


def lu(A, design="lu_v1.1", path="auto", exact=False, fused=False, devices=None):
    lib, d, p = _lib.load(), _design(LU_DESIGNS, design), _path(path, exact, fused)
    if _is_torch(A):
        import torch
        A, batch = _torch_in(A, True)
        L, U = torch.empty_like(A), torch.empty_like(A)
        P = torch.empty(A.shape[:-1], dtype=torch.int32, device=A.device)
        with torch.cuda.device(A.device):
            _lib.check(lib.hlsd_lu_dev(A.data_ptr(), L.data_ptr(), U.data_ptr(), P.data_ptr(), A.shape[-1], batch, d, p,
                                       _stream(A)))
        return L, U, P
    A, batch = _np_in(A, True)
    L, U = np.empty_like(A), np.empty_like(A)
    P = np.empty(A.shape[:-1], dtype=np.int32)
    if devices is not None:
        devs, nd = _devices(devices)
        _lib.check(lib.hlsd_lu_host_multi(A.ctypes.data, L.ctypes.data, U.ctypes.data, P.ctypes.data, A.shape[-1], batch, d, p,
                                          devs, nd))
    else:
        _lib.check(lib.hlsd_lu_host(A.ctypes.data, L.ctypes.data, U.ctypes.data, P.ctypes.data, A.shape[-1], batch, d, p))
    return L, U, P


def qr(A, design="qr_v2.1", path="auto", exact=False, fused=False, devices=None):
    lib, d, p = _lib.load(), _design(QR_DESIGNS, design), _path(path, exact, fused)
    if _is_torch(A):
        import torch
        A, batch = _torch_in(A, False)
        rows, cols = int(A.shape[-2]), int(A.shape[-1])
        Q = torch.empty(A.shape[:-2] + (rows, rows), dtype=torch.float32, device=A.device)
        R = torch.empty_like(A)
        with torch.cuda.device(A.device):
            _lib.check(lib.hlsd_qr_dev(A.data_ptr(), Q.data_ptr(), R.data_ptr(), rows, cols, batch, d, p, _stream(A)))
        return Q, R
    A, batch = _np_in(A, False)
    rows, cols = A.shape[-2:]
    Q = np.empty(A.shape[:-2] + (rows, rows), dtype=np.float32)
    R = np.empty_like(A)
    if devices is not None:
        devs, nd = _devices(devices)
        _lib.check(lib.hlsd_qr_host_multi(A.ctypes.data, Q.ctypes.data, R.ctypes.data, rows, cols, batch, d, p, devs, nd))
    else:
        _lib.check(lib.hlsd_qr_host(A.ctypes.data, Q.ctypes.data, R.ctypes.data, rows, cols, batch, d, p))
    return Q, R


OPTIONS = {"chol_slab": 1, "diag_kernel": 3, "lockstep": 4, "qr_threads": 5, "chol_tile": 6}


def set_option(name, value):
    """Measurement switch (include/hlsd.h, HLSD_OPT_*): selects the alternative kernel where two serve a request."""
    _lib.check(_lib.load().hlsd_set_option(OPTIONS[name], int(value)))


def launch_count():
    return int(_lib.load().hlsd_launch_count())


def selftest_division(num, den):
    """(mismatches, fast_pairs, unguarded_mismatches) of the kernels' shared-divisor division against IEEE
    division over two equally long float32 CUDA tensors (diagnostic; see include/hlsd.h)."""
    import torch
    assert num.is_cuda and den.is_cuda and num.dtype == den.dtype == torch.float32 and num.numel() == den.numel()
    num, den = num.contiguous(), den.contiguous()
    out = torch.zeros(3, dtype=torch.int64, device=num.device)
    with torch.cuda.device(num.device):
        _lib.check(_lib.load().hlsd_selftest_division(num.data_ptr(), den.data_ptr(), num.numel(), out.data_ptr(), _stream(num)))
    bad, fast, unguarded = out.tolist()
    return bad, fast, unguarded

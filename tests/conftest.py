import os
import resource
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

# the reference's PE functions keep N x N float arrays on the stack (oracle/_ref only)
try:
    resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
except (ValueError, OSError):
    pass


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden_chol():
    return np.load(os.path.join(GOLDEN, "golden_chol.npz"))


@pytest.fixture(scope="session")
def golden_lu():
    return np.load(os.path.join(GOLDEN, "golden_lu.npz"))


@pytest.fixture(scope="session")
def golden_qr():
    return np.load(os.path.join(GOLDEN, "golden_qr.npz"))


@pytest.fixture(scope="session")
def golden_large():
    """Reference outputs on the 256 / 512 operand files (Cholesky/op256.bin ... LU/op512.bin)."""
    return np.load(os.path.join(GOLDEN, "golden_large.npz"))


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype == np.float32:
        return np.array_equal(a.view(np.uint32), b.view(np.uint32))
    return np.array_equal(a, b)


def first_mismatch(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    bad = np.argwhere(a.view(np.uint32) != b.view(np.uint32)) if a.dtype == np.float32 else np.argwhere(a != b)
    if len(bad) == 0:
        return "identical"
    idx = tuple(bad[0])
    return f"{len(bad)} of {a.size} differ; first at {idx}: {a[idx]!r} vs {b[idx]!r}"


# ---- float64 reconstruction helpers (residuals named by BASELINE.json's north_star) -------------
def chol_residual(A, L):
    A64, L64 = A.astype(np.float64), L.astype(np.float64)
    return np.linalg.norm(A64 - L64 @ np.swapaxes(L64, -1, -2), axis=(-2, -1)) / np.linalg.norm(A64, axis=(-2, -1))


def lu_reconstruct(L, U, P):
    """Undo the reference's LINPACK-style elimination: at step k rows k and P[k] were exchanged in
    the trailing columns only, then multiples of row k subtracted (lu_v1.1/model4x4/lu.cpp:44-91)."""
    L64, X = L.astype(np.float64), U.astype(np.float64).copy()
    n = X.shape[-1]
    for k in range(n - 2, -1, -1):
        X[k + 1:, :] += np.outer(L64[k + 1:, k], X[k, :])
        p = int(P[k])
        if p != k:
            X[[k, p], :] = X[[p, k], :]
    return X


def lu_residual(A, L, U, P):
    A64 = A.astype(np.float64)
    return np.linalg.norm(A64 - lu_reconstruct(L, U, P)) / np.linalg.norm(A64)


def qr_residual(A, Q, R):
    A64 = A.astype(np.float64)
    return np.linalg.norm(A64 - Q.astype(np.float64) @ R.astype(np.float64), axis=(-2, -1)) / np.linalg.norm(A64, axis=(-2, -1))


def lu_fp64_following(A, P):
    """The reference's elimination (lu_v1.1/model4x4/lu.cpp:43-91) in float64 with the pivot sequence P imposed: the
    yardstick for element-wise errors - how far fp32 round-off alone moves L and U for this operand and these pivots."""
    n = A.shape[-1]
    w = A.astype(np.float64).copy()
    L = np.eye(n)
    for k in range(n - 1):
        p = int(P[k])
        if p != k:
            w[[k, p], k:] = w[[p, k], k:]
        w[k + 1:, k] /= w[k, k]
        L[k + 1:, k] = w[k + 1:, k]
        w[k + 1:, k + 1:] -= np.outer(w[k + 1:, k], w[k, k + 1:])
    return L, np.triu(w)


def lu_elementwise_ok(A, L, U, P, wl, wu, tol=1e-3, slack=4.0):
    """Element-wise agreement of a tolerance-exact LU with the csim's (same pivots): within `tol` (max |dL|, max |dU| /
    max |U|), or - for the few ill-conditioned trailing elements where the csim's own fp32 round-off exceeds `tol` -
    within `slack` times the csim's own distance from the float64 factors."""
    el, eu = np.abs(L - wl).max(), np.abs(U - wu).max() / np.abs(wu).max()
    if el < tol and eu < tol:
        return True, (el, eu)
    L64, U64 = lu_fp64_following(A, P)
    ref_l, ref_u = np.abs(wl - L64).max(), np.abs(wu - U64).max() / np.abs(U64).max()
    got_l, got_u = np.abs(L - L64).max(), np.abs(U - U64).max() / np.abs(U64).max()
    return (got_l < max(tol, slack * ref_l) and got_u < max(tol, slack * ref_u)), (el, eu, got_l, ref_l, got_u, ref_u)

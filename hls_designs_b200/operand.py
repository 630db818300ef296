"""Operand files and operand generators of the reference.

File format (bit-for-bit the reference's): `op<N>.bin` / `op<M>x<N>.bin` is the raw little-endian
float32 image written by MATLAB `fwrite(fp, A, 'float')` (Cholesky/genA.m:9-12, LU/genA.m:5-8,
QR/genA.m:5-8), i.e. MATLAB's column-major A, and the testbenches read it back row by row with
`fread(A[i], sizeof(float), N, fp)` (e.g. LU/lu_v1.1/model4x4/lu_tb.cpp:24-30).  The matrix the
designs factor is therefore the file's floats taken as a row-major ROWS x COLS array (the
transpose of what MATLAB held, which is irrelevant for the symmetric Cholesky operands).
"""
import os
import re

import numpy as np


def operand_name(rows, cols=None, algo="chol"):
    """File name the reference uses: op<N>.bin for Cholesky/LU, op<M>x<N>.bin for QR."""
    if algo == "qr":
        return f"op{rows}x{cols if cols is not None else rows}.bin"
    return f"op{rows}.bin"


def read_operand(path, rows=None, cols=None):
    """Read an operand file as the row-major matrix the reference testbench sees."""
    data = np.fromfile(path, dtype="<f4")
    if rows is None:
        m = re.search(r"op(\d+)(?:x(\d+))?\.bin$", os.path.basename(path))
        if not m:
            raise ValueError(f"cannot infer the size from {path}")
        rows = int(m.group(1))
        cols = int(m.group(2)) if m.group(2) else rows
    cols = rows if cols is None else cols
    if data.size != rows * cols:
        raise ValueError(f"{path}: {data.size} floats, expected {rows}x{cols}")
    return data.reshape(rows, cols)


def write_operand(path, A):
    """Write a matrix so that read_operand / the reference testbench reads back exactly A."""
    np.ascontiguousarray(A, dtype="<f4").tofile(path)


def gen_spd(n, batch=None, seed=0):
    """Cholesky/genA.m:1-7: A = rand(n); A = 0.5*(A+A'); A = A + n*eye(n).  float32, batched."""
    rng = np.random.default_rng(seed)
    shape = (n, n) if batch is None else (batch, n, n)
    A = rng.random(shape, dtype=np.float32)
    A = 0.5 * (A + np.swapaxes(A, -1, -2))
    A = A + np.float32(n) * np.eye(n, dtype=np.float32)
    return np.ascontiguousarray(A, dtype=np.float32)


def gen_full_rank(rows, cols=None, batch=None, seed=0):
    """LU/genA.m:1-3 and QR/genA.m:1-3: A = randi(100, m, n).  float32, batched."""
    cols = rows if cols is None else cols
    rng = np.random.default_rng(seed)
    shape = (rows, cols) if batch is None else (batch, rows, cols)
    return rng.integers(1, 101, size=shape).astype(np.float32)

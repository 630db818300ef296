#!/usr/bin/env python3
"""Generates the golden fixtures of tests/golden/ from the REFERENCE ITSELF.

Run in the build container (where /root/reference exists) after `python oracle/build_ref.py`:

    python tests/golden/make_golden.py

For every operand file the reference ships up to 128x128 (Cholesky/op<N>.bin, LU/op<N>.bin,
QR/op<M>x<N>.bin) it copies the operand to tests/golden/operands/<Algo>_<name> (data, not source)
and records, in tests/golden/golden_<algo>.npz, the outputs of the reference's own `top()` for
every design built at that size (oracle/_ref/lib<design>_<size>.so = the reference's C++ compiled
with g++ and the HLS header stand-ins), plus outputs on a few seeded random operands.
Arrays are stored bit-exactly (float32 / int32).
"""
import os
import resource
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import ref  # noqa: E402
from hls_designs_b200.operand import gen_full_rank, gen_spd, read_operand  # noqa: E402

REF = os.environ.get("HLS_REFERENCE_ROOT", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))


def main():
    os.makedirs(os.path.join(GOLD, "operands"), exist_ok=True)
    chol, lu, qr = {}, {}, {}
    for n in (4, 8, 16, 25, 32, 64, 128):
        src = os.path.join(REF, "Cholesky", f"op{n}.bin")
        shutil.copyfile(src, os.path.join(GOLD, "operands", f"Cholesky_op{n}.bin"))
        A = np.stack([read_operand(src)] + list(gen_spd(n, batch=2, seed=100 + n)))
        chol[f"A_{n}"] = A
        for d in ref.CHOL_DESIGNS:
            if ref.available(d, n) and not (d == "cholesky_v4.0" and n == 25):  # v4.0 @25: ap_int<5> id overflows
                chol[f"L_{d}_{n}"] = ref.ref_cholesky(A, d)
    for n in (4, 6, 8, 16, 32, 64, 128):
        src = os.path.join(REF, "LU", f"op{n}.bin")
        shutil.copyfile(src, os.path.join(GOLD, "operands", f"LU_op{n}.bin"))
        A = np.stack([read_operand(src)] + list(gen_full_rank(n, batch=2, seed=200 + n)))
        lu[f"A_{n}"] = A
        for d in ref.LU_DESIGNS:
            if ref.available(d, n):
                L, U, P = ref.ref_lu(A, d)
                lu[f"L_{d}_{n}"], lu[f"U_{d}_{n}"], lu[f"P_{d}_{n}"] = L, U, P
    for rows, cols in ((4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (128, 128), (8, 4), (16, 8), (12, 5)):
        tag = f"{rows}x{cols}"
        src = os.path.join(REF, "QR", f"op{tag}.bin")
        mats = list(gen_full_rank(rows, cols, batch=2, seed=300 + rows + cols))
        if os.path.exists(src):
            shutil.copyfile(src, os.path.join(GOLD, "operands", f"QR_op{tag}.bin"))
            mats = [read_operand(src)] + mats
        A = np.stack(mats)
        qr[f"A_{tag}"] = A
        for d in ref.QR_DESIGNS:
            if ref.available(d, tag):
                Q, R = ref.ref_qr(A, d)
                qr[f"Q_{d}_{tag}"], qr[f"R_{d}_{tag}"] = Q, R
    np.savez_compressed(os.path.join(GOLD, "golden_chol.npz"), **chol)
    np.savez_compressed(os.path.join(GOLD, "golden_lu.npz"), **lu)
    np.savez_compressed(os.path.join(GOLD, "golden_qr.npz"), **qr)
    for name, d in (("chol", chol), ("lu", lu), ("qr", qr)):
        print(name, len(d), "arrays")


if __name__ == "__main__":
    main()

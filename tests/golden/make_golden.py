#!/usr/bin/env python3
"""Generates the golden fixtures of tests/golden/ from the REFERENCE ITSELF.

Run in the build container (where /root/reference exists) after `python oracle/build_ref.py`:

    python tests/golden/make_golden.py

For every operand file the reference ships (Cholesky/op<N>.bin, LU/op<N>.bin, QR/op<M>x<N>.bin) it
copies the operand to tests/golden/operands/<Algo>_<name> (data, not source) and records, in
tests/golden/golden_<algo>.npz, the outputs of the reference's own `top()` for every design built
at that size (oracle/_ref/lib<design>_<size>.so = the reference's C++ compiled with g++ and the HLS
header stand-ins), plus outputs on a few seeded random operands.  The 256 / 512 operands
(Cholesky/op256.bin, op512.bin, LU/op256.bin, op512.bin; 1-D designs only, see oracle/build_ref.py)
go to golden_large.npz: one full output per distinct arithmetic, a SHA-256 of the output bytes for
the designs that repeat it (checked equal here).  Arrays are stored bit-exactly (float32 / int32).
"""
import hashlib
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
    large = {}
    for n in (256, 512):
        src = os.path.join(REF, "Cholesky", f"op{n}.bin")
        shutil.copyfile(src, os.path.join(GOLD, "operands", f"Cholesky_op{n}.bin"))
        A = read_operand(src)[None]
        full = {}
        for d in ref.CHOL_DESIGNS:
            if ref.available(d, n):
                L = ref.ref_cholesky(A, d)
                arith = "cholesky_v2.2" if d == "cholesky_v2.2" else "cholesky_v1.3"
                if d == arith:
                    full[d] = L
                    large[f"chol_L_{d}_{n}"] = L
                else:
                    assert np.array_equal(L.view(np.uint32), full[arith].view(np.uint32)), (d, n)
                large[f"chol_sha_{d}_{n}"] = np.array(hashlib.sha256(L.tobytes()).hexdigest())
        src = os.path.join(REF, "LU", f"op{n}.bin")
        shutil.copyfile(src, os.path.join(GOLD, "operands", f"LU_op{n}.bin"))
        B = read_operand(src)[None]
        first = None
        for d in ref.LU_DESIGNS:
            if ref.available(d, n):
                L, U, P = ref.ref_lu(B, d)
                if first is None:
                    first = (L, U, P)
                    large[f"lu_L_{n}"], large[f"lu_U_{n}"], large[f"lu_P_{n}"] = L, U, P
                else:
                    assert all(np.array_equal(x.view(np.uint32), y.view(np.uint32)) for x, y in zip((L, U, P), first)), (d, n)
                large[f"lu_sha_{d}_{n}"] = np.array(hashlib.sha256(L.tobytes() + U.tobytes() + P.tobytes()).hexdigest())
    np.savez_compressed(os.path.join(GOLD, "golden_large.npz"), **large)
    print("large", len(large), "arrays")
    np.savez_compressed(os.path.join(GOLD, "golden_chol.npz"), **chol)
    np.savez_compressed(os.path.join(GOLD, "golden_lu.npz"), **lu)
    np.savez_compressed(os.path.join(GOLD, "golden_qr.npz"), **qr)
    for name, d in (("chol", chol), ("lu", lu), ("qr", qr)):
        print(name, len(d), "arrays")


if __name__ == "__main__":
    main()

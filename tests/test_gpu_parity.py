"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libhlsd.so (host-pointer entry points with numpy buffers, device-pointer entry points with torch
tensors) and is compared BIT FOR BIT with the oracle (oracle/restate.c, itself pinned against the
reference's own code) and with the golden outputs recorded from the reference."""
import numpy as np
import pytest

import hls_designs_b200 as hd
import restate
from conftest import bits_equal, chol_residual, first_mismatch, lu_residual, qr_residual
from hls_designs_b200.operand import gen_full_rank, gen_spd

pytestmark = pytest.mark.gpu

CHOL_DESIGNS = ("cholesky_v1.3", "cholesky_v2.2", "cholesky_v3.2", "cholesky_v4.0")
LU_DESIGNS = ("lu_v1.1", "lu_v1.2", "lu_v2.1")
QR_DESIGNS = ("qr_v1.1", "qr_v1.2", "qr_v2.1")


def _paths_for(n, kind):
    paths = ["auto", "group", "global"]
    if (kind == "chol" and n in (4, 8, 16, 32, 64)) or (kind == "lu" and n in (4, 8, 16, 32, 64)) or (kind == "qr" and n in (4, 8, 16, 32)):
        paths.append("tpm")
    return paths


# ---------------------------------------------------------------- golden (reference csim outputs)
@pytest.mark.parametrize("n", (4, 8, 16, 25, 32, 64, 128))
def test_cholesky_golden(golden_chol, n):
    A = golden_chol[f"A_{n}"]
    for d in CHOL_DESIGNS:
        key = f"L_{d}_{n}"
        if key not in golden_chol:
            continue
        paths = _paths_for(n, "chol")
        if n == 128 and d != "cholesky_v2.2":
            paths.append("tpm")  # the register-tiled 128x128 kernel serves the right-looking designs only
        for path in paths:
            L = hd.cholesky(A, design=d, path=path)
            assert bits_equal(L, golden_chol[key]), f"{d} n={n} path={path}: {first_mismatch(L, golden_chol[key])}"


@pytest.mark.parametrize("n", (4, 6, 8, 16, 32, 64, 128))
def test_lu_golden(golden_lu, n):
    A = golden_lu[f"A_{n}"]
    for d in LU_DESIGNS:
        if f"L_{d}_{n}" not in golden_lu:
            continue
        for path in _paths_for(n, "lu"):
            L, U, P = hd.lu(A, design=d, path=path)
            assert bits_equal(L, golden_lu[f"L_{d}_{n}"]), f"L {d} n={n} {path}: {first_mismatch(L, golden_lu[f'L_{d}_{n}'])}"
            assert bits_equal(U, golden_lu[f"U_{d}_{n}"]), f"U {d} n={n} {path}: {first_mismatch(U, golden_lu[f'U_{d}_{n}'])}"
            assert np.array_equal(P, golden_lu[f"P_{d}_{n}"]), f"P {d} n={n} {path}"


@pytest.mark.parametrize("shape", ((4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (128, 128), (8, 4), (16, 8), (12, 5)))
def test_qr_golden(golden_qr, shape):
    tag = f"{shape[0]}x{shape[1]}"
    A = golden_qr[f"A_{tag}"]
    for d in QR_DESIGNS:
        if f"Q_{d}_{tag}" not in golden_qr:
            continue
        for path in _paths_for(shape[0] if shape[0] == shape[1] else 0, "qr"):
            Q, R = hd.qr(A, design=d, path=path)
            assert bits_equal(Q, golden_qr[f"Q_{d}_{tag}"]), f"Q {d} {tag} {path}: {first_mismatch(Q, golden_qr[f'Q_{d}_{tag}'])}"
            assert bits_equal(R, golden_qr[f"R_{d}_{tag}"]), f"R {d} {tag} {path}: {first_mismatch(R, golden_qr[f'R_{d}_{tag}'])}"


# ---------------------------------------------------------------- the 256 / 512 operands the reference ships
@pytest.mark.parametrize("n", (256, 512))
def test_reference_large_operands_exact(golden_large, n):
    """Cholesky/op256.bin, op512.bin, LU/op256.bin, op512.bin (read as T_chol_tb.cpp:19-24 reads them): the exact
    kernels - forced (`path="global"`) and as selected by HLSD_FLAG_EXACT - reproduce the outputs of the reference's
    own 1-D designs bit for bit."""
    import os
    from conftest import ROOT
    from hls_designs_b200.operand import read_operand
    A = read_operand(os.path.join(ROOT, "tests", "golden", "operands", f"Cholesky_op{n}.bin"))[None]
    for d, key in (("cholesky_v1.3", "cholesky_v1.3"), ("cholesky_v4.0", "cholesky_v1.3"), ("cholesky_v3.2", "cholesky_v1.3"),
                   ("cholesky_v2.2", "cholesky_v2.2")):
        want = golden_large[f"chol_L_{key}_{n}"]
        # global = one CTA per matrix in global memory; exact=True = HLSD_FLAG_EXACT, which selects the blocked exact
        # path (HLSD_PATH_BLOCKED) for the right-looking designs and the global kernel for cholesky_v2.2
        for kw in (dict(path="global"), dict(exact=True)) + ((dict(path="blocked"),) if key == "cholesky_v1.3" else ()):
            L = hd.cholesky(A, design=d, **kw)
            assert bits_equal(L, want), f"{d} n={n} {kw}: {first_mismatch(L, want)}"
    B = read_operand(os.path.join(ROOT, "tests", "golden", "operands", f"LU_op{n}.bin"))[None]
    for d in LU_DESIGNS:
        for kw in (dict(path="global"), dict(exact=True), dict(path="blocked")):  # exact=True selects the blocked exact LU
            L, U, P = hd.lu(B, design=d, **kw)
            assert bits_equal(L, golden_large[f"lu_L_{n}"]), f"L {d} n={n} {kw}: {first_mismatch(L, golden_large[f'lu_L_{n}'])}"
            assert bits_equal(U, golden_large[f"lu_U_{n}"]), f"U {d} n={n} {kw}: {first_mismatch(U, golden_large[f'lu_U_{n}'])}"
            assert np.array_equal(P, golden_large[f"lu_P_{n}"])


@pytest.mark.parametrize("n", (256, 512))
def test_reference_large_operands_tensor_path(golden_large, n):
    """The same operand files through the tcgen05 paths (what HLSD_PATH_AUTO picks at these sizes), against the
    reference's own outputs within the documented tolerances: Cholesky element-wise 1e-5 * max|L| and residual
    < 2e-6; LU the SAME pivot sequence as the csim, element-wise 1e-3 (L absolute, U relative to max|U|),
    residual < 5e-5."""
    import os
    from conftest import ROOT
    from hls_designs_b200.operand import read_operand
    A = read_operand(os.path.join(ROOT, "tests", "golden", "operands", f"Cholesky_op{n}.bin"))[None]
    want = golden_large[f"chol_L_cholesky_v1.3_{n}"]
    for kw in (dict(path="tensor"), dict()):
        L = hd.cholesky(A, **kw)
        assert np.abs(np.triu(L, 1)).max() == 0.0
        assert np.abs(L - want).max() / np.abs(want).max() < 1e-5
        assert chol_residual(A, L).max() < 2e-6
    B = read_operand(os.path.join(ROOT, "tests", "golden", "operands", f"LU_op{n}.bin"))[None]
    wl, wu, wp = golden_large[f"lu_L_{n}"], golden_large[f"lu_U_{n}"], golden_large[f"lu_P_{n}"]
    for kw in (dict(path="tensor"), dict()):
        L, U, P = hd.lu(B, **kw)
        assert np.array_equal(P, wp), f"n={n}: pivot sequence differs from the csim at {np.argwhere(P != wp)[:4].tolist()}"
        assert np.abs(L - wl).max() < 1e-3 and np.abs(U - wu).max() / np.abs(wu).max() < 1e-3
        assert lu_residual(B[0], L[0], U[0], P[0]) < 5e-5


def test_library_vs_handwritten_model4x4():
    """The reference's hand-written 4x4 models (<design>/model4x4/*.cpp, compiled as oracle/_ref/lib*_model4x4.so) on
    the reference's op4.bin / op4x4.bin and random operands, against the library, every design, bit for bit.
    (qr_v2.1's model carries a hand edit the template does not have; the library follows the template, see
    tests/test_oracle.py::test_restate_matches_handwritten_model4x4.)"""
    import os
    import ref
    from conftest import ROOT
    from hls_designs_b200.operand import read_operand
    if not ref.available("cholesky_v1.3", "model4x4"):
        pytest.skip("oracle/_ref not built")
    ops = os.path.join(ROOT, "tests", "golden", "operands")
    A = np.stack([read_operand(os.path.join(ops, "Cholesky_op4.bin"))] + list(gen_spd(4, batch=40, seed=4)))
    for d in CHOL_DESIGNS:
        want = ref.ref_cholesky(A, d, tag="model4x4")
        for path in ("auto", "group", "global"):
            assert bits_equal(hd.cholesky(A, design=d, path=path), want), (d, path)
    B = np.stack([read_operand(os.path.join(ops, "LU_op4.bin"))] + list(gen_full_rank(4, batch=40, seed=5)))
    for d in LU_DESIGNS:
        want = ref.ref_lu(B, d, tag="model4x4")
        for path in ("auto", "group", "global"):
            for g, w in zip(hd.lu(B, design=d, path=path), want):
                assert bits_equal(g, w), (d, path)
    C = np.stack([read_operand(os.path.join(ops, "QR_op4x4.bin"))] + list(gen_full_rank(4, 4, batch=40, seed=6)))
    for d in ("qr_v1.1", "qr_v1.2"):
        want = ref.ref_qr(C, d, tag="model4x4")
        for path in ("auto", "group", "global"):
            for g, w in zip(hd.qr(C, design=d, path=path), want):
                assert bits_equal(g, w), (d, path)
    Qt, Rt = ref.ref_qr(C, "qr_v2.1")  # template expansion at 4x4
    for path in ("auto", "group", "global"):
        Q, R = hd.qr(C, design="qr_v2.1", path=path)
        assert bits_equal(Q, Qt) and bits_equal(R, Rt), path


@pytest.mark.parametrize("n,batch", ((256, 9), (384, 5), (512, 4), (640, 2), (1024, 3)))
def test_blocked_exact_cholesky_vs_oracle(n, batch):
    """HLSD_PATH_BLOCKED: 128-column panels, SIMT trailing update in the reference's per-element order: bit-identical
    to the oracle (and to the one-CTA-per-matrix kernel) on SPD batches, on a matrix that is not positive definite
    (NaN pattern included) and through the device-pointer entry point."""
    import torch
    A = gen_spd(n, batch=batch, seed=7000 + n)
    A[batch - 1, n // 2, n // 2] = -1.0  # not positive definite from here on: NaNs must appear where the csim has them
    want = restate.cholesky(A, threads=8)
    got = hd.cholesky(A, path="blocked")
    assert bits_equal(got[:batch - 1], want[:batch - 1]), f"n={n}: {first_mismatch(got[:batch - 1], want[:batch - 1])}"
    # the indefinite matrix: NaNs in the same places (their payload bits are the hardware's), same bits elsewhere
    gn, wn = np.isnan(got[batch - 1]), np.isnan(want[batch - 1])
    assert wn.any() and np.array_equal(gn, wn)
    assert bits_equal(np.where(wn, 0.0, got[batch - 1]).astype(np.float32), np.where(wn, 0.0, want[batch - 1]).astype(np.float32))
    assert bits_equal(hd.cholesky(A, exact=True)[:batch - 1], want[:batch - 1])
    assert bits_equal(hd.cholesky(torch.from_numpy(A).cuda(), path="blocked").cpu().numpy()[:batch - 1], want[:batch - 1])
    assert bits_equal(hd.cholesky(A, path="global")[:batch - 1], got[:batch - 1])
    assert chol_residual(A[:batch - 1], got[:batch - 1]).max() < 2e-6
    with pytest.raises(hd.HlsdError):
        hd.cholesky(gen_spd(200), path="blocked")


@pytest.mark.parametrize("n,batch", ((256, 5), (384, 3), (512, 2), (1024, 1)))
def test_blocked_exact_cholesky_v22_vs_oracle(n, batch):
    """cholesky_v2.2 (root-free elimination, roots at the end) at n >= 256: the blocked exact path keeps the un-scaled
    values and the multipliers apart and applies every update in ascending k: bit-identical to the oracle and to the
    one-CTA-per-matrix kernel; auto dispatch takes it; the tensor path refuses the design."""
    A = gen_spd(n, batch=batch, seed=7100 + n)
    want = restate.cholesky(A, "cholesky_v2.2", threads=8)
    got = hd.cholesky(A, design="cholesky_v2.2", path="blocked")
    assert bits_equal(got, want), f"n={n}: {first_mismatch(got, want)}"
    assert bits_equal(hd.cholesky(A, design="cholesky_v2.2"), want)
    assert bits_equal(hd.cholesky(A[:1], design="cholesky_v2.2", path="global"), want[:1])
    assert np.abs(np.triu(got, 1)).max() == 0.0
    with pytest.raises(hd.HlsdError):
        hd.cholesky(A, design="cholesky_v2.2", path="tensor")


@pytest.mark.parametrize("n,batch", ((256, 7), (384, 4), (512, 3), (1024, 2)))
def test_blocked_exact_lu_vs_oracle(n, batch):
    """HLSD_PATH_BLOCKED for LU: pivoted panel kernel in exact arithmetic, LAPACK-style exchanges inside the panel undone
    afterwards, U12 by forward substitution and the trailing update in the reference's per-element order: L, U and P
    bit-identical to the oracle on integer (genA) and real operands, with pivot ties and a singular matrix."""
    import torch
    rng = np.random.default_rng(n)
    A = gen_full_rank(n, batch=batch, seed=8000 + n)
    B = (rng.standard_normal((batch, n, n)) * 3).astype(np.float32)
    B[0, :, 5] = B[0, :, 3]          # two equal columns: an exactly singular matrix (zero pivot -> inf / nan like the csim)
    A[1, 7, :] = A[1, 3, :]          # two equal rows: pivot ties
    for M in (A, B):
        want = restate.lu(M, threads=8)
        for kw in (dict(path="blocked"), dict(exact=True)):
            got = hd.lu(M, **kw)
            for name, g, w in zip("LUP", got, want):
                if g.dtype == np.float32:
                    gn, wn = np.isnan(g), np.isnan(w)
                    assert np.array_equal(gn, wn), f"{name} n={n} {kw}: NaN pattern differs"
                    g, w = np.where(wn, 0.0, g).astype(np.float32), np.where(wn, 0.0, w).astype(np.float32)
                assert bits_equal(g, w), f"{name} n={n} {kw}: {first_mismatch(g, w)}"
    got_d = hd.lu(torch.from_numpy(A).cuda(), path="blocked")
    for g, w in zip(got_d, restate.lu(A, threads=8)):
        assert bits_equal(g.cpu().numpy(), w)
    # HLSD_LU_NOPIVOT (this library's unpivoted variant, P[k] = k) through the same blocked exact path: bit-identical to
    # the oracle's unpivoted elimination (a diagonally dominant operand keeps it finite)
    D = (A[:2] + np.float32(200 * A.shape[-1]) * np.eye(A.shape[-1], dtype=np.float32)).astype(np.float32)
    for kw in (dict(path="blocked"), dict(exact=True)):
        for g, w in zip(hd.lu(D, design="nopivot", **kw), restate.lu(D, pivoting=False, threads=8)):
            assert bits_equal(g, w), f"nopivot {kw}: {first_mismatch(g, w)}"


# ---------------------------------------------------------------- HLSD_FLAG_FUSED (opt-in contraction)
@pytest.mark.parametrize("n", (4, 8, 16, 32, 64, 128))
def test_fused_flag_within_tolerance(n):
    """HLSD_FLAG_FUSED: same operations in the same order with a - b*c (and the Givens pair) contracted into FMAs.
    Not bit-identical; checked like the tensor path: residuals within the bounds the exact kernels meet, element-wise
    agreement with the oracle 1e-5 (Cholesky, relative to max|L|), 2e-3 (LU where the pivot sequences agree, which they
    must for the integer-valued genA operands at these sizes... see assertion), 1e-4 (QR, absolute on Q, relative on R)."""
    batch = 64
    A = gen_spd(n, batch=batch, seed=900 + n)
    for d in ("cholesky_v1.3", "cholesky_v2.2"):
        L = hd.cholesky(A, design=d, fused=True)
        want = restate.cholesky(A, d, threads=8)
        assert np.abs(np.triu(L, 1)).max() == 0.0
        assert chol_residual(A, L).max() < 2e-6
        assert np.abs(L - want).max() / np.abs(want).max() < 1e-5
    if n <= 64:
        B = gen_full_rank(n, batch=batch, seed=901 + n)
        L, U, P = hd.lu(B, fused=True)
        wl, wu, wp = restate.lu(B, threads=8)
        same = [np.array_equal(P[b], wp[b]) for b in range(batch)]
        assert sum(same) >= batch - 2, f"n={n}: {batch - sum(same)} of {batch} pivot sequences differ from the oracle"
        for b in range(batch):
            assert lu_residual(B[b], L[b], U[b], P[b]) < 5e-5
            if same[b]:
                assert np.abs(L[b] - wl[b]).max() < 2e-3 and np.abs(U[b] - wu[b]).max() / np.abs(wu[b]).max() < 2e-3
    Q, R = hd.qr(B if n <= 64 else gen_full_rank(n, batch=8, seed=7), fused=True)
    Bq = B if n <= 64 else gen_full_rank(n, batch=8, seed=7)
    wq, wr = restate.qr(Bq, threads=8)
    assert qr_residual(Bq, Q, R).max() < 1e-5
    assert np.abs(np.tril(R, -1)).max() == 0.0
    assert np.abs(Q - wq).max() < 1e-4 and np.abs(R - wr).max() / np.abs(wr).max() < 1e-4
    # the flag is a permission: kernels without a fused form ignore it and stay bit-identical
    A5 = gen_spd(5, batch=9, seed=5)
    assert bits_equal(hd.cholesky(A5, fused=True), restate.cholesky(A5))
    # and the default stays exact
    assert bits_equal(hd.cholesky(A), restate.cholesky(A, threads=8))


# ---------------------------------------------------------------- batch split over a device list
def test_multi_device_entry_points_match_single_device():
    """hlsd_*_host_multi splits one host batch into contiguous spans, one worker thread per listed device.  With a
    single GPU the device may be listed several times (several workers on one device): the spans, the offsets of every
    argument (A, L, U, P / Q, R) and the error path are the same as on eight GPUs."""
    import torch
    ndev = torch.cuda.device_count()
    lists = [[0], [0, 0, 0], list(range(ndev)), None]
    A = gen_spd(16, batch=1003, seed=77)
    want = hd.cholesky(A)
    for devs in lists:
        assert bits_equal(hd.cholesky(A, devices=devs if devs is not None else ndev), want), devs
    B = gen_full_rank(32, batch=101, seed=78)
    want = hd.lu(B)
    for devs in lists[:3]:
        for g, w in zip(hd.lu(B, devices=devs), want):
            assert bits_equal(g, w), devs
    C = gen_full_rank(24, 16, batch=37, seed=79)
    want = hd.qr(C)
    for devs in lists[:3]:
        for g, w in zip(hd.qr(C, devices=devs), want):
            assert bits_equal(g, w), devs
    # fewer matrices than devices, empty batch, bad device
    assert bits_equal(hd.cholesky(A[:2], devices=[0, 0, 0]), want_small := hd.cholesky(A[:2]))
    assert hd.cholesky(A[:0], devices=[0]).shape == (0, 16, 16)
    with pytest.raises(hd.HlsdError):
        hd.cholesky(A, devices=[ndev + 7])


def test_staging_chunks_for_large_matrices():
    """Round-1 advisor finding: the staging chunk was rounded UP to 32 matrices after the 64 MB cap (384 MB of staging
    for one 1024x1024 LU).  A batch of 1024x1024 matrices that needs several chunks of fewer than 32 matrices: same
    answer as the device-resident call, and the staging workspace stays near the cap."""
    import torch
    n, batch = 1024, 7
    B = gen_full_rank(n, batch=batch, seed=31)
    free0 = torch.cuda.mem_get_info()[0]
    got = hd.lu(B, exact=False)
    used = free0 - torch.cuda.mem_get_info()[0]
    assert used < (640 << 20), f"staging + workspaces hold {used >> 20} MiB"
    want = hd.lu(torch.from_numpy(B).cuda())
    for g, w in zip(got, want):
        assert bits_equal(g, w.cpu().numpy())
    hd.load().hlsd_release_workspace()


def test_copy_probe_runs():
    lib = hd.load()
    a = np.zeros(1 << 22, np.uint8)
    b = np.zeros(1 << 21, np.uint8)
    assert lib.hlsd_selftest_copy_host(a.ctypes.data, b.ctypes.data, 4096, 2048, 1024) == 0


# ---------------------------------------------------------------- seeded batches vs the oracle
@pytest.mark.parametrize("n,batch", ((4, 1000), (8, 777), (16, 4099), (16, 31), (5, 70), (25, 40), (32, 300), (48, 64),
                                     (64, 200), (100, 9), (128, 20), (200, 3)))
def test_cholesky_batches_vs_oracle(n, batch):
    A = gen_spd(n, batch=batch, seed=n * 1000 + batch)
    for d in ("cholesky_v1.3", "cholesky_v2.2"):
        want = restate.cholesky(A, d, threads=8)
        got = hd.cholesky(A, design=d)
        assert bits_equal(got, want), f"{d} n={n}: {first_mismatch(got, want)}"
    assert chol_residual(A, got).max() < 2e-6


@pytest.mark.parametrize("n,batch", ((4, 1000), (6, 50), (8, 777), (16, 500), (17, 33), (32, 300), (64, 130), (96, 10),
                                     (128, 12), (230, 3)))
def test_lu_batches_vs_oracle(n, batch):
    rng = np.random.default_rng(n)
    for A in (gen_full_rank(n, batch=batch, seed=n + batch),
              (rng.standard_normal((batch, n, n)) * 2).astype(np.float32)):
        for design, piv in (("lu_v1.1", True), ("nopivot", False)):
            want = restate.lu(A, pivoting=piv, threads=8)
            got = hd.lu(A, design=design)
            for name, g, w in zip("LUP", got, want):
                assert bits_equal(g, w), f"{name} {design} n={n}: {first_mismatch(g, w)}"
    L, U, P = hd.lu(A[:2], design="lu_v2.1")
    assert lu_residual(A[0], L[0], U[0], P[0]) < 5e-5


@pytest.mark.parametrize("rows,cols,batch", ((4, 4, 500), (8, 8, 300), (8, 4, 100), (12, 5, 64), (16, 16, 200),
                                             (32, 32, 100), (33, 20, 17), (64, 64, 40), (128, 128, 6), (128, 64, 5),
                                             (160, 160, 2)))
def test_qr_batches_vs_oracle(rows, cols, batch):
    A = gen_full_rank(rows, cols, batch=batch, seed=rows * cols + batch)
    A[0, rows - 1, 0] = 0.0  # exercise the "already zero" branch (c = 1, s = 0)
    for d in ("qr_v2.1", "qr_v1.1"):
        want = restate.qr(A, d, threads=8)
        got = hd.qr(A, design=d)
        for name, g, w in zip("QR", got, want):
            assert bits_equal(g, w), f"{name} {d} {rows}x{cols}: {first_mismatch(g, w)}"
    Q, R = hd.qr(A, design="qr_v2.1")
    assert qr_residual(A, Q, R).max() < 1e-5
    assert np.abs(np.tril(R, -1)).max() == 0.0


# ---------------------------------------------------------------- device-pointer entry points
def test_device_entry_points_match_host_entry_points():
    import torch
    A = gen_spd(16, batch=1025, seed=1)
    Ld = hd.cholesky(torch.from_numpy(A).cuda())
    torch.cuda.synchronize()
    assert bits_equal(Ld.cpu().numpy(), hd.cholesky(A))
    B = gen_full_rank(32, batch=65, seed=2)
    out_d = hd.lu(torch.from_numpy(B).cuda())
    out_h = hd.lu(B)
    for g, w in zip(out_d, out_h):
        assert bits_equal(g.cpu().numpy(), w)
    C = gen_full_rank(24, 16, batch=33, seed=3)
    Qd, Rd = hd.qr(torch.from_numpy(C).cuda())
    Qh, Rh = hd.qr(C)
    assert bits_equal(Qd.cpu().numpy(), Qh) and bits_equal(Rd.cpu().numpy(), Rh)
    assert hd.launch_count() > 0


# ---------------------------------------------------------------- edges
def test_edge_cases():
    assert hd.cholesky(np.zeros((0, 16, 16), np.float32)).shape == (0, 16, 16)
    assert hd.cholesky(np.array([[4.0]], np.float32))[0, 0] == 2.0
    L, U, P = hd.lu(np.array([[0.0, 1.0], [2.0, 3.0]], np.float32))
    assert list(P) == [1, 1] and U[0, 0] == 2.0
    # not positive definite: NaNs exactly where the reference's csim produces them
    A = gen_spd(16, batch=40, seed=5)
    A[3, 5, 5] = -1.0
    for path in ("tpm", "group", "global"):
        got = hd.cholesky(A, path=path)
        want = restate.cholesky(A)
        assert bits_equal(np.isnan(got), np.isnan(want)) and np.isnan(got[3]).any()
        ok = ~np.isnan(want)
        assert bits_equal(got[ok], want[ok])
    # singular / tie pivots
    B = gen_full_rank(8, batch=64, seed=6)
    B[1, :, 0] = 0.0
    B[2, 3, :] = B[2, 2, :]
    got, want = hd.lu(B), restate.lu(B)
    assert np.array_equal(got[2], want[2])
    with pytest.raises(hd.HlsdError):
        hd.cholesky(gen_spd(48), path="tpm")  # forced path that cannot serve the size
    with pytest.raises(hd.HlsdError):
        hd.qr(np.ones((3, 5), np.float32))


def _special_batch(kind, n, rng):
    """Matrices that leave the fast paths of the kernels' square root / division (zeros, signed zeros,
    subnormals, huge and tiny magnitudes, infinities, NaNs) next to ordinary ones."""
    batch = 64
    A = gen_spd(n, batch=batch, seed=n) if kind == "chol" else gen_full_rank(n, n, batch=batch, seed=n)
    eye = np.eye(n, dtype=np.float32)
    A[0] = eye                                              # zeros everywhere off the diagonal
    A[1] = np.diag(np.arange(1, n + 1, dtype=np.float32))
    A[2] = 4 * eye + np.eye(n, k=1, dtype=np.float32) + np.eye(n, k=-1, dtype=np.float32)  # banded
    i, j = rng.integers(0, n, 2 * n).reshape(2, n)
    A[3][i, j] = 0.0
    A[3][j, i] = 0.0
    A[4] *= np.float32(1e-25)
    A[5] *= np.float32(1e25)
    A[6] *= np.float32(1e-37)                               # subnormal intermediates
    A[7][i, j] = -0.0
    A[8][n // 2, n // 3] = np.nan
    A[9][n - 1, 0] = np.inf
    A[10] = 0.0
    A[11][:, 0] = 0.0                                       # zero pivot column
    A[12] *= np.float32(2.0 ** 60)                          # the guard's upper bound
    A[13] *= np.float32(2.0 ** -60)
    A[14] = -A[14]
    A[15][n - 1] = A[15][0]                                 # two equal rows: exact zeros appear during elimination
    return A


@pytest.mark.parametrize("kind", ("chol", "lu", "qr"))
def test_special_values_bit_exact(kind):
    """Zeros, signed zeros, subnormals, extreme magnitudes, infinities and NaNs through every fast kernel size:
    NaNs where the csim has NaNs, every other bit identical (the shared-reciprocal division and the
    'scratch lanes divide the pivot by itself' tricks of the kernels must not change any value)."""
    rng = np.random.default_rng(77)
    with np.errstate(all="ignore"):
        for n in (4, 8, 16, 32, 64):
            A = _special_batch(kind, n, rng)
            if kind == "chol":
                designs = ("cholesky_v1.3", "cholesky_v2.2")
                run = lambda d: ([hd.cholesky(A, design=d)], [restate.cholesky(A, d, threads=8)])
            elif kind == "lu":
                designs = ("lu_v1.1", "nopivot")
                run = lambda d: (hd.lu(A, design=d), restate.lu(A, pivoting=d != "nopivot", threads=8))
            else:
                designs = ("qr_v2.1", "qr_v1.1")
                run = lambda d: (hd.qr(A, design=d), restate.qr(A, d, threads=8))
            for d in designs:
                got, want = run(d)
                for g, w in zip(got, want):
                    if g.dtype.kind == "f":
                        gn, wn = np.isnan(g), np.isnan(w)
                        bad = np.argwhere(gn != wn)
                        assert bad.size == 0, f"{kind} n={n} {d}: NaN pattern differs first at {bad[0]}"
                        ok = ~wn
                        assert bits_equal(g[ok], w[ok]), f"{kind} n={n} {d}: {first_mismatch(np.where(ok, g, 0), np.where(ok, w, 0))}"
                    else:
                        assert np.array_equal(g, w), f"{kind} n={n} {d}: pivots differ"


# ---------------------------------------------------------------- BASELINE.json full sizes, by property
def test_full_size_cholesky_n16_batch_1m_properties():
    """configs[1]: batched Cholesky fp32 N=16, batch 1M.  Residual on the device for every matrix,
    bit-exact spot check of 4096 matrices spread over the batch against the oracle."""
    import torch
    batch = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.rand((batch, 16, 16), generator=g, device="cuda")
    A = 0.5 * (A + A.transpose(1, 2)) + 16.0 * torch.eye(16, device="cuda")
    L = hd.cholesky(A)
    res = (A - L @ L.transpose(1, 2)).flatten(1).norm(dim=1) / A.flatten(1).norm(dim=1)
    assert float(res.max()) < 2e-6
    assert float(torch.triu(L, 1).abs().max()) == 0.0
    idx = torch.arange(0, batch, batch // 4096, device="cuda")
    want = restate.cholesky(A[idx].cpu().numpy(), threads=8)
    assert bits_equal(L[idx].cpu().numpy(), want)


# ---------------------------------------------------------------- randomized sweep over sizes / batches / designs
def test_random_sweep_bit_exact():
    """60 random (kind, size, batch, design) draws through the auto dispatch, including odd sizes, batches
    that are not multiples of a slab, and unaligned device views (which must fall back to a kernel family
    without alignment requirements): all bit-identical to the oracle."""
    import torch
    rng = np.random.default_rng(2026)
    for it in range(60):
        kind = ("chol", "lu", "qr")[it % 3]
        n = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 12, 15, 16, 17, 24, 31, 32, 33, 48, 63, 64, 65, 80]))
        batch = int(rng.choice([1, 2, 15, 16, 17, 31, 33, 63, 100, 257]))
        if kind == "chol":
            d = str(rng.choice(["cholesky_v1.3", "cholesky_v2.2", "cholesky_v3.2", "cholesky_v4.0"]))
            A = gen_spd(n, batch=batch, seed=it)
            got, want = [hd.cholesky(A, design=d)], [restate.cholesky(A, d, threads=8)]
        elif kind == "lu":
            d = str(rng.choice(["lu_v1.1", "lu_v1.2", "lu_v2.1", "nopivot"]))
            A = gen_full_rank(n, batch=batch, seed=it) + rng.random((batch, n, n)).astype(np.float32)
            got, want = hd.lu(A, design=d), restate.lu(A, pivoting=d != "nopivot", threads=8)
        else:
            d = str(rng.choice(["qr_v1.1", "qr_v1.2", "qr_v2.1"]))
            rows = n + int(rng.integers(0, 4))
            A = gen_full_rank(rows, n, batch=batch, seed=it)
            got, want = hd.qr(A, design=d), restate.qr(A, d, threads=8)
        for g, w in zip(got, want):
            assert bits_equal(g, w), f"draw {it}: {kind} n={n} batch={batch} {d}: {first_mismatch(g, w)}"
    # device pointers that are only 4-byte aligned (a view one float into a buffer)
    for n in (16, 64):
        A = gen_spd(n, batch=40, seed=n)
        buf = torch.empty(A.size + 1, device="cuda")
        view = buf[1:].view(40, n, n)
        view.copy_(torch.from_numpy(A))
        L = hd.cholesky(view)
        assert bits_equal(L.cpu().numpy(), restate.cholesky(A))


def test_full_size_lu_n64_batch_262144_properties():
    """configs[2] at full size: every matrix is checked on the device through P-free properties (L unit
    lower with |l| <= 1, U upper, the pivot sequence valid), 2048 matrices spread over the batch are
    compared bit for bit with the oracle, and 64 of them are reconstructed (||A - P^T L U|| / ||A||)."""
    import torch
    batch, n = 262144, 64
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randint(1, 101, (batch, n, n), generator=g, device="cuda").float()
    L, U, P = hd.lu(A)
    assert float(torch.triu(L, 1).abs().max()) == 0.0
    assert bool((torch.diagonal(L, dim1=1, dim2=2) == 1.0).all())
    assert float(L.abs().max()) <= 1.0
    assert float(torch.tril(U, -1).abs().max()) == 0.0
    ar = torch.arange(n, device="cuda", dtype=torch.int32)
    assert bool((P >= ar).all()) and bool((P < n).all()) and bool((P[:, -1] == n - 1).all())
    idx = torch.arange(0, batch, batch // 2048, device="cuda")
    want = restate.lu(A[idx].cpu().numpy(), threads=8)
    for got, w in zip((L[idx], U[idx], P[idx]), want):
        assert bits_equal(got.cpu().numpy(), w)
    Ah, Lh, Uh, Ph = (x[idx[:64]].cpu().numpy() for x in (A, L, U, P))
    assert max(lu_residual(Ah[b], Lh[b], Uh[b], Ph[b]) for b in range(64)) < 5e-5


def test_full_size_qr_128_batch_16384_properties():
    """configs[3] at full size: ||A - QR|| / ||A|| and ||Q^T Q - I|| on the device for every matrix, R upper
    triangular exactly, 64 matrices bit for bit against the oracle."""
    import torch
    batch, n = 16384, 128
    g = torch.Generator(device="cuda").manual_seed(4)
    A = torch.randint(1, 101, (batch, n, n), generator=g, device="cuda").float()
    Q, R = hd.qr(A)
    assert float(torch.tril(R, -1).abs().max()) == 0.0
    res = (A.double() - Q.double() @ R.double()).flatten(1).norm(dim=1) / A.double().flatten(1).norm(dim=1)
    assert float(res.max()) < 1e-5
    eye = torch.eye(n, device="cuda", dtype=torch.float64)
    orth = (Q.double().transpose(1, 2) @ Q.double() - eye).flatten(1).norm(dim=1)
    assert float(orth.max()) < 1e-4
    idx = torch.arange(0, batch, batch // 64, device="cuda")
    wq, wr = restate.qr(A[idx].cpu().numpy(), threads=8)
    assert bits_equal(Q[idx].cpu().numpy(), wq) and bits_equal(R[idx].cpu().numpy(), wr)

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

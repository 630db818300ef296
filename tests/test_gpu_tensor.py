"""GPU tests of the blocked tensor-core Cholesky (tcgen05, 3xTF32): not bit-identical to the csim by
construction (different summation order), so the check is the tolerance BASELINE.json's north_star
names: relative residual ||A - L L^T|| / ||A|| and element-wise agreement with the oracle."""
import numpy as np
import pytest

import hls_designs_b200 as hd
import restate
from conftest import chol_residual
from hls_designs_b200.operand import gen_spd

pytestmark = pytest.mark.gpu

RESIDUAL_TOL = 2e-6      # same bound the exact paths meet
ELEMENT_TOL = 1e-5       # max |L - L_oracle| / max |L_oracle|  (the oracle itself carries ~1e-6 of fp32 round-off at n=1024)


@pytest.mark.parametrize("n,batch", ((256, 5), (384, 3), (512, 3), (1024, 2)))
def test_tensor_cholesky_vs_oracle(n, batch):
    A = gen_spd(n, batch=batch, seed=n)
    L = hd.cholesky(A, path="tensor")
    want = restate.cholesky(A, threads=8)
    assert np.isfinite(L).all()
    assert np.abs(np.triu(L, 1)).max() == 0.0
    err = np.abs(L - want).max() / np.abs(want).max()
    res = chol_residual(A, L).max()
    print(f"n={n}: element err {err:.2e}, residual {res:.2e} (oracle residual {chol_residual(A, want).max():.2e})")
    assert res < RESIDUAL_TOL
    assert err < ELEMENT_TOL


def test_auto_dispatch_uses_tensor_path_for_large_n():
    A = gen_spd(256, batch=2, seed=9)
    assert np.array_equal(hd.cholesky(A), hd.cholesky(A, path="tensor"))
    # the exact path stays available at any size and is bit-identical to the oracle: forced, or by HLSD_FLAG_EXACT
    for kw in (dict(path="global"), dict(exact=True)):
        exact = hd.cholesky(A, **kw)
        assert np.array_equal(exact.view(np.uint32), restate.cholesky(A).view(np.uint32))
    B = gen_spd(256, batch=1, seed=10) * np.float32(0.01) + np.arange(256 * 256, dtype=np.float32).reshape(256, 256) % 7
    for g, w in zip(hd.lu(B, exact=True), restate.lu(B)):
        assert np.array_equal(g, w, equal_nan=True) if g.dtype == np.float32 else np.array_equal(g, w)
    with pytest.raises(hd.HlsdError):
        hd.cholesky(A, path="tensor", exact=True)
    # cholesky_v2.2 has no tensor form: auto takes its blocked exact path (bit-identical)
    v22 = hd.cholesky(A[:1], design="cholesky_v2.2")
    assert np.array_equal(v22.view(np.uint32), restate.cholesky(A[:1], "cholesky_v2.2").view(np.uint32))
    with pytest.raises(hd.HlsdError):
        hd.cholesky(gen_spd(200), path="tensor")


# ---------------------------------------------------------------- blocked LU (tensor-core trailing update)
def _lu_check(A, L, U, P, want):
    from conftest import lu_residual
    n = A.shape[-1]
    wl, wu, wp = want
    out = []
    for b in range(A.shape[0]):
        assert np.isfinite(L[b]).all() and np.isfinite(U[b]).all()
        assert np.abs(np.triu(L[b], 1)).max() == 0.0 and np.allclose(np.diag(L[b]), 1.0)
        assert np.abs(np.tril(U[b], -1)).max() == 0.0
        assert np.abs(L[b]).max() <= 1.0 + 1e-6  # partial pivoting bounds the multipliers
        res = lu_residual(A[b], L[b], U[b], P[b])
        same = np.array_equal(P[b], wp[b])
        from conftest import lu_elementwise_ok
        err = lu_elementwise_ok(A[b], L[b], U[b], P[b], wl[b], wu[b]) if same else None
        out.append((res, same, err))
    return out


@pytest.mark.parametrize("n,batch", ((256, 4), (512, 3), (1024, 2)))
def test_tensor_lu_vs_oracle(n, batch):
    """Residual ||A - P^T L U|| / ||A|| for every matrix; element-wise agreement with the oracle where
    the pivot sequences coincide (a near-tie between two pivot candidates may resolve differently
    under tensor-core rounding, after which L and U legitimately differ)."""
    from hls_designs_b200.operand import gen_full_rank
    rng = np.random.default_rng(n)
    for kind, A in (("genA", gen_full_rank(n, batch=batch, seed=n)), ("normal", rng.standard_normal((batch, n, n)).astype(np.float32))):
        L, U, P = hd.lu(A, path="tensor")
        want = restate.lu(A, threads=8)
        from conftest import lu_residual
        ref_res = max(lu_residual(A[b], want[0][b], want[1][b], want[2][b]) for b in range(batch))
        checks = _lu_check(A, L, U, P, want)
        print(f"n={n}:", [(f"{r:.2e}", s, None if e is None else tuple(f"{x:.1e}" for x in e[1])) for r, s, e in checks], f"oracle residual {ref_res:.2e}")
        for res, same, err in checks:
            assert res < 5e-5
            if same:
                assert err[0], err[1]  # 1e-3, or 4x the csim's own distance from the float64 factors where that is larger
        # The csim's pivot sequence is reproduced on most operands, and on the reference's own op256.bin / op512.bin
        # (test_reference_large_operands_tensor_path asserts those exactly).  It is not a guarantee: where two
        # candidates agree to a few ulp, 3xTF32 rounding can order them differently, after which L and U differ
        # legitimately (one of three genA matrices at n = 512 did so under an earlier version of the diagonal-block
        # kernel).  So: a majority here, exact agreement on the reference's files, residuals always.
        assert sum(1 for _, same, _ in checks if same) * 2 > len(checks), f"n={n} {kind}: most pivot sequences differ from the csim"
        assert P[:, 0].tolist() == want[2][:, 0].tolist()  # first column: no arithmetic yet, pivots must agree


@pytest.mark.parametrize("n", (256, 512))
def test_tensor_lu_unpivoted(n):
    """HLSD_LU_NOPIVOT on the tensor path (auto dispatch at n >= 256): P[k] = k, residual and element-wise agreement with the
    oracle's unpivoted elimination on a diagonally dominant operand."""
    from conftest import lu_residual
    from hls_designs_b200.operand import gen_full_rank
    A = (gen_full_rank(n, batch=3, seed=n) + np.float32(100 * n) * np.eye(n, dtype=np.float32)).astype(np.float32)
    L, U, P = hd.lu(A, design="nopivot")
    wl, wu, wp = restate.lu(A, pivoting=False, threads=8)
    assert np.array_equal(P, wp) and np.array_equal(P[0], np.arange(n))
    for b in range(3):
        assert np.abs(np.triu(L[b], 1)).max() == 0.0 and np.abs(np.tril(U[b], -1)).max() == 0.0
        assert lu_residual(A[b], L[b], U[b], P[b]) < 5e-5
        assert np.abs(L[b] - wl[b]).max() < 1e-3 and np.abs(U[b] - wu[b]).max() / np.abs(wu[b]).max() < 1e-3


@pytest.mark.parametrize("n", (384, 640, 768, 896))
def test_tensor_lu_other_multiples_of_128(n):
    """Panel heights that are not powers of two (block sizes 192, 320, 384, 448 threads in the panel kernel, odd tile counts
    in the long-K kernels): same pivots as the csim on integer operands, residual, element-wise tolerance."""
    from conftest import lu_elementwise_ok, lu_residual
    from hls_designs_b200.operand import gen_full_rank
    A = gen_full_rank(n, batch=2, seed=n + 1)
    L, U, P = hd.lu(A, path="tensor")
    wl, wu, wp = restate.lu(A, threads=8)
    for b in range(2):
        assert np.abs(np.triu(L[b], 1)).max() == 0.0 and np.abs(np.tril(U[b], -1)).max() == 0.0 and np.allclose(np.diag(L[b]), 1.0)
        assert lu_residual(A[b], L[b], U[b], P[b]) < 5e-5
        if np.array_equal(P[b], wp[b]):
            ok, detail = lu_elementwise_ok(A[b], L[b], U[b], P[b], wl[b], wu[b])
            assert ok, detail
    assert any(np.array_equal(P[b], wp[b]) for b in range(2)), "no pivot sequence equal to the csim's"


def test_tensor_lu_workspace_slices():
    """The tensor LU walks a large batch in slices that bound its workspace (18 GB; lu_crout.cu): a batch that needs two
    slices gives, matrix for matrix, the results of separate calls on its parts, and valid factorisations in both."""
    import torch
    n, batch = 256, 24000  # 0.84 MB of workspace per matrix -> 23 011 matrices per slice
    g = torch.Generator(device="cuda").manual_seed(11)
    A = torch.randint(1, 101, (batch, n, n), generator=g, device="cuda").float()
    L, U, P = hd.lu(A)
    head = slice(0, 64)
    tail = slice(batch - 700, batch)  # inside the second slice
    for part in (head, tail):
        Lp, Up, Pp = hd.lu(A[part].contiguous())
        assert torch.equal(Lp, L[part]) and torch.equal(Up, U[part]) and torch.equal(Pp, P[part])
    # P[k] >= k everywhere, unit lower / upper structure on a sample from both slices
    assert bool((P >= torch.arange(n, device="cuda", dtype=P.dtype)).all()) and bool((P < n).all())
    pick = torch.tensor([0, 1, 22999, 23011, 23012, batch - 1], device="cuda")
    from conftest import lu_residual
    for b in pick.tolist():
        assert lu_residual(A[b].cpu().numpy(), L[b].cpu().numpy(), U[b].cpu().numpy(), P[b].cpu().numpy()) < 5e-5


def test_full_size_tensor_paths_n1024_properties():
    """configs[4] (N = 1024) on a 256-matrix slice of the batch: residuals on the device for EVERY matrix, structure of
    the factors, and twelve matrices spread over the slice against the oracle within the documented tolerances (tensor
    paths) and bit for bit (the exact blocked paths on the same operands)."""
    import torch
    from conftest import lu_residual
    n, batch = 1024, 256
    pick = np.unique(np.linspace(0, batch - 1, 12).astype(np.int64))
    tpick = torch.from_numpy(pick).cuda()
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.rand((batch, n, n), generator=g, device="cuda")
    S = (0.5 * (A + A.transpose(1, 2)))
    S.diagonal(dim1=1, dim2=2).add_(float(n))
    del A
    L = hd.cholesky(S)
    res = (S - L @ L.transpose(1, 2)).flatten(1).norm(dim=1) / S.flatten(1).norm(dim=1)
    assert float(res.max()) < RESIDUAL_TOL and float(torch.triu(L, 1).abs().max()) == 0.0
    Sh = S[tpick].cpu().numpy()
    want = restate.cholesky(Sh, threads=8)
    got = L[tpick].cpu().numpy()
    assert np.abs(got - want).max() / np.abs(want).max() < ELEMENT_TOL
    Lx = hd.cholesky(S[tpick].contiguous(), exact=True)  # blocked exact path on the same operands
    assert np.array_equal(Lx.cpu().numpy().view(np.uint32), want.view(np.uint32))
    del L, S, Lx
    B = torch.randint(1, 101, (batch, n, n), generator=g, device="cuda").float()
    Ll, Uu, P = hd.lu(B)
    assert float(torch.triu(Ll, 1).abs().max()) == 0.0 and float(Ll.abs().max()) <= 1.0 + 1e-6
    assert float(torch.tril(Uu, -1).abs().max()) == 0.0
    assert bool((P >= torch.arange(n, device="cuda", dtype=P.dtype)).all()) and bool((P < n).all())  # P[k] >= k: a valid exchange
    Bh, Lh, Uh, Ph = (x[tpick].cpu().numpy() for x in (B, Ll, Uu, P))
    wl, wu, wp = restate.lu(Bh, threads=8)
    same = 0
    for b in range(len(pick)):
        assert lu_residual(Bh[b], Lh[b], Uh[b], Ph[b]) < 5e-5
        if np.array_equal(Ph[b], wp[b]):
            same += 1
            from conftest import lu_elementwise_ok
            ok, detail = lu_elementwise_ok(Bh[b], Lh[b], Uh[b], Ph[b], wl[b], wu[b])
            assert ok, detail
    print(f"tensor LU 1024: {same} of {len(pick)} pivot sequences equal to the oracle's")
    xl, xu, xp = hd.lu(B[tpick].contiguous(), exact=True)  # blocked exact path: bit-identical L, U, P
    assert np.array_equal(xp.cpu().numpy(), wp)
    assert np.array_equal(xl.cpu().numpy().view(np.uint32), wl.view(np.uint32))
    assert np.array_equal(xu.cpu().numpy().view(np.uint32), wu.view(np.uint32))

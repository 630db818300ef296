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
    # the exact path stays available at any size and is bit-identical to the oracle
    exact = hd.cholesky(A, path="global")
    assert np.array_equal(exact.view(np.uint32), restate.cholesky(A).view(np.uint32))
    # cholesky_v2.2 has no blocked form: auto falls back to the exact kernel
    v22 = hd.cholesky(A[:1], design="cholesky_v2.2")
    assert np.array_equal(v22.view(np.uint32), restate.cholesky(A[:1], "cholesky_v2.2").view(np.uint32))
    with pytest.raises(hd.HlsdError):
        hd.cholesky(gen_spd(200), path="tensor")

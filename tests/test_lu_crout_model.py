"""CPU check of the data flow of the tensor-core LU (csrc/lu_crout.cu): rows never move, everything is addressed
through posmap, L is assembled from the physical-row multiplier store at the end.  scripts/lu_crout_model.py runs
those steps in numpy (float64, small blocks) and this test compares them with a plain LINPACK-style elimination, the
reference's arithmetic (LU/lu_v1.1/model4x4/lu.cpp:23-92, restated in oracle/restate.c:101-139)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
from lu_crout_model import crout_lu, reference_lu  # noqa: E402


@pytest.mark.parametrize("n,nb,sub", ((16, 8, 4), (32, 8, 4), (64, 16, 4), (96, 32, 8), (128, 32, 8)))
def test_crout_data_flow_matches_linpack_elimination(n, nb, sub):
    rng = np.random.default_rng(n)
    for trial in range(4):
        A = rng.standard_normal((n, n)) if trial < 3 else rng.integers(1, 4, (n, n)).astype(np.float64)  # last: pivot ties
        L0, U0, P0 = reference_lu(A)
        L1, U1, P1 = crout_lu(A, nb, sub)
        assert np.array_equal(P0, P1)
        assert np.allclose(L0, L1, atol=1e-9)
        assert np.allclose(U0, U1, atol=1e-9 * max(1.0, np.abs(U0).max()))


def test_model_agrees_with_the_oracle_on_a_reference_operand():
    """Same P as oracle/restate.c on LU/op16.bin (the oracle is fp32, the model fp64: the operand has no near-ties)."""
    import restate
    from hls_designs_b200.operand import read_operand
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "operands", "LU_op16.bin")
    if not os.path.exists(path):
        pytest.skip("operand file not in the fixtures")
    A = read_operand(path, 16)
    wl, wu, wp = restate.lu(A[None])
    L1, U1, P1 = crout_lu(A.astype(np.float64), 8, 4)
    assert np.array_equal(P1, wp[0])
    assert np.allclose(L1, wl[0], atol=1e-4) and np.allclose(U1, wu[0], atol=1e-3)

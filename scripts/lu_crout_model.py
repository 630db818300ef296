"""Numpy model of the blocked tensor-core LU's data flow (csrc/lu_crout.cu): rows are NEVER moved.

The kernels keep every matrix row where it lies in A and address rows through `posmap` (position in the
reference's exchanged order -> physical row).  This model runs the same steps in float64 with small blocks and
checks them against a plain LINPACK-style elimination (the reference's arithmetic, oracle/restate.c:101-139), so
the index logic of the CUDA kernels can be validated without a GPU (tests/test_lu_crout_model.py).

    M2[phys][k]   multiplier of physical row `phys` at step k (GEMM operand, row order never changes)
    UT[c][k]      U[k][c] (k-contiguous B operand)
    PWT[c][t]     panel workspace, transposed: column c of the panel, thread/row t (position o + t at panel start)
    hist[p][i]    posmap at the start of panel p
"""
import numpy as np


def reference_lu(A):
    """LINPACK-style elimination as in the reference (P[k] = position exchanged with k; L columns not exchanged)."""
    n = A.shape[0]
    w = A.astype(np.float64).copy()
    L = np.eye(n)
    U = np.zeros((n, n))
    P = np.zeros(n, dtype=np.int64)
    for k in range(n - 1):
        p = k + int(np.argmax(np.abs(w[k:, k])))
        P[k] = p
        if p != k:
            w[[k, p], k:] = w[[p, k], k:]
        w[k + 1:, k] /= w[k, k]
        L[k + 1:, k] = w[k + 1:, k]
        U[k, k:] = w[k, k:]
        w[k + 1:, k + 1:] -= np.outer(w[k + 1:, k], w[k, k + 1:])
    P[n - 1] = n - 1
    U[n - 1, n - 1] = w[n - 1, n - 1]
    return L, U, P


def crout_lu(A, NB=8, SUB=4):
    n = A.shape[0]
    assert n % NB == 0 and NB % SUB == 0
    A = A.astype(np.float64)
    panels = n // NB
    M2 = np.zeros((n, n))
    UT = np.zeros((n, n))
    U = np.zeros((n, n))
    P = np.zeros(n, dtype=np.int64)
    hist = np.zeros((panels + 1, n), dtype=np.int64)
    hist[0] = np.arange(n)
    for p in range(panels):
        o = p * NB
        R = n - o
        pm = hist[p]
        # ---- K1: column update, X[i][c] = A[pm[i]][o + c] - sum_{k < o} M2[pm[i]][k] UT[o + c][k]  -> PWT[c][i - o]
        phys = pm[o:]                                    # thread t <-> position o + t, physical row phys[t]
        X = A[phys, o:o + NB] - M2[phys, :o] @ UT[o:o + NB, :o].T
        PWT = X.T.copy()
        # ---- K2: panel kernel, thread = row, rows never move, positions tracked
        mypos = np.arange(o, n)                          # per thread
        won_step = np.full(R, -1)                        # step at which the thread became a pivot row
        for cs in range(0, NB, SUB):
            a = PWT[cs:cs + SUB, :].T.copy()             # a[t][j]
            prow = np.zeros((SUB, SUB))
            tw = np.zeros(SUB, dtype=np.int64)
            for kk in range(SUB):
                k = o + cs + kk
                cand = mypos >= k
                key = np.where(cand, np.abs(a[:, kk]), -1.0)
                best = key.max()
                gp = mypos[(key == best)].min()          # smallest POSITION among the largest |a|
                w = int(np.nonzero(mypos == gp)[0][0])   # winner thread
                P[k] = gp
                disp = np.nonzero(mypos == k)[0]
                if gp != k:
                    mypos[disp[0]] = gp
                mypos[w] = k
                won_step[w] = k
                prow[kk] = a[w]
                tw[kk] = w
                act = mypos > k
                l = a[act, kk] / a[w, kk]
                a[act, kk] = l
                a[act, kk + 1:] -= np.outer(l, a[w, kk + 1:])
            M2[phys, o + cs:o + cs + SUB] = a            # winners: multipliers left of their step, U values right of it
            rem0 = cs + SUB
            nrem = NB - rem0
            urows = np.zeros((SUB, NB))
            for kk in range(SUB):
                urows[kk, cs + kk:cs + SUB] = prow[kk, kk:]
            if nrem > 0:
                xs = PWT[rem0:, tw].T.copy()             # xs[kk][c]: the pivot rows in the remaining panel columns
                for kk in range(SUB):
                    for j in range(kk):
                        xs[kk] -= prow[kk, j] * xs[j]
                urows[:, rem0:] = xs
                act = mypos >= o + cs + SUB              # rows that are not pivots yet
                PWT[rem0:, act] -= (a[act] @ xs).T
            U[o + cs:o + cs + SUB, o:o + NB] = urows
        hist[p + 1] = hist[p]
        hist[p + 1][mypos] = phys
        pm2 = hist[p + 1]
        if p == panels - 1:
            break
        # ---- K3: W = inv(L11), L11[i][j] = M2[pm2[o + i]][o + j], j < i, unit diagonal
        piv = pm2[o:o + NB]
        L11 = np.tril(M2[piv, o:o + NB], -1) + np.eye(NB)
        W = np.linalg.inv(L11)
        # ---- K4a: X12^T[c][m] = A[piv[m]][c] - sum_{k<o} UT[c][k] M2[piv[m]][k]   (c > o + NB)
        cols = np.arange(o + NB, n)
        T0 = A[np.ix_(piv, cols)].T - UT[cols, :o] @ M2[piv, :o].T
        # ---- K4b: T1[c][k'] = sum_m T0[c][m] W[k'][m] -> UT[c][o + k'] and U[o + k'][c]
        T1 = T0 @ W.T
        UT[o + NB:, o:o + NB] = T1
        U[o:o + NB, o + NB:] = T1.T
    # ---- L assembly: L[i][k] = M2[occupant of position i after step k's exchange][k]
    L = np.eye(n)
    for p in range(panels):
        o = p * NB
        occ = hist[p].copy()
        for j in range(NB):
            k = o + j
            q = P[k]
            occ[[k, q]] = occ[[q, k]]
            L[k + 1:, k] = M2[occ[k + 1:], k]
    return L, U, P


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n, NB, SUB in ((16, 8, 4), (32, 8, 4), (64, 16, 4), (96, 32, 8)):
        for trial in range(5):
            A = rng.standard_normal((n, n))
            if trial == 4:
                A = rng.integers(1, 4, (n, n)).astype(np.float64)  # many pivot ties
            L0, U0, P0 = reference_lu(A)
            L1, U1, P1 = crout_lu(A, NB, SUB)
            assert np.array_equal(P0, P1), (n, trial)
            assert np.allclose(L0, L1, atol=1e-9), (n, trial, np.abs(L0 - L1).max())
            assert np.allclose(U0, U1, atol=1e-9 * np.abs(U0).max()), (n, trial, np.abs(U0 - U1).max())
    print("crout model ok")

import sys, os
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
import hls_designs_b200 as hd, restate
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rng = np.random.default_rng(1)
A = rng.standard_normal((1, n, n)).astype(np.float32)
L, U, P = hd.lu(A, path="tensor")
wl, wu, wp = restate.lu(A)
L, U, P, wl, wu, wp = L[0], U[0], P[0], wl[0], wu[0], wp[0]
neq = np.nonzero(P != wp)[0]
print("first pivot mismatch:", neq[:5], "of", len(neq))
def blockerr(X, Y, name):
    nb = n // 128
    for bi in range(nb):
        row = []
        for bj in range(nb):
            x = X[bi*128:(bi+1)*128, bj*128:(bj+1)*128]; y = Y[bi*128:(bi+1)*128, bj*128:(bj+1)*128]
            row.append(f"{np.abs(x-y).max():9.2e}")
        print(name, bi, ' '.join(row))
blockerr(L, wl, 'L')
blockerr(U, wu, 'U')
# finer: first panel columns in 32-wide sub-panels
for sp in range(4):
    c = slice(32*sp, 32*sp+32)
    print('subpanel', sp, 'L err', np.abs(L[:, c]-wl[:, c]).max(), 'U err', np.abs(U[:128, c]-wu[:128, c]).max())

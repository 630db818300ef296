import sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import hls_designs_b200 as hd, restate
from hls_designs_b200.operand import gen_full_rank
for n, opt in ((128, 0), (64, 3)):
    hd.set_option("qr_threads", opt)
    A = gen_full_rank(n, n, batch=5, seed=n)
    A[1, 5:9, 0] = 0.0   # skipped rotations
    A[2, :, 3] = 0.0
    Q, R = hd.qr(A)
    wq, wr = restate.qr(A)
    print(n, "Q bits", np.array_equal(Q.view(np.uint32), wq.view(np.uint32)), "R bits", np.array_equal(R.view(np.uint32), wr.view(np.uint32)),
          np.abs(Q - wq).max(), np.abs(R - wr).max())
    if not np.array_equal(R.view(np.uint32), wr.view(np.uint32)):
        bad = np.argwhere(R.view(np.uint32) != wr.view(np.uint32)); print(bad[:10].tolist(), len(bad))
hd.set_option("qr_threads", 0)

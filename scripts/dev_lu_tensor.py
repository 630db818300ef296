#!/usr/bin/env python3
"""Tensor LU (lu_crout.cu) on the GPU: correctness diagnosis against the oracle at 256/512/1024 and per-kernel-class
times at N = 1024.  Usage: python scripts/dev_lu_tensor.py [batch_for_timing] [--quick] [--opt=name:value]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import hls_designs_b200 as hd  # noqa: E402
import restate  # noqa: E402
from conftest import lu_residual  # noqa: E402
from hls_designs_b200 import _lib  # noqa: E402

for a in sys.argv[1:]:
    if a.startswith("--opt="):
        name, val = a[6:].split(":")
        hd.set_option(name, int(val))
quick = "--quick" in sys.argv
args = [a for a in sys.argv[1:] if not a.startswith("--")]
tb = int(args[0]) if args else 1024

for n, batch in ((256, 3), (512, 2), (1024, 2)) if not quick else ((256, 2), (1024, 1)):
    rng = np.random.default_rng(n)
    for kind in ("int", "normal"):
        A = rng.integers(1, 101, (batch, n, n)).astype(np.float32) if kind == "int" else rng.standard_normal((batch, n, n)).astype(np.float32)
        L, U, P = hd.lu(A, path="tensor")
        wl, wu, wp = restate.lu(A, threads=8)
        for b in range(batch):
            same = np.array_equal(P[b], wp[b])
            first = int(np.argmax(P[b] != wp[b])) if not same else -1
            res = lu_residual(A[b], L[b], U[b], P[b]) if np.isfinite(L[b]).all() and np.isfinite(U[b]).all() else float("nan")
            eu = np.abs(U[b] - wu[b]).max() / np.abs(wu[b]).max()
            el = np.abs(L[b] - wl[b]).max()
            struct = (np.abs(np.triu(L[b], 1)).max(), np.abs(np.tril(U[b], -1)).max(), np.abs(np.diag(L[b]) - 1).max())
            print(f"n={n} {kind} b={b}: pivots same={same} first_diff={first} residual={res:.2e} errU={eu:.2e} errL={el:.2e} struct={struct}")
            if not same or not (eu < 1e-3) or not (el < 1e-3):
                # where do L / U go wrong first?
                bu = np.argwhere(np.abs(U[b] - wu[b]) > 1e-3 * np.abs(wu[b]).max())
                bl = np.argwhere(np.abs(L[b] - wl[b]) > 1e-3)
                print("   first bad U:", bu[:3].tolist(), len(bu), " first bad L:", bl[:3].tolist(), len(bl))

lib = _lib.load()
n = 1024
g = torch.Generator(device="cuda").manual_seed(5)
B = torch.randint(1, 101, (tb, n, n), generator=g, device="cuda").float()
for _ in range(2):
    out = hd.lu(B)
torch.cuda.synchronize()
ms = []
for _ in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    out = hd.lu(B)
    e.record()
    torch.cuda.synchronize()
    ms.append(s.elapsed_time(e))
t = sorted(ms)[1]
print(f"lu n=1024 batch={tb}: {t:.2f} ms -> {tb / t * 1e3:.0f} fact/s; scaled to 4096: {t * 4096 / tb:.1f} ms")
names = ["other", "update", "diag", "trsm", "panel", "permute", "exact"]
arr = (ctypes.c_double * 7)()
cnt = (ctypes.c_int64 * 7)()
lib.hlsd_profile_begin()
hd.lu(B)
torch.cuda.synchronize()
lib.hlsd_profile_end(arr, cnt, 7)
print("per class (ms, launches):", {names[i]: (round(arr[i], 2), cnt[i]) for i in range(7) if cnt[i]})

#!/usr/bin/env python3
"""Time the exact (bit-identical) paths at N > 128: HLSD_FLAG_EXACT vs the tensor path.  python scripts/time_exact_large.py"""
import os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hls_designs_b200 as hd  # noqa: E402

dev = torch.device("cuda", 0)
for kind, n, batch in (("chol", 256, 8192), ("chol", 512, 2048), ("chol", 1024, 512), ("lu", 256, 8192), ("lu", 512, 2048), ("lu", 1024, 1024)):
    g = torch.Generator(device=dev).manual_seed(1)
    if kind == "chol":
        A = torch.rand((batch, n, n), generator=g, device=dev)
        A = 0.5 * (A + A.transpose(1, 2))
        A.diagonal(dim1=1, dim2=2).add_(float(n))
        A = A.contiguous()
        fn = hd.cholesky
    else:
        A = torch.randint(1, 101, (batch, n, n), generator=g, device=dev).float()
        fn = hd.lu
    for label, kw in (("exact", dict(exact=True)), ("tensor", dict())):
        fn(A, **kw)
        torch.cuda.synchronize()
        ms = []
        for _ in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(A, **kw); e.record(); torch.cuda.synchronize()
            ms.append(s.elapsed_time(e))
        t = statistics.median(ms)
        fl = (n ** 3 / 3 if kind == "chol" else 2 * n ** 3 / 3) * batch / (t * 1e-3) / 1e12
        print(f"{kind} n={n} batch={batch} {label:7s} {t:9.3f} ms {batch / t * 1e3:10.4g} fact/s {fl:7.2f} TFLOP/s", flush=True)

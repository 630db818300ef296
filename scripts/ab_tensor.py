#!/usr/bin/env python3
"""A/B timing of the tensor-core paths under the measurement switches (HLSD_OPT_*), with the per-kernel-class
breakdown of hlsd_profile_begin/_end.  Operands resident, CUDA events.
    python scripts/ab_tensor.py [chol|lu] [n] [batch]"""
import ctypes
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import hls_designs_b200 as hd  # noqa: E402
from bench import PROF_CLASSES  # noqa: E402

CASES = [("chol", 1024, 4096), ("chol", 512, 8192), ("chol", 256, 16384), ("lu", 1024, 2048), ("lu", 512, 4096)]
VARIANTS = [("default", {}, {}), ("diag_kernel=1", {"diag_kernel": 1}, {}), ("exact (blocked)", {}, {"exact": True})]


def main():
    dev = torch.device("cuda", 0)
    lib = hd.load()
    cases = CASES
    if len(sys.argv) > 3:
        cases = [(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))]
    for kind, n, batch in cases:
        g = torch.Generator(device=dev).manual_seed(3)
        if kind == "chol":
            A = torch.rand((batch, n, n), generator=g, device=dev)
            A = 0.5 * (A + A.transpose(1, 2))
            A.diagonal(dim1=1, dim2=2).add_(float(n))
            A = A.contiguous()
            fn = hd.cholesky
        else:
            A = torch.randint(1, 101, (batch, n, n), generator=g, device=dev).float()
            fn = hd.lu
        ref = None
        for label, opts, kw in VARIANTS:
            for k in hd.api.OPTIONS:
                hd.set_option(k, 0)
            for k, v in opts.items():
                hd.set_option(k, v)
            for _ in range(2):
                out = fn(A, **kw)
            torch.cuda.synchronize()
            ms = []
            for _ in range(5):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                out = fn(A, **kw)
                e.record()
                torch.cuda.synchronize()
                ms.append(s.elapsed_time(e))
            t = statistics.median(ms)
            prof = (ctypes.c_double * len(PROF_CLASSES))()
            cnt = (ctypes.c_int64 * len(PROF_CLASSES))()
            lib.hlsd_profile_begin()
            fn(A, **kw)
            torch.cuda.synchronize()
            lib.hlsd_profile_end(prof, cnt, len(PROF_CLASSES))
            first = (out[0] if isinstance(out, tuple) else out)[:4].float().cpu().numpy()
            diff = "" if ref is None else f"  max|diff| vs first variant {np.abs(first - ref).max():.2e} (max|x| {np.abs(ref).max():.2e})"
            ref = first if ref is None else ref
            print(f"{kind}{n} batch={batch} {label:16s} {t:8.3f} ms {batch / t * 1e3:10.4g} fact/s  "
                  + " ".join(f"{c}={prof[i]:.2f}" for i, c in enumerate(PROF_CLASSES) if cnt[i]) + diff, flush=True)
        del A, out
        torch.cuda.empty_cache()
    for k in hd.api.OPTIONS:
        hd.set_option(k, 0)


if __name__ == "__main__":
    main()

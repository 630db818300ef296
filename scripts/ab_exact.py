#!/usr/bin/env python3
"""A/B timing of the exact register kernels under the measurement switches of include/hlsd.h (HLSD_OPT_*), each
setting first checked bit for bit against the oracle on a small batch.  Operands resident, CUDA events, median of 5.
    python scripts/ab_exact.py [case ...]        cases: chol32 chol64 lu8 lu16 lu32 lu64 lu128 qr16 qr32 ..."""
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import hls_designs_b200 as hd  # noqa: E402
import restate  # noqa: E402
from bench import algorithmic_bytes, flops  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6548.5) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6548.5
CASES = {
    # name: (kind, n, [(label, {option: value}, fused)])
    "chol32": ("chol", 32, [("dense6", {}, False), ("packed12", {"chol_slab": 1}, False), ("packed8", {"chol_slab": 2}, False),
                            ("dense6+lockstep", {"lockstep": 1}, False), ("dense6+fused", {}, True)]),
    "chol64": ("chol", 64, [("tile128", {}, False), ("tile64", {"chol_tile": 2}, False), ("regs packed16", {"chol_tile": 1}, False),
                            ("regs dense12", {"chol_tile": 1, "chol_slab": 1}, False), ("tile128+fused", {}, True)]),
    "chol128": ("chol", 128, [("tile", {}, False), ("regs (blk_diag)", {"chol_tile": 1}, False), ("tile+fused", {}, True)]),
    "lu8": ("lu", 8, [("default", {}, False)]),
    "lu16": ("lu", 16, [("default", {}, False), ("fused", {}, True)]),
    "lu32": ("lu", 32, [("lane", {}, False), ("lane+fused", {}, True)]),
    "lu64": ("lu", 64, [("default", {}, False), ("lockstep", {"lockstep": 1}, False), ("fused", {}, True)]),
    "lu128": ("lu", 128, [("lane", {}, False), ("lockstep", {"lockstep": 1}, False), ("fused", {}, True)]),
    "qr8": ("qr", 8, [("default", {}, False)]),
    "qr16": ("qr", 16, [("default", {}, False), ("lockstep", {"lockstep": 1}, False), ("fused", {}, True)]),
    "qr32": ("qr", 32, [("default", {}, False), ("lockstep", {"lockstep": 1}, False), ("fused", {}, True)]),
    "qr64": ("qr", 64, [("default", {}, False), ("two-role", {"qr_threads": 1}, False), ("fused", {}, True)]),
    "qr128": ("qr", 128, [("default", {}, False), ("two-role", {"qr_threads": 1}, False), ("fused", {}, True), ("two-role+fused", {"qr_threads": 1}, True)]),
    "chol16": ("chol", 16, [("default", {}, False)]),
}


def operand(kind, n, batch, dev):
    g = torch.Generator(device=dev).manual_seed(7)
    if kind == "chol":
        A = torch.rand((batch, n, n), generator=g, device=dev)
        A = 0.5 * (A + A.transpose(1, 2))
        A.diagonal(dim1=1, dim2=2).add_(float(n))
        return A.contiguous()
    return torch.randint(1, 101, (batch, n, n), generator=g, device=dev).float()


def main():
    dev = torch.device("cuda", 0)
    names = sys.argv[1:] or list(CASES)
    for name in names:
        kind, n, variants = CASES[name]
        per = algorithmic_bytes(kind, n, n)
        batch = int(max(64, min(1 << 20, (3 << 30) // per)))
        A = operand(kind, n, batch, dev)
        fn = {"chol": hd.cholesky, "lu": hd.lu, "qr": hd.qr}[kind]
        small = A[:257].cpu().numpy()
        want = {"chol": restate.cholesky, "lu": restate.lu, "qr": restate.qr}[kind](small, threads=8)
        want = want if isinstance(want, tuple) else (want,)
        for label, opts, fused in variants:
            for k in hd.api.OPTIONS:
                hd.set_option(k, 0)
            for k, v in opts.items():
                hd.set_option(k, v)
            got = fn(A[:257], fused=fused)
            got = got if isinstance(got, tuple) else (got,)
            exact = all(np.array_equal(g.cpu().numpy().view(np.uint32), w.view(np.uint32)) for g, w in zip(got, want))
            err = max(float(np.abs(g.cpu().numpy().astype(np.float64) - w).max() / max(1e-30, np.abs(w).max())) for g, w in zip(got, want))
            for _ in range(3):
                fn(A, fused=fused)
            torch.cuda.synchronize()
            ms = []
            for _ in range(5):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn(A, fused=fused)
                e.record()
                torch.cuda.synchronize()
                ms.append(s.elapsed_time(e))
            t = statistics.median(ms) * 1e-3
            print(f"{name:8s} {label:16s} batch={batch:8d} {t * 1e3:8.3f} ms {batch / t:10.4g} fact/s {flops(kind, n, n) * batch / t / 1e12:6.2f} TFLOP/s "
                  f"{per * batch / t / 1e9:6.0f} GB/s = {per * batch / t / 1e9 / PEAK:.3f} of HBM peak  "
                  f"{'bit-exact' if exact else f'max rel diff {err:.2e}'}{'' if exact or fused else '   <-- MISMATCH'}", flush=True)
        del A
        torch.cuda.empty_cache()
    for k in hd.api.OPTIONS:
        hd.set_option(k, 0)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""BASELINE.json's metric in full: factorizations/s, GFLOP/s and fraction of the HBM roofline per N
(4..1024) for Cholesky, LU and QR, operands resident in HBM, CUDA-event timing, one GPU.
Writes profiles/r01_sweep.json and prints a markdown table.  Usage: python scripts/sweep.py"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hls_designs_b200 as hd  # noqa: E402
from bench import algorithmic_bytes, flops  # noqa: E402

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
PEAK = float(peaks.get("hbm_gbs", 6650.0))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
rows_out = []
for kind, sizes in (("chol", (4, 8, 16, 32, 64, 128, 256, 512, 1024)), ("lu", (4, 8, 16, 32, 64, 128, 256, 512, 1024)),
                    ("qr", (4, 8, 16, 32, 64, 128))):
    for n in sizes:
        per = algorithmic_bytes(kind, n, n)
        # ~3 GiB of traffic per step, far beyond L2; the smallest sizes are capped at 8M matrices (a 1M batch of
        # 4x4 matrices is a 40 us kernel, mostly launch ramp)
        batch = int(max(64, min(1 << 23, (3 << 30) // per)))
        if kind == "qr" and n >= 64:
            batch = min(batch, 16384)
        if kind == "chol":
            A = torch.rand((batch, n, n), generator=g, device=dev)
            A = 0.5 * (A + A.transpose(1, 2))
            A.diagonal(dim1=1, dim2=2).add_(float(n))
            A = A.contiguous()
            fn = hd.cholesky
        else:
            A = torch.randint(1, 101, (batch, n, n), generator=g, device=dev).float()
            fn = hd.lu if kind == "lu" else hd.qr
        for _ in range(3):
            fn(A)
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn(A)
            e.record()
            torch.cuda.synchronize()
            ms.append(s.elapsed_time(e))
        t = statistics.median(ms) * 1e-3
        gbs = per * batch / t / 1e9
        rec = {"kind": kind, "n": n, "batch": batch, "ms": t * 1e3, "fact_per_s": batch / t,
               "gflops": flops(kind, n, n) * batch / t / 1e9, "gbs": gbs, "hbm_frac": gbs / PEAK}
        rows_out.append(rec)
        del A
        torch.cuda.empty_cache()
json.dump({"hbm_peak_gbs": PEAK, "rows": rows_out}, open(os.path.join(ROOT, "gpurun_out", "r01_sweep.json"), "w"), indent=1)
print("| op | N | batch | ms | fact/s | GFLOP/s | GB/s | of HBM peak |")
print("|---|---|---|---|---|---|---|---|")
for r in rows_out:
    print(f"| {r['kind']} | {r['n']} | {r['batch']} | {r['ms']:.3f} | {r['fact_per_s']:.4g} | {r['gflops']:.0f} | {r['gbs']:.0f} | {r['hbm_frac']:.3f} |")

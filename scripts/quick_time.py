#!/usr/bin/env python3
"""Time one decomposition at one size on the GPU (operands resident in HBM, CUDA events, median of 5).
Usage: python scripts/quick_time.py <chol|lu|qr> <n> [batch] [--fused] [--opt name=value ...]"""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hls_designs_b200 as hd  # noqa: E402
from bench import algorithmic_bytes  # noqa: E402

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
fused = "--fused" in sys.argv
opts = [a for i, a in enumerate(sys.argv) if i > 0 and sys.argv[i - 1] == "--opt"]
argv = [a for a in argv if a not in opts]
for o in opts:
    hd.set_option(o.split("=")[0], int(o.split("=")[1]))
kind, n = argv[0], int(argv[1])
per = algorithmic_bytes(kind, n, n)
batch = int(argv[2]) if len(argv) > 2 else int(max(64, min(1 << 20, (3 << 30) // per)))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(7)
if kind == "chol":
    A = torch.rand((batch, n, n), generator=g, device=dev)
    A = (0.5 * (A + A.transpose(1, 2)))
    A.diagonal(dim1=1, dim2=2).add_(float(n))
    A = A.contiguous()
    fn0 = hd.cholesky
else:
    A = torch.randint(1, 101, (batch, n, n), generator=g, device=dev).float()
    fn0 = hd.lu if kind == "lu" else hd.qr


def fn(x):
    return fn0(x, fused=fused)


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
print(f"{kind} n={n} batch={batch} fused={fused} opts={opts}: "
      f"{t * 1e3:.3f} ms  {batch / t:.4g} fact/s  {per * batch / t / 1e9:.0f} GB/s")

#!/usr/bin/env python3
"""The host staging pipeline without kernels (hlsd_selftest_copy_host): record shapes, and input buffers allocated as
plain pinned vs write-combined pinned memory.  python scripts/dev_copytest.py"""
import ctypes
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hls_designs_b200 import _lib  # noqa: E402

lib = _lib.load()
lib.hlsd_selftest_copy_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int64]
torch.cuda.init()
rt = ctypes.CDLL("libcudart.so.12")  # the runtime torch has already loaded
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaFreeHost.argtypes = [ctypes.c_void_p]
size = 1 << 30


def alloc(flags):
    ptr = ctypes.c_void_p()
    err = rt.cudaHostAlloc(ctypes.byref(ptr), size, flags)
    assert err == 0, err
    return ptr.value


po = alloc(0)
for name, flags in (("pinned", 0), ("write-combined", 4), ("pinned", 0), ("write-combined", 4)):
    pa = alloc(flags)
    ctypes.memset(pa, 1, size)
    lib.hlsd_selftest_copy_host(pa, po, 2048, 2048, size // 2048)
    ts = []
    for it in range(5):
        t0 = time.perf_counter()
        rc = lib.hlsd_selftest_copy_host(pa, po, 2048, 2048, size // 2048)
        ts.append(time.perf_counter() - t0)
    dt = sorted(ts)[2]
    print(f"input {name:15s}: {dt * 1e3:7.1f} ms per GiB each way -> {size / dt / 1e9:.1f} GB/s per direction", flush=True)
    rt.cudaFreeHost(pa)

import sys, time, ctypes, numpy as np
sys.path.insert(0, "/root/repo")
import torch
import hls_designs_b200 as hd
from hls_designs_b200 import _lib
lib = _lib.load()
lib.hlsd_host_alloc.restype = ctypes.c_void_p
lib.hlsd_selftest_copy_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int64]
torch.cuda.init()
n, batch = 1024, 682
in_mat, out_mat = n * n * 4, 2 * n * n * 4 + n * 4
pa = lib.hlsd_host_alloc(ctypes.c_size_t(in_mat * batch)); po = lib.hlsd_host_alloc(ctypes.c_size_t(out_mat * batch))
for shape in ((in_mat, out_mat), (in_mat, in_mat), (in_mat // 2, out_mat // 2), (64 << 10, 128 << 10)):
    units = batch * in_mat // shape[0]
    for it in range(4):
        t0 = time.perf_counter()
        rc = lib.hlsd_selftest_copy_host(pa, po, shape[0], shape[1], units)
        dt = time.perf_counter() - t0
        print(shape, units, rc, f"{dt*1e3:.1f} ms  h2d {shape[0]*units/dt/1e9:.1f} GB/s d2h {shape[1]*units/dt/1e9:.1f} GB/s", flush=True)

"""Builds hls_designs_b200/libhlsd.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m hls_designs_b200.build [--force]

nvcc cross-compiles without a GPU.  No fast-math: the exact paths rely on IEEE division and
square root and on un-contracted multiplies/adds (see csrc/common.cuh).
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libhlsd.so")
SOURCES = ["chol.cu", "lu.cu", "qr.cu", "blocked.cu", "lu_crout.cu", "blocked_exact.cu", "selftest.cu", "cabi.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _newest_header():
    t = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def _compile(src, force):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(path), _newest_header()):
        return obj, ""
    res = subprocess.run([NVCC] + FLAGS + ["-c", path, "-o", obj], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
    return obj, res.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(lambda s: _compile(s, force), SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        res = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-lcuda"], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stderr)
    return LIB


HOST = os.path.join(HERE, "host")                       # the drop-in top() headers (product)
TB_DIR = os.path.join(HERE, "..", "tests", "host")      # the testbench that exercises them (test code)
HOST_BIN = os.path.join(TB_DIR, "bin")
# (name, macros): host-side testbenches in the shape of the reference's *_tb.cpp, linked to libhlsd.so
HOST_TBS = [("chol_tb_256", ["-DTB_CHOL", "-Dmatrix_size=256"]), ("lu_tb_256", ["-DTB_LU", "-Dmatrix_size=256"]),
            ("chol_tb_4", ["-DTB_CHOL", "-Dmatrix_size=4"]), ("chol_tb_16", ["-DTB_CHOL", "-Dmatrix_size=16"]),
            ("chol_tb_64", ["-DTB_CHOL", "-Dmatrix_size=64"]),
            ("lu_tb_16", ["-DTB_LU", "-Dmatrix_size=16"]), ("lu_tb_64", ["-DTB_LU", "-Dmatrix_size=64"]),
            ("qr_tb_16x16", ["-DTB_QR", "-DROWS=16", "-DCOLS=16"]), ("qr_tb_64x64", ["-DTB_QR", "-DROWS=64", "-DCOLS=64"])]


def host_testbenches():
    """g++ builds of tests/host/tb.cpp against the drop-in headers (C++ host side above the C ABI)."""
    os.makedirs(HOST_BIN, exist_ok=True)
    src = os.path.join(TB_DIR, "tb.cpp")
    for name, macros in HOST_TBS:
        out = os.path.join(HOST_BIN, name)
        if os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(LIB)):
            continue
        cmd = ["g++", "-O2", "-std=c++14", "-ffp-contract=off"] + macros + ["-I" + HOST, "-I" + os.path.join(HERE, "..", "include"),
               src, "-o", out, "-L" + HERE, "-lhlsd", "-Wl,-rpath,$ORIGIN/../../../hls_designs_b200"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("host testbench build failed:\n" + res.stderr)
    return HOST_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(host_testbenches())

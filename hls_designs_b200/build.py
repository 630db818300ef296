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
SOURCES = ["chol.cu", "lu.cu", "qr.cu", "blocked.cu", "cabi.cu"]
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


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

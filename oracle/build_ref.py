#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY (oracle/): build the reference's own designs into oracle/_ref/.

For every design directory of /root/reference (Cholesky/cholesky_v{1.3,2.2,3.2,4.0},
LU/lu_v{1.1,1.2,2.1}, QR/qr_v{1.1,1.2,2.1}) this compiles, with plain g++ and the HLS header
stand-ins in oracle/shim/:

  * model4x4/<algo>.cpp            -> oracle/_ref/lib<design>_model4x4.so   (read in place)
  * template/T_<algo>.cpp @ size N -> oracle/_ref/lib<design>_<N>.so        (expanded by
    expand_template.py into oracle/_ref/gen/, which is git-ignored like the rest of _ref)
  * the matching testbench         -> oracle/_ref/tb_<design>_<N>           (run by the tests
    here against the reference's op<N>.bin to pin the shim: the reference's own pass/fail)

Each library also contains oracle/ref_wrap.cpp (a cast-and-loop C-ABI).  The reference's build
system (Vivado HLS `script.tcl`, `runit.csh`, `src_config`) is not run.  Floating point: g++ for
baseline x86-64 (no FMA instructions) with -ffp-contract=off, i.e. every float operation is
rounded separately, as in Vivado's csim (`csim_design -compiler gcc`, common/script.tcl:12).

Only this container has /root/reference; the GPU box uses the .so files built here.
"""
import concurrent.futures as cf
import math
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HLS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(OUT, "gen")
sys.path.insert(0, HERE)
from expand_template import expand  # noqa: E402

DESIGNS = {
    # design dir (relative to REF)       algo   file stem  1-D? (1-D designs stay small at large N)
    "cholesky_v1.3": ("Cholesky", "chol", True),
    "cholesky_v2.2": ("Cholesky", "chol", True),
    "cholesky_v3.2": ("Cholesky", "chol", False),
    "cholesky_v4.0": ("Cholesky", "chol", True),
    "lu_v1.1": ("LU", "lu", True),
    "lu_v1.2": ("LU", "lu", True),
    "lu_v2.1": ("LU", "lu", False),
    "qr_v1.1": ("QR", "qr", True),
    "qr_v1.2": ("QR", "qr", True),
    "qr_v2.1": ("QR", "qr", False),
}
SIZES = {
    "chol": [4, 8, 16, 25, 32, 64, 128],
    "lu": [4, 6, 8, 16, 32, 64, 128],
    "qr": [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (128, 128), (8, 4), (16, 8), (12, 5)],
}
LARGE_1D = {"chol": [256, 512], "lu": [256, 512], "qr": []}
# A 2-D array design instantiates ~N*N/2 PEs and FIFOs inside one `top()`; g++ needs tens of
# minutes on that function at N=128, so the 2-D designs are built up to MAX_2D only.
MAX_2D = 64
COMPILE_TIMEOUT_S = 600
CXX = ["g++", "-std=c++14", "-O2", "-w", "-fPIC", "-ffp-contract=off", "-fopenmp", "-I" + os.path.join(HERE, "shim")]


def bits(n):
    """Loop-counter width the reference uses: enough to hold the value N itself
    (cfg.xml: SIZE=4/BIT=3, SIZE=128/BIT=8; design_files/3x3: BIT=2)."""
    return int(math.floor(math.log2(n))) + 1


def run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=COMPILE_TIMEOUT_S)
    if res.returncode != 0:
        raise RuntimeError("FAILED: " + " ".join(cmd) + "\n" + res.stderr[-4000:])


def build_one(design, tag, src_dir, stem, with_tb=True):
    macro = {"chol": "REF_CHOL", "lu": "REF_LU", "qr": "REF_QR"}[stem]
    lib = os.path.join(OUT, f"lib{design}_{tag}.so")
    run(CXX + ["-shared", "-D" + macro, "-I" + src_dir, os.path.join(src_dir, stem + ".cpp"),
               os.path.join(HERE, "ref_wrap.cpp"), "-o", lib])
    tb = os.path.join(src_dir, stem + "_tb.cpp")
    if with_tb and os.path.exists(tb):
        run(CXX + ["-I" + src_dir, "-include", os.path.join(HERE, "shim", "tb_fopen_redirect.h"),
                   os.path.join(src_dir, stem + ".cpp"), tb, "-o", os.path.join(OUT, f"tb_{design}_{tag}")])
    return lib


FORCE = False


def job(design, size):
    algo, stem, _ = DESIGNS[design]
    ddir = os.path.join(REF, algo, design)
    tag0 = size if isinstance(size, str) else (f"{size[0]}x{size[1]}" if isinstance(size, tuple) else str(size))
    lib0 = os.path.join(OUT, f"lib{design}_{tag0}.so")
    if not FORCE and os.path.exists(lib0) and os.path.getmtime(lib0) >= os.path.getmtime(os.path.join(HERE, "ref_wrap.cpp")):
        return lib0  # already built (the reference tree is read-only and does not change)
    if size == "model4x4":
        return build_one(design, "model4x4", os.path.join(ddir, "model4x4"), stem)
    if stem == "qr":
        rows, cols = size
        tag = f"{rows}x{cols}"
        params = {"ROW": rows, "COL": cols, "BIT_R": bits(rows), "BIT_C": bits(cols)}
    else:
        tag = str(size)
        params = {"SIZE": size, "BIT": bits(size)}
    gdir = os.path.join(GEN, design, tag)
    os.makedirs(gdir, exist_ok=True)
    for suffix in (".h", ".cpp", "_tb.cpp"):
        expand(os.path.join(ddir, "template", "T_" + stem + suffix), os.path.join(gdir, stem + suffix), params)
    # The reference ships operand files only for some sizes; a testbench is built when one exists.
    opname = f"op{tag}.bin"
    return build_one(design, tag, gdir, stem, with_tb=os.path.exists(os.path.join(REF, algo, opname)))


def main():
    if not os.path.isdir(REF):
        print(f"build_ref: {REF} not present; keeping prebuilt oracle/_ref as is")
        return 0
    os.makedirs(GEN, exist_ok=True)
    global FORCE
    FORCE = "--force" in sys.argv
    only = set(a for a in sys.argv[1:] if not a.startswith("--"))
    jobs = []
    for design, (algo, stem, one_d) in DESIGNS.items():
        if only and design not in only:
            continue
        jobs.append((design, "model4x4"))
        for size in SIZES[stem] + (LARGE_1D[stem] if one_d else []):
            if not one_d and max(size if isinstance(size, tuple) else (size,)) > MAX_2D:
                continue
            jobs.append((design, size))
    failed = []
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        futs = {pool.submit(job, d, s): (d, s) for d, s in jobs}
        for fut in cf.as_completed(futs):
            try:
                print("built", os.path.basename(fut.result()), flush=True)
            except Exception as exc:  # noqa: BLE001
                failed.append((futs[fut], str(exc)))
    for (d, s), msg in failed:
        print(f"build_ref: {d} @ {s} failed:\n{msg}", file=sys.stderr)
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())

"""GPU test of the C++ host side: tests/host/tb.cpp built against the drop-in headers of hls_designs_b200/host/ (same flow as the
reference's *_tb.cpp: fread op<N>.bin, call top() five times, compare with the testbench's own
textbook loop, exit code = verdict) run against the reference's own operand files."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

BIN = os.path.join(ROOT, "tests", "host", "bin")
OPS = os.path.join(ROOT, "tests", "golden", "operands")


@pytest.mark.parametrize("tb,operand", (("chol_tb_4", "Cholesky_op4.bin"), ("chol_tb_16", "Cholesky_op16.bin"),
                                        ("chol_tb_64", "Cholesky_op64.bin"), ("lu_tb_16", "LU_op16.bin"),
                                        ("lu_tb_64", "LU_op64.bin"), ("qr_tb_16x16", "QR_op16x16.bin"),
                                        ("qr_tb_64x64", "QR_op64x64.bin"),
                                        # the drop-in top() sets HLSD_FLAG_EXACT: bit-identical at 256 too (no tensor path)
                                        ("chol_tb_256", "Cholesky_op256.bin"), ("lu_tb_256", "LU_op256.bin")))
def test_host_testbench_passes(tb, operand):
    exe = os.path.join(BIN, tb)
    assert os.path.exists(exe), "run `python -m hls_designs_b200.build` (build() does)"
    res = subprocess.run([exe, os.path.join(OPS, operand)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    # bit-identical arithmetic: the testbench's own float loop reproduces the GPU result exactly
    assert "error = 0.000000" in res.stdout, res.stdout

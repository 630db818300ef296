"""hls_designs_b200 — batched Cholesky / LU / Givens-QR for NVIDIA B200 (sm_100a) behind the
top-level interface of Sibylau/HLS_designs.  The compute lives in libhlsd.so (csrc/, C ABI in
include/hlsd.h); this package is the Python host-side mirror of the reference's functions and
operand files.  There is no CPU implementation here."""
from .api import CHOL_DESIGNS, FLAG_EXACT, FLAG_FUSED, LU_DESIGNS, PATHS, QR_DESIGNS, cholesky, launch_count, lu, qr, selftest_division, set_option  # noqa: F401
from .operand import gen_full_rank, gen_spd, operand_name, read_operand, write_operand  # noqa: F401
from ._lib import HlsdError, LIB_PATH, load  # noqa: F401

"""ctypes binding of libhlsd.so (include/hlsd.h).  The library is the product: if it is missing
there is nothing to fall back to and the import of any compute entry point fails loudly."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhlsd.so")

_fp = ctypes.POINTER(ctypes.c_float)
_ip = ctypes.POINTER(ctypes.c_int32)
_vp = ctypes.c_void_p
_i, _i64 = ctypes.c_int, ctypes.c_int64

# symbol -> (restype, argtypes); exactly the declarations of include/hlsd.h
SIGNATURES = {
    "hlsd_cholesky_host": (_i, [_vp, _vp, _i, _i64, _i, _i]),
    "hlsd_lu_host": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _i, _i]),
    "hlsd_qr_host": (_i, [_vp, _vp, _vp, _i, _i, _i64, _i, _i]),
    "hlsd_cholesky_dev": (_i, [_vp, _vp, _i, _i64, _i, _i, _vp]),
    "hlsd_lu_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _i, _i, _vp]),
    "hlsd_qr_dev": (_i, [_vp, _vp, _vp, _i, _i, _i64, _i, _i, _vp]),
    "hlsd_cholesky_host_multi": (_i, [_vp, _vp, _i, _i64, _i, _i, _vp, _i]),
    "hlsd_lu_host_multi": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _i, _i, _vp, _i]),
    "hlsd_qr_host_multi": (_i, [_vp, _vp, _vp, _i, _i, _i64, _i, _i, _vp, _i]),
    "hlsd_host_alloc": (_vp, [ctypes.c_size_t]),
    "hlsd_host_free": (None, [_vp]),
    "hlsd_release_workspace": (None, []),
    "hlsd_set_device": (_i, [_i]),
    "hlsd_device_count": (_i, []),
    "hlsd_launch_count": (_i64, []),
    "hlsd_last_error": (ctypes.c_char_p, []),
    "hlsd_version": (ctypes.c_char_p, []),
    "hlsd_set_option": (_i, [_i, _i]),
    "hlsd_profile_begin": (_i, []),
    "hlsd_profile_end": (_i, [_vp, _vp, _i]),
    "hlsd_selftest_division": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "hlsd_selftest_copy_host": (_i, [_vp, _vp, ctypes.c_size_t, ctypes.c_size_t, _i64]),
}

_lib = None


class HlsdError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"hlsd status {status}: {message}")
        self.status = status


def load():
    """Load libhlsd.so and declare every prototype.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m hls_designs_b200.build` "
                "(there is no CPU fallback behind this package)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(status):
    if status != 0:
        raise HlsdError(status, load().hlsd_last_error().decode())

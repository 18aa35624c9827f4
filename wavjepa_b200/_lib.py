"""ctypes binding of libwavjepa_b200.so (the C ABI in include/wavjepa_b200.h).

The library is the product: if it is missing, or the device is not an sm_100 part, every op raises --
there is no PyTorch / CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwavjepa_b200.so")

WJ_OK = 0


class Operand(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p),
        ("dim", C.c_int64 * 4),
        ("stride_bytes", C.c_int64 * 3),
        ("seg_width", C.c_int32),
        ("seg_q", C.c_int32 * 4),
        ("seg_p", C.c_int32 * 4),
    ]


class Epilogue(C.Structure):
    _fields_ = [
        ("out", C.c_void_p),
        ("ld_out", C.c_int64),
        ("out_f32", C.c_int32),
        ("accumulate", C.c_int32),
        ("out2", C.c_void_p),
        ("ld_out2", C.c_int64),
        ("bias", C.c_void_p),
        ("resid", C.c_void_p),
        ("resid_f32", C.c_int32),
        ("resid_mod", C.c_int32),
        ("ld_resid", C.c_int64),
        ("aux", C.c_void_p),
        ("ld_aux", C.c_int64),
        ("act", C.c_int32),
        ("_pad", C.c_int32),
        ("out_rows", C.c_void_p),
    ]


class WavJepaLibError(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = False) -> C.CDLL:
    """Loads the shared library (once).  Raises loudly when it does not exist."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build as _b

            _b.build()
        else:
            raise WavJepaLibError(
                f"{LIB_PATH} not found: build it with `python -m wavjepa_b200.build` "
                "(wavjepa_b200 has no CPU / PyTorch fallback)"
            )
    lib = C.CDLL(LIB_PATH)
    lib.wj_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != WJ_OK:
        raise WavJepaLibError(f"libwavjepa_b200 error {rc}: {load().wj_last_error().decode()}")


def require_device() -> None:
    check(load().wj_check_device())

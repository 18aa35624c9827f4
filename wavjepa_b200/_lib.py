"""ctypes binding of libwavjepa_b200.so (the C ABI in include/wavjepa_b200.h).

The library is the product: if it is missing, or the device is not an sm_100 part, every op raises --
there is no PyTorch / CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# WJ_LIB: an alternative build of the same C ABI (A/B timing of two builds on one box; the default is the in-tree library)
LIB_PATH = os.environ.get("WJ_LIB") or os.path.join(_HERE, "libwavjepa_b200.so")

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
        ("colsum", C.c_void_p),
    ]


class WavJepaLibError(RuntimeError):
    pass


_lib = None


class KernelProfile:
    """Optional per-entry-point device timing (CUDA events on the current stream around every C-ABI call), used by
    bench.py AFTER the timed region to attribute the step time to kernel families and to measure the dominant
    kernel's average duration for the roofline.  Not active unless `with KernelProfile() as kp:` is entered."""

    def __init__(self):
        self.records = []   # (entry point, start event, end event, meta, algorithmic bytes | None)
        self.meta = None
        self.nbytes = None  # set by ops.* for the HBM-bound kernels: the bytes the op must move (reads + writes)

    def __enter__(self):
        global _profile
        _profile = self
        return self

    def __exit__(self, *exc):
        global _profile
        _profile = None

    def summary(self):
        """-> {entry point: (calls, total ms)} (synchronises)."""
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1, _, _ in self.records:
            c, t = out.get(name, (0, 0.0))
            out[name] = (c + 1, t + e0.elapsed_time(e1))
        return out

    def hbm_summary(self):
        """-> {entry point: (calls, total ms, total algorithmic bytes)} over the calls that declared their bytes."""
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1, _, nb in self.records:
            if nb is None:
                continue
            c, t, b = out.get(name, (0, 0.0, 0))
            out[name] = (c + 1, t + e0.elapsed_time(e1), b + int(nb))
        return out


_profile = None


class _Proxy:
    """Attribute proxy over the CDLL so that calls can be bracketed by events when a KernelProfile is active."""

    def __init__(self, cdll):
        self._cdll = cdll

    def __getattr__(self, name):
        fn = getattr(self._cdll, name)
        if not name.startswith("wj_") or name in ("wj_last_error", "wj_version", "wj_check_device", "wj_kernel_launches"):
            setattr(self, name, fn)
            return fn

        def call(*args):
            if _profile is None:
                return fn(*args)
            import torch
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            _profile.records.append((name, e0, e1, _profile.meta, _profile.nbytes))
            _profile.meta = None
            _profile.nbytes = None
            return rc

        setattr(self, name, call)
        return call


def kernel_launches() -> int:
    """Kernels launched by libwavjepa_b200.so in this process so far."""
    return int(load().wj_kernel_launches())


def load(build_if_missing: bool = False) -> "_Proxy":
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
    cdll = C.CDLL(LIB_PATH)
    cdll.wj_last_error.restype = C.c_char_p
    cdll.wj_kernel_launches.restype = C.c_longlong
    _lib = _Proxy(cdll)
    if os.environ.get("WJ_GEMM_PAIR_GRADS") == "0" and hasattr(cdll, "wj_gemm_option"):
        # measurement switch (same-box A/B): keep the weight- / data-gradient GEMMs on single CTAs
        cdll.wj_gemm_option(1, 0)
        cdll.wj_gemm_option(2, 0)
    return _lib


def check(rc: int) -> None:
    if rc != WJ_OK:
        raise WavJepaLibError(f"libwavjepa_b200 error {rc}: {load().wj_last_error().decode()}")


def require_device() -> None:
    check(load().wj_check_device())

"""ctypes binding of libgodsp_host.so, the C++ host-side mirror of the Go API (host/godsp.hpp).
Go panics surface as GoPanic with the reference's message."""
import ctypes as C
import os

from . import _capi

LIB_PATH = os.path.join(os.path.dirname(_capi.LIB_PATH), "libgodsp_host.so")
_dp, _i64 = C.POINTER(C.c_double), C.c_int64
WINDOW_CB = C.CFUNCTYPE(None, _i64, _dp, C.c_void_p)
_lib = None


class GoPanic(RuntimeError):
    """panic(...) of the reference API (or a failed C-ABI call: there is no CPU fallback)."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libgodsp_host.so is not built (make -C go-dsp_b200); there is no CPU fallback")
        _capi.lib()                                  # dependency, resolved through rpath as well
        L = C.CDLL(LIB_PATH)
        vp, ci, u64 = C.c_void_p, C.c_int, C.c_uint64
        sig = {
            "gdh_last_panic": (C.c_char_p, []),
            "gdh_fft": (ci, [vp, _i64, vp, ci, ci]),
            "gdh_convolve": (ci, [vp, _i64, vp, _i64, vp]),
            "gdh_fft2": (ci, [C.POINTER(vp), C.POINTER(_i64), _i64, C.POINTER(vp), ci, ci]),
            "gdh_fftn": (ci, [vp, C.POINTER(_i64), ci, vp, ci]),
            "gdh_ensure_radix2_factors": (ci, [_i64]),
            "gdh_set_worker_pool_size": (None, [ci]), "gdh_worker_pool_size": (ci, []),
            "gdh_reverse_bits": (u64, [u64, u64]),
            "gdh_next_power_of2": (_i64, [_i64]), "gdh_is_power_of2": (ci, [_i64]),
            "gdh_window": (ci, [ci, _i64, vp]),
            "gdh_segment_count": (_i64, [_i64, _i64, _i64]),
            "gdh_dsputils_segment": (ci, [_i64, _i64, C.c_double, C.POINTER(_i64), C.POINTER(_i64)]),
            "gdh_pwelch": (ci, [vp, _i64, C.c_double, ci, _i64, _i64, _i64, ci, ci, WINDOW_CB, vp, vp, vp, _i64, C.POINTER(_i64)]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise GoPanic(lib().gdh_last_panic().decode("utf-8", "replace"))

"""ctypes binding of the C ABI in include/godsp_b200.h (libgodsp_b200.so).

This is the same boundary the cgo shim binds; nothing here computes.
The library has no CPU fallback: loading works anywhere, every compute call needs a B200.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libgodsp_b200.so")

_dp = C.POINTER(C.c_double)
_i64, _u64, _int, _vp, _sz = C.c_int64, C.c_uint64, C.c_int, C.c_void_p, C.c_size_t

# every symbol include/godsp_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gd_init": (_int, [_int]), "gd_shutdown": (_int, []), "gd_last_error": (C.c_char_p, []),
    "gd_device_count": (_int, []), "gd_use_device": (_int, [_int]), "gd_set_option": (_int, [C.c_char_p, _i64]),
    "gd_fft_c2c": (_int, [_vp, _vp, _i64, _int]), "gd_fft_r2c_full": (_int, [_vp, _vp, _i64, _int]),
    "gd_fft_batch_c2c": (_int, [_vp, _vp, _i64, _i64, _int]), "gd_convolve_c2c": (_int, [_vp, _vp, _vp, _i64]),
    "gd_fft2_c2c": (_int, [_vp, _vp, _i64, _i64, _int]),
    "gd_fftn_c2c": (_int, [_vp, _vp, C.POINTER(_i64), _int, _int]),
    "gd_plan_warm": (_int, [_i64]), "gd_bluestein_padded_len": (_i64, [_i64]),
    "gd_pwelch_f64": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, C.c_double, _vp]),
    "gd_pwelch_samples": (_int, [_vp, _int, _i64, _i64, _i64, _i64, _i64, _i64, _vp, C.c_double, _vp]),
    "gd_pwelch_stream_begin": (_int, [C.POINTER(_vp), _int, _i64, _i64, _i64, _i64, _vp]),
    "gd_pwelch_stream_push": (_int, [_vp, _vp, _i64]),
    "gd_pwelch_stream_end": (_int, [_vp, C.c_double, _vp, C.POINTER(_i64)]),
    "gd_stft_f64": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp]),
    "gd_fft_segments_c2c": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _vp]),
    "gd_convolve_linear_c2c": (_int, [_vp, _i64, _vp, _i64, _vp]),
    "gd_pwelch_partial_samples_dev": (_int, [_vp, _int, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp]),
    "gd_stft_f64_dev": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp]),
    "gd_convolve_linear_c2c_dev": (_int, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "gd_pinned_alloc": (_vp, [_sz]), "gd_pinned_free": (None, [_vp]),
    "gd_dev_alloc": (_int, [C.POINTER(_vp), _sz]), "gd_dev_free": (_int, [_vp]),
    "gd_memcpy_h2d": (_int, [_vp, _vp, _sz]), "gd_memcpy_d2h": (_int, [_vp, _vp, _sz]),
    "gd_stream_sync": (_int, [_vp]),
    "gd_fill_splitmix_dev": (_int, [_vp, _i64, _u64, _u64, _vp]),
    "gd_fft_batch_c2c_dev": (_int, [_vp, _vp, _i64, _i64, _int, _vp]),
    "gd_fft_batch_r2c_full_dev": (_int, [_vp, _vp, _i64, _i64, _int, _vp]),
    "gd_convolve_c2c_dev": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "gd_fftn_c2c_dev": (_int, [_vp, _vp, C.POINTER(_i64), _int, _int, _vp]),
    "gd_fft_strided_c2c_dev": (_int, [_vp, _vp, _i64, _i64, _i64, _int, _vp]),
    "gd_fourstep_twiddle_dev": (_int, [_vp, _i64, _i64, _i64, _i64, _int, _vp]),
    "gd_repack_gkw_dev": (_int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "gd_fourstep_exchange_dev": (_int, [_vp, C.POINTER(_vp), _i64, _i64, _int, _int, _int, _vp]),
    "gd_fourstep_lines_exchange_dev": (_int, [_vp, _vp, C.POINTER(_vp), _i64, _i64, _int, _int, _int, _vp]),
    "gd_fourstep_fused_supported": (_int, [_i64, _i64, _int]),
    "gd_fourstep_lines_peer_dev": (_int, [_vp, C.POINTER(_vp), _i64, _i64, _int, _int, _int, _int, _vp]),
    "gd_fourstep_rows_seg_dev": (_int, [_vp, _vp, _i64, _i64, _int, _int, _vp]),
    "gd_peer_block_copy_dev": (_int, [_vp, C.POINTER(_vp), _int, _int, _i64, _i64, _i64, _i64, _i64, _i64, _vp]),
    "gd_ipc_alloc": (_int, [C.POINTER(_vp), _sz, C.c_char_p]), "gd_ipc_open": (_int, [C.c_char_p, C.POINTER(_vp)]),
    "gd_ipc_close": (_int, [_vp]),
    "gd_transpose_batched_dev": (_int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "gd_pwelch_partial_dev": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp]),
    "gd_pwelch_finalize_dev": (_int, [_vp, _i64, _i64, C.c_double, _vp, _vp]),
    "gd_kernel_launches": (_i64, []),
    "gd_tma_profile_read": (_int, [_vp, _int]),
}

_lib = None


class GodspError(RuntimeError):
    """What the Go shim turns into panic(): a non-zero gd_status."""

    def __init__(self, status, message):
        super().__init__("go-dsp_b200 status %d: %s" % (status, message))
        self.status = status


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libgodsp_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `make -C go-dsp_b200`); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise GodspError(status, lib().gd_last_error().decode("utf-8", "replace"))


def ptr(a):
    return a.ctypes.data if isinstance(a, np.ndarray) else a

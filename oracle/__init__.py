"""ctypes binding of the CPU parity oracle (oracle/godsp_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by the product
package (go-dsp_b200/).  Names mirror the reference Go API they restate
(fft.FFT -> fft(), spectral.Pwelch -> pwelch(), ...).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgodsp_oracle.so")

WINDOWS = {"rectangular": 0, "hamming": 1, "hann": 2, "bartlett": 3, "flattop": 4, "blackman": 5}


def build(force=False):
    src = os.path.join(_HERE, "godsp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        i64, u64, dbl, ci = C.c_int64, C.c_uint64, C.c_double, C.c_int
        sig = {
            "gdo_is_pow2": (ci, [i64]), "gdo_next_pow2": (i64, [i64]),
            "gdo_reverse_bits": (u64, [u64, u64]),
            "gdo_ensure_radix2_factors": (None, [i64]), "gdo_radix2_factors": (None, [i64, _dp]),
            "gdo_bluestein_factors": (None, [i64, _dp, _dp]), "gdo_bluestein_padded_len": (i64, [i64]),
            "gdo_fft": (None, [_dp, _dp, i64]), "gdo_ifft": (None, [_dp, _dp, i64]),
            "gdo_fft_real": (None, [_dp, _dp, i64]), "gdo_ifft_real": (None, [_dp, _dp, i64]),
            "gdo_convolve": (None, [_dp, _dp, _dp, i64]),
            "gdo_fft2": (None, [_dp, _dp, i64, i64, ci]),
            "gdo_fftn": (None, [_dp, _dp, C.POINTER(i64), ci, ci]),
            "gdo_window": (ci, [ci, i64, _dp]),
            "gdo_segment_count": (i64, [i64, i64, i64]), "gdo_segment": (None, [_dp, i64, i64, i64, _dp]),
            "gdo_pwelch": (i64, [_dp, i64, dbl, i64, i64, i64, _dp, _dp, ci, _dp, _dp, ci]),
            "gdo_fft_batch": (None, [_dp, _dp, i64, i64, ci]),
            "gdo_fill_splitmix": (None, [_dp, i64, u64, u64]),
            "gdo_max_threads": (ci, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp)


def _c(x):
    return np.ascontiguousarray(x, dtype=np.complex128)


def _r(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def is_pow2(x): return bool(lib().gdo_is_pow2(int(x)))
def next_pow2(x): return int(lib().gdo_next_pow2(int(x)))
def reverse_bits(v, s): return int(lib().gdo_reverse_bits(int(v), int(s)))
def bluestein_padded_len(n): return int(lib().gdo_bluestein_padded_len(int(n)))
def max_threads(): return int(lib().gdo_max_threads())


def radix2_factors(n):
    out = np.empty(n, np.complex128)
    lib().gdo_radix2_factors(n, _p(out))
    return out


def bluestein_factors(n):
    f, inv = np.empty(n, np.complex128), np.empty(n, np.complex128)
    lib().gdo_bluestein_factors(n, _p(f), _p(inv))
    return f, inv


def _unary(fn, x, rin):
    x = _r(x) if rin else _c(x)
    out = np.empty(x.shape[0], np.complex128)
    fn(_p(x), _p(out), x.shape[0])
    return out


def fft(x): return _unary(lib().gdo_fft, x, False)
def ifft(x): return _unary(lib().gdo_ifft, x, False)
def fft_real(x): return _unary(lib().gdo_fft_real, x, True)
def ifft_real(x): return _unary(lib().gdo_ifft_real, x, True)


def convolve(x, y):
    x, y = _c(x), _c(y)
    if x.shape[0] != y.shape[0]:
        raise ValueError("arrays not of equal size")       # fft/fft.go:57
    out = np.empty_like(x)
    lib().gdo_convolve(_p(x), _p(y), _p(out), x.shape[0])
    return out


def fft2(x, inverse=False):
    x = _c(x)
    out = np.empty_like(x)
    lib().gdo_fft2(_p(x), _p(out), x.shape[0], x.shape[1], int(inverse))
    return out


def fftn(x, inverse=False):
    x = _c(x)
    out = np.empty_like(x)
    dims = (C.c_int64 * x.ndim)(*x.shape)
    lib().gdo_fftn(_p(x), _p(out), dims, x.ndim, int(inverse))
    return out


def fft_batch(x, threads=1):
    x = _c(x)
    out = np.empty_like(x)
    lib().gdo_fft_batch(_p(x), _p(out), x.shape[1], x.shape[0], threads)
    return out


def window(name, L):
    out = np.empty(max(L, 0), np.float64)
    if lib().gdo_window(WINDOWS[name], L, _p(out)) != 0:
        raise ValueError(name)
    return out


def segment_count(lx, size, noverlap): return int(lib().gdo_segment_count(lx, size, noverlap))


def segment(x, size, noverlap):
    x = _r(x)
    n = segment_count(x.shape[0], size, noverlap)
    out = np.empty((n, size), np.float64)
    if n:
        lib().gdo_segment(_p(x), x.shape[0], size, noverlap, _p(out))
    return out


def pwelch(x, fs, nfft=0, pad=0, noverlap=0, window_fn="hann", scale_off=False, threads=1):
    """spectral.Pwelch (spectral/pwelch.go:74-145). window_fn: a name from WINDOWS or a
    callable L -> array (PwelchOptions.Window); None means the default (Hann)."""
    x = _r(x)
    if x.shape[0] == 0:
        return np.empty(0), np.empty(0)
    n_eff = nfft or 256
    p_eff = pad or n_eff
    wf = (lambda L: window(window_fn or "hann", L)) if (window_fn is None or isinstance(window_fn, str)) else window_fn
    wa, wn = _r(wf(max(p_eff, n_eff))), _r(wf(n_eff))
    lp = p_eff // 2 + 1
    pxx, freqs = np.empty(lp), np.empty(lp)
    got = lib().gdo_pwelch(_p(x), x.shape[0], float(fs), nfft, pad, noverlap, _p(wa), _p(wn), int(scale_off),
                           _p(pxx), _p(freqs), threads)
    assert got == lp, (got, lp)
    return pxx, freqs


def fill_splitmix(n, seed, offset=0):
    out = np.empty(n, np.float64)
    lib().gdo_fill_splitmix(_p(out), n, seed, offset)
    return out


def splitmix_complex(n, seed, offset=0):
    """complex element i uses counters 2i (re), 2i+1 (im)."""
    return fill_splitmix(2 * n, seed, 2 * offset).view(np.complex128)

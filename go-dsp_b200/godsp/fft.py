"""Package fft (fft/fft.go, fft/radix2.go) over the B200 engine. Same names as the Go API."""
import ctypes as C

import numpy as np

from . import _host
from .dsputils import Matrix


def _c(x):
    return np.ascontiguousarray(x, dtype=np.complex128)


def _run(x, direction, real_in):
    x = np.ascontiguousarray(x, dtype=np.float64) if real_in else _c(x)
    if x.ndim != 1:
        raise ValueError("1-D input expected")
    out = np.empty(x.shape[0], np.complex128)
    _host.check(_host.lib().gdh_fft(x.ctypes.data, x.shape[0], out.ctypes.data, direction, int(real_in)))
    return out


def FFT(x): return _run(x, 1, False)            # fft/fft.go:72
def IFFT(x): return _run(x, -1, False)          # fft/fft.go:35
def FFTReal(x): return _run(x, 1, True)         # fft/fft.go:25
def IFFTReal(x): return _run(x, -1, True)       # fft/fft.go:30


def Convolve(x, y):                             # fft/fft.go:55
    x, y = _c(x), _c(y)
    out = np.empty(max(x.shape[0], 1), np.complex128)
    _host.check(_host.lib().gdh_convolve(x.ctypes.data, x.shape[0], y.ctypes.data, y.shape[0], out.ctypes.data))
    return out[: x.shape[0]]


def _fft2(x, direction, real_in):
    """x: a list of rows (like [][]complex128 -- rows may be ragged, which panics as in the reference)."""
    rows = [np.ascontiguousarray(r, dtype=np.float64 if real_in else np.complex128) for r in x]
    n = len(rows)
    outs = [np.empty(len(rows[0]) if n else 0, np.complex128) for _ in rows]
    VP = C.c_void_p * max(n, 1)
    I64 = C.c_int64 * max(n, 1)
    rp, op = VP(*[r.ctypes.data for r in rows]), VP(*[o.ctypes.data for o in outs])
    ln = I64(*[r.shape[0] for r in rows])
    _host.check(_host.lib().gdh_fft2(rp, ln, n, op, direction, int(real_in)))
    return outs


def FFT2(x): return _fft2(x, 1, False)          # fft/fft.go:109
def IFFT2(x): return _fft2(x, -1, False)        # fft/fft.go:119
def FFT2Real(x): return _fft2(x, 1, True)       # fft/fft.go:104
def IFFT2Real(x): return _fft2(x, -1, True)     # fft/fft.go:114


def _fftn(m, direction):
    out = np.empty_like(m.list)
    dims = (C.c_int64 * len(m.dims))(*m.dims)
    _host.check(_host.lib().gdh_fftn(m.list.ctypes.data, dims, len(m.dims), out.ctypes.data, direction))
    return Matrix(out, list(m.dims))


def FFTN(m): return _fftn(m, 1)                 # fft/fft.go:157
def IFFTN(m): return _fftn(m, -1)               # fft/fft.go:162


def FFTBatch(x, n, direction=1):
    """Additive batched call (go/fft/fft_b200.go FFTBatch): len(x)/n transforms back to back."""
    from . import _capi
    x = _c(x).reshape(-1)
    if n <= 0 or x.shape[0] % n:
        raise _host.GoPanic("FFTBatch: len(x) must be a multiple of n")
    out = np.empty_like(x)
    if x.shape[0]:
        _capi.check(_capi.lib().gd_fft_batch_c2c(x.ctypes.data, out.ctypes.data, n, x.shape[0] // n, direction))
    return out


def ConvolveLinear(x, h):
    """Linear convolution, len(x) + len(h) - 1 outputs, by overlap-save on the circular Convolve (fft/fft.go:55-69); what a
    go-dsp user computes as Convolve(ZeroPad(x, m), ZeroPad(h, m))[:n], n = len(x) + len(h) - 1, m = NextPowerOf2(n)."""
    from . import _capi
    x, h = _c(x).reshape(-1), _c(h).reshape(-1)
    if x.shape[0] == 0 or h.shape[0] == 0:
        return np.empty(0, np.complex128)
    out = np.empty(x.shape[0] + h.shape[0] - 1, np.complex128)
    _capi.check(_capi.lib().gd_convolve_linear_c2c(x.ctypes.data, x.shape[0], h.ctypes.data, h.shape[0], out.ctypes.data))
    return out


def FFTSegments(x, segs, noverlap):
    """fft.FFT(dsputils.ZeroPad2(s)) for every slice s of dsputils.Segment(x, segs, noverlap) (dsputils/dsputils.go:72-75,
    89-115), one batched launch; the slices are described to the GPU (offset, length), never copied."""
    from . import _capi, dsputils
    x = _c(x).reshape(-1)
    sl = dsputils.Segment(x, segs, noverlap)
    length = len(sl[0])
    step = (sl[1].ctypes.data - sl[0].ctypes.data) // 16 if segs > 1 else max(length, 1)
    fftlen = dsputils.NextPowerOf2(length)
    out = np.empty((segs, fftlen), np.complex128)
    _capi.check(_capi.lib().gd_fft_segments_c2c(x.ctypes.data, x.shape[0], length, max(step, 1), segs, fftlen, out.ctypes.data))
    return out


def SetWorkerPoolSize(n): _host.lib().gdh_set_worker_pool_size(int(n))          # fft/fft.go:95 (no effect on the GPU)
def EnsureRadix2Factors(n): _host.check(_host.lib().gdh_ensure_radix2_factors(int(n)))   # fft/radix2.go:35
def reverseBits(v, s): return int(_host.lib().gdh_reverse_bits(int(v), int(s)))  # fft/radix2.go:184

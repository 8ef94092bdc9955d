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
            "gdo_stft": (i64, [_dp, i64, i64, i64, i64, i64, _dp, _dp]),
            "gdo_dsputils_segment": (ci, [i64, i64, dbl, C.POINTER(i64), C.POINTER(i64)]),
            "gdo_wav_new": (ci, [C.c_char_p, i64, C.POINTER(i64)]),
            "gdo_wav_read_floats": (ci, [C.c_char_p, ci, i64, C.POINTER(C.c_float)]),
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


def stft(x, nfft, noverlap=0, pad=0, window_fn="hann"):
    """The segment loop of spectral.Pwelch without the accumulate (spectral/pwelch.go:104-113): [nsegs][pad/2+1] complex."""
    x = _r(x)
    pad = pad or nfft
    fftlen = max(pad, nfft)
    lp = pad // 2 + 1
    wf = (lambda L: window(window_fn or "hann", L)) if (window_fn is None or isinstance(window_fn, str)) else window_fn
    wa = _r(wf(fftlen))
    n = segment_count(x.shape[0], nfft, noverlap)
    out = np.empty((n, lp), np.complex128)
    got = lib().gdo_stft(_p(x), x.shape[0], nfft, noverlap, fftlen, lp, _p(wa), _p(out))
    assert got == n
    return out


def dsputils_segment(lx, segs, noverlap):
    """(length, step) of dsputils.Segment's slices (dsputils/dsputils.go:89-115); raises like the reference panics."""
    ln, st = C.c_int64(0), C.c_int64(0)
    if lib().gdo_dsputils_segment(lx, segs, float(noverlap), C.byref(ln), C.byref(st)) != 0:
        raise ValueError("too many segments")
    return ln.value, st.value


WAV_ERRORS = {1: "unexpected EOF", 2: "wav: missing RIFF", 3: "wav: missing WAVE", 4: "wav: bad fmt size",
              5: "wav: unknown audio format", 6: "wav: unexpected fmt chunk"}


def wav_new(data):
    """wav.New (wav/wav.go:59-110) on the bytes of a file: dict of header fields, Samples, Duration (ns), data offset/size."""
    hdr = (C.c_int64 * 10)()
    rc = lib().gdo_wav_new(bytes(data), len(data), hdr)
    if rc:
        raise ValueError(WAV_ERRORS[rc])
    keys = ("AudioFormat", "NumChannels", "SampleRate", "ByteRate", "BlockAlign", "BitsPerSample", "Samples", "Duration", "data_offset", "data_size")
    return dict(zip(keys, [int(v) for v in hdr]))


def wav_read_floats(raw, fmt, n):
    """wav.ReadFloats (wav/wav.go:138-161) on raw little-endian sample bytes; fmt: 1 float32, 2 int16, 3 uint8."""
    out = np.empty(n, np.float32)
    if lib().gdo_wav_read_floats(bytes(raw), fmt, n, out.ctypes.data_as(C.POINTER(C.c_float))) != 0:
        raise ValueError("wav: unknown type")
    return out


def convolve_linear(x, h):
    """Linear convolution the way a go-dsp user gets it: fft.Convolve (fft/fft.go:55-69) of both operands zero-padded with
    dsputils.ZeroPad to the next power of two >= len(x) + len(h) - 1 (dsputils.NextPowerOf2, as ZeroPad2 pads), truncated.
    (Padding to exactly len(x) + len(h) - 1 would route a non power of two through Bluestein, whose chirp phase is only good
    to ~N*eps rad, fft/bluestein.go:53 -- 3e-11 at N = 10^5.)"""
    x, h = _c(x), _c(h)
    n = x.shape[0] + h.shape[0] - 1
    m = next_pow2(n)
    xp, hp = np.zeros(m, np.complex128), np.zeros(m, np.complex128)
    xp[: x.shape[0]], hp[: h.shape[0]] = x, h
    return convolve(xp, hp)[:n]

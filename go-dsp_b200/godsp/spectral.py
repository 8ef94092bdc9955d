"""Package spectral (spectral/pwelch.go, spectral/spectral.go) over the B200 engine."""
import ctypes as C
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from . import _host


@dataclass
class PwelchOptions:                             # spectral/pwelch.go:28-65 (zero values = defaults)
    NFFT: int = 0
    Window: Optional[Callable[[int], np.ndarray]] = None
    Pad: int = 0
    Noverlap: int = 0
    Scale_off: bool = False


def Pwelch(x, Fs, o):                            # spectral/pwelch.go:74
    """Returns (Pxx, freqs). o=None mirrors a nil *PwelchOptions (panics for non-empty x, as in Go)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    L = _host.lib()
    nfft = (o.NFFT or 256) if o else 256
    pad = (o.Pad or nfft) if o else nfft
    cap = max(pad // 2 + 1, 1)
    pxx, freqs, lp = np.empty(cap), np.empty(cap), C.c_int64(0)
    wid, cb = -2, _host.WINDOW_CB()
    if o is not None and o.Window is not None:
        wid = getattr(o.Window, "_window_id", -1)
        if wid == -1:
            wf = o.Window

            def _cb(n, out, _ctx):
                w = np.asarray(wf(int(n)), dtype=np.float64)
                C.memmove(out, w.ctypes.data, 8 * min(len(w), int(n)))
            cb = _host.WINDOW_CB(_cb)
    rc = L.gdh_pwelch(x.ctypes.data, x.shape[0], float(Fs), int(o is not None), o.NFFT if o else 0, o.Pad if o else 0,
                      o.Noverlap if o else 0, int(o.Scale_off) if o else 0, wid, cb, None,
                      pxx.ctypes.data, freqs.ctypes.data, cap, C.byref(lp))
    _host.check(rc)
    return pxx[: lp.value].copy(), freqs[: lp.value].copy()


def Segment(x, size, noverlap):                  # spectral/spectral.go:22: deep copies
    x = np.asarray(x, dtype=np.float64)
    n = _host.lib().gdh_segment_count(len(x), size, noverlap)
    if n < 0:
        _host.check(-1)
    stride = size - noverlap
    return [x[i * stride: i * stride + size].copy() for i in range(n)]


# ---------------------------------------------------------------- additive API: formats and callers around Pwelch
SAMPLE_FORMATS = {np.dtype("float64"): 0, np.dtype("float32"): 1, np.dtype("int16"): 2, np.dtype("uint8"): 3}


def _resolve(o, n_samples):
    """option defaults, window evaluations, norm and freqs exactly as spectral.Pwelch forms them (spectral/pwelch.go:79-102,124-142)"""
    from . import window as _w
    nfft = o.NFFT or 256
    pad = o.Pad or nfft
    wf = o.Window or _w.Hann
    fftlen = max(pad, nfft)
    lp = pad // 2 + 1
    win_apply = np.ascontiguousarray(wf(fftlen), dtype=np.float64)
    norm = 0.0
    for v in np.asarray(wf(nfft), dtype=np.float64):
        norm += v * v
    return nfft, pad, fftlen, lp, win_apply, norm


def _freqs(Fs, pad, lp):
    coef = Fs / float(pad)
    return np.array([float(i) * coef for i in range(lp)])


def PwelchSamples(samples, Fs, o):
    """spectral.Pwelch on raw samples as wav.ReadSamples returns them ([]uint8, []int16, []float32) or float64: the
    wav.ReadFloats conversion (wav/wav.go:138-161) happens on the GPU in the segment load, so the signal crosses PCIe at
    1, 2 or 4 bytes per sample. Equals Pwelch(float64(ReadFloats(...)), Fs, o)."""
    from . import _capi
    x = np.ascontiguousarray(samples)
    if x.dtype not in SAMPLE_FORMATS:
        raise _host.GoPanic("PwelchSamples: unsupported sample type %s" % x.dtype)
    if x.shape[0] == 0:
        return np.empty(0), np.empty(0)
    nfft, pad, fftlen, lp, win, norm = _resolve(o, x.shape[0])
    if x.shape[0] < nfft:                        # pwelch.go:97-99: short signals are zero-padded before segmenting (host side: tiny)
        return Pwelch(_decode(x), Fs, o)
    if not o.Scale_off:
        norm *= Fs
    nsegs = _host.lib().gdh_segment_count(x.shape[0], nfft, o.Noverlap)
    pxx = np.empty(lp)
    _capi.check(_capi.lib().gd_pwelch_samples(x.ctypes.data, SAMPLE_FORMATS[x.dtype], x.shape[0], nfft, o.Noverlap, fftlen, lp, nsegs,
                                              win.ctypes.data, norm, pxx.ctypes.data))
    return pxx, _freqs(Fs, pad, lp)


def _decode(x):
    if x.dtype == np.uint8:
        return (x.astype(np.float32) / np.float32(255)).astype(np.float64)
    if x.dtype == np.int16:
        return ((x.astype(np.float32) - np.float32(-32768)) / np.float32(65535)).astype(np.float64)
    return x.astype(np.float64)


def PwelchWav(w, Fs, o, block=1 << 20):
    """spectral.Pwelch over a wav.Wav, read block by block (wav.ReadSamples) and pushed to the GPU as it arrives."""
    st = PwelchStream(o, {8: np.uint8, 16: np.int16, 32: np.float32}[w.BitsPerSample])
    left = w.Samples
    while left > 0:
        n = min(block, left)
        st.Push(w.ReadSamples(n))
        left -= n
    return st.Finish(Fs)


class PwelchStream:
    """Streaming spectral.Pwelch (gd_pwelch_stream_*): Push chunks of any size, Finish returns (Pxx, freqs) of the whole signal."""

    def __init__(self, o, dtype=np.float64):
        from . import _capi
        self._capi, self.o, self.dtype = _capi, o, np.dtype(dtype)
        self.nfft, self.pad, self.fftlen, self.lp, win, self.norm = _resolve(o, 0)
        self.h = C.c_void_p()
        _capi.check(_capi.lib().gd_pwelch_stream_begin(C.byref(self.h), SAMPLE_FORMATS[self.dtype], self.nfft, o.Noverlap, self.fftlen,
                                                       self.lp, win.ctypes.data))

    def Push(self, chunk):
        c = np.ascontiguousarray(chunk, dtype=self.dtype)
        self._capi.check(self._capi.lib().gd_pwelch_stream_push(self.h, c.ctypes.data, c.shape[0]))

    def Finish(self, Fs):
        norm = self.norm if self.o.Scale_off else self.norm * Fs
        pxx, n = np.empty(self.lp), C.c_int64(0)
        h, self.h = self.h, None
        self._capi.check(self._capi.lib().gd_pwelch_stream_end(h, norm, pxx.ctypes.data, C.byref(n)))
        self.nsegs = n.value
        if n.value == 0:
            pxx[:] = 0.0
        return pxx, _freqs(Fs, self.pad, self.lp)


def Spectrogram(x, Fs, o):
    """STFT: the segment loop of Pwelch without the accumulate (spectral/pwelch.go:104-113). Returns (S, freqs, times):
    S[c][j] = FFT(window * segment c, zero-padded)[j], j < pad/2 + 1; times = segment start / Fs."""
    from . import _capi
    x = np.ascontiguousarray(x, dtype=np.float64)
    nfft, pad, fftlen, lp, win, _ = _resolve(o, x.shape[0])
    if x.shape[0] < nfft:
        x = np.concatenate([x, np.zeros(nfft - x.shape[0])])
    nsegs = _host.lib().gdh_segment_count(x.shape[0], nfft, o.Noverlap)
    out = np.empty((max(nsegs, 0), lp), np.complex128)
    if nsegs > 0:
        _capi.check(_capi.lib().gd_stft_f64(x.ctypes.data, x.shape[0], nfft, o.Noverlap, fftlen, lp, nsegs, win.ctypes.data, out.ctypes.data))
    times = np.array([float(c * (nfft - o.Noverlap)) / Fs for c in range(max(nsegs, 0))])
    return out, _freqs(Fs, pad, lp), times

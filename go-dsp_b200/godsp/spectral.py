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

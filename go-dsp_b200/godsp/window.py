"""Package window (window/window.go): host-side generators (O(L) tables; spectral.Pwelch uploads them)."""
import numpy as np

from . import _host

_IDS = {"Rectangular": 0, "Hamming": 1, "Hann": 2, "Bartlett": 3, "FlatTop": 4, "Blackman": 5}


def _gen(name):
    def f(L):
        out = np.empty(max(int(L), 0), np.float64)
        _host.check(_host.lib().gdh_window(_IDS[name], int(L), out.ctypes.data))
        return out
    f.__name__ = name
    f._window_id = _IDS[name]
    return f


Rectangular, Hamming, Hann = _gen("Rectangular"), _gen("Hamming"), _gen("Hann")      # window.go:32,44,62
Bartlett, FlatTop, Blackman = _gen("Bartlett"), _gen("FlatTop"), _gen("Blackman")    # window.go:80,103,138


def Apply(x, windowFunction):                    # window.go:25: in place
    x *= np.asarray(windowFunction(len(x)), dtype=np.float64)[: len(x)]

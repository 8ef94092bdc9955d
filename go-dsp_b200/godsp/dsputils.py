"""Package dsputils (dsputils/dsputils.go, matrix.go, compare.go): helpers and the Matrix container.
Pure host code, as in the Go drop-in; the integer helpers go through the C++ mirror so the
Python, C++ and Go shims share one definition."""
import ctypes as C

import numpy as np

from . import _host
from ._host import GoPanic

closeFactor = 1e-8                               # compare.go:23-25


def ToComplex(x): return np.asarray(x, dtype=np.float64).astype(np.complex128)      # dsputils.go:25
def ToComplex2(x): return [ToComplex(r) for r in x]                                  # dsputils.go:77
def IsPowerOf2(x): return bool(_host.lib().gdh_is_power_of2(int(x)))                 # dsputils.go:34
def NextPowerOf2(x): return int(_host.lib().gdh_next_power_of2(int(x)))              # dsputils.go:39


def ZeroPad(x, length):                          # dsputils.go:49: returns x itself when already long enough
    x = np.asarray(x, dtype=np.complex128)
    if len(x) >= length:
        return x
    r = np.zeros(length, np.complex128)
    r[: len(x)] = x
    return r


def ZeroPadF(x, length):                         # dsputils.go:61
    x = np.asarray(x, dtype=np.float64)
    if len(x) >= length:
        return x
    r = np.zeros(length, np.float64)
    r[: len(x)] = x
    return r


def ZeroPad2(x): return ZeroPad(x, NextPowerOf2(len(x)))                             # dsputils.go:72


def Segment(x, segs, noverlap):                  # dsputils.go:89: returns views into x (aliasing, like Go sub-slices)
    offs = (C.c_int64 * max(segs, 1))()
    length = C.c_int64(0)
    _host.check(_host.lib().gdh_dsputils_segment(len(x), segs, float(noverlap), offs, C.byref(length)))
    return [x[offs[i]: offs[i] + length.value] for i in range(segs)]


def Float64Equal(a, b):                          # compare.go:94
    if abs(a - b) <= closeFactor:
        return True
    return b != 0 and abs(1 - a / b) <= closeFactor


def ComplexEqual(a, b): return Float64Equal(a.real, b.real) and Float64Equal(a.imag, b.imag)   # compare.go:84


def PrettyClose(a, b): return len(a) == len(b) and all(Float64Equal(float(p), float(q)) for p, q in zip(a, b))
def PrettyCloseC(a, b): return len(a) == len(b) and all(ComplexEqual(complex(p), complex(q)) for p, q in zip(a, b))
def PrettyClose2(a, b): return len(a) == len(b) and all(PrettyCloseC(p, q) for p, q in zip(a, b))
def PrettyClose2F(a, b): return len(a) == len(b) and all(PrettyClose(p, q) for p, q in zip(a, b))


class Matrix:
    """dsputils/matrix.go:21-216: flat row-major storage, last dimension fastest."""

    def __init__(self, x, dims):
        dims = [int(d) for d in dims]
        if any(d < 1 for d in dims):
            raise GoPanic("invalid dimensions")
        self.offsets, length = [0] * len(dims), 1
        for i in range(len(dims) - 1, -1, -1):
            self.offsets[i] = length
            length *= dims[i]
        x = np.ascontiguousarray(x, dtype=np.complex128).ravel()
        if x.shape[0] != length:
            raise GoPanic("incorrect dimensions")
        self.list, self.dims = x, dims

    def Copy(self): return Matrix(self.list.copy(), self.dims)
    def Dimensions(self): return list(self.dims)

    def _indexes(self, idx):
        i = -1
        for n, v in enumerate(idx):
            if v == -1:
                if i >= 0:
                    raise GoPanic("only one dimension index allowed")
                i = n
            elif v >= self.dims[n]:
                raise GoPanic("dimension out of bounds")
        if i == -1:
            raise GoPanic("must specify one dimension index")
        x = sum(self.offsets[n] * v for n, v in enumerate(idx) if v >= 0)
        return x + self.offsets[i] * np.arange(self.dims[i])

    def Dim(self, idx): return self.list[self._indexes(idx)].copy()

    def SetDim(self, x, idx):
        inds = self._indexes(idx)
        if len(x) != len(inds):
            raise GoPanic("incorrect array length")
        self.list[inds] = x

    def _offset(self, idx):
        if len(idx) != len(self.dims):
            raise GoPanic("incorrect dimensions")
        if any(v > self.dims[n] for n, v in enumerate(idx)):
            raise GoPanic("incorrect dimensions")
        return sum(v * self.offsets[n] for n, v in enumerate(idx))

    def Value(self, idx): return complex(self.list[self._offset(idx)])
    def SetValue(self, x, idx): self.list[self._offset(idx)] = x

    def To2D(self):
        if len(self.dims) != 2:
            raise GoPanic("can only convert 2-D Matrixes")
        return [self.list[i * self.dims[1]:(i + 1) * self.dims[1]].copy() for i in range(self.dims[0])]

    def PrettyClose(self, n): return list(self.dims) == list(n.dims) and PrettyCloseC(self.list, n.list)


def MakeMatrix(x, dims): return Matrix(x, dims)                   # matrix.go:37


def MakeMatrix2(x):                                                # matrix.go:60
    if any(len(r) != len(x[0]) for r in x):
        raise GoPanic("ragged array")
    return Matrix(np.concatenate([np.asarray(r, np.complex128) for r in x]), [len(x), len(x[0])])


def MakeEmptyMatrix(dims): return Matrix(np.zeros(int(np.prod(dims)), np.complex128), dims)   # matrix.go:83

"""Handle-based device buffers: data that stays in HBM between calls, so a chain of transforms pays
the PCIe copy once. Mirrors the DeviceBuffer type of the Go shim (go/fft/device_b200.go): an opaque handle + length."""
import ctypes as C

import numpy as np

from . import _capi


class DeviceBuffer:
    """n complex128 elements in the memory of the calling thread's GPU (gd_dev_alloc / gd_dev_free)."""

    def __init__(self, n):
        self.n = int(n)
        p = C.c_void_p()
        _capi.check(_capi.lib().gd_dev_alloc(C.byref(p), max(self.n, 1) * 16))
        self.ptr = p.value

    @classmethod
    def FromHost(cls, x):
        x = np.ascontiguousarray(x, dtype=np.complex128).reshape(-1)
        b = cls(x.shape[0])
        b.Upload(x)
        return b

    def Upload(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex128).reshape(-1)
        if x.shape[0] != self.n:
            raise ValueError("DeviceBuffer.Upload: length mismatch")
        if self.n:
            _capi.check(_capi.lib().gd_memcpy_h2d(self.ptr, x.ctypes.data, self.n * 16))

    def Download(self):
        out = np.empty(self.n, np.complex128)
        if self.n:
            _capi.check(_capi.lib().gd_memcpy_d2h(out.ctypes.data, self.ptr, self.n * 16))
        return out

    def _unary(self, fn_name, *args):
        out = DeviceBuffer(self.n)
        L = _capi.lib()
        _capi.check(getattr(L, fn_name)(self.ptr, out.ptr, *args, None))
        _capi.check(L.gd_stream_sync(None))
        return out

    def FFT(self, n=None, direction=1):
        """len/n transforms of n points back to back (n defaults to the whole buffer): fft.FFT / fft.IFFT on resident data"""
        n = self.n if n is None else int(n)
        if n <= 0 or self.n % n:
            raise ValueError("DeviceBuffer.FFT: length must be a multiple of n")
        return self._unary("gd_fft_batch_c2c_dev", n, self.n // n, direction)

    def IFFT(self, n=None):
        return self.FFT(n, -1)

    def FFTN(self, dims, direction=1):
        dims = [int(v) for v in dims]
        if int(np.prod(dims)) != self.n:
            raise ValueError("DeviceBuffer.FFTN: dims do not match the length")
        arr = (C.c_int64 * len(dims))(*dims)
        out = DeviceBuffer(self.n)
        L = _capi.lib()
        _capi.check(L.gd_fftn_c2c_dev(self.ptr, out.ptr, arr, len(dims), direction, None))
        _capi.check(L.gd_stream_sync(None))
        return out

    def Convolve(self, other):
        if other.n != self.n:
            raise ValueError("arrays not of equal size")          # fft/fft.go:57
        out = DeviceBuffer(self.n)
        L = _capi.lib()
        _capi.check(L.gd_convolve_c2c_dev(self.ptr, other.ptr, out.ptr, self.n, None))
        _capi.check(L.gd_stream_sync(None))
        return out

    def Free(self):
        if self.ptr:
            _capi.check(_capi.lib().gd_dev_free(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.Free()
        except Exception:
            pass

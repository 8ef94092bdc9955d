"""godsp -- host-side mirror of go-dsp's exported API over the B200 C ABI: packages fft, spectral,
window and dsputils with the Go names. Importing never touches the GPU; every transform call
does, and fails loudly without one (there is no CPU fallback)."""
from . import _capi, _host  # noqa: F401
from . import dsputils, fft, spectral, wav, window  # noqa: F401
from .device import DeviceBuffer  # noqa: F401
from ._host import GoPanic  # noqa: F401

"""godsp -- host-side mirror of go-dsp's exported API over the B200 C ABI (see fft.py,
spectral.py, window.py, dsputils.py). Importing never touches the GPU; every transform
call does, and fails loudly without one (no CPU fallback)."""
from . import _capi  # noqa: F401

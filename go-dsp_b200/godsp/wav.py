"""Package wav (wav/wav.go): RIFF/WAVE reader with the reference's API -- New, Wav.ReadSamples, Wav.ReadFloats. Header
parsing and byte shuffling stay on the host, as in the Go drop-in (which keeps go-dsp's own wav.go); what is new is that
the raw samples can go to the GPU as they are on disk: spectral.PwelchWav / PwelchSamples decode them in the segment load
of the Pwelch kernel exactly as ReadFloats does."""
import io
import struct

import numpy as np

wavFormatPCM, wavFormatIEEEFloat = 1, 3          # wav/wav.go:34-37


class WavError(Exception):
    """the error value the reference returns"""


class Wav:                                       # wav/wav.go:50-58
    def __init__(self):
        self.AudioFormat = self.NumChannels = self.SampleRate = self.ByteRate = self.BlockAlign = self.BitsPerSample = 0
        self.Samples = 0
        self.Duration = 0                        # nanoseconds (time.Duration)
        self._r = None
        self._left = 0

    def header(self):
        return {k: getattr(self, k) for k in ("AudioFormat", "NumChannels", "SampleRate", "ByteRate", "BlockAlign", "BitsPerSample")}

    def _dtype(self):                            # wav/wav.go:115-130
        if self.AudioFormat == wavFormatPCM:
            if self.BitsPerSample == 8:
                return np.dtype("u1")
            if self.BitsPerSample == 16:
                return np.dtype("<i2")
            raise WavError("wav: unknown bits per sample: %d" % self.BitsPerSample)
        if self.AudioFormat == wavFormatIEEEFloat:
            return np.dtype("<f4")
        raise WavError("wav: unknown audio format")

    def ReadSamples(self, n):                    # wav/wav.go:113-136: []uint8, []int16 or []float32
        dt = self._dtype()
        want = n * dt.itemsize
        raw = self._r.read(min(want, self._left))
        self._left -= len(raw)
        if len(raw) < want:
            raise WavError("EOF" if not raw else "unexpected EOF")
        return np.frombuffer(raw, dtype=dt).copy()

    def ReadFloats(self, n):                     # wav/wav.go:138-161, float32 arithmetic
        d = self.ReadSamples(n)
        if d.dtype == np.uint8:
            return d.astype(np.float32) / np.float32(255)
        if d.dtype == np.dtype("<i2"):
            return (d.astype(np.float32) - np.float32(-32768)) / np.float32(65535)
        return d


def _read_full(r, n):
    b = r.read(n)
    if len(b) < n:
        raise WavError("EOF" if not b else "unexpected EOF")
    return b


def New(r):                                      # wav/wav.go:59-110
    if isinstance(r, (bytes, bytearray, memoryview)):
        r = io.BytesIO(bytes(r))
    w = Wav()
    head = _read_full(r, 12)
    if head[0:4] != b"RIFF":
        raise WavError("wav: missing RIFF")
    if head[8:12] != b"WAVE":
        raise WavError("wav: missing WAVE")
    has_fmt = False
    while True:
        ch = _read_full(r, 8)
        sz = struct.unpack("<I", ch[4:])[0]
        typ = ch[:4]
        if typ == b"fmt ":
            if sz < 16:
                raise WavError("wav: bad fmt size")
            f = _read_full(r, sz)
            (w.AudioFormat, w.NumChannels, w.SampleRate, w.ByteRate, w.BlockAlign, w.BitsPerSample) = struct.unpack("<HHIIHH", f[:16])
            if w.AudioFormat not in (wavFormatPCM, wavFormatIEEEFloat):
                raise WavError("wav: unknown audio format: %02x" % w.AudioFormat)
            has_fmt = True
        elif typ == b"data":
            if not has_fmt:
                raise WavError("wav: unexpected fmt chunk")
            w.Samples = sz // w.BitsPerSample * 8                                  # int(sz) / int(BitsPerSample) * 8
            w.Duration = w.Samples * 1000000000 // w.SampleRate // w.NumChannels     # Duration(Samples) * Second / rate / channels
            w._r, w._left = r, sz
            return w
        else:
            r.read(sz)                           # io.CopyN(ioutil.Discard, r, sz)

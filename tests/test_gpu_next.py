"""GPU parity tests of the SURVEY.md 8(f) rows -- the callers and data formats either side of the hot path -- through the
C ABI and the mirror of the Go API, against the CPU oracle: wav ingest fused into the Pwelch segment load, streaming
Pwelch, STFT / spectrogram, dsputils.Segment descriptors, overlap-save linear convolution, device-resident buffers."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gd():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: these tests must run on the GPU box (there is no CPU fallback)")
    import godsp
    from godsp import _capi
    L = _capi.lib()
    _capi.check(L.gd_use_device(0))
    return godsp, _capi, L


def _raw(dtype, n, seed):
    r = np.random.default_rng(seed)
    if dtype == np.uint8:
        return r.integers(0, 256, n, dtype=np.uint8)
    if dtype == np.int16:
        return r.integers(-32768, 32768, n, dtype=np.int16)
    return (r.random(n, dtype=np.float32) * 2 - 1).astype(np.float32)


def _floats(raw):
    fmt = {np.dtype("float32"): 1, np.dtype("int16"): 2, np.dtype("uint8"): 3}[raw.dtype]
    return oracle.wav_read_floats(raw.astype(raw.dtype.newbyteorder("<")).tobytes(), fmt, raw.size).astype(np.float64)


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.float32, np.float64])
@pytest.mark.parametrize("nx,nfft,nov,pad", [(200000, 4096, 2048, 0), (50001, 1024, 100, 2048), (30000, 100, 30, 0), (9000, 256, -64, 0)])
def test_pwelch_samples(gd, dtype, nx, nfft, nov, pad):      # wav.ReadFloats fused into the segment load (wav/wav.go:138-161)
    godsp = gd[0]
    raw = oracle.fill_splitmix(nx, 11) if dtype == np.float64 else _raw(dtype, nx, 7)
    x = raw if dtype == np.float64 else _floats(raw)
    o = godsp.spectral.PwelchOptions(NFFT=nfft, Noverlap=nov, Pad=pad)
    p, f = godsp.spectral.PwelchSamples(raw, 44100.0, o)
    pw, fw = oracle.pwelch(x, 44100.0, nfft=nfft, noverlap=nov, pad=pad)
    assert len(p) == len(pw) and np.array_equal(f.view(np.uint64), fw.view(np.uint64))
    assert rel_l2(p, pw) <= TOL


def test_pwelch_negative_noverlap_mirror(gd):    # spectral.Segment accepts noverlap < 0 (gapped segments)
    godsp = gd[0]
    x = oracle.fill_splitmix(40000, 3)
    for nfft, nov in ((256, -100), (4096, -1), (1000, -3000)):
        p, f = godsp.spectral.Pwelch(x, 1.0, godsp.spectral.PwelchOptions(NFFT=nfft, Noverlap=nov))
        pw, fw = oracle.pwelch(x, 1.0, nfft=nfft, noverlap=nov)
        assert rel_l2(p, pw) <= TOL and np.array_equal(f, fw)


def test_pwelch_wav_fixtures(gd):                # the reference's own wav fixtures, streamed block by block
    godsp = gd[0]
    for fn, nread in (("small.wav", None), ("float_head.wav", 16384)):
        data = open(os.path.join(GOLDEN, fn), "rb").read()
        w = godsp.wav.New(data)
        if nread is not None:
            w.Samples = nread                     # the cut fixture holds its first 16384 samples only
        o = godsp.spectral.PwelchOptions(NFFT=1024, Noverlap=512)
        p, f = godsp.spectral.PwelchWav(w, float(w.SampleRate), o, block=5000)
        h = oracle.wav_new(data)
        n = h["Samples"] if nread is None else nread
        fmt = 1 if h["AudioFormat"] == 3 else (2 if h["BitsPerSample"] == 16 else 3)
        x = oracle.wav_read_floats(data[h["data_offset"]:], fmt, n).astype(np.float64)
        pw, fw = oracle.pwelch(x, float(h["SampleRate"]), nfft=1024, noverlap=512)
        assert rel_l2(p, pw) <= TOL and np.array_equal(f, fw)
        # ReadFloats of the mirror is bit-identical to the oracle's restatement
        w2 = godsp.wav.New(data)
        assert np.array_equal(w2.ReadFloats(1000), oracle.wav_read_floats(data[h["data_offset"]:], fmt, 1000))


@pytest.mark.parametrize("nfft,nov", [(4096, 2048), (256, 0), (300, 77), (512, -200)])
def test_pwelch_stream_equals_one_shot(gd, nfft, nov):
    godsp = gd[0]
    x = oracle.fill_splitmix(300000, 13)
    o = godsp.spectral.PwelchOptions(NFFT=nfft, Noverlap=nov)
    st = godsp.spectral.PwelchStream(o)
    r = np.random.default_rng(5)
    pos = 0
    while pos < len(x):
        n = int(r.integers(1, 40000))
        st.Push(x[pos:pos + n])
        pos += n
    p, f = st.Finish(8000.0)
    pw, fw = oracle.pwelch(x, 8000.0, nfft=nfft, noverlap=nov)
    assert st.nsegs == oracle.segment_count(len(x), nfft, nov)
    assert rel_l2(p, pw) <= TOL and np.array_equal(f, fw)


@pytest.mark.parametrize("nx,nfft,nov,pad,win", [(20000, 256, 128, 0, None), (100000, 4096, 2048, 0, "Hann"), (5000, 100, 30, 0, "Blackman"),
                                                 (9000, 512, 0, 1024, "Hamming"), (70000, 8192, 4096, 0, None), (3000, 300, 100, 1000, None),
                                                 (40000, 1024, -500, 0, "FlatTop")])
def test_spectrogram(gd, nx, nfft, nov, pad, win):          # STFT: Pwelch's segment loop without the accumulate
    godsp = gd[0]
    x = oracle.fill_splitmix(nx, 17)
    o = godsp.spectral.PwelchOptions(NFFT=nfft, Noverlap=nov, Pad=pad, Window=getattr(godsp.window, win) if win else None)
    S, f, t = godsp.spectral.Spectrogram(x, 2.0, o)
    want = oracle.stft(x, nfft, nov, pad, (win or "hann").lower())
    assert S.shape == want.shape and rel_l2(S, want) <= TOL
    assert len(t) == S.shape[0] and t[1] == (nfft - nov) / 2.0


@pytest.mark.parametrize("n,segs,nov", [(16, 3, 0.5), (100000, 37, 0.25), (5000, 5, 0.0), (70000, 9, 0.9)])
def test_fft_segments(gd, n, segs, nov):         # dsputils.Segment slices as gather descriptors, ZeroPad2, one batched transform
    godsp = gd[0]
    x = oracle.splitmix_complex(n, 21)
    got = godsp.fft.FFTSegments(x, segs, nov)
    length, step = oracle.dsputils_segment(n, segs, nov)
    fl = oracle.next_pow2(length)
    for i in range(segs):
        s = np.zeros(fl, np.complex128)
        s[:length] = x[i * step: i * step + length]
        assert rel_l2(got[i], oracle.fft(s)) <= TOL


@pytest.mark.parametrize("nx,nh", [(1, 1), (5, 3), (3, 50), (1000, 17), (100000, 1000), (1 << 20, 4097), (300000, 70000)])
def test_convolve_linear(gd, nx, nh):            # overlap-save on top of fft.Convolve (fft/fft.go:55-69)
    godsp = gd[0]
    x, h = oracle.splitmix_complex(nx, 1), oracle.splitmix_complex(nh, 2)
    got = godsp.fft.ConvolveLinear(x, h)
    want = oracle.convolve_linear(x, h)
    assert got.shape == want.shape and rel_l2(got, want) <= TOL


def test_device_buffer_chain(gd):                # resident data: FFT -> IFFT -> Convolve without leaving HBM
    godsp = gd[0]
    n = 1 << 16
    x, y = oracle.splitmix_complex(n, 1), oracle.splitmix_complex(n, 2)
    bx, by = godsp.DeviceBuffer.FromHost(x), godsp.DeviceBuffer.FromHost(y)
    X = bx.FFT()
    assert rel_l2(X.Download(), oracle.fft(x)) <= TOL
    assert rel_l2(X.IFFT().Download(), x) <= TOL
    assert rel_l2(bx.Convolve(by).Download(), oracle.convolve(x, y)) <= TOL
    rows = godsp.DeviceBuffer.FromHost(oracle.splitmix_complex(8 * 4096, 3))
    assert rel_l2(rows.FFT(4096).Download(), oracle.fft_batch(oracle.splitmix_complex(8 * 4096, 3).reshape(8, 4096)).ravel()) <= TOL
    m = oracle.splitmix_complex(64 * 32, 4).reshape(64, 32)
    assert rel_l2(godsp.DeviceBuffer.FromHost(m).FFTN([64, 32]).Download(), oracle.fft2(m).ravel()) <= TOL
    for b in (bx, by, X, rows):
        b.Free()

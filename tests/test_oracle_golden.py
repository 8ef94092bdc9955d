"""Pins the CPU oracle (oracle/godsp_oracle.c) against every golden vector the
reference's own tests hold for the FFT / Pwelch path (tests/golden/reference_vectors.json,
extracted from go-dsp's *_test.go files), with the reference's own tolerance
(dsputils.Float64Equal, 1e-8 abs-or-rel).  CPU only."""
import math
import os

import numpy as np
import pytest

import oracle
from conftest import cplx, float64_equal, pretty_close, rel_l2


def test_fft_vectors(golden):           # fft/fft_test.go:197-209 TestFFT
    for c in golden["fft"]["cases"]:
        x, want = np.array(c["in"], float), cplx(c["out"])
        assert pretty_close(oracle.fft_real(x), want), c
        assert pretty_close(oracle.ifft(want), x.astype(complex)), c


def test_fft2_vectors(golden):          # fft/fft_test.go:211-223 TestFFT2
    for c in golden["fft2"]["cases"]:
        x, want = np.array(c["in"], float), cplx(c["out"])
        assert pretty_close(oracle.fft2(x.astype(complex)), want)
        assert pretty_close(oracle.fft2(want, inverse=True), x.astype(complex))


def test_fftn_vectors(golden):          # fft/fft_test.go:225-239 TestFFTN
    for c in golden["fftn"]["cases"]:
        x = np.array(c["in"], float).reshape(c["dim"]).astype(complex)
        want = cplx(c["out"]).reshape(c["dim"])
        assert pretty_close(oracle.fftn(x), want)
        assert pretty_close(oracle.fftn(want, inverse=True), x)


def test_reverse_bits(golden):          # fft/fft_test.go:241-249 TestReverseBits (bit-exact)
    for c in golden["reverse_bits"]["cases"]:
        assert oracle.reverse_bits(c["in"], c["sz"]) == c["out"]


def test_example_fft_real(golden):      # fft/fft_test.go:283-320 ExampleFFTReal
    g = golden["example_fft_real"]
    x = [math.sin(2 * math.pi * n / 8) + 0.5 * math.sin(2 * math.pi * n / 4 + 3 * math.pi / 4) for n in range(8)]
    X = oracle.fft_real(x)
    for i in range(8):
        r, th = abs(X[i]), math.degrees(math.atan2(X[i].imag, X[i].real))
        if float64_equal(r, 0):
            th = 0.0
        assert "%.1f" % r == "%.1f" % g["mag"][i]
        assert "%.1f" % th == "%.1f" % g["phase_deg"][i]


def test_fft_multi():                   # fft/fft_test.go:251-259 TestFFTMulti
    n = 256
    x = np.arange(n) / n
    assert rel_l2(oracle.fft(x.astype(complex)), np.fft.fft(x)) < 1e-14


def test_pwelch_vectors(golden):        # spectral/pwelch_test.go:48-60 TestPwelch
    for c in golden["pwelch"]["cases"]:
        p, f = oracle.pwelch(np.array(c["x"], float), c["fs"])
        assert pretty_close(p, c["p"]) and pretty_close(f, c["freqs"])
        assert len(p) == len(c["p"])


def test_segment_vectors(golden):       # spectral/spectral_test.go:58-67 TestSegment
    g = golden["spectral_segment"]
    for c in g["cases"]:
        got = oracle.segment(np.array(g["x"], float), c["size"], c["noverlap"])
        assert got.shape == np.array(c["out"]).shape and np.array_equal(got, np.array(c["out"], float))


def test_window_vectors(golden):        # window/window_test.go:61-94 TestWindowFunctions
    for c in golden["window"]["cases"]:
        for name in ("hamming", "hann", "bartlett", "flattop", "blackman"):
            assert pretty_close(oracle.window(name, c["L"]), c[name]), (name, c["L"])
        assert pretty_close(oracle.window("rectangular", c["L"]) * oracle.window("hamming", c["L"]), c["hamming"])


def test_next_pow2_bit_exact():         # dsputils/dsputils.go:39-45 (float formula) vs integer ceil-pow2
    for x in list(range(1, 5000)) + [2**k + d for k in range(3, 40) for d in (-1, 0, 1)] + [1999999, 2000005]:
        want = 1 << (x - 1).bit_length() if x > 1 else 1
        assert oracle.next_pow2(x) == want, x
    assert oracle.bluestein_padded_len(1000003) == 1 << 21
    assert oracle.next_pow2(0) == 0       # IsPowerOf2(0) is true in the reference (dsputils.go:34-36)


def test_radix2_factor_seeds():         # fft/radix2.go:27-29,56-58: exact 1,-i,-1,i propagate to every table
    for n in (4, 8, 64, 1024):
        f = oracle.radix2_factors(n)
        assert f[0] == 1 and f[n // 4] == -1j and f[n // 2] == -1 and f[3 * n // 4] == 1j
        k = np.arange(n)
        assert np.max(np.abs(f - np.exp(-2j * np.pi * k / n))) < 1e-15


@pytest.mark.parametrize("n", [2, 4, 8, 16, 64, 256, 1024, 1 << 14, 3, 5, 6, 7, 12, 100, 1000, 4099])
def test_fft_matches_numpy(n):          # sanity beyond the reference's vectors (N > 8 is unpinned there)
    x = oracle.splitmix_complex(n, 7)
    tol = 1e-14 if oracle.is_pow2(n) else 1e-11    # Bluestein chirp-phase rounding grows ~N
    assert rel_l2(oracle.fft(x), np.fft.fft(x)) < tol
    assert rel_l2(oracle.ifft(x), np.fft.ifft(x)) < tol


def test_convolve_and_roundtrips():
    x, y = oracle.splitmix_complex(48, 1), oracle.splitmix_complex(48, 2)
    want = np.fft.ifft(np.fft.fft(x) * np.fft.fft(y))
    assert rel_l2(oracle.convolve(x, y), want) < 1e-12
    r = oracle.fill_splitmix(1003, 3)
    assert rel_l2(oracle.ifft(oracle.fft_real(r)), r.astype(complex)) < 1e-11
    assert rel_l2(oracle.fft(oracle.ifft_real(r)), r.astype(complex)) < 1e-11


def test_pwelch_multisegment_against_direct_numpy():
    x = oracle.fill_splitmix(5000, 5)
    nfft, nov = 256, 128
    p, f = oracle.pwelch(x, 1.0, nfft=nfft, noverlap=nov)
    w = oracle.window("hann", nfft)
    nseg = (len(x) - nfft) // (nfft - nov) + 1
    acc = np.zeros(nfft // 2 + 1)
    for s in range(nseg):
        X = np.fft.fft(x[s * (nfft - nov): s * (nfft - nov) + nfft] * w)[: nfft // 2 + 1]
        d = np.abs(X) ** 2 / nseg
        d[1:-1] *= 2
        acc += d
    acc /= np.sum(w ** 2)
    assert rel_l2(p, acc) < 1e-13 and len(f) == nfft // 2 + 1 and f[1] == 1.0 / nfft
    p2, _ = oracle.pwelch(x, 1.0, nfft=nfft, noverlap=nov, threads=4)
    assert rel_l2(p2, p) < 1e-14


def test_splitmix_matches_numpy():
    n, seed = 1000, 5
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (i + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    want = (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53 * 2 - 1
    assert np.array_equal(oracle.fill_splitmix(n, seed), want)


# ----------------------------------------------------------------- wav ingest, dsputils.Segment
def test_wav_header_golden(golden):              # wav/wav_test.go:66-98 TestWav, on the copied fixtures
    g = golden["wav"]
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name, fn in g["files"].items():
        data = open(os.path.join(here, fn), "rb").read()
        h = oracle.wav_new(data)
        for k, v in g["cases"][name].items():
            assert h[k] == v, (name, k, h[k], v)
        assert h["data_offset"] == 44


def test_wav_header_errors():                    # wav/wav.go:62-70,80-82,93,99: the reference's error returns
    import pytest
    riff = b"RIFF\0\0\0\0WAVE"
    for data, msg in ((b"RIF", "EOF"), (b"XXXX\0\0\0\0WAVE", "missing RIFF"), (b"RIFF\0\0\0\0WAVX", "missing WAVE"),
                      (riff + b"fmt \x08\0\0\0" + b"\0" * 8, "bad fmt size"),
                      (riff + b"fmt \x10\0\0\0" + b"\x07\0" + b"\0" * 14, "unknown audio format"),
                      (riff + b"data\x04\0\0\0\0\0\0\0", "unexpected fmt chunk"), (riff + b"JUNK\x02\0\0\0ab", "EOF")):
        with pytest.raises(ValueError, match=msg):
            oracle.wav_new(data)


def test_wav_read_floats_exact():                # wav/wav.go:138-161: float32 arithmetic, every int16 / uint8 value
    i16 = np.arange(-32768, 32768, dtype=np.int16)
    got = oracle.wav_read_floats(i16.astype("<i2").tobytes(), 2, i16.size)
    want = (i16.astype(np.float32) - np.float32(-32768)) / np.float32(65535)
    assert np.array_equal(got, want) and got[0] == 0.0 and got[-1] == 1.0
    u8 = np.arange(256, dtype=np.uint8)
    assert np.array_equal(oracle.wav_read_floats(u8.tobytes(), 3, 256), u8.astype(np.float32) / np.float32(255))
    f = np.array([0.5, -1.25, 3e-8], np.float32)
    assert np.array_equal(oracle.wav_read_floats(f.tobytes(), 1, 3), f)


def test_dsputils_segment_golden(golden):        # dsputils/dsputils_test.go:28-57
    g = golden["dsputils_segment"]
    for c in g["cases"]:
        length, step = oracle.dsputils_segment(g["n"], c["segs"], c["noverlap"])
        assert [[i * step, i * step + length] for i in range(c["segs"])] == c["slices"]
    import pytest
    with pytest.raises(ValueError):
        oracle.dsputils_segment(4, 9, 0.0)       # dsputils.go:103-105 panic("too many segments")


def test_stft_is_pwelch_without_the_sum():       # spectral/pwelch.go:104-122: Pxx from the per-segment spectra
    x = oracle.fill_splitmix(5000, 9)
    S = oracle.stft(x, 256, 128, pad=512, window_fn="hamming")
    p, _ = oracle.pwelch(x, 2.0, nfft=256, pad=512, noverlap=128, window_fn="hamming")
    d = (S.real ** 2 + S.imag ** 2) / S.shape[0]
    d[:, 1:-1] *= 2.0
    norm = float(np.sum(oracle.window("hamming", 256) ** 2)) * 2.0
    assert np.allclose(d.sum(0) / norm, p, rtol=1e-13, atol=0)


def test_convolve_linear_matches_direct():
    x, h = oracle.splitmix_complex(300, 1), oracle.splitmix_complex(17, 2)
    assert np.allclose(oracle.convolve_linear(x, h), np.convolve(x, h), rtol=0, atol=1e-12)

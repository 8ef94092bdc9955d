"""GPU parity tests (pytest -m gpu, on the B200 box): the CUDA path, called through the C ABI
(include/godsp_b200.h) and through the mirror of the Go API, against the CPU oracle on the same
seeded inputs -- bit-exact for integer results, relative L2 <= 1e-12 for spectra and PSDs
(BASELINE.json north_star) -- plus the reference's own golden vectors with its own tolerance
(dsputils.Float64Equal, 1e-8) and size-independent properties at the benchmark sizes."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle
from conftest import cplx, float64_equal, pretty_close, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12                                       # north_star: relative L2 <= 1e-12 (complex128)


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


# ----------------------------------------------------------------- the reference's own test tables
def test_reference_TestFFT(gd, golden):          # fft/fft_test.go:197-209
    godsp = gd[0]
    for c in golden["fft"]["cases"]:
        x, want = np.array(c["in"], float), cplx(c["out"])
        assert pretty_close(godsp.fft.FFTReal(x), want), c
        assert pretty_close(godsp.fft.IFFT(want), x.astype(complex)), c


def test_reference_TestFFT2(gd, golden):         # fft/fft_test.go:211-223
    godsp = gd[0]
    for c in golden["fft2"]["cases"]:
        want = cplx(c["out"])
        v = godsp.fft.FFT2Real(c["in"])
        assert godsp.dsputils.PrettyClose2(v, list(want))
        vi = godsp.fft.IFFT2(list(want))
        assert godsp.dsputils.PrettyClose2(vi, godsp.dsputils.ToComplex2(c["in"]))


def test_reference_TestFFTN(gd, golden):         # fft/fft_test.go:225-239
    godsp = gd[0]
    for c in golden["fftn"]["cases"]:
        m = godsp.dsputils.MakeMatrix(godsp.dsputils.ToComplex(c["in"]), c["dim"])
        o = godsp.dsputils.MakeMatrix(cplx(c["out"]), c["dim"])
        assert godsp.fft.FFTN(m).PrettyClose(o)
        assert godsp.fft.IFFTN(o).PrettyClose(m)


def test_reference_TestFFTMulti_and_Example(gd, golden):   # fft/fft_test.go:251-259, 283-320
    godsp = gd[0]
    n = 256
    a = (np.arange(n) / n).astype(complex)
    assert rel_l2(godsp.fft.FFT(a), oracle.fft(a)) <= TOL
    g = golden["example_fft_real"]
    x = [math.sin(2 * math.pi * i / 8) + 0.5 * math.sin(2 * math.pi * i / 4 + 3 * math.pi / 4) for i in range(8)]
    X = godsp.fft.FFTReal(x)
    for i in range(8):
        r, th = abs(X[i]), math.degrees(math.atan2(X[i].imag, X[i].real))
        if float64_equal(r, 0):
            th = 0.0
        assert "%.1f" % r == "%.1f" % g["mag"][i] and "%.1f" % th == "%.1f" % g["phase_deg"][i]


def test_reference_TestPwelch(gd, golden):       # spectral/pwelch_test.go:48-60
    godsp = gd[0]
    for c in golden["pwelch"]["cases"]:
        p, f = godsp.spectral.Pwelch(np.array(c["x"], float), c["fs"], godsp.spectral.PwelchOptions())
        assert pretty_close(p, c["p"]) and pretty_close(f, c["freqs"]) and len(p) == len(c["p"])


def test_benchmark_input(gd):                    # fft/fft_test.go:262-280 BenchmarkFFT input, N = 2^20
    godsp = gd[0]
    n = 1 << 20
    a = (np.arange(n) / n).astype(complex)
    godsp.fft.EnsureRadix2Factors(n)
    assert rel_l2(godsp.fft.FFT(a), oracle.fft(a)) <= TOL


# ----------------------------------------------------------------- 1-D transforms vs the oracle
@pytest.mark.parametrize("lg", list(range(1, 22)))
def test_pow2_sizes(gd, lg):
    godsp = gd[0]
    n = 1 << lg
    x = oracle.splitmix_complex(n, 1)
    want = oracle.fft(x)
    assert rel_l2(godsp.fft.FFT(x), want) <= TOL
    assert rel_l2(godsp.fft.IFFT(want), x) <= TOL
    assert rel_l2(godsp.fft.IFFT(x), oracle.ifft(x)) <= TOL


@pytest.mark.parametrize("lg", [21, 22, 23, 24, 25])
def test_pow2_sizes_large(gd, lg):               # single transforms of 2^21 .. 2^25 points: outer four-step over the fused family
    godsp, capi, L = gd
    n = 1 << lg
    x = oracle.splitmix_complex(n, 1)
    want = oracle.fft(x)

    def both():
        assert rel_l2(godsp.fft.FFT(x), want) <= TOL
        assert rel_l2(godsp.fft.IFFT(want), x) <= TOL

    capi.check(L.gd_set_option(b"huge_min_log2n", 21))
    try:
        both()                                              # two sweeps: columns with the twiddle on their stores, rows with the transposed store
        for l1 in (13, 14, 15, 17):                         # every column length of the family
            capi.check(L.gd_set_option(b"huge_l1", l1))
            assert rel_l2(godsp.fft.FFT(x), want) <= TOL
        capi.check(L.gd_set_option(b"huge_l1", 0))
        for sweeps in (3, 4):                               # columns, transpose, columns / columns, twiddle, rows, transpose
            capi.check(L.gd_set_option(b"huge_sweeps", sweeps))
            both()
        capi.check(L.gd_set_option(b"huge_sweeps", 0))
        if lg <= 24:
            capi.check(L.gd_set_option(b"huge_min_log2n", 25))  # and the two-pass schedule
            both()
    finally:
        capi.check(L.gd_set_option(b"huge_sweeps", 0))
        capi.check(L.gd_set_option(b"huge_l1", 0))
        capi.check(L.gd_set_option(b"huge_min_log2n", 22))


@pytest.mark.parametrize("lg,b", [(21, 5), (22, 3), (23, 2)])
def test_pow2_large_batched(gd, lg, b):           # a batch of large transforms: every sweep of the outer four-step in one launch
    godsp, capi, L = gd
    n = 1 << lg
    x = oracle.splitmix_complex(n * b, 2)
    want = np.concatenate([oracle.fft(x[i * n:(i + 1) * n]) for i in range(b)])
    try:
        got = godsp.fft.FFTBatch(x, n)
        assert max(rel_l2(got[i * n:(i + 1) * n], want[i * n:(i + 1) * n]) for i in range(b)) <= TOL
        assert rel_l2(godsp.fft.FFTBatch(want, n, -1), x) <= TOL
    finally:
        capi.check(L.gd_set_option(b"huge_min_log2n", 22))


@pytest.mark.parametrize("n", [1, 3, 5, 6, 7, 9, 12, 100, 255, 1000, 4097, 65537, 100003, 1000003])
def test_bluestein_sizes(gd, n):                 # config C2 is n = 1,000,003 (la = 2^21)
    godsp = gd[0]
    x = oracle.splitmix_complex(n, 2)
    assert rel_l2(godsp.fft.FFT(x), oracle.fft(x)) <= TOL          # incl. the reference's chirp-phase rounding
    assert rel_l2(godsp.fft.IFFT(x), oracle.ifft(x)) <= TOL
    r = oracle.fill_splitmix(n, 4)
    assert rel_l2(godsp.fft.FFTReal(r), oracle.fft_real(r)) <= TOL
    assert rel_l2(godsp.fft.IFFTReal(r), oracle.ifft_real(r)) <= TOL


def test_bluestein_above_2p24(gd):
    """Non-power-of-two lengths whose padded length exceeds 2^24 (n > 2^23; the reference has no length limit,
    fft/fft.go:72-87): chirp products and padding as streaming kernels around the outer four-step transforms."""
    godsp, capi, L = gd
    n = 8500003
    assert L.gd_bluestein_padded_len(n) == 1 << 25
    x = oracle.splitmix_complex(n, 2)
    got = godsp.fft.FFT(x)
    assert rel_l2(got, oracle.fft(x)) <= TOL
    assert rel_l2(godsp.fft.IFFT(x), oracle.ifft(x)) <= TOL
    # real input is the same transform of (r, 0): identical bits, no second oracle run needed
    r = np.ascontiguousarray(x.real)
    assert np.array_equal(godsp.fft.FFTReal(r), godsp.fft.FFT(r.astype(np.complex128)))


@pytest.mark.parametrize("n", [3, 5, 6, 7, 9, 17, 33, 100, 129, 257, 1000, 1025, 2047])
def test_bluestein_fused_small(gd, n):           # padded length <= 4096: one kernel per transform (bluestein_small.cuh)
    """Batched lines (ragged batch: not a multiple of the lines per CTA) against the oracle, forward and inverse, complex
    and real input; and the two-launch path it replaces gives the same bits (same operations in the same order)."""
    godsp, capi, L = gd
    assert L.gd_bluestein_padded_len(n) <= 4096
    b = 131 if n < 300 else 7
    x = oracle.splitmix_complex(b * n, 11).reshape(b, n)
    want = np.stack([oracle.fft(x[i]) for i in range(b)])
    wanti = np.stack([oracle.ifft(x[i]) for i in range(b)])
    got, goti = np.empty_like(x), np.empty_like(x)
    capi.check(L.gd_fft_batch_c2c(x.ctypes.data, got.ctypes.data, n, b, 1))
    capi.check(L.gd_fft_batch_c2c(x.ctypes.data, goti.ctypes.data, n, b, -1))
    assert rel_l2(got, want) <= TOL and rel_l2(goti, wanti) <= TOL
    r = oracle.fill_splitmix(n, 4)
    gr, gri = godsp.fft.FFTReal(r), godsp.fft.IFFTReal(r)
    assert rel_l2(gr, oracle.fft_real(r)) <= TOL and rel_l2(gri, oracle.ifft_real(r)) <= TOL
    capi.check(L.gd_set_option(b"bluestein_fused", 0))
    try:
        g2, g2i = np.empty_like(x), np.empty_like(x)
        capi.check(L.gd_fft_batch_c2c(x.ctypes.data, g2.ctypes.data, n, b, 1))
        capi.check(L.gd_fft_batch_c2c(x.ctypes.data, g2i.ctypes.data, n, b, -1))
        r2, r2i = godsp.fft.FFTReal(r), godsp.fft.IFFTReal(r)
    finally:
        capi.check(L.gd_set_option(b"bluestein_fused", 1))
    assert np.array_equal(got, g2) and np.array_equal(goti, g2i)
    assert np.array_equal(gr, r2) and np.array_equal(gri, r2i)


@pytest.mark.parametrize("n", [2049, 3000, 4095])
def test_bluestein_fused_8192(gd, n):            # padded length 8192: still one kernel per transform (512 threads, 16 x 16 x 16 x 2)
    godsp, capi, L = gd
    assert L.gd_bluestein_padded_len(n) == 8192
    b = 5
    x = oracle.splitmix_complex(b * n, 12).reshape(b, n)
    want = np.stack([oracle.fft(x[i]) for i in range(b)])
    wanti = np.stack([oracle.ifft(x[i]) for i in range(b)])
    got, goti = np.empty_like(x), np.empty_like(x)
    capi.check(L.gd_fft_batch_c2c(x.ctypes.data, got.ctypes.data, n, b, 1))      # builds the plan (chirp, FFT(b)) as well
    l0 = L.gd_kernel_launches()
    capi.check(L.gd_fft_batch_c2c(x.ctypes.data, goti.ctypes.data, n, b, -1))
    assert L.gd_kernel_launches() - l0 == 1
    assert rel_l2(got, want) <= TOL and rel_l2(goti, wanti) <= TOL
    r = oracle.fill_splitmix(n, 4)
    assert rel_l2(godsp.fft.FFTReal(r), oracle.fft_real(r)) <= TOL and rel_l2(godsp.fft.IFFTReal(r), oracle.ifft_real(r)) <= TOL
    capi.check(L.gd_set_option(b"bluestein_fused", 0))      # against the two-transform path (different arithmetic: not bit-equal)
    try:
        g2 = np.empty_like(x)
        capi.check(L.gd_fft_batch_c2c(x.ctypes.data, g2.ctypes.data, n, b, 1))
    finally:
        capi.check(L.gd_set_option(b"bluestein_fused", 1))
    assert rel_l2(got, g2) <= 1e-13


def test_real_roundtrips(gd):                    # C2: IFFT(FFTReal(x)) ~ x and FFT(IFFTReal(x)) ~ x
    godsp = gd[0]
    for n in (4096, 1000003, 1 << 16):
        r = oracle.fill_splitmix(n, 7)
        # Bluestein's chirp phase is only good to ~N*eps rad (fft/bluestein.go:53), so the round trip of the
        # reference itself is ~1e-10 at N = 1,000,003; the power-of-two sizes round-trip at rounding level
        tol = 1e-9 if n == 1000003 else TOL * 10
        assert rel_l2(godsp.fft.IFFT(godsp.fft.FFTReal(r)), r.astype(complex)) <= tol
        assert rel_l2(godsp.fft.FFT(godsp.fft.IFFTReal(r)), r.astype(complex)) <= tol
        # and each half against the oracle (power-of-two FFTReal / IFFTReal beyond the golden N <= 8)
        assert rel_l2(godsp.fft.FFTReal(r), oracle.fft_real(r)) <= TOL
        assert rel_l2(godsp.fft.IFFTReal(r), oracle.ifft_real(r)) <= TOL


def test_above_2p24_single_gpu(gd):              # the reference has no length limit (fft/fft.go:72-87): outer four-step above 2^24
    godsp, capi, L = gd
    n = 1 << 25
    x = oracle.splitmix_complex(n, 1)
    X = godsp.fft.FFT(x)
    assert rel_l2(X, oracle.fft(x)) <= TOL
    assert rel_l2(godsp.fft.IFFT(X), x) <= TOL
    # a Bluestein length whose padded size exceeds 2^30 is refused loudly (documented limit), never computed on the CPU
    assert L.gd_plan_warm((1 << 29) + 1) < 0
    assert "2^30" in L.gd_last_error().decode()


@pytest.mark.parametrize("lg", [13, 16, 20, 22, 25])
def test_real_input_large(gd, lg):               # FFTReal / IFFTReal = transform of dsputils.ToComplex(x): widening sweep + plain transform
    godsp, capi, L = gd
    import torch
    n = 1 << lg
    r = oracle.fill_splitmix(n, 9)
    assert rel_l2(godsp.fft.FFTReal(r), oracle.fft_real(r)) <= TOL         # 2^25: failed before (the outer four-step takes no load operator)
    if lg <= 22:
        assert rel_l2(godsp.fft.IFFTReal(r), oracle.ifft_real(r)) <= TOL
    if lg <= 20:                                                            # a batch of real lines on the device, against the complex path's bits
        b = max(2, (1 << 22) // n) + 1
        x = torch.empty(b * n, dtype=torch.float64, device="cuda")
        capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), b * n, 5, 0, None))
        xc = torch.complex(x, torch.zeros_like(x))
        y1, y2 = torch.empty(b * n, dtype=torch.complex128, device="cuda"), torch.empty(b * n, dtype=torch.complex128, device="cuda")
        capi.check(L.gd_fft_batch_r2c_full_dev(x.data_ptr(), y1.data_ptr(), n, b, 1, None))
        capi.check(L.gd_fft_batch_c2c_dev(xc.data_ptr(), y2.data_ptr(), n, b, 1, None))
        capi.check(L.gd_stream_sync(None))
        assert torch.equal(y1, y2)
        row = y1.view(b, n)[b - 1].cpu().numpy()
        assert rel_l2(row, oracle.fft_real(x.view(b, n)[b - 1].cpu().numpy())) <= TOL


def test_bluestein_padding_lengths(gd):
    L = gd[2]
    for n in (3, 5, 1000, 65537, 1000003):
        assert L.gd_bluestein_padded_len(n) == oracle.bluestein_padded_len(n)


@pytest.mark.parametrize("n", [8, 48, 1000, 4096, 1 << 15, 1 << 21, 1 << 22, 5000])      # 2^22: plain transforms + product sweep
def test_convolve(gd, n):
    godsp = gd[0]
    x, y = oracle.splitmix_complex(n, 1), oracle.splitmix_complex(n, 2)
    assert rel_l2(godsp.fft.Convolve(x, y), oracle.convolve(x, y)) <= TOL


@pytest.mark.parametrize("b,lg", [(37, 10), (5, 14), (7, 16), (3, 18), (5, 20), (300, 5), (1000, 1), (70, 16)])
def test_batched(gd, b, lg):
    _, capi, L = gd
    n = 1 << lg
    x = oracle.splitmix_complex(b * n, 3).reshape(b, n)
    want = oracle.fft_batch(x, threads=8)
    out = np.empty_like(x)
    capi.check(L.gd_fft_batch_c2c(x.ctypes.data, out.ctypes.data, n, b, 1))
    assert rel_l2(out, want) <= TOL
    capi.check(L.gd_fft_batch_c2c(want.ctypes.data, out.ctypes.data, n, b, -1))
    assert rel_l2(out, x) <= TOL


SCHEDULE_DEFAULTS = {"tma": 1, "tma_delay": 2, "tma_slots": 3, "tma_opt": 0, "fused": 0, "wide_tiles": 0, "pass_scratch_mb": 1024, "w32": 2}


@pytest.mark.parametrize("opts", [{"tma": 1}, {"tma": 1, "tma_delay": 1}, {"tma": 1, "tma_delay": 0}, {"tma": 1, "tma_delay": 3, "tma_slots": 4}, {"tma": 1, "tma_delay": 1, "tma_slots": 2}, {"tma": 1, "tma_opt": 2}, {"tma": 0},
                                  {"tma": 0, "fused": 1}, {"tma": 0, "wide_tiles": 1}, {"tma": 0, "pass_scratch_mb": 16},
                                  {"tma": 0, "w32": 0}, {"tma": 0, "w32": 5}])
def test_alternative_schedules_agree(gd, opts):   # every planner variant of the 2^20-point transform must give the same result
    _, capi, L = gd
    x = oracle.splitmix_complex(6 << 20, 3).reshape(6, 1 << 20)
    want = oracle.fft_batch(x, threads=8)
    out, back = np.empty_like(x), np.empty_like(x)
    for k, v in opts.items():
        capi.check(L.gd_set_option(k.encode(), v))
    try:
        capi.check(L.gd_fft_batch_c2c(x.ctypes.data, out.ctypes.data, 1 << 20, 6, 1))
        capi.check(L.gd_fft_batch_c2c(out.ctypes.data, back.ctypes.data, 1 << 20, 6, -1))     # conj / scale hooks of the same schedule
    finally:
        for k, v in SCHEDULE_DEFAULTS.items():
            capi.check(L.gd_set_option(k.encode(), v))
    assert rel_l2(out, want) <= TOL
    assert rel_l2(back, x) <= TOL


@pytest.mark.parametrize("n,b", [(1 << 20, 3), (1000, 7), (4096, 33), (1 << 13, 5)])
def test_fft_batch_api(gd, n, b):                 # additive FFTBatch of the shim
    godsp = gd[0]
    x = oracle.splitmix_complex(n * b, 11)
    got = godsp.fft.FFTBatch(x, n).reshape(b, n)
    want = np.stack([oracle.fft(np.ascontiguousarray(r)) for r in x.reshape(b, n)])
    assert rel_l2(got, want) <= TOL
    back = godsp.fft.FFTBatch(got.reshape(-1), n, -1)
    assert rel_l2(back, x) <= TOL
    with pytest.raises(Exception):
        godsp.fft.FFTBatch(x[:-1], n)


@pytest.mark.parametrize("delay", [2, 1])
def test_tma_fused_stress(gd, delay):             # race hunt: every row of every repetition must keep its energy (Parseval)
    _, capi, L = gd
    import torch
    nb, n, reps = 64, 1 << 20, 60
    x = torch.empty(nb * n * 2, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), nb * n * 2, 3, 0, None))
    capi.check(L.gd_stream_sync(None))
    ex = (x.view(nb, -1) ** 2).sum(1)
    capi.check(L.gd_set_option(b"tma_delay", delay))
    try:
        for _ in range(reps):
            y.zero_()
            torch.cuda.synchronize()
            capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, nb, 1, None))
            capi.check(L.gd_stream_sync(None))
            ey = (y.view(nb, -1) ** 2).sum(1)
            assert float(((ey / n - ex).abs() / ex).max()) < 1e-13
    finally:
        capi.check(L.gd_set_option(b"tma_delay", SCHEDULE_DEFAULTS["tma_delay"]))


def test_tma14_fused_stress(gd):                  # the same race hunt for the fused 2^14 kernel, both modes (Parseval per line)
    _, capi, L = gd
    import torch
    n, nb, reps = 1 << 14, 2048, 40
    x = torch.empty(nb * n * 2, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), nb * n * 2, 3, 0, None))
    capi.check(L.gd_stream_sync(None))
    xc = torch.view_as_complex(x.view(-1, 2))
    ex_rows = (xc.view(nb, n).abs() ** 2).sum(1)            # rows: nb transforms of 2^14 points
    ex_cols = (xc.view(n, nb).abs() ** 2).sum(0)            # columns of a 2^14 x nb matrix
    for _ in range(reps):
        y.zero_()
        torch.cuda.synchronize()
        capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, nb, 1, None))
        capi.check(L.gd_stream_sync(None))
        ey = (torch.view_as_complex(y.view(-1, 2)).view(nb, n).abs() ** 2).sum(1)
        assert float(((ey / n - ex_rows).abs() / ex_rows).max()) < 1e-13
        y.zero_()
        torch.cuda.synchronize()
        capi.check(L.gd_fft_strided_c2c_dev(x.data_ptr(), y.data_ptr(), 1, n, nb, 1, None))
        capi.check(L.gd_stream_sync(None))
        ey = (torch.view_as_complex(y.view(-1, 2)).view(n, nb).abs() ** 2).sum(0)
        assert float(((ey / n - ex_cols).abs() / ex_cols).max()) < 1e-13


def test_tma_fused_chunking(gd):                  # more than one 128-transform launch, a partial last chunk, in place
    _, capi, L = gd
    import torch
    nb, n = 131, 1 << 20
    x = torch.empty(nb * n * 2, dtype=torch.float64, device="cuda")
    capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), nb * n * 2, 3, 0, None))
    torch.cuda.synchronize()
    rows = [0, 1, 127, 128, 130]
    src = {r: x.view(nb, n * 2)[r].cpu().numpy().view(np.complex128).copy() for r in rows}
    capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), x.data_ptr(), n, nb, 1, None))             # in == out
    capi.check(L.gd_stream_sync(None))
    for r in rows:
        got = x.view(nb, n * 2)[r].cpu().numpy().view(np.complex128)
        assert rel_l2(got, oracle.fft(src[r])) <= TOL, r


# ----------------------------------------------------------------- FFT2 / FFTN
@pytest.mark.parametrize("shape", [(2, 3), (3, 5), (64, 32), (300, 7), (8192, 16), (16, 8192), (512, 512),
                                   (4096, 128), (2048, 96), (3000, 70), (2, 4096, 64), (5000, 33),     # blocked four-step on 4096-point strided lines; tiled gather / scatter around Bluestein lines
                                   (2, 2, 3), (4, 6, 8, 5), (16, 16, 16), (3, 1, 4)])
def test_fftn_shapes(gd, shape):
    godsp = gd[0]
    x = oracle.splitmix_complex(int(np.prod(shape)), 4).reshape(shape)
    m = godsp.dsputils.MakeMatrix(x.ravel(), list(shape))
    assert rel_l2(godsp.fft.FFTN(m).list, oracle.fftn(x).ravel()) <= TOL
    assert rel_l2(godsp.fft.IFFTN(m).list, oracle.fftn(x, inverse=True).ravel()) <= TOL
    if len(shape) == 2:
        got = np.stack(godsp.fft.FFT2(list(x)))
        assert rel_l2(got, oracle.fft2(x)) <= TOL


def test_fft2_large_rows_cols(gd):               # 2^14-long lines in both axes (the C3 matrix is 16384^2)
    _, capi, L = gd
    for rows, cols in ((16384, 32), (32, 16384)):
        x = oracle.splitmix_complex(rows * cols, 4).reshape(rows, cols)
        out = np.empty_like(x)
        capi.check(L.gd_fft2_c2c(x.ctypes.data, out.ctypes.data, rows, cols, 1))
        assert rel_l2(out, oracle.fft2(x)) <= TOL


@pytest.mark.parametrize("rows,cols", [(16384, 128), (256, 16384), (16384, 320), (192, 16384)])
def test_fft2_fused_2p14_lines(gd, rows, cols):  # the fused 2^14 kernel (fft_tma14.cuh): columns of a 2^14-row matrix, batched 2^14 rows
    _, capi, L = gd
    x = oracle.splitmix_complex(rows * cols, 4).reshape(rows, cols)
    out, back = np.empty_like(x), np.empty_like(x)
    capi.check(L.gd_fft2_c2c(x.ctypes.data, out.ctypes.data, rows, cols, 1))
    assert rel_l2(out, oracle.fft2(x)) <= TOL
    capi.check(L.gd_fft2_c2c(out.ctypes.data, back.ctypes.data, rows, cols, -1))
    assert rel_l2(back, x) <= TOL
    capi.check(L.gd_set_option(b"tma14", 0))     # and the two-launch schedule gives the same result
    try:
        o2 = np.empty_like(x)
        capi.check(L.gd_fft2_c2c(x.ctypes.data, o2.ctypes.data, rows, cols, 1))
    finally:
        capi.check(L.gd_set_option(b"tma14", 1))
    assert rel_l2(o2, out) <= 1e-14


@pytest.mark.parametrize("phases", [1, 2])
@pytest.mark.parametrize("lg", [13, 14, 15, 16, 17, 18, 19])
def test_fused_family_rows_and_columns(gd, lg, phases):
    """The TMA-fed fused four-step (fft_tma14.cuh) for every size of its family, N = LA x LB = 2^13 .. 2^19: batched contiguous
    transforms and the columns of an N-row matrix, forward and inverse, with a batch / column count that is not a multiple
    of a phase (the remainder goes through the two-launch path); sampled lines against the oracle, everything against the
    two-launch schedule."""
    _, capi, L = gd
    import torch
    n = 1 << lg
    unit = (1 << 20) // n
    nb = phases * unit + 3                                    # whole phases + a remainder
    x = torch.empty(nb * n * 2, dtype=torch.float64, device="cuda")
    capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), nb * n * 2, 30 + lg, 0, None))
    capi.check(L.gd_stream_sync(None))
    xh = x.cpu().numpy().view(np.complex128)
    outs = {}
    for fused in (1, 0):
        capi.check(L.gd_set_option(b"tma14", fused)); capi.check(L.gd_set_option(b"tma16", fused)); capi.check(L.gd_set_option(b"tma19", fused))
        try:
            y, z, yc, zc = (torch.empty_like(x) for _ in range(4))
            l0 = L.gd_kernel_launches()
            capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, nb, 1, None))
            capi.check(L.gd_fft_batch_c2c_dev(y.data_ptr(), z.data_ptr(), n, nb, -1, None))
            nl_rows = L.gd_kernel_launches() - l0
            ok_cols = lg <= 17
            if ok_cols:                                       # the same memory as an n x nb matrix: every column is a transform
                capi.check(L.gd_fft_strided_c2c_dev(x.data_ptr(), yc.data_ptr(), 1, n, nb, 1, None))
                capi.check(L.gd_fft_strided_c2c_dev(yc.data_ptr(), zc.data_ptr(), 1, n, nb, -1, None))
            if fused:                                         # in place: pass 2 of a phase starts only after all its pass-1 tiles were read
                w = x.clone()
                capi.check(L.gd_fft_batch_c2c_dev(w.data_ptr(), w.data_ptr(), n, nb, 1, None))
                capi.check(L.gd_stream_sync(None))
                assert torch.equal(w, y), "in-place rows differ"
                if ok_cols:
                    w = x.clone()
                    capi.check(L.gd_fft_strided_c2c_dev(w.data_ptr(), w.data_ptr(), 1, n, nb, 1, None))
                    capi.check(L.gd_stream_sync(None))
                    assert torch.equal(w, yc), "in-place columns differ"
            capi.check(L.gd_stream_sync(None))
            outs[fused] = (y.cpu().numpy().view(np.complex128), z.cpu().numpy().view(np.complex128),
                           yc.cpu().numpy().view(np.complex128) if ok_cols else None, zc.cpu().numpy().view(np.complex128) if ok_cols else None, nl_rows)
        finally:
            capi.check(L.gd_set_option(b"tma14", 1)); capi.check(L.gd_set_option(b"tma16", 1)); capi.check(L.gd_set_option(b"tma19", 1))
    y, z, yc, zc, nl_fused = outs[1]
    y0, z0, yc0, zc0, nl_plain = outs[0]
    if phases == 2:
        assert nl_fused < nl_plain                            # one persistent launch per direction (+ the remainder) instead of chunks
    for r in sorted({0, 1, unit - 1, unit, phases * unit - 1, phases * unit, nb - 1}):
        assert rel_l2(y.reshape(nb, n)[r], oracle.fft(np.ascontiguousarray(xh.reshape(nb, n)[r]))) <= TOL, ("row", r)
    assert rel_l2(z, xh) <= TOL
    assert rel_l2(y, y0) <= 1e-14 and rel_l2(z, z0) <= 1e-14
    if yc is not None:
        for c in sorted({0, unit - 1, unit, phases * unit - 1, phases * unit, nb - 1}):
            assert rel_l2(yc.reshape(n, nb)[:, c], oracle.fft(np.ascontiguousarray(xh.reshape(n, nb)[:, c]))) <= TOL, ("column", c)
        assert rel_l2(zc, xh) <= TOL
        assert rel_l2(yc, yc0) <= 1e-14 and rel_l2(zc, zc0) <= 1e-14


@pytest.mark.parametrize("rows,cols", [(65536, 32), (32, 65536), (65536, 80), (48, 65536)])
def test_fft2_fused_2p16_lines(gd, rows, cols):  # 256 x 256 sub-lines: columns of a 2^16-row matrix, batched 2^16 rows (host-pointer API)
    _, capi, L = gd
    x = oracle.splitmix_complex(rows * cols, 4).reshape(rows, cols)
    out, back = np.empty_like(x), np.empty_like(x)
    capi.check(L.gd_fft2_c2c(x.ctypes.data, out.ctypes.data, rows, cols, 1))
    assert rel_l2(out, oracle.fft2(x)) <= TOL
    capi.check(L.gd_fft2_c2c(out.ctypes.data, back.ctypes.data, rows, cols, -1))
    assert rel_l2(back, x) <= TOL


def test_tma16_fused_stress(gd):                  # race hunt for the 256 x 256 variant, both modes (Parseval per line)
    _, capi, L = gd
    import torch
    n, nb, reps = 1 << 16, 512, 25
    x = torch.empty(nb * n * 2, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), nb * n * 2, 3, 0, None))
    capi.check(L.gd_stream_sync(None))
    xc = torch.view_as_complex(x.view(-1, 2))
    ex_rows = (xc.view(nb, n).abs() ** 2).sum(1)
    ex_cols = (xc.view(n, nb).abs() ** 2).sum(0)
    for _ in range(reps):
        y.zero_()
        torch.cuda.synchronize()
        capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, nb, 1, None))
        capi.check(L.gd_stream_sync(None))
        ey = (torch.view_as_complex(y.view(-1, 2)).view(nb, n).abs() ** 2).sum(1)
        assert float(((ey / n - ex_rows).abs() / ex_rows).max()) < 1e-13
        y.zero_()
        torch.cuda.synchronize()
        capi.check(L.gd_fft_strided_c2c_dev(x.data_ptr(), y.data_ptr(), 1, n, nb, 1, None))
        capi.check(L.gd_stream_sync(None))
        ey = (torch.view_as_complex(y.view(-1, 2)).view(n, nb).abs() ** 2).sum(0)
        assert float(((ey / n - ex_cols).abs() / ex_cols).max()) < 1e-13


def test_fft2_16384_square_sampled(gd):          # config C3: the full 16384 x 16384 matrix, device resident
    """Output rows / columns of the full-size result against oracle.fft of the matching single-bin DFT of the other axis
    (a float64 matrix-vector product with exactly reduced phasors, computed by torch -- not by this library):
    out[r, :] = FFT(f_r^T . src), out[:, c] = FFT(src . f_c)   (fft/fft.go:138-151: columns, then rows)."""
    import torch
    _, capi, L = gd
    R = Cc = 16384
    st = C.c_void_p(0)
    src = torch.empty(R * Cc, dtype=torch.complex128, device="cuda")
    out = torch.empty_like(src)
    torch.cuda.synchronize()
    capi.check(L.gd_fill_splitmix_dev(src.data_ptr(), 2 * R * Cc, 4, 0, st))
    dims = (C.c_int64 * 2)(R, Cc)
    capi.check(L.gd_fftn_c2c_dev(src.data_ptr(), out.data_ptr(), dims, 2, 1, st))
    capi.check(L.gd_stream_sync(st))
    s2, o2 = src.view(R, Cc), out.view(R, Cc)

    def phasors(length, k):
        i = torch.arange(length, dtype=torch.int64, device="cuda")
        ang = (-2.0 * math.pi / length) * ((i * k) % length).double()
        return torch.complex(torch.cos(ang), torch.sin(ang))

    for r in (0, 1, 4097, R - 1):
        u = torch.mv(s2.t(), phasors(R, r))
        assert rel_l2(o2[r].cpu().numpy(), oracle.fft(u.cpu().numpy())) <= TOL, r
    for c in (0, 2, 8191, Cc - 1):
        v = torch.mv(s2, phasors(Cc, c))
        assert rel_l2(o2[:, c].contiguous().cpu().numpy(), oracle.fft(v.cpu().numpy())) <= TOL, c
    # inverse round trip, in place
    capi.check(L.gd_fftn_c2c_dev(out.data_ptr(), out.data_ptr(), dims, 2, -1, st))
    capi.check(L.gd_stream_sync(st))
    assert float((out - src).abs().max()) < 1e-12


# ----------------------------------------------------------------- Pwelch
PW_CASES = [  # (nx, NFFT, Noverlap, Pad, window, scale_off)
    (100, 0, 0, 0, None, False), (5000, 256, 128, 0, None, False), (100000, 4096, 2048, 0, "Hann", False),
    (50000, 1024, 512, 2048, "Hamming", False), (5000, 100, 30, 0, "Blackman", True), (5000, 256, 0, 128, "FlatTop", False),
    (9000, 4096, 2048, 0, "Bartlett", False), (70000, 64, 63, 0, "Rectangular", False), (33000, 2048, 1024, 0, None, True),
    (40000, 512, 100, 0, None, False), (4096, 4096, 0, 0, None, False), (20000, 32, 16, 0, None, False),
    (20000, 6000, 3000, 0, None, False), (3000, 300, 0, 1000, None, False),
]


@pytest.mark.parametrize("case", PW_CASES)
def test_pwelch_options(gd, case):
    godsp = gd[0]
    nx, nfft, nov, pad, wname, soff = case
    x = oracle.fill_splitmix(nx, 5)
    w = getattr(godsp.window, wname) if wname else None
    p, f = godsp.spectral.Pwelch(x, 2.0, godsp.spectral.PwelchOptions(NFFT=nfft, Window=w, Pad=pad, Noverlap=nov, Scale_off=soff))
    pw, fw = oracle.pwelch(x, 2.0, nfft=nfft, pad=pad, noverlap=nov, window_fn=(wname or "hann").lower(), scale_off=soff)
    assert len(p) == len(pw)                                    # bin count, bit-exact
    assert np.array_equal(f.view(np.uint64), fw.view(np.uint64))   # frequency vector, bit-exact
    assert rel_l2(p, pw) <= TOL


def test_pwelch_custom_window_closure(gd):       # PwelchOptions.Window is an arbitrary function (pwelch.go:41)
    godsp = gd[0]
    wf = lambda n: np.cos(np.arange(n) / max(n, 1)) ** 2 + 0.1
    x = oracle.fill_splitmix(30000, 5)
    p, f = godsp.spectral.Pwelch(x, 3.0, godsp.spectral.PwelchOptions(NFFT=1024, Window=wf, Noverlap=256))
    pw, fw = oracle.pwelch(x, 3.0, nfft=1024, noverlap=256, window_fn=wf)
    assert rel_l2(p, pw) <= TOL and np.array_equal(f, fw)


@pytest.mark.parametrize("case", [(1 << 21, 8192, 4096, 0, "Hann"), ((1 << 21) + 5000, 8192, 1000, 0, "Hamming"),
                                  (1 << 22, 16384, 8192, 32768, "Hann"), (3 * (1 << 20), 65536, 0, 0, "Rectangular"),
                                  (1 << 22, 262144, 131072, 0, "Hann"), (1 << 22, 1 << 20, 1 << 19, 0, "Hann"),
                                  (1 << 20, 6000, 3000, 8192, "Hann")])
def test_pwelch_large_nfft(gd, case):            # power-of-two transform lengths above 4096: pairs of segments per complex transform
    """NFFT / Pad beyond the fused kernel's 4096 points: segment pairs packed into complex transforms, batched through the fused
    size family (2^13 .. 2^18) or the chunked passes, per-bin sums in a fixed order; even and odd segment counts."""
    godsp = gd[0]
    nx, nfft, nov, pad, wname = case
    x = oracle.fill_splitmix(nx, 5)
    p, f = godsp.spectral.Pwelch(x, 2.0, godsp.spectral.PwelchOptions(NFFT=nfft, Window=getattr(godsp.window, wname), Pad=pad, Noverlap=nov))
    pw, fw = oracle.pwelch(x, 2.0, nfft=nfft, pad=pad, noverlap=nov, window_fn=wname.lower(), threads=8)
    assert len(p) == len(pw) and np.array_equal(f.view(np.uint64), fw.view(np.uint64))
    assert rel_l2(p, pw) <= TOL


def test_pwelch_config4_sample(gd):              # C4 shape (NFFT 4096, 50% overlap, Hann) on 2^24 samples vs the oracle
    godsp = gd[0]
    x = oracle.fill_splitmix(1 << 24, 5)
    p, f = godsp.spectral.Pwelch(x, 1.0, godsp.spectral.PwelchOptions(NFFT=4096, Noverlap=2048, Window=godsp.window.Hann))
    pw, fw = oracle.pwelch(x, 1.0, nfft=4096, noverlap=2048, threads=8)
    assert len(p) == 2049 and np.array_equal(f, fw) and rel_l2(p, pw) <= TOL


# ----------------------------------------------------------------- properties at benchmark sizes (device-resident)
def test_full_size_properties(gd):
    """2^20-point batch of 256 on the device: impulse known answers, linearity, Parseval, and
    rows checked against the oracle; Pwelch on 2^28 samples: segment count, Parseval-type checksum."""
    import torch
    _, capi, L = gd
    n, b = 1 << 20, 256
    x = torch.empty(b * n * 2, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    st = C.c_void_p(0)
    torch.cuda.synchronize()
    capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), b * n * 2, 3, 0, st))
    capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, b, 1, st))
    capi.check(L.gd_stream_sync(st))
    xc, yc = torch.view_as_complex(x.view(-1, 2)).view(b, n), torch.view_as_complex(y.view(-1, 2)).view(b, n)
    # Parseval per row: sum|X|^2 = N sum|x|^2
    ex, ey = (xc.abs() ** 2).sum(1), (yc.abs() ** 2).sum(1)
    assert float(((ey / n - ex).abs() / ex).max()) < 1e-13
    # sampled rows against the oracle (row r uses counters from 2*r*n)
    for r in (0, 100, 255):
        want = oracle.fft(oracle.splitmix_complex(n, 3, r * n))
        assert rel_l2(yc[r].cpu().numpy(), want) <= TOL
    # inverse round trip on the device
    z = torch.empty_like(x)
    capi.check(L.gd_fft_batch_c2c_dev(y.data_ptr(), z.data_ptr(), n, b, -1, st))
    capi.check(L.gd_stream_sync(st))
    assert float((z - x).norm() / x.norm()) < 1e-14
    # impulse at p -> exp(-2 pi i p k / N)
    x.zero_()
    xv = x.view(b, n, 2)
    ps = [0, 1, 12345, n - 1]
    for i, p in enumerate(ps):
        xv[i, p, 0] = 1.0
    torch.cuda.synchronize()                     # torch's stream and the library's non-blocking stream are not ordered
    capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, b, 1, st))
    capi.check(L.gd_stream_sync(st))
    k = torch.arange(n, device="cuda", dtype=torch.float64)
    for i, p in enumerate(ps):
        ang = -2 * math.pi * ((k * p) % n) / n
        want = torch.complex(torch.cos(ang), torch.sin(ang))
        assert float((yc[i] - want).abs().max()) < 1e-12
    assert float(yc[len(ps):].abs().max()) == 0.0
    del x, y, z
    # Pwelch, 2^28 samples resident: total power check  sum(Pxx)*Fs/NFFT ~ var(x) and sharded == unsharded
    ns, nfft, nov = 1 << 28, 4096, 2048
    s = torch.empty(ns, dtype=torch.float64, device="cuda")
    capi.check(L.gd_fill_splitmix_dev(s.data_ptr(), ns, 5, 0, st))
    win = oracle.window("hann", nfft)
    norm = float(np.sum(win ** 2))
    dwin = torch.from_numpy(win).cuda()
    lp, nsegs = nfft // 2 + 1, oracle.segment_count(ns, nfft, nov)
    assert nsegs == 131071
    raw, raw2, pxx = (torch.empty(lp, dtype=torch.float64, device="cuda") for _ in range(3))
    capi.check(L.gd_pwelch_partial_dev(s.data_ptr(), nfft, nov, nfft, lp, 0, nsegs, dwin.data_ptr(), raw.data_ptr(), st))
    capi.check(L.gd_pwelch_finalize_dev(raw.data_ptr(), lp, nsegs, norm, pxx.data_ptr(), st))
    capi.check(L.gd_stream_sync(st))
    total_power = float(pxx.sum()) / nfft
    assert abs(total_power - 1.0 / 3.0) < 2e-4                   # uniform [-1,1): variance 1/3
    # two "ranks": segment ranges [0, h) and [h, nsegs) summed in order equal the single range
    h = nsegs // 2
    capi.check(L.gd_pwelch_partial_dev(s.data_ptr(), nfft, nov, nfft, lp, 0, h, dwin.data_ptr(), raw2.data_ptr(), st))
    capi.check(L.gd_pwelch_partial_dev(s.data_ptr(), nfft, nov, nfft, lp, h, nsegs - h, dwin.data_ptr(), pxx.data_ptr(), st))
    capi.check(L.gd_stream_sync(st))
    assert float(((raw2 + pxx) - raw).norm() / raw.norm()) < 1e-13


def test_edge_cases(gd):
    godsp = gd[0]
    assert len(godsp.fft.FFT(np.zeros(0))) == 0                   # fft/fft.go:76-80
    assert godsp.fft.FFT(np.array([3 + 4j]))[0] == 3 + 4j
    assert godsp.fft.IFFT(np.array([3 + 4j]))[0] == 3 + 4j
    with pytest.raises(godsp.GoPanic):
        godsp.fft.IFFT(np.zeros(0))                               # index out of range in the reference (fft.go:40)
    x = oracle.splitmix_complex(64, 9)
    keep = x.copy()
    godsp.fft.FFT(x)
    assert np.array_equal(x, keep)                                # inputs are never modified
    assert gd[2].gd_kernel_launches() > 0


def test_two_streams_share_the_device_scratch(gd):
    """Calls on different streams use the same scratch slots and dependency counters: the library orders them on the
    device, so back-to-back batches on two streams must both come out right."""
    _, capi, L = gd
    import torch
    n, b = 1 << 20, 24
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    x1 = torch.empty(b * n * 2, dtype=torch.float64, device="cuda")
    x2 = torch.empty_like(x1)
    capi.check(L.gd_fill_splitmix_dev(x1.data_ptr(), b * n * 2, 21, 0, None))
    capi.check(L.gd_fill_splitmix_dev(x2.data_ptr(), b * n * 2, 22, 0, None))
    capi.check(L.gd_stream_sync(None))
    y1, y2 = torch.empty_like(x1), torch.empty_like(x2)
    for _ in range(5):
        capi.check(L.gd_fft_batch_c2c_dev(x1.data_ptr(), y1.data_ptr(), n, b, 1, C.c_void_p(s1.cuda_stream)))
        capi.check(L.gd_fft_batch_c2c_dev(x2.data_ptr(), y2.data_ptr(), n, b, 1, C.c_void_p(s2.cuda_stream)))
    torch.cuda.synchronize()
    for x, y in ((x1, y1), (x2, y2)):
        ex, ey = (x.view(b, -1) ** 2).sum(1), (y.view(b, -1) ** 2).sum(1)
        assert float(((ey / n - ex).abs() / ex).max()) < 1e-13
        row = x.view(b, -1)[b - 1].cpu().numpy().view(np.complex128)
        assert rel_l2(y.view(b, -1)[b - 1].cpu().numpy().view(np.complex128), oracle.fft(row)) <= TOL


def test_concurrent_callers_one_device(gd):      # goroutines calling into one GPU: lanes (own streams / scratch / staging) instead of one mutex
    import threading
    _, capi, L = gd
    jobs = [(1 << 20, 4, 3), (1 << 14, 128, 5), (4096, 300, 7), (1000003, 1, 9), (1 << 16, 16, 11), (1 << 20, 3, 13)]
    res, errs = {}, []

    def work(i, n, b, seed):
        try:
            capi.check(L.gd_use_device(0))
            x = oracle.splitmix_complex(n * b, seed).reshape(b, n)
            out = np.empty_like(x)
            for _ in range(3):
                capi.check(L.gd_fft_batch_c2c(x.ctypes.data, out.ctypes.data, n, b, 1))
            res[i] = (x, out)
        except Exception as e:                     # noqa: BLE001
            errs.append((i, repr(e)))

    th = [threading.Thread(target=work, args=(i,) + j) for i, j in enumerate(jobs)]
    for t_ in th:
        t_.start()
    for t_ in th:
        t_.join(timeout=300)
    assert not errs, errs
    assert len(res) == len(jobs)
    for i, (n, b, seed) in enumerate(jobs):
        x, out = res[i]
        assert rel_l2(out, oracle.fft_batch(x, threads=8)) <= TOL, (n, b)
    # Pwelch and FFT2 from two threads at once
    sig = oracle.fill_splitmix(1 << 22, 5)
    godsp = gd[0]
    outp = {}

    def pw(k):
        outp[k] = godsp.spectral.Pwelch(sig, 1.0, godsp.spectral.PwelchOptions(NFFT=4096, Noverlap=2048))[0]
    th = [threading.Thread(target=pw, args=(k,)) for k in range(3)]
    for t_ in th:
        t_.start()
    for t_ in th:
        t_.join(timeout=300)
    want, _ = oracle.pwelch(sig, 1.0, nfft=4096, noverlap=2048, threads=8)
    for k in range(3):
        assert rel_l2(outp[k], want) <= TOL

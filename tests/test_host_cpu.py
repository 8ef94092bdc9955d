"""CPU-only checks: the C-ABI library loads and exports every symbol include/godsp_b200.h declares,
the host-side mirror reproduces the reference's host logic (windows, segments, Matrix layout,
panics) against the golden vectors, and the product path fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest

import oracle
from conftest import ROOT, pretty_close

godsp = pytest.importorskip("godsp")
from godsp import _capi, dsputils, fft, spectral, window  # noqa: E402


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "godsp_b200.h")).read()
    declared = set(re.findall(r"GD_API\s+[\w\s\*]+?\b(gd_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    L = _capi.lib()
    for name in declared:
        assert getattr(L, name) is not None


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(godsp.GoPanic, match="no CUDA device|no CPU fallback"):
        fft.FFT(np.ones(8))
    L = _capi.lib()
    out = np.empty(8, np.complex128)
    assert L.gd_fft_c2c(np.ones(8, np.complex128).ctypes.data, out.ctypes.data, 8, 1) < 0
    assert b"no CPU fallback" in L.gd_last_error()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "go-dsp_b200")):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".go")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "godsp_oracle" not in src and "gdo_" not in src, f


def test_bluestein_padded_len_bit_exact():      # dsputils.NextPowerOf2(2N-1), dsputils/dsputils.go:39-45
    L = _capi.lib()
    for n in [2, 3, 5, 6, 7, 9, 100, 1000, 4099, 65537, 1000003, (1 << 20) + 1, (1 << 23) - 1]:
        assert L.gd_bluestein_padded_len(n) == oracle.bluestein_padded_len(n) == dsputils.NextPowerOf2(2 * n - 1)


def test_next_power_of2_and_reverse_bits(golden):
    for x in list(range(0, 3000)) + [2**k + d for k in range(3, 40) for d in (-1, 0, 1)]:
        assert dsputils.NextPowerOf2(x) == oracle.next_pow2(x)
        assert dsputils.IsPowerOf2(x) == oracle.is_pow2(x)
    for c in golden["reverse_bits"]["cases"]:       # fft/fft_test.go:241-249
        assert fft.reverseBits(c["in"], c["sz"]) == c["out"]
    for s in range(1, 20):
        for v in np.random.default_rng(s).integers(0, 1 << s, 50):
            assert fft.reverseBits(int(v), s) == oracle.reverse_bits(int(v), s)


def test_window_vectors(golden):                # window/window_test.go:61-94
    for c in golden["window"]["cases"]:
        L = c["L"]
        assert pretty_close(window.Hamming(L), c["hamming"]) and pretty_close(window.Hann(L), c["hann"])
        assert pretty_close(window.Bartlett(L), c["bartlett"]) and pretty_close(window.FlatTop(L), c["flattop"])
        assert pretty_close(window.Blackman(L), c["blackman"])
        o = window.Rectangular(L)
        window.Apply(o, window.Hamming)
        assert pretty_close(o, c["hamming"])
    for name, f in [("hamming", window.Hamming), ("hann", window.Hann), ("bartlett", window.Bartlett),
                    ("flattop", window.FlatTop), ("blackman", window.Blackman), ("rectangular", window.Rectangular)]:
        for L in (1, 2, 3, 64, 255, 4096):
            assert np.array_equal(f(L), oracle.window(name, L)), (name, L)      # same expressions -> same bits


def test_spectral_segment(golden):              # spectral/spectral_test.go:58-67
    g = golden["spectral_segment"]
    x = np.array(g["x"], float)
    for c in g["cases"]:
        got = spectral.Segment(x, c["size"], c["noverlap"])
        assert len(got) == len(c["out"]) and all(np.array_equal(a, np.array(b, float)) for a, b in zip(got, c["out"]))
    for lx, size, nov in [(10, 10, 3), (9, 10, 0), (1 << 30, 4096, 2048), (100, 7, 6)]:
        assert len(spectral.Segment(np.zeros(min(lx, 200)), size, nov)) == oracle.segment_count(min(lx, 200), size, nov)
    from godsp import sharding
    assert sharding.segment_count(1 << 30, 4096, 2048) == 524287 == oracle.segment_count(1 << 30, 4096, 2048)


def test_dsputils_segment_aliases(golden):      # dsputils/dsputils_test.go:40-57
    g = golden["dsputils_segment"]
    x = np.arange(g["n"]).astype(np.complex128)
    for c in g["cases"]:
        v = dsputils.Segment(x, c["segs"], c["noverlap"])
        for seg, (a, b) in zip(v, c["slices"]):
            assert np.array_equal(seg, x[a:b]) and np.shares_memory(seg, x)
    with pytest.raises(godsp.GoPanic, match="too many segments"):
        dsputils.Segment(x[:2], 5, 0.0)


def test_matrix_layout(golden):                 # dsputils/matrix_test.go:23-46
    g = golden["matrix"]
    m = dsputils.MakeMatrix(np.array(g["list"], float), g["dims"])
    for c in g["dim_cases"]:
        assert pretty_close(m.Dim(c["idx"]), np.array(c["out"], float))
    m.SetDim(np.array(g["setdim"]["values"], complex), g["setdim"]["idx"])
    assert pretty_close(m.Dim(g["setdim"]["idx"]), g["setdim"]["values"])
    m.SetValue(14, [1, 2, 3])
    assert m.Value([1, 2, 3]) == 14
    for bad, msg in [([1, -1, -1], "only one dimension"), ([0, 0, 0], "must specify one"), ([2, -1, 0], "out of bounds")]:
        with pytest.raises(godsp.GoPanic, match=msg):
            m.Dim(bad)
    with pytest.raises(godsp.GoPanic, match="incorrect dimensions"):
        dsputils.MakeMatrix(np.zeros(5), [2, 3])


def test_reference_panics():
    with pytest.raises(godsp.GoPanic, match="arrays not of equal size"):      # fft/fft.go:57
        fft.Convolve(np.ones(3), np.ones(2))
    with pytest.raises(godsp.GoPanic, match="empty input array"):             # fft/fft.go:126
        fft.FFT2([])
    with pytest.raises(godsp.GoPanic, match="ragged input array"):            # fft/fft.go:133
        fft.FFT2([[1, 2], [3]])
    assert len(fft.FFT(np.zeros(0))) == 0                                     # fft/fft.go:76-80
    p, f = spectral.Pwelch(np.zeros(0), 0, spectral.PwelchOptions())          # spectral/pwelch.go:75-77
    assert len(p) == 0 and len(f) == 0
    fft.SetWorkerPoolSize(-3)
    from godsp import _host
    assert _host.lib().gdh_worker_pool_size() == 0                            # fft/fft.go:96-98


def test_sharding_partitions():
    from godsp import sharding
    for world in (1, 2, 3, 8):
        rows = [sharding.batch_rows(r, world, 4096) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == 4096 and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        segs = [sharding.pwelch_segment_range(r, world, 1 << 20, 4096, 2048) for r in range(world)]
        assert segs[0][0] == 0 and segs[-1][1] == sharding.segment_count(1 << 20, 4096, 2048)
        assert all(a[1] == b[0] for a, b in zip(segs, segs[1:])) and segs[-1][3] <= 1 << 20


def test_generated_codelets_are_current_and_correct():
    """tools/gen_codelets.py: the committed fft_codelets.cuh is what the generator emits, every codelet's DAG evaluates to
    the DFT (checked in Python against exactly reduced roots), and the operation counts DESIGN.md quotes hold."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_codelets", os.path.join(ROOT, "tools", "gen_codelets.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    want = {4: 16, 8: 52, 16: 144, 32: 376}
    plans = {4: [4], 8: [2, 4], 16: [4, 4], 32: [4, 2, 4]}
    for n, ops in want.items():
        assert g.check(n, plans[n]) < 1e-13 * n
        _, total, nops = g.codelet(n, plans[n], "x", n < 32)
        assert total == ops and nops["mul"] == 0
    text = open(os.path.join(ROOT, "go-dsp_b200", "csrc", "fft_codelets.cuh")).read()
    for n, ops in want.items():
        assert "forward %d-point DFT, natural order in and out: %d FP64 instructions" % (n, ops) in text

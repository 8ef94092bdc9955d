#!/usr/bin/env python3
"""Times the BASELINE.json configs that bench.py's headline line does not carry (device-resident, CUDA
events on the launching stream): C1 fft.FFT 2^16, C2 Bluestein N=1,000,003 (+ FFTReal), C3b fft.FFT2
16384x16384, plus cuFFT (torch.fft, complex128) on the C3 batch as a comparison only. Writes one JSON."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "go-dsp_b200"))
from godsp import _capi as capi  # noqa: E402

L = capi.lib()
capi.check(L.gd_use_device(0))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sp = C.c_void_p(stream.cuda_stream)


def timed(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()                     # > L2: evicts the previous iteration's lines
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


out = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

# C1: fft.FFT 2^16
n = 1 << 16
x = torch.empty(2 * n, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), 2 * n, 1, 0, sp))
med, best = timed(lambda: capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, 1, 1, sp)), flush=flush)
out["C1_fft_2^16"] = {"median_us": med * 1e3, "best_us": best * 1e3, "note": "one transform, L2 flushed between iterations"}

# C2: Bluestein N = 1,000,003 complex and real input
n = 1000003
x = torch.empty(2 * n, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), 2 * n, 2, 0, sp))
capi.check(L.gd_plan_warm(n))
med, best = timed(lambda: capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, 1, 1, sp)), flush=flush)
out["C2_bluestein_1000003"] = {"median_ms": med, "best_ms": best, "la": int(L.gd_bluestein_padded_len(n)),
                               "note": "chirp + FFT(b) cached per N; prep, forward 2^21 transform, product, inverse transform, post (plain transforms between streaming kernels)"}
med, best = timed(lambda: capi.check(L.gd_fft_batch_r2c_full_dev(x.data_ptr(), y.data_ptr(), n, 1, 1, sp)), flush=flush)
out["C2_fftreal_1000003"] = {"median_ms": med, "best_ms": best}

# C3b: fft.FFT2 on 16384 x 16384 (4 GiB in, 4 GiB out)
r = c = 16384
x = torch.empty(2 * r * c, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), 2 * r * c, 4, 0, sp))
dims = (C.c_int64 * 2)(r, c)
med, best = timed(lambda: capi.check(L.gd_fftn_c2c_dev(x.data_ptr(), y.data_ptr(), dims, 2, 1, sp)), iters=5, warm=2)
out["C3_fft2_16384x16384"] = {"median_ms": med, "best_ms": best, "Gelem_per_s": r * c / (med * 1e-3) / 1e9,
                              "algorithmic_GBps": 64.0 * r * c / (med * 1e-3) / 1e9,
                              "note": "64 B/element algorithmic (two sweeps); columns first, then rows"}
# spot check against torch on a sub-block is too large here; Parseval instead
ex = float((x.view(-1, 2) ** 2).sum()); ey = float((y.view(-1, 2) ** 2).sum())
out["C3_fft2_16384x16384"]["parseval_rel_err"] = abs(ey / (r * c) - ex) / ex
del x, y

# comparison only: cuFFT Z2Z through torch.fft on the C3 batch shape (256 x 2^20)
b, n = 256, 1 << 20
xc = torch.randn(b, n, dtype=torch.complex128, device="cuda")
med, best = timed(lambda: torch.fft.fft(xc, dim=1), iters=10)
out["cufft_comparison_256x2^20"] = {"median_ms": med, "GS_per_s": b * n / (med * 1e-3) / 1e9,
                                    "note": "torch.fft.fft (cuFFT Z2Z), comparison only, not on the product path"}
xg = torch.view_as_real(xc).contiguous().view(-1)
yg = torch.empty_like(xg)
med, best = timed(lambda: capi.check(L.gd_fft_batch_c2c_dev(xg.data_ptr(), yg.data_ptr(), n, b, 1, sp)), iters=10)
out["ours_256x2^20"] = {"median_ms": med, "GS_per_s": b * n / (med * 1e-3) / 1e9}
ref = torch.fft.fft(xc, dim=1)
got = torch.view_as_complex(yg.view(-1, 2)).view(b, n)
out["ours_vs_cufft_rel_l2"] = float((got - ref).norm() / ref.norm())
print(json.dumps(out, indent=1))

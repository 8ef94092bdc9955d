#!/usr/bin/env python3
"""FFT2 16384 x 16384 (device resident) under different inter-pass scratch budgets. usage: exp_fft2.py 1024 128 64 32"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "go-dsp_b200"))
from godsp import _capi as capi
L = capi.lib(); capi.check(L.gd_use_device(0))
R = Cc = 16384
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sp = C.c_void_p(st.cuda_stream)
x = torch.empty(R * Cc * 2, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), R * Cc * 2, 4, 0, sp)); torch.cuda.synchronize()
dims = (C.c_int64 * 2)(R, Cc)
ref_sum = None
for mb in [int(a) for a in sys.argv[1:]] or [1024]:
    capi.check(L.gd_set_option(b"pass_scratch_mb", mb))
    ts = []
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); capi.check(L.gd_fftn_c2c_dev(x.data_ptr(), y.data_ptr(), dims, 2, 1, sp)); e1.record(st)
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    chk = float((y.view(-1)[:: 65537] ** 2).sum())
    if ref_sum is None: ref_sum = chk
    ms = float(np.median(ts[1:]))
    print("pass_scratch_mb %5d: %.3f ms  %.1f Gelem/s  %.0f GB/s algorithmic  checksum rel diff %.1e" % (mb, ms, R * Cc / ms / 1e6, 64.0 * R * Cc / ms / 1e6, abs(chk - ref_sum) / ref_sum), flush=True)

#!/usr/bin/env python3
"""fft.FFT2 over matrix shapes, device resident (gd_fftn_c2c_dev): ms, Gelem/s and the fraction of the HBM roofline at the
algorithmic 64 B per element (two sweeps). usage: exp_fft2_shapes.py ["opt=val,..."] [RxC ...]"""
import ctypes as C, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "go-dsp_b200"))
from godsp import _capi as capi
L = capi.lib(); capi.check(L.gd_use_device(0))
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6542.1
args = sys.argv[1:]
opts = ""
if args and "=" in args[0]:
    opts = args.pop(0)
    for kv in opts.split(","):
        k, v = kv.split("="); capi.check(L.gd_set_option(k.encode(), int(v)))
shapes = [tuple(int(v) for v in a.split("x")) for a in args] or [(1024, 1024), (2048, 2048), (4096, 4096), (8192, 8192), (16384, 16384),
                                                                  (512, 131072), (131072, 512), (4096, 16384), (16384, 4096), (3000, 5000)]
st = torch.cuda.Stream(); sp = st.cuda_stream
for R, Cc in shapes:
    # batch of matrices so that small shapes are not one short launch: at least 2^26 elements per timed call
    src = torch.empty(R * Cc, dtype=torch.complex128, device="cuda")
    out = torch.empty_like(src)
    capi.check(L.gd_fill_splitmix_dev(src.data_ptr(), 2 * R * Cc, 4, 0, None)); capi.check(L.gd_stream_sync(None))
    dims = (C.c_int64 * 2)(R, Cc)
    reps = max(1, (1 << 26) // (R * Cc))
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for _ in range(2):
            capi.check(L.gd_fftn_c2c_dev(src.data_ptr(), out.data_ptr(), dims, 2, 1, sp))
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(reps):
                capi.check(L.gd_fftn_c2c_dev(src.data_ptr(), out.data_ptr(), dims, 2, 1, sp))
            e1.record(st)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / reps)
    ms = min(ts)
    ge = R * Cc / ms / 1e6
    print(json.dumps({"opts": opts, "rows": R, "cols": Cc, "ms": round(ms, 4), "gelem_s": round(ge, 2), "hbm_frac": round(ge * 64 / PEAK, 3)}), flush=True)
    del src, out

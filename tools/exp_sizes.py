#!/usr/bin/env python3
"""Device-resident batched complex128 FFT over sizes: total 2^28 points per call (inputs larger than L2), GS/s and the
fraction of the HBM roofline (32 B per point / MEASURED_PEAKS). usage: exp_sizes.py [log2n ...] (default 8..24 and a few Bluestein sizes)"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "go-dsp_b200"))
from godsp import _capi as capi
L = capi.lib(); capi.check(L.gd_use_device(0))
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6542.1
TOTAL = 1 << 28
x = torch.empty(TOTAL * 2, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), TOTAL * 2, 3, 0, None)); capi.check(L.gd_stream_sync(None))
st = torch.cuda.Stream(); sp = st.cuda_stream
args = sys.argv[1:]
ONE = "--one" in args          # a single transform per call (latency of one large transform) instead of 2^28 points
REAL = "--real" in args        # float64 input (fft.FFTReal = the transform of dsputils.ToComplex(x)); the roofline is then 24 B per point
args = [a for a in args if a not in ("--one", "--real")]
opts = ""
if args and "=" in args[0]:
    opts = args.pop(0)
    for kv in opts.split(","):
        k, v = kv.split("="); capi.check(L.gd_set_option(k.encode(), int(v)))
sizes = [int(a[1:]) if a.startswith("n") else 1 << int(a) for a in args] or [1 << k for k in range(8, 25)] + [1000, 4095, 100003, 1000003]
import time
for n in sizes:
    batch = 1 if ONE else max(1, TOTAL // n)
    if REAL:
        fn = lambda: capi.check(L.gd_fft_batch_r2c_full_dev(x.data_ptr(), y.data_ptr(), n, batch, 1, sp))
    else:
        fn = lambda: capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, batch, 1, sp))
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for _ in range(2): fn()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); host_ms = (time.perf_counter() - t0) * 1e3; torch.cuda.synchronize()
    gs = n * batch / ms / 1e6
    print(json.dumps({"opts": opts, "host_issue_ms": round(host_ms, 3), "n": n, "batch": batch, "ms": round(ms, 4), "gs": round(gs, 2), "hbm_frac": round(gs * (24 if REAL else 32) / PEAK, 3), **({"real_input": True} if REAL else {})}), flush=True)

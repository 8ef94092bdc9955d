#!/usr/bin/env python3
"""FFT2 16384 x 16384 timing through the C ABI under option sets. usage: exp_fft2.py "opt=val,..." ..."""
import ctypes as C, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "go-dsp_b200"))
from godsp import _capi as capi
L = capi.lib(); capi.check(L.gd_use_device(0))
R = Cc = 16384
src = torch.empty(R * Cc, dtype=torch.complex128, device="cuda")
out = torch.empty_like(src)
capi.check(L.gd_fill_splitmix_dev(src.data_ptr(), 2 * R * Cc, 4, 0, None)); capi.check(L.gd_stream_sync(None))
dims = (C.c_int64 * 2)(R, Cc)
st = torch.cuda.Stream(); sp = st.cuda_stream
DEFAULTS = {"l2_block_mb": 24, "chunk_streams": 2, "pass_scratch_mb": 1024, "l2_block_window": 1, "tma14": 1}
for combo in (sys.argv[1:] or [""]):
    for k0, v0 in DEFAULTS.items():
        capi.check(L.gd_set_option(k0.encode(), v0))
    for kv in combo.split(","):
        if kv:
            k, v = kv.split("="); capi.check(L.gd_set_option(k.encode(), int(v)))
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for _ in range(2):
            capi.check(L.gd_fftn_c2c_dev(src.data_ptr(), out.data_ptr(), dims, 2, 1, sp))
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            capi.check(L.gd_fftn_c2c_dev(src.data_ptr(), out.data_ptr(), dims, 2, 1, sp))
            e1.record(st)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    print(json.dumps({"opts": combo, "ms_best": min(ts), "ms_all": ts}), flush=True)

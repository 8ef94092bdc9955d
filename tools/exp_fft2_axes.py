#!/usr/bin/env python3
"""Per-axis timing of the 16384 x 16384 FFT2: columns (strided lines) and rows (batched) separately. usage: exp_fft2_axes.py "opt=val,..." ..."""
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
st = torch.cuda.Stream(); sp = st.cuda_stream
DEFAULTS = {"tma14": 1, "tma_delay": 2, "tma_slots": 3}
def timeit(fn):
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for _ in range(2): fn()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
for combo in (sys.argv[1:] or [""]):
    for k0, v0 in DEFAULTS.items(): capi.check(L.gd_set_option(k0.encode(), v0))
    for kv in combo.split(","):
        if kv:
            k, v = kv.split("="); capi.check(L.gd_set_option(k.encode(), int(v)))
    cols = timeit(lambda: capi.check(L.gd_fft_strided_c2c_dev(src.data_ptr(), out.data_ptr(), 1, R, Cc, 1, sp)))
    rows = timeit(lambda: capi.check(L.gd_fft_batch_c2c_dev(src.data_ptr(), out.data_ptr(), Cc, R, 1, sp)))
    rows_inplace = timeit(lambda: capi.check(L.gd_fft_batch_c2c_dev(out.data_ptr(), out.data_ptr(), Cc, R, 1, sp)))
    prof = {}
    if "--prof" in os.environ.get("EXP_FLAGS", ""):
        import numpy as np
        capi.check(L.gd_set_option(b"tma_prof", 1))
        for name, fn in (("cols", lambda: capi.check(L.gd_fft_strided_c2c_dev(src.data_ptr(), out.data_ptr(), 1, R, Cc, 1, sp))),
                         ("rows", lambda: capi.check(L.gd_fft_batch_c2c_dev(src.data_ptr(), out.data_ptr(), Cc, R, 1, sp)))):
            with torch.cuda.stream(st):
                fn()
            torch.cuda.synchronize()
            buf = np.zeros(148 * 32, np.int64)
            nc = L.gd_tma_profile_read(buf.ctypes.data, 148)
            b = buf[: nc * 32].reshape(nc, 32).astype(np.float64)
            m = b[:, :16].reshape(nc * 2, 8)
            tot = m[:, 6].mean() + 1e-9
            prof[name] = {k: round(float(m[:, i].mean() / tot), 4) for i, k in enumerate(["full0_p1", "full0_p2", "full1", "drained", "rd", "group_bar"])}
            prof[name]["tiles_per_group"] = float(m[:, 7].mean())
            sm = b[:, 16:24].reshape(nc * 2, 4)             # storer lanes: fractions of the consumer's total time
            prof[name].update({k: round(float(sm[:, i].mean() / tot), 4) for i, k in enumerate(["st_staged", "st_slot", "st_read", "st_publish"])})
        capi.check(L.gd_set_option(b"tma_prof", 0))
    print(json.dumps({"opts": combo, "prof": prof, "cols_ms": cols, "rows_ms": rows, "rows_inplace_ms": rows_inplace,
                      "cols_frac_of_6542": 32.0 * R * Cc / cols / 1e6 / 6542.1, "rows_frac_of_6542": 32.0 * R * Cc / rows / 1e6 / 6542.1}), flush=True)

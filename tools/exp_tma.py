#!/usr/bin/env python3
"""Timing / wait-profile experiments on the fused 2^20 kernel through the C ABI.
usage: exp_tma.py [--batch 512] [--iters 5] [--prof] "opt=val,opt=val" ...      (one line of JSON per option set)"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "go-dsp_b200"))
from godsp import _capi as capi

args = sys.argv[1:]; batch, iters, prof = 512, 5, False
while args and args[0].startswith("--"):
    k = args.pop(0)
    if k == "--prof": prof = True; continue
    v = int(args.pop(0))
    if k == "--batch": batch = v
    if k == "--iters": iters = v
L = capi.lib(); capi.check(L.gd_use_device(0))
DEFAULTS = {"tma": 1, "tma_delay": 2, "tma_slots": 3, "tma_opt": 0, "tma_prof": 0, "tma_grid_cap": 0}
n = 1 << 20
x = torch.empty(batch * n * 2, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), batch * n * 2, 3, 0, None)); capi.check(L.gd_stream_sync(None))
ex = (x.view(batch, -1) ** 2).sum(1)
st = torch.cuda.Stream(); sp = st.cuda_stream
NAMES = {0: "g0_full0_p1", 1: "g0_full0_p2", 2: "g0_full1", 3: "g0_wbuf", 4: "g0_rd", 5: "g0_total", 6: "g0_n_p1", 7: "g0_n_p2",
         8: "g1_full0_p1", 9: "g1_full0_p2", 10: "g1_full1", 11: "g1_wbuf", 12: "g1_rd", 13: "g1_total", 14: "g1_n_p1", 15: "g1_n_p2",
         16: "ld_claim", 17: "ld_done1", 18: "ld_done2", 19: "ld_freed", 20: "ld_total", 21: "ld_n_done1_waits",
         24: "st0_staged", 25: "st0_read", 26: "st0_publish", 27: "st0_slot", 28: "st1_staged", 29: "st1_read", 30: "st1_publish", 31: "st1_slot"}
for combo in (args or [""]):
    for k0, v0 in DEFAULTS.items():
        capi.check(L.gd_set_option(k0.encode(), v0))
    for kv in combo.split(","):
        if kv:
            k, v = kv.split("="); capi.check(L.gd_set_option(k.encode(), int(v)))
    if prof:
        capi.check(L.gd_set_option(b"tma_prof", 1))
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for _ in range(2):
            capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, batch, 1, sp))
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, batch, 1, sp))
            e1.record(st)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    ey = (y.view(batch, -1) ** 2).sum(1)
    bad = int((((ey / n - ex).abs() / ex) > 1e-12).sum())
    out = {"opts": combo, "batch": batch, "gs_best": batch * n / min(ts) / 1e6, "gs_median": batch * n / float(np.median(ts)) / 1e6,
           "gs_last_third": batch * n / float(np.mean(ts[-max(1, len(ts) // 3):])) / 1e6, "bad_rows": bad}
    if prof:
        buf = np.zeros(148 * 32, np.int64)
        nc = L.gd_tma_profile_read(buf.ctypes.data, 148)
        if nc > 0:
            m = buf[: nc * 32].reshape(nc, 32).astype(np.float64)
            tot = m[:, 5].mean() + 1e-9
            out["prof_frac_of_consumer_time"] = {NAMES[i]: round(float(m[:, i].mean() / tot), 4) for i in NAMES if "_n_" not in NAMES[i]}
            out["prof_counts"] = {NAMES[i]: float(m[:, i].mean()) for i in NAMES if "_n_" in NAMES[i]}
            out["consumer_cycles_mean"] = float(tot)
    print(json.dumps(out), flush=True)

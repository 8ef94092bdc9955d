#!/usr/bin/env python3
"""Pwelch kernel timing through the C ABI (device-resident). usage: exp_pwelch.py [--log2 30] [--nfft 4096] "opt=val,..." ..."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "go-dsp_b200"))
from godsp import _capi as capi
from godsp import window as gw
args = sys.argv[1:]; lg = 30
nfft = 4096
while args and args[0] in ("--log2", "--nfft"):
    if args[0] == "--log2": lg = int(args[1])
    else: nfft = int(args[1])
    args = args[2:]
L = capi.lib(); capi.check(L.gd_use_device(0))
ns, nov = 1 << lg, nfft // 2
stride = nfft - nov
nsegs = (ns - nfft) // stride + 1
lp = nfft // 2 + 1
x = torch.empty(ns, dtype=torch.float64, device="cuda")
capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), ns, 5, 0, None)); capi.check(L.gd_stream_sync(None))
dwin = torch.from_numpy(gw.Hann(nfft)).cuda()
raw = torch.empty(lp, dtype=torch.float64, device="cuda")
st = torch.cuda.Stream(); sp = st.cuda_stream
ref = None
for combo in (args or [""]):
    capi.check(L.gd_set_option(b"pwelch_bulk", 1))
    for kv in combo.split(","):
        if kv:
            k, v = kv.split("="); capi.check(L.gd_set_option(k.encode(), int(v)))
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        for _ in range(2):
            capi.check(L.gd_pwelch_partial_dev(x.data_ptr(), nfft, nov, nfft, lp, 0, nsegs, dwin.data_ptr(), raw.data_ptr(), sp))
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            capi.check(L.gd_pwelch_partial_dev(x.data_ptr(), nfft, nov, nfft, lp, 0, nsegs, dwin.data_ptr(), raw.data_ptr(), sp))
            e1.record(st)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    r = raw.cpu().numpy().copy()
    if ref is None:
        ref = r
    print(json.dumps({"opts": combo, "nfft": nfft, "log2_samples": lg, "msamples_per_s_best": ns / min(ts) / 1e3, "ms_best": min(ts),
                      "rel_diff_vs_first": float(np.linalg.norm(r - ref) / np.linalg.norm(ref))}), flush=True)

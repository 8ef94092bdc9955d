#!/usr/bin/env python3
"""Experiment driver: batched 2^20 complex128 FFT under several planner options, device-resident,
CUDA-event timing + a correctness check against torch.fft (cuFFT, comparison only).
usage: exp_fft.py [--batch 256] [--iters 5] "w32=0" "w32=1,tiled_scratch=1" ..."""
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

args = sys.argv[1:]
batch, iters, log2n = 256, 5, 20
while args and args[0].startswith("--"):
    k = args.pop(0)
    v = int(args.pop(0))
    if k == "--batch": batch = v
    elif k == "--iters": iters = v
    elif k == "--log2n": log2n = v
combos = args or ["w32=0"]
DEFAULTS = {"w32": 1, "fused": 0, "tiled_scratch": 0, "wide_tiles": 0, "pass_scratch_mb": 1024, "fused_delay": 2,
            "fused_slot_mb": 16, "l2_window": 1, "debug_alias": 0, "tma": 1, "tma_delay": 2, "tma_slots": 3, "tma_opt": 0}

L = capi.lib()
capi.check(L.gd_use_device(0))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sp = C.c_void_p(stream.cuda_stream)
n = 1 << log2n
x = torch.empty(batch * n * 2, dtype=torch.float64, device="cuda")
y = torch.empty(batch * n * 2, dtype=torch.float64, device="cuda")
capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), batch * n * 2, 3, 0, sp))
torch.cuda.synchronize()
nref = min(batch, 8)
xc = torch.view_as_complex(x.view(batch, n, 2))
ref = torch.fft.fft(xc[:nref], dim=1)
refn = torch.linalg.vector_norm(ref).item()
out = []
for combo in combos:
    opts = dict(DEFAULTS)
    for kv in combo.split(","):
        if kv:
            k, v = kv.split("=")
            opts[k] = int(v)
    for k, v in opts.items():
        capi.check(L.gd_set_option(k.encode(), v))
    y.zero_()
    def step():
        capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, batch, 1, sp))
    try:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); step(); e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        yc = torch.view_as_complex(y.view(batch, n, 2))
        err = (torch.linalg.vector_norm(yc[:nref] - ref) / refn).item()
        # last rows too (partial groups / tails)
        ref_last = torch.fft.fft(xc[batch - 1:], dim=1)
        err_last = (torch.linalg.vector_norm(yc[batch - 1:] - ref_last) / torch.linalg.vector_norm(ref_last)).item()
        ms = float(np.median(ts))
        r = {"opts": combo, "ms": ms, "best_ms": float(np.min(ts)), "GS/s": batch * n / ms / 1e6, "rel_err": err, "rel_err_last": err_last}
    except Exception as e:  # noqa
        r = {"opts": combo, "error": str(e)}
    print(json.dumps(r), flush=True)
    out.append(r)

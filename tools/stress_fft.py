#!/usr/bin/env python3
"""Race hunt for the fused 2^20 kernel: repeat a batched transform and check Parseval on every row (a wrong tile
changes a row's energy by O(1)). usage: stress_fft.py [--batch 256] [--reps 20] "opt=val,..." ..."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "go-dsp_b200"))
from godsp import _capi as capi
args = sys.argv[1:]; batch, reps = 256, 20
while args and args[0].startswith("--"):
    k = args.pop(0); v = int(args.pop(0))
    if k == "--batch": batch = v
    if k == "--reps": reps = v
L = capi.lib(); capi.check(L.gd_use_device(0))
n = 1 << 20
x = torch.empty(batch * n * 2, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), batch * n * 2, 3, 0, None)); capi.check(L.gd_stream_sync(None))
ex = (x.view(batch, -1) ** 2).sum(1)
for combo in (args or [""]):
    for k0, v0 in {"tma": 1, "tma_delay": 2, "tma_slots": 3, "tma_opt": 0}.items():
        capi.check(L.gd_set_option(k0.encode(), v0))
    for kv in combo.split(","):
        if kv:
            k, v = kv.split("="); capi.check(L.gd_set_option(k.encode(), int(v)))
    bad_runs, bad_rows = 0, set()
    for r in range(reps):
        y.zero_()
        torch.cuda.synchronize()      # torch's stream and the library's non-blocking stream are not ordered
        capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, batch, 1, None)); capi.check(L.gd_stream_sync(None))
        ey = (y.view(batch, -1) ** 2).sum(1)
        rel = ((ey / n - ex).abs() / ex)
        bad = (rel > 1e-12).nonzero().flatten().tolist()
        if bad:
            bad_runs += 1; bad_rows.update(bad)
            if bad_runs <= 3:
                for rr in bad[:2]:
                    xc = torch.view_as_complex(x.view(batch, n, 2)[rr]); yc = torch.view_as_complex(y.view(batch, n, 2)[rr])
                    d = (yc - torch.fft.fft(xc)).abs().view(1024, 1024)          # [k2][k1]
                    badk1 = (d.max(0).values > 1e-6).nonzero().flatten().tolist()
                    badk2 = (d.max(1).values > 1e-6).nonzero().flatten().tolist()
                    print("   rep %d row %d: %d bad elements; bad k1 columns: %d (%s...) bad k2 rows: %d (%s...)" % (r, rr, int((d > 1e-6).sum()),
                          len(badk1), badk1[:12], len(badk2), badk2[:12]), flush=True)
    print("opts [%s]: %d of %d runs had wrong rows; rows %s" % (combo, bad_runs, reps, sorted(bad_rows)[:20]), flush=True)

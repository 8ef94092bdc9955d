#!/usr/bin/env python3
"""One short invocation of the kernels added after tools/prof_targets.py was written, for `ncu --set full` (profiles/):
the 2^19 member of the fused family (64 transforms), one 2^24-point transform through the outer four-step (the fused column
kernel with the outer twiddle on its stores, then the row pass with the transposed store), the one-kernel Bluestein with a
padded length of 8192 (n = 4095, 4096 transforms) and the streaming Bluestein kernels around plain transforms (n = 30000,
512 transforms). Single process, single GPU."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "go-dsp_b200"))
from godsp import _capi as capi
L = capi.lib(); capi.check(L.gd_use_device(0))
def dev(n): return torch.empty(2 * n, dtype=torch.float64, device="cuda")
def fill(t, seed): capi.check(L.gd_fill_splitmix_dev(t.data_ptr(), t.numel(), seed, 0, None))
def sync(): capi.check(L.gd_stream_sync(None)); torch.cuda.synchronize()
x, y = dev(1 << 25), dev(1 << 25); fill(x, 3); sync()
for n, b in ((1 << 19, 64), (1 << 24, 1), (1 << 24, 2), (4095, 4096), (30000, 512)):
    for _ in range(2):      # the first call of a Bluestein length builds its plan
        capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, b, 1, None)); sync()
print("prof_targets_large ok; launches:", L.gd_kernel_launches())

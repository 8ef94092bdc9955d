#!/usr/bin/env python3
"""One short invocation of every hot kernel, for `ncu --set full` (profiles/): the fused 2^20 kernel (128 transforms), the
fused 2^14 kernel in both modes and the 2^15 / 2^16 / 2^18 members of its family, the bulk-fed Pwelch kernel (2^28 samples), the 32-point-per-thread pass kernel, the
GENERIC Bluestein passes (N = 1,000,003), an FFT2 strided axis of 4096-point lines, and both peer-memory exchange
kernels (world = 1: the stores go to this GPU's own buffer). Single process, single GPU."""
import ctypes as C, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "go-dsp_b200"))
from godsp import _capi as capi
L = capi.lib(); capi.check(L.gd_use_device(0))
def dev(n): return torch.empty(2 * n, dtype=torch.float64, device="cuda")
def fill(t, seed): capi.check(L.gd_fill_splitmix_dev(t.data_ptr(), t.numel(), seed, 0, None))
def sync(): capi.check(L.gd_stream_sync(None)); torch.cuda.synchronize()
# 1. fused 2^20 kernel, 128 transforms (4.29 GB algorithmic)
n, b = 1 << 20, 128
x, y = dev(n * b), dev(n * b); fill(x, 3); sync()
capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, b, 1, None)); sync()
del x, y
# 3. fused 2^14 kernel: columns of a 16384 x 2048 matrix, rows of a 2048 x 16384 matrix
m, o = dev(16384 * 2048), dev(16384 * 2048); fill(m, 4); sync()
capi.check(L.gd_fft_strided_c2c_dev(m.data_ptr(), o.data_ptr(), 1, 16384, 2048, 1, None)); sync()
capi.check(L.gd_fft_batch_c2c_dev(m.data_ptr(), o.data_ptr(), 16384, 2048, 1, None)); sync()
# 3b. the other sizes of the fused family: 2^16 = 256 x 256 (columns of a 65536 x 512 matrix, rows of 512 x 65536), 2^15 = 256 x 128
#     and 2^18 = 512 x 512 as batched rows (the same 2^25 points each)
capi.check(L.gd_fft_strided_c2c_dev(m.data_ptr(), o.data_ptr(), 1, 65536, 512, 1, None)); sync()
capi.check(L.gd_fft_batch_c2c_dev(m.data_ptr(), o.data_ptr(), 65536, 512, 1, None)); sync()
capi.check(L.gd_fft_batch_c2c_dev(m.data_ptr(), o.data_ptr(), 32768, 1024, 1, None)); sync()
capi.check(L.gd_fft_batch_c2c_dev(m.data_ptr(), o.data_ptr(), 262144, 128, 1, None)); sync()
# 4. an FFT2 axis of 4096-point strided lines (single pass kernel, column mode)
capi.check(L.gd_fft_strided_c2c_dev(m.data_ptr(), o.data_ptr(), 1, 4096, 8192, 1, None)); sync()
del m, o
# 5. Pwelch, 2^28 samples (2.15 GB algorithmic)
ns, nfft, nov = 1 << 28, 4096, 2048
s = torch.empty(ns, dtype=torch.float64, device="cuda"); fill(s, 5)
from godsp import window as gw
dwin = torch.from_numpy(gw.Hann(nfft)).cuda()
raw = torch.empty(nfft // 2 + 1, dtype=torch.float64, device="cuda")
sync()
capi.check(L.gd_pwelch_partial_dev(s.data_ptr(), nfft, nov, nfft, nfft // 2 + 1, 0, (ns - nfft) // (nfft - nov) + 1, dwin.data_ptr(), raw.data_ptr(), None)); sync()
del s
# 2. 32-point-per-thread pass kernel (the two-launch schedule of the same size, one transform)
capi.check(L.gd_set_option(b"tma", 0))
x2, y2 = dev(n), dev(n); fill(x2, 3); sync()
capi.check(L.gd_fft_batch_c2c_dev(x2.data_ptr(), y2.data_ptr(), n, 1, 1, None)); sync()
del x2, y2
capi.check(L.gd_set_option(b"tma", 1))
# 6. Bluestein, N = 1,000,003 (GENERIC passes with fused chirp / product / truncation), batch 8
nb, bb = 1000003, 1
xb, yb = dev(nb * bb), dev(nb * bb); fill(xb, 2); sync()
capi.check(L.gd_fft_batch_c2c_dev(xb.data_ptr(), yb.data_ptr(), nb, bb, 1, None)); sync()
del xb, yb
# 7. exchange kernels with world = 1 (own buffer): four-step exchange of a [4096][4096] slab, block copy of a 4096 x 4096 block
slab, recv = dev(4096 * 4096), dev(4096 * 4096); fill(slab, 6); sync()
ptrs = (C.c_void_p * 1)(recv.data_ptr())
capi.check(L.gd_fourstep_exchange_dev(slab.data_ptr(), ptrs, 4096, 4096, 0, 1, 24, None)); sync()
capi.check(L.gd_peer_block_copy_dev(slab.data_ptr(), ptrs, 1, 0, 4096, 4096, 4096, 4096, 0, 4096, None)); sync()
print("prof_targets ok; launches:", L.gd_kernel_launches())

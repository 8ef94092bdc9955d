"""One process driving several GPUs through the C ABI ("multi-GPU calls fan out inside C"): after gd_init(ndev)
the batched host-pointer calls split over the devices, one host thread each. Needs >= 2 GPUs (`gpurun --gpus 2`); also
covers what a single process touching device 1 needs (per-device kernel attributes)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import sys, os
import numpy as np
ROOT = sys.argv[1]; ndev = int(sys.argv[2])
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "go-dsp_b200"))
import oracle
from godsp import _capi as capi
L = capi.lib()
def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
# 1. a process that never calls gd_init but works on device 1 (dynamic shared-memory opt-in is per device)
capi.check(L.gd_use_device(1))
for n, b in ((1 << 20, 2), (1 << 14, 4), (4096, 8), (1000, 3)):
    x = oracle.splitmix_complex(n * b, 3).reshape(b, n)
    out = np.empty_like(x)
    capi.check(L.gd_fft_batch_c2c(x.ctypes.data, out.ctypes.data, n, b, 1))
    assert rel(out, oracle.fft_batch(x, threads=4)) <= 1e-12, ("device 1", n)
capi.check(L.gd_use_device(0))
# 2. fan-out over ndev devices
capi.check(L.gd_init(ndev))
n, b = 1 << 16, 64
x = oracle.splitmix_complex(n * b, 5).reshape(b, n)
out, back = np.empty_like(x), np.empty_like(x)
l0 = L.gd_kernel_launches()
capi.check(L.gd_fft_batch_c2c(x.ctypes.data, out.ctypes.data, n, b, 1))
capi.check(L.gd_fft_batch_c2c(out.ctypes.data, back.ctypes.data, n, b, -1))
assert rel(out, oracle.fft_batch(x, threads=8)) <= 1e-12 and rel(back, x) <= 1e-12
assert L.gd_device_count() >= ndev and L.gd_kernel_launches() > l0
# Pwelch: segment ranges per device, partial sums added in device order
xs = oracle.fill_splitmix(1 << 24, 5)
nfft, nov = 4096, 2048
want, _ = oracle.pwelch(xs, 1.0, nfft=nfft, noverlap=nov, threads=8)
win = oracle.window("hann", nfft); norm = 0.0
for v in win: norm += v * v
nsegs = oracle.segment_count(len(xs), nfft, nov)
pxx = np.empty(nfft // 2 + 1)
capi.check(L.gd_pwelch_f64(xs.ctypes.data, len(xs), nfft, nov, nfft, len(pxx), nsegs, win.ctypes.data, norm, pxx.ctypes.data))
assert rel(pxx, want) <= 1e-12
# FFT2: row blocks per device, both exchanges over peer memory
R, Cc = 2048, 4096
m = oracle.splitmix_complex(R * Cc, 4).reshape(R, Cc)
o2, b2 = np.empty_like(m), np.empty_like(m)
capi.check(L.gd_fft2_c2c(m.ctypes.data, o2.ctypes.data, R, Cc, 1))
assert rel(o2, oracle.fft2(m)) <= 1e-12
capi.check(L.gd_fft2_c2c(o2.ctypes.data, b2.ctypes.data, R, Cc, -1))
assert rel(b2, m) <= 1e-12
# ONE large power-of-two transform: sharded four-step over the devices (lines, fused twiddle + transpose + peer stores, lines),
# forward against the oracle and inverse as a round trip; first below the default threshold (2^26 points), then at it
capi.check(L.gd_set_option(b"fanout_min_log2n", 20))
for lg in (20, 23):
    n = 1 << lg
    x = oracle.splitmix_complex(n, 8)
    out, back = np.empty_like(x), np.empty_like(x)
    l0 = L.gd_kernel_launches()
    capi.check(L.gd_fft_c2c(x.ctypes.data, out.ctypes.data, n, 1))
    nl = L.gd_kernel_launches() - l0
    assert nl >= 3 * ndev, ("the transform was not sharded", nl)
    assert rel(out, oracle.fft(x)) <= 1e-12, ("sharded 1-D forward", lg)
    capi.check(L.gd_fft_c2c(out.ctypes.data, back.ctypes.data, n, -1))
    assert rel(back, x) <= 1e-12, ("sharded 1-D inverse", lg)
    assert rel(back, oracle.ifft(out)) <= 1e-12
capi.check(L.gd_set_option(b"fanout_min_log2n", 26))
n = 1 << 26
x = oracle.splitmix_complex(n, 9)
out = np.empty_like(x)
capi.check(L.gd_fft_c2c(x.ctypes.data, out.ctypes.data, n, 1))
assert rel(out, oracle.fft(x)) <= 1e-12, "sharded 1-D forward 2^26"
print("MULTI_DEVICE_OK", ndev)
'''


def test_one_process_two_devices():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: these tests must run on the GPU box (there is no CPU fallback)")
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("one GPU visible; driving two devices from one process needs `gpurun --gpus 2`")
    r = subprocess.run([sys.executable, "-c", WORKER, ROOT, str(min(ndev, 4))], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_DEVICE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]

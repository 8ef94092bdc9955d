import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/go-dsp_b200')
import numpy as np
import oracle
from godsp import _capi as c
L = c.lib()
c.check(L.gd_init(1))
def fft(x, dir=1):
    x = np.ascontiguousarray(x, np.complex128); out = np.empty_like(x)
    c.check(L.gd_fft_c2c(c.ptr(x), c.ptr(out), x.shape[0], dir)); return out
def rel(a, b): return np.linalg.norm(a-b)/max(np.linalg.norm(b), 1e-300)
worst = 0
for lg in range(1, 22):
    n = 1 << lg
    x = oracle.splitmix_complex(n, 1)
    t=time.time(); got = fft(x); dt=time.time()-t
    want = oracle.fft(x) if lg <= 20 else np.fft.fft(x)
    e = rel(got, want); ei = rel(fft(want, -1), x)
    worst = max(worst, e, ei)
    print("n=2^%d fwd %.2e inv %.2e  (%.1f ms)" % (lg, e, ei, dt*1e3), flush=True)
for n in [1, 3, 5, 6, 7, 12, 100, 1000, 4099, 65537, 1000003]:
    x = oracle.splitmix_complex(n, 2)
    got = fft(x); want = oracle.fft(x)
    e = rel(got, want); ei = rel(fft(x, -1), oracle.ifft(x))
    print("n=%d fwd %.2e inv %.2e vs numpy %.2e" % (n, e, ei, rel(got, np.fft.fft(x))), flush=True)
# batch
x = oracle.splitmix_complex(37*1024, 3).reshape(37, 1024); out = np.empty_like(x)
c.check(L.gd_fft_batch_c2c(c.ptr(x), c.ptr(out), 1024, 37, 1)); print("batch 37x1024", rel(out, np.fft.fft(x, axis=1)))
x = oracle.splitmix_complex(5*(1<<14), 3).reshape(5, 1<<14); out = np.empty_like(x)
c.check(L.gd_fft_batch_c2c(c.ptr(x), c.ptr(out), 1<<14, 5, 1)); print("batch 5x2^14", rel(out, np.fft.fft(x, axis=1)))
for (b, lg) in [(7, 16), (3, 18), (5, 20), (70, 16)]:
    x = oracle.splitmix_complex(b*(1<<lg), 3).reshape(b, 1<<lg); out = np.empty_like(x)
    c.check(L.gd_fft_batch_c2c(c.ptr(x), c.ptr(out), 1<<lg, b, 1)); e = rel(out, np.fft.fft(x, axis=1))
    c.check(L.gd_fft_batch_c2c(c.ptr(x), c.ptr(out), 1<<lg, b, -1)); print("batch %dx2^%d" % (b, lg), e, rel(out, np.fft.ifft(x, axis=1)))
# real
r = oracle.fill_splitmix(1000, 4); out = np.empty(1000, np.complex128)
c.check(L.gd_fft_r2c_full(c.ptr(r), c.ptr(out), 1000, 1)); print("fftreal 1000", rel(out, oracle.fft_real(r)))
c.check(L.gd_fft_r2c_full(c.ptr(r), c.ptr(out), 1000, -1)); print("ifftreal 1000", rel(out, oracle.ifft_real(r)))
r = oracle.fill_splitmix(8192, 4); out = np.empty(8192, np.complex128)
c.check(L.gd_fft_r2c_full(c.ptr(r), c.ptr(out), 8192, 1)); print("fftreal 8192", rel(out, oracle.fft_real(r)))
# convolve
a, b = oracle.splitmix_complex(48, 1), oracle.splitmix_complex(48, 2); out = np.empty_like(a)
c.check(L.gd_convolve_c2c(c.ptr(a), c.ptr(b), c.ptr(out), 48)); print("convolve 48", rel(out, oracle.convolve(a, b)))
# fft2 / fftn
for shape in [(2,3), (3,5), (64, 32), (300, 7), (8192, 16), (16, 8192), (2,2,3), (4, 6, 8, 5)]:
    x = oracle.splitmix_complex(int(np.prod(shape)), 5).reshape(shape); out = np.empty_like(x)
    dims = (c.C.c_int64 * len(shape))(*shape)
    c.check(L.gd_fftn_c2c(c.ptr(x), c.ptr(out), dims, len(shape), 1)); e = rel(out, np.fft.fftn(x))
    c.check(L.gd_fftn_c2c(c.ptr(x), c.ptr(out), dims, len(shape), -1)); print("fftn", shape, e, rel(out, np.fft.ifftn(x)))
# pwelch
for (nx, nfft, nov, pad) in [(100, 256, 0, 0), (5000, 256, 128, 0), (100000, 4096, 2048, 0), (50000, 1024, 512, 2048), (5000, 100, 30, 0), (5000, 256, 0, 128), (9000, 4096, 2048, 0)]:
    x = oracle.fill_splitmix(nx, 5)
    pw, fw = oracle.pwelch(x, 2.0, nfft=nfft, pad=pad, noverlap=nov)
    n_eff = nfft or 256; p_eff = pad or n_eff
    xx = x if nx >= n_eff else np.concatenate([x, np.zeros(n_eff-nx)])
    fftlen = max(p_eff, n_eff); lp = p_eff//2+1
    nsegs = oracle.segment_count(len(xx), n_eff, nov)
    win = oracle.window("hann", fftlen); norm = float(np.sum(oracle.window("hann", n_eff)**2))*2.0
    # reference norm accumulates sequentially
    nrm = 0.0
    for v in oracle.window("hann", n_eff): nrm += v*v
    nrm *= 2.0
    pxx = np.empty(lp)
    c.check(L.gd_pwelch_f64(c.ptr(xx), len(xx), n_eff, nov, fftlen, lp, nsegs, c.ptr(win), nrm, c.ptr(pxx)))
    print("pwelch", (nx, nfft, nov, pad), "nsegs", nsegs, rel(pxx, pw))
print("launches", L.gd_kernel_launches())

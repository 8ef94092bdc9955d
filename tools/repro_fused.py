import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/go-dsp_b200')
import numpy as np, oracle
from godsp import _capi as c
L = c.lib(); c.check(L.gd_use_device(0))
b, lg = int(sys.argv[1]), int(sys.argv[2])
if len(sys.argv) > 3: c.check(L.gd_set_option(b"fused_slot_mb", int(sys.argv[3])))
x = oracle.splitmix_complex(b*(1<<lg), 3).reshape(b, 1<<lg); out = np.empty_like(x)
c.check(L.gd_fft_batch_c2c(c.ptr(x), c.ptr(out), 1<<lg, b, 1))
print(b, lg, "rel", np.linalg.norm(out-np.fft.fft(x,axis=1))/np.linalg.norm(out))

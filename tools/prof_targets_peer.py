#!/usr/bin/env python3
"""The fused lines + twiddle + exchange kernel (TW2 = 2) and the segmented-row pass on ONE GPU (world = 1: the peer stores
land in this GPU's own receive buffer), 2^28 points = 2^14 x 2^14, for `ncu --set full` (profiles/r2_ncu_peer_summary.json)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "go-dsp_b200"))
from godsp import _capi as capi
L = capi.lib(); capi.check(L.gd_use_device(0))
n1 = n2 = 1 << 14
slab = torch.empty(2 * n1 * n2, dtype=torch.float64, device="cuda")
recv = torch.empty_like(slab); out = torch.empty_like(slab)
capi.check(L.gd_fill_splitmix_dev(slab.data_ptr(), slab.numel(), 6, 0, None)); capi.check(L.gd_stream_sync(None))
assert L.gd_fourstep_fused_supported(n1, n2, 1) == 1
ptrs = (C.c_void_p * 1)(recv.data_ptr())
for _ in range(2):
    capi.check(L.gd_fourstep_lines_peer_dev(slab.data_ptr(), ptrs, n1, n2, 0, 1, 28, 1, None))
    capi.check(L.gd_fourstep_rows_seg_dev(recv.data_ptr(), out.data_ptr(), n2, n1, 1, 1, None))
    capi.check(L.gd_stream_sync(None))
print("prof_targets_peer ok; launches:", L.gd_kernel_launches())

#!/usr/bin/env python3
"""Sharded paths under torchrun: one 2^(29 + log2 G)-point transform and FFT2 16384^2 over G ranks, timed per option set.
usage: torchrun ... tools/exp_sharded.py "opt=val,..." ...        (rank 0 prints one JSON line per option set)"""
import json, os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "go-dsp_b200"))
from godsp import _capi as capi
from godsp import distributed as D
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = capi.lib(); capi.check(L.gd_use_device(local))
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
ops = D.DeviceOps()
lg = 29 + (world.bit_length() - 1)
n = 1 << lg
n1, n2, k, w = D.split_1d(n, world)
src = torch.empty(n1 * w, dtype=torch.complex128, device="cuda")
work = torch.empty(n1 * w, dtype=torch.complex128, device="cuda")
capi.check(L.gd_fill_splitmix_dev(src.data_ptr(), 2 * n1 * w, 6, 2 * rank * n1 * w, st.cuda_stream))
px = D.PeerExchange(n1 * w, ops)
R = Cc = 16384
rg = R // world
m = torch.empty(rg * Cc, dtype=torch.complex128, device="cuda")
res = torch.empty_like(m)
capi.check(L.gd_fill_splitmix_dev(m.data_ptr(), 2 * rg * Cc, 4, 2 * rank * rg * Cc, st.cuda_stream))
peers = (D.PeerExchange(rg * Cc, ops), D.PeerExchange(rg * Cc, ops))
DEFAULTS = {"l2_block_mb": 24, "chunk_streams": 2, "fourstep_pipeline": 0, "fourstep_pipeline_mb": 256, "fourstep_exchange_ctas": 0, "fourstep_lines_sms": 0, "tma14": 1}

def timeit(fn, reps=3):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st)
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

for combo in (sys.argv[1:] or [""]):
    for k0, v0 in DEFAULTS.items(): capi.check(L.gd_set_option(k0.encode(), v0))
    for kv in combo.split(","):
        if kv:
            kk, vv = kv.split("="); capi.check(L.gd_set_option(kk.encode(), int(vv)))
    t1 = timeit(lambda: D.fft_1d_sharded(src, n, ops, work=work, peer=px))
    # phases of the 1-D transform, each timed alone (max over ranks)
    tl1 = timeit(lambda: ops.fft_strided(src, work, 1, n1, w, 1))
    tex = timeit(lambda: (px.fence(), px.exchange(work, n1, w, lg), px.fence()))
    tl2 = timeit(lambda: ops.fft_strided(px.recv, work, 1, n2, k, 1))
    t2 = timeit(lambda: D.fft2_sharded(m, R, Cc, ops, peers=peers, out=res, fused=False), reps=5)
    t2f = timeit(lambda: D.fft2_sharded(m, R, Cc, ops, peers=peers, out=res, fused=True), reps=5)
    # the exchange fused into the first line pass (TMA stores into the peers' receive buffers), segmented rows
    fused = {}
    if px.fused_supported(n1, n2):
        fused["fft1d_fused_ms"] = timeit(lambda: D.fft_1d_sharded(src, n, ops, work=work, peer=px, fused=True))
        fused["lines_peer_alone_ms"] = timeit(lambda: (px.fence(), px.lines_peer(src, n1, w, lg), px.fence()))
        fused["rows_seg_alone_ms"] = timeit(lambda: px.rows_seg(work, n2, k))
    if rank == 0:
        print(json.dumps({"opts": combo, "world": world, "fft1d_log2n": lg, "fft1d_ms": t1, "lines1_alone_ms": tl1, "exchange_alone_ms": tex,
                          "lines2_alone_ms": tl2, "fft2_sharded_ms": t2, "fft2_sharded_fused_ms": t2f, **fused}), flush=True)
px.close(); peers[0].close(); peers[1].close()
dist.destroy_process_group()

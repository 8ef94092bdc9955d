"""world_size-2 gloo test of the N>1 path's host logic: batch rows and Pwelch segment ranges are
sharded with godsp.sharding, each rank produces its share with the ORACLE standing in for the
device (no GPU here), partial PSD sums are all-gathered and added in rank order, and the result
must equal the single-rank answer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle

NFFT, NOV, NS, BATCH, N = 256, 128, 1 << 14, 6, 1 << 10


def _raw_partial(x, nseg, win):
    acc = np.zeros(NFFT // 2 + 1)
    stride = NFFT - NOV
    for s in range(nseg):
        X = oracle.fft_real(x[s * stride: s * stride + NFFT] * win)[: NFFT // 2 + 1]
        acc += X.real ** 2 + X.imag ** 2
    return acc


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "go-dsp_b200"))
    from godsp import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # batched FFT rows: no collective in the data path, only the result gather for the check
    r0, r1 = sharding.batch_rows(rank, world, BATCH)
    rows = np.stack([oracle.fft(oracle.splitmix_complex(N, 3, b << 21)) for b in range(r0, r1)])
    got = [None] * world
    dist.all_gather_object(got, (r0, rows))
    # Pwelch: segment range + halo, partial sums, all-gather, rank-order sum
    s0, s1, x0, x1 = sharding.pwelch_segment_range(rank, world, NS, NFFT, NOV)
    x = oracle.fill_splitmix(x1 - x0, 5, x0)
    win = oracle.window("hann", NFFT)
    part = torch.from_numpy(_raw_partial(x, s1 - s0, win))
    parts = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(parts, part)
    tot = sharding.reduce_partials(parts).numpy()
    if rank == 0:
        out["rows"] = np.concatenate([g[1] for g in sorted(got, key=lambda t: t[0])])
        out["raw"] = tot
    dist.barrier()
    dist.destroy_process_group()


def test_world2_matches_single_rank():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    want_rows = np.stack([oracle.fft(oracle.splitmix_complex(N, 3, b << 21)) for b in range(BATCH)])
    assert np.array_equal(out["rows"], want_rows)
    x = oracle.fill_splitmix(NS, 5)
    win = oracle.window("hann", NFFT)
    nsegs = oracle.segment_count(NS, NFFT, NOV)
    raw = out["raw"]
    pxx = raw / nsegs
    pxx[1:-1] *= 2
    pxx /= np.sum(win ** 2)
    want, _ = oracle.pwelch(x, 1.0, nfft=NFFT, noverlap=NOV)
    assert np.linalg.norm(pxx - want) / np.linalg.norm(want) < 1e-13

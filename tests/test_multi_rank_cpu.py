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


# ---------------------------------------------------------------------------------------------
# the two exchange paths (godsp.distributed): one sharded 1-D transform (four-step, one all-to-all)
# and FFT2 on row blocks (two all-to-alls).  The ORACLE stands in for the device kernels; what is
# checked here is the data movement: slab layouts, all-to-all splits, twiddle offsets, repacks.
class OracleOps:
    def empty(self, nelem):
        return torch.empty(nelem, dtype=torch.complex128)

    @staticmethod
    def _fft(line, direction):
        return oracle.fft(line) if direction > 0 else oracle.ifft(line)

    def fft_strided(self, src, dst, outer, length, stride, direction=1):
        a = src.numpy().reshape(outer, length, stride).copy()
        for o in range(outer):
            for c in range(stride):
                a[o, :, c] = self._fft(np.ascontiguousarray(a[o, :, c]), direction)
        dst.copy_(torch.from_numpy(a.reshape(-1)))

    def fft_rows(self, src, dst, n, batch, direction=1):
        a = src.numpy().reshape(batch, n)
        dst.copy_(torch.from_numpy(np.stack([self._fft(np.ascontiguousarray(r), direction) for r in a]).reshape(-1)))

    def fourstep_twiddle(self, blk, rows, cols, row0, col0, log2n):
        r = np.arange(row0, row0 + rows, dtype=np.int64)[:, None]
        c = np.arange(col0, col0 + cols, dtype=np.int64)[None, :]
        e = (r * c) % (1 << log2n)
        blk.copy_(torch.from_numpy((blk.numpy().reshape(rows, cols) * np.exp(-2j * np.pi * e / (1 << log2n))).reshape(-1)))

    def swap_leading(self, src, dst, a, b, w):
        dst.copy_(torch.from_numpy(np.ascontiguousarray(src.numpy().reshape(a, b, w).transpose(1, 0, 2)).reshape(-1)))

    def transpose_batched(self, src, dst, batch, rows, cols):
        dst.copy_(torch.from_numpy(np.ascontiguousarray(src.numpy().reshape(batch, rows, cols).transpose(0, 2, 1)).reshape(-1)))


N1D, R2D, C2D = 1 << 10, 8, 16


def _worker_exchange(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "go-dsp_b200"))
    from godsp import distributed as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ops = OracleOps()
    x = torch.from_numpy(oracle.splitmix_complex(N1D, 6))
    slab = D.scatter_signal(x, N1D, rank, world)
    spec = D.fft_1d_sharded(slab, N1D, ops)
    got = [None] * world
    dist.all_gather_object(got, spec.numpy())
    m = oracle.splitmix_complex(R2D * C2D, 4).reshape(R2D, C2D)
    rg = R2D // world
    blk = torch.from_numpy(m[rank * rg:(rank + 1) * rg].copy().reshape(-1))
    res = D.fft2_sharded(blk, R2D, C2D, ops)
    back = D.fft2_sharded(res.clone(), R2D, C2D, ops, direction=-1)
    got2 = [None] * world
    dist.all_gather_object(got2, (res.numpy(), back.numpy()))
    if rank == 0:
        out["spec"] = D.gather_spectrum([torch.from_numpy(g) for g in got], N1D).numpy()
        out["fft2"] = np.concatenate([g[0] for g in got2]).reshape(R2D, C2D)
        out["ifft2"] = np.concatenate([g[1] for g in got2]).reshape(R2D, C2D)
    dist.barrier()
    dist.destroy_process_group()


def test_world2_exchange_paths():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_exchange, args=(2, port, out), nprocs=2, join=True)

    def rel(a, b):
        return np.linalg.norm(a - b) / np.linalg.norm(b)

    want = oracle.fft(oracle.splitmix_complex(N1D, 6))
    assert rel(out["spec"], want) < 1e-13
    m = oracle.splitmix_complex(R2D * C2D, 4).reshape(R2D, C2D)
    want2 = oracle.fft2(m)
    assert rel(out["fft2"], want2) < 1e-13
    assert rel(out["ifft2"], m) < 1e-13


def test_split_1d_shapes():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "go-dsp_b200"))
    from godsp import distributed as D
    assert D.split_1d(1 << 32, 8) == (1 << 16, 1 << 16, 1 << 13, 1 << 13)      # BASELINE config 5
    assert D.split_1d(1 << 21, 2) == (1 << 11, 1 << 10, 1 << 10, 1 << 9)
    with pytest.raises(ValueError):
        D.split_1d(1000, 2)
    with pytest.raises(ValueError):
        D.split_1d(1 << 4, 8)


def test_gather_spectrum_layouts():
    """The two result layouts of the sharded transform map back to the same natural-order spectrum: rank h's [N2][K] slab
    (separate exchange kernel) and its transposed [K][N2] block (exchange fused into the first line pass)."""
    import sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "go-dsp_b200"))
    from godsp import distributed as D
    n, world = 1 << 10, 4
    n1, n2, k, w = D.split_1d(n, world)
    spec = torch.arange(n, dtype=torch.float64).to(torch.complex128)          # X[k1 + N1 k2] at natural index k2 * N1 + k1
    nat = spec.view(n2, n1)
    slabs = [nat[:, h * k:(h + 1) * k].contiguous().view(-1) for h in range(world)]
    blocks = [nat[:, h * k:(h + 1) * k].t().contiguous().view(-1) for h in range(world)]
    assert torch.equal(D.gather_spectrum(slabs, n), spec)
    assert torch.equal(D.gather_spectrum(blocks, n, fused=True), spec)

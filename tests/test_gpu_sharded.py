"""GPU tests of the exchange paths (godsp.distributed) through the C ABI: world size 1 on one GPU
(every building-block kernel: strided lines, four-step twiddle, batched transpose, repack) against
the CPU oracle, and -- when the box has two GPUs -- two NCCL ranks against the same oracle result."""
import os
import socket

import numpy as np
import pytest

import oracle
from conftest import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, log2n, rows, cols, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "go-dsp_b200"))
    import torch
    import torch.distributed as dist
    from godsp import distributed as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ops = D.DeviceOps()
    n = 1 << log2n
    x = torch.from_numpy(oracle.splitmix_complex(n, 6))
    slab = D.scatter_signal(x, n, rank, world).cuda()
    spec = D.fft_1d_sharded(slab, n, ops)
    torch.cuda.synchronize()
    # the same transform with the exchange step as one kernel over peer memory (CUDA IPC + NVLink stores)
    px = D.PeerExchange(n // world, ops)
    slab2 = D.scatter_signal(x, n, rank, world).cuda()
    spec_p = D.fft_1d_sharded(slab2, n, ops, peer=px).clone()
    slab2 = D.scatter_signal(x, n, rank, world).cuda()
    spec_p2 = D.fft_1d_sharded(slab2, n, ops, peer=px).clone()     # buffer reuse across calls
    torch.cuda.synchronize()
    assert torch.equal(spec_p, spec_p2)
    px.close()
    m = oracle.splitmix_complex(rows * cols, 4).reshape(rows, cols)
    rg = rows // world
    blk = torch.from_numpy(m[rank * rg:(rank + 1) * rg].copy().reshape(-1)).cuda()
    res = D.fft2_sharded(blk, rows, cols, ops)
    back = D.fft2_sharded(res.clone(), rows, cols, ops, direction=-1)
    torch.cuda.synchronize()
    # the same through peer memory, twice (buffer reuse)
    pp = (D.PeerExchange(rg * cols, ops), D.PeerExchange(rg * cols, ops))
    blk2 = torch.from_numpy(m[rank * rg:(rank + 1) * rg].copy().reshape(-1)).cuda()
    res_p = D.fft2_sharded(blk2, rows, cols, ops, peers=pp).clone()
    back_p = D.fft2_sharded(res_p.clone(), rows, cols, ops, direction=-1, peers=pp).clone()
    torch.cuda.synchronize()
    assert torch.equal(res_p, res) and torch.equal(back_p, back)     # same kernels, same order: bit-identical
    pp[0].close()
    pp[1].close()
    out[rank] = (spec.cpu().numpy(), res.cpu().numpy(), back.cpu().numpy(), spec_p.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def _run(world, log2n, rows, cols):
    import torch
    import torch.multiprocessing as mp
    from godsp import distributed as D
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), log2n, rows, cols, out), nprocs=world, join=True)
    n = 1 << log2n
    spec = D.gather_spectrum([torch.from_numpy(out[r][0]) for r in range(world)], n).numpy()
    want = oracle.fft(oracle.splitmix_complex(n, 6))
    assert rel_l2(spec, want) <= TOL
    spec_p = D.gather_spectrum([torch.from_numpy(out[r][3]) for r in range(world)], n).numpy()
    assert rel_l2(spec_p, want) <= TOL
    m = oracle.splitmix_complex(rows * cols, 4).reshape(rows, cols)
    got = np.concatenate([out[r][1] for r in range(world)]).reshape(rows, cols)
    assert rel_l2(got, oracle.fft2(m)) <= TOL
    back = np.concatenate([out[r][2] for r in range(world)]).reshape(rows, cols)
    assert rel_l2(back, m) <= TOL


@pytest.mark.parametrize("log2n,rows,cols", [(12, 8, 16), (20, 512, 1024), (23, 8192, 64)])
def test_sharded_paths_one_rank(log2n, rows, cols):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: these tests must run on the GPU box (there is no CPU fallback)")
    _run(1, log2n, rows, cols)


@pytest.mark.parametrize("log2n,rows,cols", [(12, 8, 16), (22, 2048, 512)])
def test_sharded_paths_two_ranks(log2n, rows, cols):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: these tests must run on the GPU box (there is no CPU fallback)")
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible; the 2-rank NCCL run needs `gpurun --gpus 2`")
    _run(2, log2n, rows, cols)


def _fused_worker(rank, world, port, log2n, path, out):
    """fft_1d_sharded(fused=True): lines + twiddle + NVLink stores in one kernel, segmented rows; this rank's [K][N2] block against
    the oracle spectrum (a .npy the parent wrote), twice (receive-buffer reuse), and against the unfused peer-memory formulation."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "go-dsp_b200"))
    import torch
    import torch.distributed as dist
    from godsp import distributed as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ops = D.DeviceOps()
    n = 1 << log2n
    n1, n2, k, w = D.split_1d(n, world)
    x = torch.from_numpy(oracle.splitmix_complex(n, 6))
    slab = D.scatter_signal(x, n, rank, world).cuda()
    keep = slab.clone()
    px = D.PeerExchange(n // world, ops)
    assert px.fused_supported(n1, n2)
    got = D.fft_1d_sharded(slab, n, ops, peer=px, fused=True).clone()
    got2 = D.fft_1d_sharded(slab, n, ops, peer=px, fused=True).clone()
    plain = D.fft_1d_sharded(slab, n, ops, peer=px).clone()            # [N2][K]
    torch.cuda.synchronize()
    assert torch.equal(slab, keep), "the input slab was modified"
    assert torch.equal(got, got2)
    want = torch.from_numpy(np.load(path, mmap_mode="r").reshape(n2, n1)[:, rank * k:(rank + 1) * k].T.copy()).cuda()
    err = float((got.view(k, n2) - want).norm() / want.norm())
    err_plain = float((got.view(k, n2) - plain.view(n2, k).t()).norm() / want.norm())
    px.close()
    # fft.FFT2 on row blocks, 8192 x 8192: the column pass whose stores are the second exchange + segmented rows give the bits of
    # the block-copy formulation (same kernels, same arithmetic), forward and inverse
    rows = cols = 8192
    rg = rows // world
    blk = torch.empty(rg * cols, dtype=torch.complex128, device="cuda")
    from godsp import _capi as capi
    capi.check(ops.L.gd_fill_splitmix_dev(blk.data_ptr(), 2 * rg * cols, 4, 2 * rank * rg * cols, ops._sp()))
    pp = (D.PeerExchange(rg * cols, ops), D.PeerExchange(rg * cols, ops))
    assert pp[1].fused_supported(rows, cols) or world == 1
    keep2 = blk.clone()
    res_f = D.fft2_sharded(blk.clone(), rows, cols, ops, peers=pp, fused=True).clone()
    res_u = D.fft2_sharded(blk.clone(), rows, cols, ops, peers=pp, fused=False).clone()
    back_f = D.fft2_sharded(res_f.clone(), rows, cols, ops, direction=-1, peers=pp, fused=True).clone()
    torch.cuda.synchronize()
    same = bool(torch.equal(res_f, res_u))
    rt = float((back_f - keep2).norm() / keep2.norm())
    pp[0].close()
    pp[1].close()
    out[rank] = (err, err_plain, same, rt)
    dist.barrier()
    dist.destroy_process_group()


def _run_fused(world, log2n, tmp_path):
    import torch.multiprocessing as mp
    n = 1 << log2n
    path = str(tmp_path / "want.npy")
    np.save(path, oracle.fft(oracle.splitmix_complex(n, 6)))
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_fused_worker, args=(world, _free_port(), log2n, path, out), nprocs=world, join=True)
    for r in range(world):
        assert out[r][0] <= TOL and out[r][1] <= 1e-13 and out[r][2] and out[r][3] <= TOL, (r, out[r])


def test_sharded_fused_exchange_one_rank(tmp_path):       # 2^26 = 2^13 x 2^13: the smallest size both line passes are fused for
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: these tests must run on the GPU box (there is no CPU fallback)")
    _run_fused(1, 26, tmp_path)


@pytest.mark.parametrize("log2n", [26, 27])
def test_sharded_fused_exchange_two_ranks(log2n, tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: these tests must run on the GPU box (there is no CPU fallback)")
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible; the 2-rank run needs `gpurun --gpus 2`")
    _run_fused(2, log2n, tmp_path)

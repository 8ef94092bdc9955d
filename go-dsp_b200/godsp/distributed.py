"""The two paths of SURVEY.md 8(e) that need an exchange step, one process per GPU over torch.distributed:

* `fft_1d_sharded` -- ONE transform of N = N1*N2 points over G ranks (BASELINE config 5: 2^32 points on
  8 GPUs) as a four-step decomposition with a single all-to-all.  Index maps (same as the single-GPU
  four-step in csrc/engine.cu, which replaces the log2 N sweeps of fft/radix2.go:131-151):
      n = n1*N2 + n2,   k = k1 + N1*k2,
      X[k1 + N1*k2] = sum_n2 w_N2^(n2 k2) * w_N^(n2 k1) * sum_n1 w_N1^(n1 k1) x[n1*N2 + n2].
  Layout: the signal viewed as a row-major [N1][N2] matrix, rank g holds the column slab
  [N1][g*W, (g+1)*W), W = N2/G; the spectrum viewed as a row-major [N2][N1] matrix, rank g holds the
  column slab [N2][g*K, (g+1)*K), K = N1/G.  Input and output have the same kind of layout, so no extra
  transposes are needed: strided length-N1 lines -> twiddle -> all-to-all -> per-source [K][W]->[W][K]
  transpose -> strided length-N2 lines.  `scatter_signal` / `gather_spectrum` give the maps from/to natural order.

* `fft2_sharded` -- fft.FFT2 (fft/fft.go:123-154: every column, then every row) on a matrix whose row
  blocks are spread over the ranks; columns become local through one all-to-all, and a second one
  restores the row-block layout.

The arithmetic is done by an `ops` object.  `DeviceOps` is the product: every method is one call into
the C ABI (include/godsp_b200.h) on CUDA tensors, and raises if the library or the GPU is missing.
The CPU test suite passes its own stand-in (tests/test_multi_rank_cpu.py) to check the data movement
under gloo; nothing in this module computes on the host.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _capi


def _ilog2(n):
    lg = n.bit_length() - 1
    if n < 1 or (1 << lg) != n:
        raise ValueError("power of two expected, got %d" % n)
    return lg


class DeviceOps:
    """C-ABI-backed building blocks; tensors are 1-D complex128 CUDA tensors."""

    def __init__(self, device=None, stream=None):
        if not torch.cuda.is_available():
            raise RuntimeError("go-dsp_b200: no CUDA device; there is no CPU fallback")
        self.L = _capi.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        _capi.check(self.L.gd_use_device(self.device.index or 0))
        self.stream = stream

    def _sp(self):
        s = self.stream if self.stream is not None else torch.cuda.current_stream(self.device)
        # torch's default stream is CUDA's legacy default stream (handle 0); the C ABI reads NULL as "the library's
        # own stream", which would not be ordered with the collectives, so name the legacy stream explicitly
        return C.c_void_p(s.cuda_stream if s.cuda_stream else 1)          # 1 = cudaStreamLegacy

    def stream_context(self):
        """context under which torch collectives are ordered with this object's kernels (no-op when they already run on
        torch's current stream)"""
        import contextlib
        if self.stream is None or self.stream == torch.cuda.current_stream(self.device):
            return contextlib.nullcontext()
        return torch.cuda.stream(self.stream)

    def empty(self, nelem):
        return torch.empty(nelem, dtype=torch.complex128, device=self.device)

    def fft_strided(self, src, dst, outer, length, stride, direction=1):
        _capi.check(self.L.gd_fft_strided_c2c_dev(src.data_ptr(), dst.data_ptr(), outer, length, stride, direction, self._sp()))

    def fft_rows(self, src, dst, n, batch, direction=1):
        _capi.check(self.L.gd_fft_batch_c2c_dev(src.data_ptr(), dst.data_ptr(), n, batch, direction, self._sp()))

    def fourstep_twiddle(self, blk, rows, cols, row0, col0, log2n):
        _capi.check(self.L.gd_fourstep_twiddle_dev(blk.data_ptr(), rows, cols, row0, col0, log2n, self._sp()))

    def swap_leading(self, src, dst, a, b, w):
        """src[a][b][w] -> dst[b][a][w]"""
        _capi.check(self.L.gd_repack_gkw_dev(src.data_ptr(), dst.data_ptr(), a, b, w, self._sp()))

    def transpose_batched(self, src, dst, batch, rows, cols):
        """src[batch][rows][cols] -> dst[batch][cols][rows]"""
        _capi.check(self.L.gd_transpose_batched_dev(src.data_ptr(), dst.data_ptr(), batch, rows, cols, self._sp()))


class _DevArray:
    """A raw device pointer as a CUDA-array-interface object (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr, nelem):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": "<c16", "data": (ptr, False), "version": 3, "strides": None}


class PeerExchange:
    """Receive buffers of all ranks mapped into this process (CUDA IPC), for the fused exchange kernel
    (`gd_fourstep_exchange_dev`: twiddle + transpose + NVLink P2P stores in one pass, no NCCL data movement).
    One instance per (group, slab size); collective to construct and to close."""

    def __init__(self, nelem, ops, group=None):
        self.L, self.ops, self.group = ops.L, ops, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        own = C.c_void_p()
        handle = C.create_string_buffer(64)
        _capi.check(self.L.gd_ipc_alloc(C.byref(own), nelem * 16, handle))
        self.own = own.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.ptrs, self.opened = [], []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.own)
            else:
                p = C.c_void_p()
                _capi.check(self.L.gd_ipc_open(C.c_char_p(h), C.byref(p)))
                self.ptrs.append(p.value)
                self.opened.append(p.value)
        self.ptr_array = (C.c_void_p * self.world)(*self.ptrs)
        self.recv = torch.as_tensor(_DevArray(self.own, nelem), device=ops.device)
        self.token = torch.zeros(1, device=ops.device)

    def fence(self):
        """stream-ordered rendezvous of all ranks (a 1-element all-reduce): nobody passes before everybody's kernels
        enqueued so far are done. torch.distributed orders a collective against torch's CURRENT stream, so it is issued
        under the stream the kernels of `ops` run on."""
        with self.ops.stream_context():
            dist.all_reduce(self.token, group=self.group)

    def exchange(self, slab, n1, w, log2n):
        _capi.check(self.L.gd_fourstep_exchange_dev(slab.data_ptr(), self.ptr_array, n1, w, self.rank, self.world, log2n, self.ops._sp()))

    def lines_exchange(self, slab, tmp, n1, w, log2n):
        """length-n1 lines of the slab into `tmp`, pipelined with the exchange of the finished column blocks"""
        _capi.check(self.L.gd_fourstep_lines_exchange_dev(slab.data_ptr(), tmp.data_ptr(), self.ptr_array, n1, w, self.rank, self.world, log2n,
                                                          self.ops._sp()))

    def fused_supported(self, n1, n2):
        """both line lengths are in the range of the fused TMA kernel for this world size (gd_fourstep_fused_supported)"""
        return self.L.gd_fourstep_fused_supported(n1, n2, self.world) == 1

    def lines_peer(self, slab, n1, w, log2n, direction=1):
        """length-n1 lines of the slab, outer twiddle (log2n = 0: none), and the exchange as the kernel's own TMA stores into
        every rank's receive buffer over NVLink: receive buffers become [world][K][w] (gd_fourstep_lines_peer_dev)"""
        _capi.check(self.L.gd_fourstep_lines_peer_dev(slab.data_ptr(), self.ptr_array, n1, w, self.rank, self.world, log2n, direction, self.ops._sp()))

    def rows_seg(self, out, n2, k, direction=1):
        """the K rows of the receive buffer (n2 points each, in `world` segments) -> out[K][n2] (gd_fourstep_rows_seg_dev)"""
        _capi.check(self.L.gd_fourstep_rows_seg_dev(self.recv.data_ptr(), out.data_ptr(), n2, k, self.world, direction, self.ops._sp()))

    def block_copy(self, src, rows, cols, src_step, src_pitch, dst_off, dst_pitch):
        """every peer h: peer_buffer[h][dst_off + r*dst_pitch + c] = src[h*src_step + r*src_pitch + c]"""
        _capi.check(self.L.gd_peer_block_copy_dev(src.data_ptr(), self.ptr_array, self.world, self.rank, rows, cols, src_step, src_pitch,
                                                  dst_off, dst_pitch, self.ops._sp()))

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for p in self.opened:
            _capi.check(self.L.gd_ipc_close(p))
        self.opened = []
        dist.barrier(group=self.group)              # nobody has this rank's buffer mapped any more
        del self.recv
        _capi.check(self.L.gd_dev_free(self.own))


def _all_to_all(recv, send, group, ops=None):
    # equal contiguous splits; complex128 travels as pairs of float64; issued under the stream of `ops` (see PeerExchange.fence)
    import contextlib
    ctx = ops.stream_context() if ops is not None and hasattr(ops, "stream_context") else contextlib.nullcontext()
    with ctx:
        dist.all_to_all_single(torch.view_as_real(recv).view(-1), torch.view_as_real(send).view(-1), group=group)


def split_1d(n, world):
    """(N1, N2, K, W) of the sharded four-step: N = N1*N2, K = N1/world rows and W = N2/world columns per rank."""
    lg = _ilog2(n)
    l1 = (lg + 1) // 2
    n1, n2 = 1 << l1, 1 << (lg - l1)
    if n1 % world or n2 % world:
        raise ValueError("world size %d must divide both factors %d x %d" % (world, n1, n2))
    return n1, n2, n1 // world, n2 // world


def fft_1d_sharded(slab, n, ops, group=None, work=None, peer=None, fused=False):
    """Forward transform of one n-point signal.  `slab`: this rank's [N1][W] column slab (flattened; overwritten by the
    NCCL formulation, preserved by the peer-memory ones).
    Returns this rank's [N2][K] slab of the spectrum (a new tensor, or `work` if given: n/world elements).
    peer: a PeerExchange of n/world elements -> the exchange step is ONE kernel storing into the peers' buffers over
    NVLink (twiddle + transpose fused in); otherwise twiddle kernel + NCCL all-to-all + transpose kernel.
    fused (with peer): the exchange is the store phase of the first line pass itself -- one fused TMA kernel does lines,
    twiddle and NVLink stores -- and the result comes back TRANSPOSED: this rank's [K][N2] block, out[k1 local][k2] =
    X[k1 + N1 k2] (`gather_spectrum(..., fused=True)`)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n1, n2, k, w = split_1d(n, world)
    if slab.numel() != n1 * w:
        raise ValueError("slab has %d elements, expected %d" % (slab.numel(), n1 * w))
    if peer is not None and fused:
        out = work if work is not None else ops.empty(n1 * w)
        peer.fence()                                      # every rank is done reading its receive buffer (previous call)
        peer.lines_peer(slab, n1, w, _ilog2(n))           # lines over n1 + twiddle + NVLink stores: receive buffers [world][K][W]
        peer.fence()                                      # every rank's stores have landed
        peer.rows_seg(out, n2, k)                         # lines over n2 on segmented rows -> [K][N2]
        return out
    if peer is not None:
        out = work if work is not None else ops.empty(n1 * w)
        peer.fence()                                      # every rank is done reading its receive buffer (previous call)
        # lines over n1 into `out` (the slab is left untouched) and, block of columns by block of columns behind them, the
        # exchange: peer h gets [rank*W + c][k] <- out[h*K + k][c] * w_N^(k1 n2)
        peer.lines_exchange(slab, out, n1, w, _ilog2(n))
        peer.fence()                                      # every rank's stores have landed
        ops.fft_strided(peer.recv, out, 1, n2, k, 1)      # lines over n2
        return out
    recv = work if work is not None else ops.empty(n1 * w)
    # 1. lines over n1 (length N1, element stride W), in place
    ops.fft_strided(slab, slab, 1, n1, w, 1)
    # 2. slab[k1][c] *= w_N^(k1 * (rank*W + c))
    ops.fourstep_twiddle(slab, n1, w, 0, rank * w, _ilog2(n))
    # 3. rows [h*K, (h+1)*K) go to rank h; received: [source g][K][W]
    _all_to_all(recv, slab, group, ops)
    # 4. per source [K][W] -> [W][K]: the buffer becomes [N2][K] (n2 = g*W + c)
    ops.transpose_batched(recv, slab, world, k, w)
    # 5. lines over n2 (length N2, element stride K): out[k2][k1_local]
    ops.fft_strided(slab, recv, 1, n2, k, 1)
    return recv


def scatter_signal(x, n, rank, world):
    """Rank's [N1][W] slab of a natural-order signal (torch tensor, any device), flattened copy."""
    n1, n2, _, w = split_1d(n, world)
    return x.view(n1, n2)[:, rank * w:(rank + 1) * w].contiguous().view(-1)


def gather_spectrum(slabs, n, fused=False):
    """Natural-order spectrum from the per-rank [N2][K] slabs (list in rank order); fused: from the [K][N2] blocks of
    fft_1d_sharded(fused=True)."""
    world = len(slabs)
    n1, n2, k, _ = split_1d(n, world)
    if fused:
        return torch.cat([s.view(k, n2) for s in slabs], dim=0).t().contiguous().view(-1)
    return torch.cat([s.view(n2, k) for s in slabs], dim=1).contiguous().view(-1)


def fft2_sharded(block, rows, cols, ops, group=None, direction=1, peers=None, out=None, fused=True):
    """fft.FFT2 / IFFT2 of a rows x cols matrix; `block` is this rank's [rows/world][cols] row block
    (flattened complex128, overwritten).  Returns the rank's row block of the result.
    peers: (PeerExchange, PeerExchange) of rows*cols/world elements each -> both exchanges are block copies into the
    peers' buffers over NVLink (gd_peer_block_copy_dev), no repack kernels and no NCCL data movement; with `fused` (and
    shapes in the fused kernel's range) the second exchange disappears into the store phase of the column pass."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if rows % world or cols % world:
        raise ValueError("world size %d must divide %d x %d" % (world, rows, cols))
    rg, wc = rows // world, cols // world
    if block.numel() != rg * cols:
        raise ValueError("block has %d elements, expected %d" % (block.numel(), rg * cols))
    if peers is not None:
        colp, rowp = peers
        res = out if out is not None else ops.empty(rg * cols)
        # my columns [h*wc, (h+1)*wc) of my rows -> rank h's column slab [rows][wc], rows [rank*rg, (rank+1)*rg)
        colp.block_copy(block, rg, wc, wc, cols, rank * rg * wc, wc)
        colp.fence()
        if fused and rowp.fused_supported(rows, cols):
            # every column (fft/fft.go:138-144) through the fused TMA kernel whose stores ARE the second exchange: rows
            # [h*rg, (h+1)*rg) of my column slab land in rank h's buffer as block `rank` of [world][rg][wc]; then every row
            # (fft.go:146-151) on segmented rows -> this rank's [rg][cols] row block
            rowp.lines_peer(colp.recv, rows, wc, 0, direction)
            rowp.fence()
            rowp.rows_seg(res, cols, rg, direction)
            return res
        ops.fft_strided(colp.recv, colp.recv, 1, rows, wc, direction)    # every column (fft/fft.go:138-144)
        # rows [h*rg, (h+1)*rg) of my column slab -> rank h's row block [rg][cols], columns [rank*wc, (rank+1)*wc)
        rowp.block_copy(colp.recv, rg, wc, rg * wc, wc, rank * wc, cols)
        rowp.fence()
        ops.fft_rows(rowp.recv, res, cols, rg, direction)                # every row (fft/fft.go:146-151)
        return res
    tmp = ops.empty(rg * cols)
    # [rg][world][wc] -> [world][rg][wc]: destination-major send buffer
    ops.swap_leading(block, tmp, rg, world, wc)
    _all_to_all(block, tmp, group, ops)            # received [source][rg][wc] = [rows][wc]: my column slab
    ops.fft_strided(block, block, 1, rows, wc, direction)      # every column (fft/fft.go:138-144)
    _all_to_all(tmp, block, group, ops)            # rows [h*rg, (h+1)*rg) back to rank h: [source][rg][wc]
    ops.swap_leading(tmp, block, world, rg, wc)    # -> [rg][cols]
    ops.fft_rows(block, tmp, cols, rg, direction)  # every row (fft/fft.go:146-151)
    return tmp

#!/usr/bin/env python3
"""bench.py -- throughput of the go-dsp hot path on B200: batched 2^20-point complex128 FFT
(GS/s, the headline) and spectral.Pwelch (Msamples/s, reported in the same JSON line).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (C ABI)
    python bench.py --impl reference --gpus N ...            # go-dsp's CPU algorithm (oracle port) on host cores

One process per GPU (torchrun for N > 1, one rank per GPU); batches / segment ranges are
sharded across ranks with no data-path collective (Pwelch adds one 2049-double all-gather).
A "step" is one pass of the hot path over one batch of synthetic input (SplitMix64 counter
generator). Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks. Inputs are far larger than the 126 MB L2.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "go-dsp_b200"))

LOG2N = 20
FFT_SEED, PW_SEED = 3, 5
PW_NFFT, PW_NOVERLAP = 4096, 2048


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="2^20-point transforms per GPU per step")
    ap.add_argument("--pw-log2-samples", type=int, default=30, help="Pwelch samples per GPU per step (log2)")
    ap.add_argument("--e2e-batch", type=int, default=256, help="transforms per GPU in the host-buffer (e2e) run")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--skip", default="", help="comma list: pwelch,e2e,cpu,extra")
    ap.add_argument("--scratch-mb", type=int, default=0, help="override the inter-pass scratch budget")
    ap.add_argument("--wide-tiles", type=int, default=-1)
    ap.add_argument("--fused", type=int, default=-1)
    ap.add_argument("--fused-slot-mb", type=int, default=0)
    return ap.parse_args()


# --------------------------------------------------------------------------- helpers
def ncu_traffic(kernel_substr):
    """dram read+write bytes per launch of a kernel, from the committed ncu --set full summary (profiles/)."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_full_summary.json")
    try:
        with open(p) as f:
            rows = [r for r in json.load(f) if kernel_substr in r.get("Kernel Name", "")]
        vals = []
        for r in rows:
            tot = 0.0
            for k, v in r.items():
                if k.startswith("dram__bytes_read.sum") or k.startswith("dram__bytes_write.sum"):
                    unit = k[k.index("[") + 1:k.index("]")]
                    tot += float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
            vals.append(tot)
        return (sum(vals) / len(vals)) if vals else None
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def fp64_peak():
    """measured DFMA issue rate (thread-instructions per second), tools/ubench/fp64_peak.cu -> profiles/r2_fp64_peak.json"""
    p = os.path.join(ROOT, "profiles", "r2_fp64_peak.json")
    try:
        with open(p) as f:
            return float(json.load(f)["dfma_peak_tinst_per_s"]), "measured (profiles/r2_fp64_peak.json: DFMA, 32 warps per SM, 8 independent chains)"
    except Exception:
        return 148 * 64 * 1.965e9, "nominal (148 SMs x 64 lanes x 1.965 GHz)"


# FP64 instructions per unit of the hot kernels, counted in their SASS (profiles/r2_sass_summary.txt): two radix-32 steps
# (generated codelet, 376 each) + intra-line twiddle chain per pass, + the four-step twiddle in pass 1
FP64_INST_PER_POINT_FFT = (2 * (2 * 376 + 240) + 265) / 32.0          # 70.3
FP64_INST_PER_SAMPLE_PWELCH = 734 / 16.0                              # 45.9 (one 4096-point complex transform per 4096 new samples)


def pcie_ceiling(torch, dist, world, barrier, nbytes, hin=None, hout=None):
    """What the box's PCIe / host memory gives when every rank copies nbytes up and nbytes down AT THE SAME TIME (pinned
    memory, two streams, no kernels): the ceiling of the end-to-end numbers at this GPU count. hin / hout: addresses of
    pinned buffers to copy from / to (the e2e run's own NUMA-local gd_pinned_alloc buffers); torch-pinned memory otherwise."""
    try:
        if hin and hout:
            hp = torch.frombuffer((C.c_ubyte * nbytes).from_address(hin), dtype=torch.uint8)
            hq = torch.frombuffer((C.c_ubyte * nbytes).from_address(hout), dtype=torch.uint8)
        else:
            hp, hq = torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        dp, dq = torch.empty(nbytes, dtype=torch.uint8, device="cuda"), torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def both():
            with torch.cuda.stream(s1):
                dp.copy_(hp, non_blocking=True)
            with torch.cuda.stream(s2):
                hq.copy_(dq, non_blocking=True)
            s1.synchronize(); s2.synchronize()
        both()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            both()
        dt = (time.perf_counter() - t0) / 3
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return {"each_way_gbs_per_gpu": nbytes / dt / 1e9, "each_way_gbs_aggregate": world * nbytes / dt / 1e9,
                "e2e_gs_per_s_at_this_ceiling": world * nbytes / dt / 1e9 / 16.0,
                "how": "every rank copies %d MiB host->device and device->host concurrently from %s, max over ranks"
                       % (nbytes >> 20, "the e2e run's own pinned buffers (gd_pinned_alloc, NUMA-local)" if hin and hout else "torch-pinned memory")}
    except Exception as ex:
        return {"unavailable": str(ex)[:120]}


def cufft_compare(torch, n, batch=256, reps=5):
    """cuFFT Z2Z through torch.fft.fft on the same kind of batch: a comparison only, never on the product path"""
    try:
        x = torch.randn(batch, n, dtype=torch.complex128, device="cuda")
        y = torch.empty_like(x)
        torch.fft.fft(x, out=y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.fft.fft(x, out=y)
        e1.record()
        torch.cuda.synchronize()
        del y
        return {"gs_per_s": batch * n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9, "batch": batch, "api": "torch.fft.fft (cuFFT Z2Z), out of place"}
    except Exception as ex:
        return {"unavailable": str(ex)[:100]}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.ok = index, [], set(), False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if self.ok:
            self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}



# --------------------------------------------------------------------------- parity (outside every timed region)
PARITY_TOL = 1e-12            # BASELINE.json north_star: relative L2 <= 1e-12 for spectra and PSDs


def rel_l2(a, b):
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    d = float(np.linalg.norm(b))
    return float(np.linalg.norm(a - b)) / d if d else float(np.linalg.norm(a - b))


def allmax(torch, dist, world, v):
    if world > 1:
        t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(v)


def dft_phasors(torch, length, k, sign=-1.0):
    """exp(sign * 2 pi i * (i*k mod length) / length), i < length, exponent reduced exactly in int64"""
    i = torch.arange(length, dtype=torch.int64, device="cuda")
    ang = (sign * 2.0 * np.pi / length) * ((i * int(k)) % length).double()
    return torch.complex(torch.cos(ang), torch.sin(ang))


def parity_fft_rows(torch, y, n, row0, rows):
    """sampled rows of the timed batch against oracle.fft of the same SplitMix64 rows (fft/fft.go:72-87)"""
    import oracle
    worst, per = 0.0, {}
    for r in rows:
        got = y[2 * n * r: 2 * n * (r + 1)].cpu().numpy().view(np.complex128)
        x = oracle.splitmix_complex(n, FFT_SEED, (row0 << 21) + (r << 20))
        e = rel_l2(got, oracle.fft(x))
        per[str(r)] = e
        worst = max(worst, e)
    return worst, per


def parity_fft2(torch, dist, world, rank, src_blk, out_blk, R, Cc, rows, cols):
    """fft.FFT2 output (row blocks `out_blk` of the R x Cc result) on sampled rows and columns. One output line is the
    oracle's 1-D transform (CPU) of one bin of the other axis, and that bin is a plain float64 matrix-vector product
    with exactly reduced phasors (torch, not this library): out[r, :] = FFT(f_r^T . src), out[:, c] = FFT(src . f_c)."""
    import oracle
    rg = R // world
    s2, o2 = src_blk.view(rg, Cc), out_blk.view(rg, Cc)
    worst = 0.0
    for r in rows:
        f = dft_phasors(torch, R, r)[rank * rg:(rank + 1) * rg]
        u = torch.mv(s2.t(), f)                                   # partial sum over this rank's rows
        if world > 1:
            dist.all_reduce(u)
        if r // rg == rank:
            e = rel_l2(o2[r - rank * rg].cpu().numpy(), oracle.fft(u.cpu().numpy()))
            worst = max(worst, e)
    for c in cols:
        v = torch.mv(s2, dft_phasors(torch, Cc, c))               # this rank's rows of the bin-c vector
        got = o2[:, c].contiguous()
        if world > 1:
            vs = [torch.empty_like(v) for _ in range(world)]
            gs = [torch.empty_like(got) for _ in range(world)]
            dist.all_gather(vs, v)
            dist.all_gather(gs, got)
            v, got = torch.cat(vs), torch.cat(gs)
        if rank == 0:
            worst = max(worst, rel_l2(got.cpu().numpy(), oracle.fft(v.cpu().numpy())))
    return allmax(torch, dist, world, worst)


def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


# --------------------------------------------------------------------------- reference arm (CPU)
def cpu_fft_sample(threads, budget_s):
    """oracle port of fft.FFT (fft/radix2.go) on `threads` host threads; returns (GS/s, description)."""
    import oracle
    n = 1 << LOG2N
    x1 = oracle.splitmix_complex(n, FFT_SEED)
    oracle.fft(x1)                                   # builds the twiddle tables (EnsureRadix2Factors)
    t0 = time.perf_counter()
    oracle.fft(x1)
    t1 = time.perf_counter() - t0
    batch = int(max(threads, min(16 * threads, budget_s * threads / max(t1, 1e-3))))
    xs = np.empty((batch, n), np.complex128)
    for b in range(batch):
        xs[b] = oracle.splitmix_complex(n, FFT_SEED, b << 21) if b < 4 else xs[b % 4]
    t0 = time.perf_counter()
    oracle.fft_batch(xs, threads=threads)
    dt = time.perf_counter() - t0
    return batch * n / dt / 1e9, "%d transforms of 2^20 points, one transform per thread at a time" % batch, dt


def cpu_pwelch_sample(threads, budget_s):
    import oracle
    log2s = 24
    x = oracle.fill_splitmix(1 << log2s, PW_SEED)
    t0 = time.perf_counter()
    oracle.pwelch(x, 1.0, nfft=PW_NFFT, noverlap=PW_NOVERLAP, threads=threads)
    dt = time.perf_counter() - t0
    reps = int(max(1, min(8, budget_s / max(dt, 1e-3))))
    if reps > 1:
        x = np.tile(x, reps)
        t0 = time.perf_counter()
        oracle.pwelch(x, 1.0, nfft=PW_NFFT, noverlap=PW_NOVERLAP, threads=threads)
        dt = time.perf_counter() - t0
    return x.shape[0] / dt / 1e6, "%d samples, NFFT 4096, Noverlap 2048, Hann; segment ranges split over threads" % x.shape[0], dt


def run_reference(args):
    world, rank, _ = dist_setup(args)
    if rank != 0:
        return
    import oracle
    threads = os.cpu_count() or 1
    per_step = max(2.0, min(10.0, 150.0 / max(1, args.steps + args.warmup) / 2))
    vals, pvals, times, desc, pdesc = [], [], [], "", ""
    for i in range(args.warmup + args.steps):
        v, desc, dt = cpu_fft_sample(threads, per_step)
        pv, pdesc, _ = cpu_pwelch_sample(threads, per_step)
        if i >= args.warmup:
            vals.append(v)
            pvals.append(pv)
            times.append(dt)
    v = float(np.mean(vals))
    pv = float(np.mean(pvals))
    n = 1 << LOG2N
    line = {
        "impl": "reference", "metric": "FFT GS/s (complex128, 2^20-pt batched)", "value": v, "unit": "GS/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": {"workload": "batched 2^20-point complex128 FFT (BASELINE.json configs[2]); CPU arm times a bounded sample per step",
                   "n": n, "sample": desc},
        "cpu_baseline": {"value": v, "unit": "GS/s", "cores": threads, "kind": "port",
                         "sample": desc + "; C restatement of go-dsp (oracle/godsp_oracle.c), not the Go binary (no Go toolchain)",
                         "thread_model": "one whole transform per OpenMP thread (every transform is the sequential radix-2 loop of fft/radix2.go); "
                                         "go-dsp itself runs ONE transform at a time with a goroutine pool and a WaitGroup barrier per stage "
                                         "(radix2.go:126-151), whose last log2(P) stages under-fill the pool: this arm is kinder to the CPU"},
        "e2e": {"value": v, "unit": "GS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pwelch": {"metric": "Pwelch Msamples/s", "value": pv, "unit": "Msamples/s",
                   "cpu_baseline": {"value": pv, "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": pdesc},
                   "e2e": {"value": pv, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm (GPU)
def run_ours(args):
    import torch
    import torch.distributed as dist
    from godsp import _capi as capi

    world, rank, local = dist_setup(args)
    skip = set(s for s in args.skip.split(",") if s)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; go-dsp_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = capi.lib()
    capi.check(L.gd_use_device(local))
    if args.scratch_mb > 0:
        capi.check(L.gd_set_option(b"pass_scratch_mb", args.scratch_mb))
    if args.wide_tiles >= 0:
        capi.check(L.gd_set_option(b"wide_tiles", args.wide_tiles))
    if args.fused >= 0:
        capi.check(L.gd_set_option(b"fused", args.fused))
    if args.fused_slot_mb > 0:
        capi.check(L.gd_set_option(b"fused_slot_mb", args.fused_slot_mb))
    # a dedicated (non-default) stream: the events below and every kernel of the library share it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0
    hbm_peak, peak_src = peaks()
    fp64_pk, fp64_src = fp64_peak()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """returns (ms per step, max over ranks), launches in the timed region, clocks"""
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        l0 = L.gd_kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = L.gd_kernel_launches() - l0
        clocks = sampler.result()
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, launches, clocks

    def timed_host(fn, steps, warmup):
        """wall clock around synchronous host-buffer calls (copies inside), max over ranks"""
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        barrier()
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    n = 1 << LOG2N
    # ---------------- FFT: device-resident
    free_b, _ = torch.cuda.mem_get_info()
    batch = args.batch
    max_batch = int((free_b - (6 << 30)) // (2 * n * 16))
    if batch > max_batch:
        batch = max(1, max_batch)
    x = torch.empty(batch * n * 2, dtype=torch.float64, device="cuda")
    y = torch.empty(batch * n * 2, dtype=torch.float64, device="cuda")
    row0 = rank * batch                 # global batch row of this rank's first transform
    capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), batch * n * 2, FFT_SEED, (row0 << 21) * 2, sp))
    torch.cuda.synchronize()

    def fft_step():
        capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, batch, 1, sp))

    ms, launches, clocks = timed(fft_step, args.steps, args.warmup)
    parity = {}
    # sampled rows of the timed batch vs the oracle: first / last row of the first / last 128-transform launch
    prow = sorted(set(r for r in (0, min(127, batch - 1), max(0, batch - 128), batch - 1)))
    w_fft, per_row = parity_fft_rows(torch, y, n, row0, prow)
    parity["fft_batch"] = {"max_rel_l2": allmax(torch, dist, world, w_fft), "rows_per_rank": prow,
                           "vs": "oracle.fft (C restatement of fft/radix2.go) on the same SplitMix64 rows, every rank"}
    pts = batch * n * world
    value = pts / (ms * 1e-3) / 1e9
    per_gpu_bytes = 32.0 * batch * n
    achieved = per_gpu_bytes / (ms * 1e-3) / 1e9
    launches_per_step = launches / max(1, args.steps)
    line = {
        "metric": "FFT GS/s (complex128, 2^20-pt batched)", "value": value, "unit": "GS/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": {"workload": "batched 2^20-point complex128 FFT (BASELINE.json configs[2]), %d transforms per GPU per step" % batch,
                   "n": n, "batch_per_gpu": batch, "global_batch": batch * world, "direction": "forward",
                   "residency": "inputs and outputs resident in HBM (generated on device, SplitMix64 seed 3)",
                   "l2": "inputs (%.1f GiB per GPU) are far larger than L2; no flush needed" % (batch * n * 16 / 2**30),
                   "parallelism": "batch rows sharded over %d GPU(s), no collective" % world},
        "roofline": {"bound": "hbm", "kernel": "gd::fft_tma_fused_kernel<false,false> (both four-step passes of up to 512 transforms per launch; the intermediate stays in L2)",
                     "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "hbm_frac": achieved / hbm_peak,
                     "fp64_frac": value / world * 1e9 * FP64_INST_PER_POINT_FFT / fp64_pk, "fp64_peak_tinst_per_s": fp64_pk, "fp64_peak_source": fp64_src,
                     "fp64_inst_per_point": FP64_INST_PER_POINT_FFT,
                     "peak_source": peak_src,
                     "traffic": (ncu_traffic("fft_tma_fused_kernel") or 0) * (batch / max(1.0, launches_per_step)) / 128.0 or None,
                     "traffic_profiled_launch": ncu_traffic("fft_tma_fused_kernel"),
                     "traffic_note": "profiles/r2_ncu_full_summary.json (ncu --set full) holds one launch of 128 transforms: 4.27 GB of dram read+write against 4.29 GB algorithmic; `traffic` scales that to the transforms per timed launch (the inter-pass array never reaches HBM, so traffic is proportional to the transform count)",
                     "algorithmic_bytes_per_launch": per_gpu_bytes / max(1.0, launches_per_step),
                     "avg_launch_us": ms * 1e3 / max(1.0, launches_per_step),
                     "note": "32 B/point (16 read + 16 written once, SURVEY.md 8d) x points per launch / CUDA-event time per launch; FP64 issue is the co-limiting roof (fp64_frac), and the sustained run sits at the 1000 W power cap (clocks), see DESIGN.md"},
        "gpu_launches": int(launches), "clocks": clocks,
    }

    # ---------------- FFT: end to end through the host-buffer C ABI
    if "e2e" not in skip:
        eb = min(args.e2e_batch, batch)
        nbytes = eb * n * 16
        L.gd_pinned_alloc.restype = C.c_void_p
        hin, hout = L.gd_pinned_alloc(nbytes), L.gd_pinned_alloc(nbytes)
        if not hin or not hout:
            raise SystemExit("pinned allocation failed: " + L.gd_last_error().decode())
        hin_np = np.ctypeslib.as_array((C.c_double * (eb * n * 2)).from_address(hin))
        capi.check(L.gd_memcpy_d2h(hin, x.data_ptr(), nbytes))        # same synthetic rows, now in pinned host memory

        def e2e_step():
            capi.check(L.gd_fft_batch_c2c(hin, hout, n, eb, 1))

        ems = timed_host(e2e_step, args.e2e_steps, 1)
        hout_keep = np.ctypeslib.as_array((C.c_double * (2 * n)).from_address(hout)).copy()      # the ceiling run overwrites hout
        line["e2e_pcie_ceiling"] = pcie_ceiling(torch, dist, world, barrier, min(nbytes, 1 << 30), hin, hout)
        line["e2e"] = {"value": eb * n * world / (ems * 1e-3) / 1e9, "unit": "GS/s", "h2d_bytes_per_step": nbytes,
                       "d2h_bytes_per_step": nbytes, "ms_per_step": ems, "batch_per_gpu": eb,
                       "api": "gd_fft_batch_c2c (pinned host in/out; chunked H2D / kernels / D2H overlap on 3 streams)"}
        # spot check: e2e output equals the device-resident output for the same rows
        ref = y[: 2 * n].cpu().numpy()
        line["e2e"]["matches_device_path"] = bool(np.array_equal(hout_keep, ref))
        del hin_np
        L.gd_pinned_free(hin)
        L.gd_pinned_free(hout)
        # the same call on PAGEABLE host memory -- what fft.FFT(x) on a plain Go slice hands the library: chunks go through
        # the library's pinned staging ring, filled / drained by host threads
        pb = min(eb, 64)
        pin_np = np.empty(pb * n * 2)
        capi.check(L.gd_memcpy_d2h(pin_np.ctypes.data, x.data_ptr(), pb * n * 16))
        pout_np = np.empty(pb * n * 2)

        def e2e_pageable_step():
            capi.check(L.gd_fft_batch_c2c(pin_np.ctypes.data, pout_np.ctypes.data, n, pb, 1))

        pms = timed_host(e2e_pageable_step, args.e2e_steps, 1)
        line["e2e_pageable"] = {"value": pb * n * world / (pms * 1e-3) / 1e9, "unit": "GS/s", "ms_per_step": pms, "batch_per_gpu": pb,
                                "h2d_bytes_per_step": pb * n * 16, "d2h_bytes_per_step": pb * n * 16,
                                "matches_device_path": bool(np.array_equal(pout_np[: 2 * n], ref)),
                                "api": "gd_fft_batch_c2c on pageable numpy buffers (pinned staging ring + host copy threads inside the library)"}
        del pin_np, pout_np

    # ---------------- FFT: CPU baseline (rank 0, N = 1 only)
    if "cpu" not in skip and world == 1:
        threads = os.cpu_count() or 1
        v, desc, _ = cpu_fft_sample(threads, 1.5)
        line["cpu_baseline"] = {"value": v, "unit": "GS/s", "cores": threads, "kind": "port",
                                "sample": desc + "; C restatement of go-dsp (oracle/godsp_oracle.c), not the Go binary"}
    del x, y
    torch.cuda.empty_cache()
    if "cpu" not in skip and world == 1:
        line["cufft_comparison"] = cufft_compare(torch, n)
        torch.cuda.empty_cache()

    # ---------------- Pwelch
    if "pwelch" not in skip:
        line["pwelch"] = run_pwelch(args, torch, dist, capi, L, sp, world, rank, local, timed, timed_host, skip, hbm_peak, peak_src)
        line["gpu_launches"] += line["pwelch"].pop("_launches")

    # ---------------- the other configs of BASELINE.json that shard with an exchange step
    if "extra" not in skip:
        line["fft2"] = run_fft2(args, torch, dist, L, sp, world, rank, timed, hbm_peak)
        line["gpu_launches"] += line["fft2"].pop("_launches")
        if world > 1:
            line["fft_1d_sharded"] = run_fft1d_sharded(args, torch, dist, L, sp, world, rank, timed)
            line["gpu_launches"] += line["fft_1d_sharded"].pop("_launches")
            px = line["fft_1d_sharded"].pop("_peer", None)
            if px is not None:
                px.close()

        line["fft_sizes"] = run_fft_sizes(args, torch, dist, L, sp, world, rank, timed, hbm_peak)
        line["gpu_launches"] += line["fft_sizes"].pop("_launches")

    for key in ("pwelch", "fft2", "fft_1d_sharded", "fft_sizes"):
        if key in line and "_parity" in line[key]:
            parity[key] = line[key].pop("_parity")
    parity["max_rel_l2"] = max([v["max_rel_l2"] for v in parity.values()] or [0.0])
    parity["tolerance"] = PARITY_TOL
    parity["ok"] = bool(parity["max_rel_l2"] <= PARITY_TOL)
    line["parity"] = parity
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not parity["ok"]:
        raise SystemExit("bench.py: parity check failed: %r" % (parity,))


def run_fft_sizes(args, torch, dist, L, sp, world, rank, timed, hbm_peak):
    """Batched transforms of the other sizes: the fused size family (2^13 .. 2^19, fft_tma14.cuh), the outer four-step over it
    (2^21 .. 2^24: two sweeps) and Bluestein lengths (one kernel per transform up to a padded length of 8192, streaming kernels
    around plain transforms above). 2^28 points per GPU and call, device resident (inputs larger than L2); one transform of
    every size against the oracle."""
    from godsp import _capi as capi
    import oracle
    total = 1 << 28
    x = torch.empty(total * 2, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), total * 2, 11, (rank * total) * 2, sp))
    torch.cuda.synchronize()
    rows, worst, launches = {}, 0.0, 0
    sizes = [("2^%d" % lg, 1 << lg) for lg in range(13, 25) if lg != 20] + [("%d" % n, n) for n in (1000, 4095, 30000, 1000003)]
    for name, n in sizes:
        b = total // n
        ms, nl, _ = timed(lambda: capi.check(L.gd_fft_batch_c2c_dev(x.data_ptr(), y.data_ptr(), n, b, 1, sp)), 3, 3)
        launches += nl
        r = b - 1                                            # the last transform of the batch against the oracle
        xin = x[: 2 * n * b].view(b, 2 * n)[r].cpu().numpy().view(np.complex128)
        got = y[: 2 * n * b].view(b, 2 * n)[r].cpu().numpy().view(np.complex128)
        err = rel_l2(got, oracle.fft(np.ascontiguousarray(xin)))
        worst = max(worst, err)
        rows[name] = {"gs_per_gpu": n * b / (ms * 1e-3) / 1e9, "ms": ms, "hbm_frac": 32.0 * n * b / (ms * 1e-3) / 1e9 / hbm_peak}
    del x, y
    torch.cuda.empty_cache()
    return {"metric": "FFT GS/s per GPU (complex128, batched, 2^28 points per call) by transform size", "sizes": rows,
            "kernel": "2^13 .. 2^19: gd::fft_tma14_kernel<LA, LB, ROWS> (both four-step passes in one persistent TMA-fed launch, N = LA x LB); "
                      "2^21 .. 2^24: the same kernel on columns with the outer twiddle on its stores + one row pass with the transposed store; "
                      "other lengths: Bluestein (hbm_frac at 32 B per point of the unpadded length)",
            "_parity": {"max_rel_l2": allmax(torch, dist, world, worst), "vs": "oracle.fft on the last transform of every batch, every rank"},
            "_launches": int(launches)}


def run_fft2(args, torch, dist, L, sp, world, rank, timed, hbm_peak):
    """fft.FFT2 on a 16384 x 16384 matrix (BASELINE.json configs[2]); N > 1: row blocks sharded, strong scaling."""
    from godsp import _capi as capi
    from godsp import distributed as D
    R = Cc = 16384
    rg = R // world
    src = torch.empty(rg * Cc, dtype=torch.complex128, device="cuda")
    capi.check(L.gd_fill_splitmix_dev(src.data_ptr(), 2 * rg * Cc, 4, 2 * rank * rg * Cc, sp))
    steps, warmup = max(2, min(args.steps, 5)), max(1, min(args.warmup, 3))
    extra, peers = {}, None
    if world == 1:
        out = torch.empty_like(src)
        dims = (C.c_int64 * 2)(R, Cc)

        def step():
            capi.check(L.gd_fftn_c2c_dev(src.data_ptr(), out.data_ptr(), dims, 2, 1, sp))
        api = "gd_fftn_c2c_dev (columns then rows, fft/fft.go:138-151)"
    else:
        ops = D.DeviceOps()
        blk = torch.empty(rg * Cc, dtype=torch.complex128, device="cuda")
        res = torch.empty(rg * Cc, dtype=torch.complex128, device="cuda")

        def step_nccl():
            blk.copy_(src)
            D.fft2_sharded(blk, R, Cc, ops)
        extra["nccl_all_to_all_variant_ms"], _, _ = timed(step_nccl, steps, warmup)
        peers = (D.PeerExchange(rg * Cc, ops), D.PeerExchange(rg * Cc, ops))

        def step():
            D.fft2_sharded(src, R, Cc, ops, peers=peers, out=res)
        api = "godsp.distributed.fft2_sharded(peers=...): block copy into the peers' column slabs over NVLink, column lines, block copy back into the peers' row blocks, row lines (gd_peer_block_copy_dev; no NCCL data movement)"
    ms, launches, clocks = timed(step, steps, warmup)
    result = out if world == 1 else res
    w2 = parity_fft2(torch, dist, world, rank, src, result, R, Cc, rows=(0, 1, 5461, R - 1), cols=(0, 3, 8192, Cc - 1))
    par = {"max_rel_l2": w2, "vs": "sampled output rows and columns of the timed 16384 x 16384 result vs oracle.fft of the matching "
                                   "single-bin DFT of the other axis (float64 matrix-vector product with exactly reduced phasors)"}
    if peers is not None:
        peers[0].close()
        peers[1].close()
    return {"_parity": par, "metric": "FFT2 Gelem/s (complex128, 16384 x 16384)", "value": R * Cc / (ms * 1e-3) / 1e9, "unit": "Gelem/s",
            "ms_per_step": ms, "scaling": "strong", "api": api, **extra,
            "roofline": {"bound": "hbm", "achieved": 64.0 * R * Cc / world / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": 64.0 * R * Cc / world / (ms * 1e-3) / 1e9 / hbm_peak,
                         "note": "64 B per element algorithmic: two sweeps, the 4 GiB matrix is far larger than L2 (SURVEY.md 8d)"},
            "clocks": clocks, "_launches": int(launches)}


def sharded_1d_kat(torch, dist, D, ops, px, n, world, rank, src, work, fused=False):
    """Known-answer test at the full size: x[n] = delta[n - p] + sum_t a exp(2 pi i f_t n / N) (integer bins f_t), so
    X[k] = exp(-2 pi i p k / N) + N a [k == f_t]. Built and checked on the device in exact integer phase arithmetic."""
    n1, n2, k, w = D.split_1d(n, world)
    mask, p, amp = n - 1, 1234567891 % n, 2.0 ** -16
    tones = [f % n for f in (3, 987654321, n // 2 + 12345)]
    s2 = src.view(n1, w)
    rows = max(1, (1 << 24) // w)
    for a in range(0, n1, rows):
        r = torch.arange(a, min(n1, a + rows), dtype=torch.int64, device="cuda")[:, None]
        idx = r * n2 + (rank * w + torch.arange(w, dtype=torch.int64, device="cuda"))[None, :]
        acc = (idx == p).to(torch.complex128)
        for f in tones:
            ang = (2.0 * np.pi / n) * ((idx * f) & mask).double()
            acc += amp * torch.complex(torch.cos(ang), torch.sin(ang))
        s2[a:a + rows] = acc
    D.fft_1d_sharded(src, n, ops, work=work, peer=px, fused=fused)
    torch.cuda.synchronize()
    o2 = work.view(k, n2) if fused else work.view(n2, k)       # fused: [k1 local][k2], else [k2][k1 local]
    num = torch.zeros((), dtype=torch.float64, device="cuda")
    den = torch.zeros((), dtype=torch.float64, device="cuda")
    rows = max(1, (1 << 24) // (n2 if fused else k))
    for a in range(0, k if fused else n2, rows):
        if fused:
            k1 = (rank * k + torch.arange(a, min(k, a + rows), dtype=torch.int64, device="cuda"))[:, None]
            kk = k1 + n1 * torch.arange(n2, dtype=torch.int64, device="cuda")[None, :]
        else:
            k2 = torch.arange(a, min(n2, a + rows), dtype=torch.int64, device="cuda")[:, None]
            kk = (rank * k + torch.arange(k, dtype=torch.int64, device="cuda"))[None, :] + n1 * k2
        ang = (-2.0 * np.pi / n) * ((kk * p) & mask).double()
        want = torch.complex(torch.cos(ang), torch.sin(ang))
        for f in tones:
            want += (kk == f).to(torch.complex128) * (amp * n)
        d = o2[a:a + rows] - want
        num += (d.real ** 2 + d.imag ** 2).sum()
        den += (want.real ** 2 + want.imag ** 2).sum()
    t = torch.stack([num, den])
    dist.all_reduce(t)
    return float(torch.sqrt(t[0] / t[1]).item())


def sharded_1d_vs_oracle(torch, dist, D, ops, world, rank, lg=24, fused=False):
    """the same peer-memory code path at 2^lg points against oracle.fft (rank 0 compares the gathered spectrum)"""
    import oracle
    n = 1 << lg
    n1, n2, k, w = D.split_1d(n, world)
    x = oracle.splitmix_complex(n, 6)
    slab = D.scatter_signal(torch.from_numpy(x), n, rank, world).cuda()
    px = D.PeerExchange(n1 * w, ops)
    fused = fused and px.fused_supported(n1, n2)
    out = D.fft_1d_sharded(slab, n, ops, peer=px, fused=fused)
    torch.cuda.synchronize()
    slabs = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(slabs, out)
    err = 0.0
    if rank == 0:
        err = rel_l2(D.gather_spectrum(slabs, n, fused=fused).cpu().numpy(), oracle.fft(x))
    px.close()
    return allmax(torch, dist, world, err)


def run_fft1d_sharded(args, torch, dist, L, sp, world, rank, timed):
    """ONE transform of 2^29 points per GPU (2^32 at 8 GPUs = BASELINE.json configs[4]): four-step, one all-to-all."""
    from godsp import _capi as capi
    from godsp import distributed as D
    lg = 29 + (world.bit_length() - 1)
    if (1 << (world.bit_length() - 1)) != world:
        return {"skipped": "world size %d is not a power of two" % world, "_launches": 0}
    n = 1 << lg
    n1, n2, k, w = D.split_1d(n, world)
    ops = D.DeviceOps()
    slab = torch.empty(n1 * w, dtype=torch.complex128, device="cuda")
    src = torch.empty(n1 * w, dtype=torch.complex128, device="cuda")
    work = torch.empty(n1 * w, dtype=torch.complex128, device="cuda")
    # a synthetic slab (the layout is [N1][W]; values are SplitMix64 seed 6 at this rank's offset)
    capi.check(L.gd_fill_splitmix_dev(src.data_ptr(), 2 * n1 * w, 6, 2 * rank * n1 * w, sp))
    steps, warmup = max(2, min(args.steps, 3)), max(1, min(args.warmup, 2))

    def step_nccl():
        slab.copy_(src)
        D.fft_1d_sharded(slab, n, ops, work=work)
    ms_nccl, _, _ = timed(step_nccl, steps, warmup)
    px = D.PeerExchange(n1 * w, ops)

    fused = px.fused_supported(n1, n2)

    def step_separate():
        D.fft_1d_sharded(src, n, ops, work=work, peer=px)       # the peer-memory paths leave their input untouched

    def step():
        D.fft_1d_sharded(src, n, ops, work=work, peer=px, fused=fused)
    ms_sep = None
    if fused:
        ms_sep, _, _ = timed(step_separate, steps, warmup)
    ms, launches, clocks = timed(step, steps, warmup)
    # per phase, outside the timed region (CUDA events on the ops stream, every phase between two rendezvous, max over ranks)
    def phase_ms(fn):
        px.fence()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        return allmax(torch, dist, world, e0.elapsed_time(e1))
    if fused:
        phases = {"lines_n1_twiddle_nvlink_stores_ms": phase_ms(lambda: px.lines_peer(src, n1, w, lg)),
                  "lines_n2_segmented_rows_ms": phase_ms(lambda: px.rows_seg(work, n2, k))}
        t_x, t_l = phases["lines_n1_twiddle_nvlink_stores_ms"], phases["lines_n2_segmented_rows_ms"]
    else:
        phases = {"lines_n1_ms": phase_ms(lambda: ops.fft_strided(src, work, 1, n1, w, 1)),
                  "twiddle_transpose_nvlink_stores_ms": phase_ms(lambda: px.exchange(work, n1, w, lg)),
                  "lines_n2_ms": phase_ms(lambda: ops.fft_strided(px.recv, work, 1, n2, k, 1))}
        t_x, t_l = phases["twiddle_transpose_nvlink_stores_ms"], phases["lines_n1_ms"]
    nv_bytes = 16 * (n // world) * (world - 1) // world
    phases["nvlink_out_gbs_per_gpu"] = nv_bytes / (t_x * 1e-3) / 1e9 if world > 1 else 0.0
    phases["lines_hbm_frac"] = 32.0 * (n // world) / (t_l * 1e-3) / 1e9 / peaks()[0]
    step()                                                      # `work` holds the whole transform again
    torch.cuda.synchronize()
    # Parseval on the last step: sum |X|^2 = n * sum |x|^2 over all ranks
    e = torch.stack([(work.real ** 2 + work.imag ** 2).sum(), (src.real ** 2 + src.imag ** 2).sum()])
    dist.all_reduce(e)
    parseval = abs(float(e[0].item()) / (n * float(e[1].item())) - 1.0)
    kat = sharded_1d_kat(torch, dist, D, ops, px, n, world, rank, src, work, fused=fused)
    small = sharded_1d_vs_oracle(torch, dist, D, ops, world, rank, 24)
    small_fused = sharded_1d_vs_oracle(torch, dist, D, ops, world, rank, 26, fused=True) if fused else 0.0
    par = {"max_rel_l2": max(kat, small, small_fused), "kat_full_size_rel_l2": kat, "oracle_2p24_rel_l2": small,
           "oracle_2p26_fused_rel_l2": small_fused, "parseval_rel_err": parseval,
           "vs": "impulse + three integer-bin tones at the full 2^%d points through the timed path (closed form, exact integer phases); the "
                 "separate-exchange path at 2^24 points and the fused path at 2^26 points vs oracle.fft" % lg}
    return {"_parity": par, "metric": "single 1-D FFT GS/s (complex128, 2^%d points over %d GPUs)" % (lg, world), "value": n / (ms * 1e-3) / 1e9,
            "unit": "GS/s", "ms_per_step": ms, "scaling": "weak", "log2n": lg,
            "api": ("godsp.distributed.fft_1d_sharded(peer=PeerExchange, fused=True): ONE fused TMA kernel for the length-N1 lines, the outer twiddle and the exchange "
                    "(its pass-2 stores go through one tensor map per rank into the ranks' receive buffers over NVLink; gd_fourstep_lines_peer_dev), then the "
                    "length-N2 lines on segmented rows (gd_fourstep_rows_seg_dev); result layout [K][N2]") if fused else
                   "godsp.distributed.fft_1d_sharded(peer=PeerExchange): strided lines, ONE kernel for twiddle + transpose + NVLink P2P stores into the peers' buffers (gd_fourstep_exchange_dev), strided lines",
            "nccl_all_to_all_variant_ms": ms_nccl, "separate_exchange_kernel_variant_ms": ms_sep, "phases": phases,
            "all_to_all_bytes_per_gpu": nv_bytes,
            "parseval_rel_err": parseval,
            "clocks": clocks, "_launches": int(launches), "_peer": px}


def run_pwelch(args, torch, dist, capi, L, sp, world, rank, local, timed, timed_host, skip, hbm_peak, peak_src):
    from godsp import window as gwindow     # host-side mirror of window/window.go (what the Go shim evaluates)
    nfft, nov = PW_NFFT, PW_NOVERLAP
    stride = nfft - nov
    ns_local = 1 << args.pw_log2_samples                   # samples owned by this rank
    total = ns_local * world
    nsegs = (total - nfft) // stride + 1                   # spectral.Segment count (spectral/spectral.go:22-33)
    lp = nfft // 2 + 1
    # rank r owns segments [s0, s1): those starting inside its sample range
    s0 = rank * ns_local // stride
    s1 = min(nsegs, (rank + 1) * ns_local // stride)
    halo = nov if rank < world - 1 else 0
    nloc = ns_local + halo
    x = torch.empty(nloc, dtype=torch.float64, device="cuda")
    capi.check(L.gd_fill_splitmix_dev(x.data_ptr(), nloc, PW_SEED, rank * ns_local, sp))
    win = gwindow.Hann(nfft)
    norm = 0.0
    for v in win:
        norm += v * v                                       # spectral/pwelch.go:124-128 (Fs = 1)
    dwin = torch.from_numpy(win).cuda()
    raw = torch.empty(lp, dtype=torch.float64, device="cuda")
    pxx = torch.empty(lp, dtype=torch.float64, device="cuda")
    gathered = [torch.empty(lp, dtype=torch.float64, device="cuda") for _ in range(world)] if world > 1 else None

    def step():
        capi.check(L.gd_pwelch_partial_dev(x.data_ptr(), nfft, nov, nfft, lp, 0, s1 - s0, dwin.data_ptr(), raw.data_ptr(), sp))
        if world > 1:
            dist.all_gather(gathered, raw)
            tot = gathered[0].clone()
            for g in gathered[1:]:
                tot += g                                    # fixed rank order: reproducible
            capi.check(L.gd_pwelch_finalize_dev(tot.data_ptr(), lp, nsegs, norm, pxx.data_ptr(), sp))
        else:
            capi.check(L.gd_pwelch_finalize_dev(raw.data_ptr(), lp, nsegs, norm, pxx.data_ptr(), sp))

    ms, launches, clocks = timed(step, args.steps, args.warmup)
    # ---- parity, outside the timed region
    # (1) the full-size result against a size-independent property (Parseval over every windowed segment):
    #     sum_j pxx[j] = L * sum_segs sum_n (w[n] x[n])^2 / (nsegs * norm)          (spectral/pwelch.go:113-121,134-136)
    tsum = torch.zeros((), dtype=torch.float64, device="cuda")
    segs = x.unfold(0, nfft, stride)                       # view, no copy: [local segments][nfft]
    for a in range(0, s1 - s0, 8192):
        blk = segs[a:min(a + 8192, s1 - s0)] * dwin
        tsum += (blk * blk).sum()
    if world > 1:
        dist.all_reduce(tsum)
    want_sum = nfft * float(tsum.item()) / (nsegs * norm)
    parseval = abs(float(pxx.sum().item()) / want_sum - 1.0)
    # (2) the same kernel on a 2^24-sample prefix of this rank's signal against oracle.pwelch
    import oracle
    npre = min(1 << 24, ns_local)
    xo = oracle.fill_splitmix(npre, PW_SEED, rank * ns_local)
    want, _ = oracle.pwelch(xo, 1.0, nfft=nfft, noverlap=nov, threads=min(8, os.cpu_count() or 1))
    nsegs_pre = (npre - nfft) // stride + 1
    raw2 = torch.empty(lp, dtype=torch.float64, device="cuda")
    pxx2 = torch.empty(lp, dtype=torch.float64, device="cuda")
    capi.check(L.gd_pwelch_partial_dev(x.data_ptr(), nfft, nov, nfft, lp, 0, nsegs_pre, dwin.data_ptr(), raw2.data_ptr(), sp))
    capi.check(L.gd_pwelch_finalize_dev(raw2.data_ptr(), lp, nsegs_pre, norm, pxx2.data_ptr(), sp))
    torch.cuda.synchronize()
    e_pre = rel_l2(pxx2.cpu().numpy(), want)
    par = {"max_rel_l2": allmax(torch, dist, world, max(e_pre, parseval)), "prefix_rel_l2": e_pre, "parseval_rel_err_full_size": parseval,
           "vs": "oracle.pwelch on a 2^%d-sample prefix of every rank's signal; Parseval over all %d windowed segments of the timed run" % (npre.bit_length() - 1, nsegs)}
    value = total / (ms * 1e-3) / 1e6
    achieved = 8.0 * ns_local / (ms * 1e-3) / 1e9
    out = {
        "metric": "Pwelch Msamples/s", "value": value, "unit": "Msamples/s", "ms_per_step": ms, "scaling": "weak",
        "config": {"workload": "spectral.Pwelch, %d float64 samples per GPU, NFFT 4096, Noverlap 2048, Hann (BASELINE.json configs[3])" % ns_local,
                   "samples_per_gpu": ns_local, "segments_total": int(nsegs), "bins": lp,
                   "parallelism": "segment ranges sharded over %d GPU(s); one %d-double all-gather, summed in rank order" % (world, lp)},
        "roofline": {"bound": "hbm", "kernel": "gd::pwelch_bulk_kernel", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "hbm_frac": achieved / hbm_peak,
                     "fp64_frac": value / world * 1e6 * FP64_INST_PER_SAMPLE_PWELCH / fp64_peak()[0], "fp64_inst_per_sample": FP64_INST_PER_SAMPLE_PWELCH,
                     "peak_source": peak_src, "traffic": (ncu_traffic("pwelch_bulk_kernel") or 0) * (ns_local / float(1 << 28)) or None,
                     "traffic_profiled_launch": ncu_traffic("pwelch_bulk_kernel"),
                     "traffic_note": "profiles/r2_ncu_full_summary.json holds one launch over 2^28 samples: 2.15 GB of dram traffic = the algorithmic 8 B per sample; `traffic` scales that to the samples of the timed launch",
                     "note": "8 B per input sample (SURVEY.md 8d); FP64 issue and the shared-memory pipe, not HBM, are the tighter roofs for this kernel (DESIGN.md 3.4)"},
        "clocks": clocks, "_launches": int(launches), "_parity": par,
        "pxx_checksum": float(pxx.sum().item()),
    }
    if "e2e" not in skip:
        nbytes = nloc * 8
        L.gd_pinned_alloc.restype = C.c_void_p
        hx = L.gd_pinned_alloc(nbytes)
        if not hx:
            raise SystemExit("pinned allocation failed: " + L.gd_last_error().decode())
        capi.check(L.gd_memcpy_d2h(hx, x.data_ptr(), nbytes))
        hp = np.empty(lp)
        nseg_loc = s1 - s0

        def e2e_step():
            capi.check(L.gd_pwelch_f64(hx, nloc, nfft, nov, nfft, lp, nseg_loc, win.ctypes.data, norm, hp.ctypes.data))

        ems = timed_host(e2e_step, args.e2e_steps, 1)
        out["e2e"] = {"value": total / (ems * 1e-3) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": nbytes,
                      "d2h_bytes_per_step": lp * 8, "ms_per_step": ems,
                      "api": "gd_pwelch_f64 (pinned host signal streamed in 256 MiB ranges, H2D overlapped with the fused kernel)"}
        L.gd_pinned_free(hx)
        # the same PSD from 16-bit PCM as wav.ReadSamples returns it: the ReadFloats conversion (wav/wav.go:138-161) runs in the
        # kernel's segment load, so the signal crosses PCIe at 2 bytes per sample instead of 8
        hx16 = L.gd_pinned_alloc(nloc * 2)
        if hx16:
            pcm = np.ctypeslib.as_array((C.c_int16 * nloc).from_address(hx16))
            blk = 1 << 24                                  # deterministic 16-bit signal: one hashed block, re-keyed per block and rank
            i = np.arange(blk, dtype=np.int64)
            base = (((i * 2654435761) >> 7) & 0xFFFF).astype(np.uint16).view(np.int16)
            for k, a in enumerate(range(0, nloc, blk)):
                e = min(nloc, a + blk)
                pcm[a:e] = base[: e - a] ^ np.int16(((k + 1 + 64 * rank) * 7919) & 0x7FFF)
            del i, base
            hp16 = np.empty(lp)

            def e2e16_step():
                capi.check(L.gd_pwelch_samples(hx16, 2, nloc, nfft, nov, nfft, lp, nseg_loc, win.ctypes.data, norm, hp16.ctypes.data))

            ems16 = timed_host(e2e16_step, args.e2e_steps, 1)
            # parity of the PCM path: a 2^22-sample prefix against oracle.pwelch of the oracle's ReadFloats restatement
            import oracle
            npre = min(1 << 22, nloc)
            nsp = (npre - nfft) // stride + 1
            hpre = np.empty(lp)
            capi.check(L.gd_pwelch_samples(hx16, 2, npre, nfft, nov, nfft, lp, nsp, win.ctypes.data, norm, hpre.ctypes.data))
            xf = oracle.wav_read_floats(pcm[:npre].astype("<i2").tobytes(), 2, npre).astype(np.float64)
            want16, _ = oracle.pwelch(xf, 1.0, nfft=nfft, noverlap=nov, threads=min(8, os.cpu_count() or 1))
            e16 = rel_l2(hpre, want16)
            out["e2e_pcm16"] = {"value": total / (ems16 * 1e-3) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": nloc * 2,
                                "d2h_bytes_per_step": lp * 8, "ms_per_step": ems16, "prefix_rel_l2_vs_oracle": e16,
                                "api": "gd_pwelch_samples(GD_SAMPLE_S16): int16 PCM over PCIe, decoded in the Pwelch kernel's segment load"}
            out["_parity"]["max_rel_l2"] = allmax(torch, dist, world, max(out["_parity"]["max_rel_l2"], e16))
            out["_parity"]["pcm16_prefix_rel_l2"] = e16
            del pcm
            L.gd_pinned_free(hx16)
    if "cpu" not in skip and world == 1:
        threads = os.cpu_count() or 1
        v, desc, _ = cpu_pwelch_sample(threads, 1.5)
        out["cpu_baseline"] = {"value": v, "unit": "Msamples/s", "cores": threads, "kind": "port",
                               "sample": desc + "; C restatement of go-dsp, not the Go binary"}
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

import os, sys, time, ctypes as C
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/go-dsp_b200")
import torch, torch.distributed as dist
from godsp import _capi as capi
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
use_aff = int(sys.argv[1])
torch.cuda.set_device(local)
if use_aff:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
    if cpus: os.sched_setaffinity(0, cpus)
    if rank == 0: print("cpus for gpu0:", len(cpus), cpus[:8], flush=True)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = capi.lib(); capi.check(L.gd_use_device(local))
n, eb = 1 << 20, 256
nbytes = eb * n * 16
L.gd_pinned_alloc.restype = C.c_void_p
hin, hout = L.gd_pinned_alloc(nbytes), L.gd_pinned_alloc(nbytes)
C.memset(hin, 1, nbytes)
capi.check(L.gd_fft_batch_c2c(hin, hout, n, eb, 1))
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): capi.check(L.gd_fft_batch_c2c(hin, hout, n, eb, 1))
dt = (time.perf_counter() - t0) / 3
t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0: print("affinity", use_aff, "e2e GS/s total", eb * n * world / t.item() / 1e9, "per-direction GB/s per GPU", nbytes / t.item() / 1e9, flush=True)
dist.destroy_process_group()

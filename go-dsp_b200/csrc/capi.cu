// capi.cu -- the C ABI (include/godsp_b200.h): device contexts, staging, error convention.
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/godsp_b200.h"
#include "engine.h"

using namespace gd;

namespace {

constexpr int kMaxDev = 16;
// Lanes: independent execution contexts of one GPU (own streams, events, scratch, staging, plan caches). A call takes a free
// lane of its device, so concurrent host threads (goroutines) working on the same GPU overlap their copies and kernels
// instead of queueing on one mutex; a single-threaded caller always gets lane 0.
constexpr int kLanes = 2;
Device g_dev[kMaxDev * kLanes];
inline Device& lane_of(int dev, int lane) { return g_dev[dev * kLanes + lane]; }
thread_local int t_lane_held[kMaxDev] = {0};            // 1 + lane this thread currently holds on a device (nested entry points)
std::atomic<int> g_ndev{0};
std::mutex g_init_mu;
thread_local int t_dev = 0;

struct StageEvents {
    cudaEvent_t h2d[2] = {nullptr, nullptr}, comp[2] = {nullptr, nullptr}, d2h[2] = {nullptr, nullptr};
    bool ready = false;
};
StageEvents g_ev[kMaxDev * kLanes];

Status ensure_events(int slot) {
    StageEvents& e = g_ev[slot];
    if (e.ready) return ::gd::GD_OK;
    for (int i = 0; i < 2; i++) {
        GD_CUDA(cudaEventCreateWithFlags(&e.h2d[i], cudaEventDisableTiming));
        GD_CUDA(cudaEventCreateWithFlags(&e.comp[i], cudaEventDisableTiming));
        GD_CUDA(cudaEventCreateWithFlags(&e.d2h[i], cudaEventDisableTiming));
    }
    e.ready = true;
    return ::gd::GD_OK;
}

// devices are initialised one by one, on first use (a torchrun rank only ever touches its own GPU)
Status ensure_device(int dev) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    if (g_ndev.load() == 0) {
        int have = 0;
        cudaError_t e = cudaGetDeviceCount(&have);
        if (e != cudaSuccess || have < 1) {
            set_error(std::string("no CUDA device: go-dsp_b200 has no CPU fallback (") +
                      (e != cudaSuccess ? cudaGetErrorString(e) : "0 devices") + ")");
            return ::gd::GD_ERR_CUDA;
        }
        g_ndev.store(have > kMaxDev ? kMaxDev : have);
    }
    if (dev < 0 || dev >= g_ndev.load()) { set_error("device index out of range"); return ::gd::GD_ERR_INVALID; }
    if (!lane_of(dev, 0).ready) GD_TRY(lane_of(dev, 0).init(dev, 0));
    return ::gd::GD_OK;
}

// RAII: select + lock the calling thread's device
struct DevLock {
    Device* d = nullptr;
    Status st = ::gd::GD_OK;
    std::unique_lock<std::recursive_mutex> lk;
    int dev = 0, prev_held = 0;
    explicit DevLock(int only_lane = -1) {
        dev = t_dev;
        st = ensure_device(dev);
        if (st != ::gd::GD_OK) return;
        prev_held = t_lane_held[dev];
        if (only_lane >= 0) {                               // gd_set_option walks every lane
            d = &lane_of(dev, only_lane);
            lk = std::unique_lock<std::recursive_mutex>(d->mu);
        } else if (prev_held) {                             // nested entry point: stay on the lane this thread already holds
            d = &lane_of(dev, prev_held - 1);
            lk = std::unique_lock<std::recursive_mutex>(d->mu);
        } else {
            for (int k = 0; k < kLanes && !d; k++) {
                std::unique_lock<std::recursive_mutex> l(lane_of(dev, k).mu, std::try_to_lock);
                if (l.owns_lock()) { d = &lane_of(dev, k); lk = std::move(l); }
            }
            if (!d) {                                       // all busy: queue on one of them
                d = &lane_of(dev, (int)(std::hash<std::thread::id>()(std::this_thread::get_id()) % kLanes));
                lk = std::unique_lock<std::recursive_mutex>(d->mu);
            }
        }
        const int lane = (int)(d - &lane_of(dev, 0));
        t_lane_held[dev] = 1 + lane;
        if (!d->ready) { st = d->init(dev, lane); if (st != ::gd::GD_OK) return; }
        cudaError_t e = cudaSetDevice(d->dev);
        if (e != cudaSuccess) st = cuda_fail(e, "cudaSetDevice");
    }
    ~DevLock() { if (lk.owns_lock()) t_lane_held[dev] = prev_held; }
};

#define GD_ENTER()        \
    DevLock L__;          \
    if (L__.st != ::gd::GD_OK) return (int)L__.st; \
    Device& d = *L__.d

inline cudaStream_t pick(Device& d, void* stream) { return stream ? (cudaStream_t)stream : d.stream; }

Status invalid_arg(const char* m) { set_error(m); return ::gd::GD_ERR_INVALID; }

// whole-buffer upload / download on the compute stream (pinned memory is truly async)
Status up(Device& d, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    (void)d;
    GD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return ::gd::GD_OK;
}
Status down(Device& d, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    (void)d;
    GD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    return ::gd::GD_OK;
}

__global__ void add_inplace_kernel(double* tot, const double* part, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) tot[i] += part[i];
}

// ---- multi-device calls ("multi-GPU calls fan out inside C") ----
// After gd_init(ndev > 1) the batched host-pointer entry points split their work over devices 0 .. ndev-1, one host thread
// per device, each running the single-device path on its share. Threads of one process: no torch, no NCCL.
std::atomic<int> g_fanout{1};
std::atomic<int> g_fanout_min_log2n{26};     // single power-of-two transforms of at least this size are sharded over the devices
thread_local bool t_in_fanout = false;

// CPUs next to a GPU (sysfs local_cpulist of its PCI function); false when the container hides the topology
bool device_cpuset(int dev, cpu_set_t* set) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) return false;
    for (char* p = bus; *p; p++) if (*p >= 'A' && *p <= 'F') *p = (char)(*p - 'A' + 'a');
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return false;
    char buf[4096] = {0};
    const size_t n = fread(buf, 1, sizeof(buf) - 1, f);
    fclose(f);
    if (n == 0) return false;
    CPU_ZERO(set);
    int count = 0;
    for (char* p = buf; *p && *p != '\n';) {
        char* e;
        long a = strtol(p, &e, 10), b = a;
        if (e == p) break;
        if (*e == '-') { p = e + 1; b = strtol(p, &e, 10); }
        for (long k = a; k <= b && k < CPU_SETSIZE; k++) { CPU_SET((int)k, set); count++; }
        p = *e == ',' ? e + 1 : e;
    }
    return count > 0;
}
// RAII: run the calling thread on the CPUs next to `dev` (NUMA-local pinned allocations and staging copies)
struct NearDevice {
    cpu_set_t old;
    bool changed = false;
    explicit NearDevice(int dev) {
        cpu_set_t want;
        if (sched_getaffinity(0, sizeof(old), &old) != 0 || !device_cpuset(dev, &want)) return;
        cpu_set_t both;
        CPU_AND(&both, &old, &want);
        if (CPU_COUNT(&both) > 0 && sched_setaffinity(0, sizeof(both), &both) == 0) changed = true;
    }
    ~NearDevice() { if (changed) sched_setaffinity(0, sizeof(old), &old); }
};

template <class F>
int fan_out(int ndev, F&& fn) {
    std::vector<std::thread> th;
    std::vector<int> rc((size_t)ndev, 0);
    std::vector<std::string> err((size_t)ndev);
    for (int i = 0; i < ndev; i++)
        th.emplace_back([&, i] {
            t_dev = i;
            t_in_fanout = true;
            NearDevice near(i);
            rc[(size_t)i] = fn(i);
            if (rc[(size_t)i]) err[(size_t)i] = last_error();
        });
    for (auto& t : th) t.join();
    for (int i = 0; i < ndev; i++)
        if (rc[(size_t)i]) { set_error("device " + std::to_string(i) + ": " + err[(size_t)i]); return rc[(size_t)i]; }
    return ::gd::GD_OK;
}
inline int fanout_width() { return t_in_fanout ? 1 : g_fanout.load(); }
inline long long share_begin(long long total, int parts, int i) { return total * i / parts; }

// host barrier for the threads of one fan-out
struct HostBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int n, waiting = 0, phase = 0;
    explicit HostBarrier(int count) : n(count) {}
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        const int ph = phase;
        if (++waiting == n) { waiting = 0; phase++; cv.notify_all(); }
        else cv.wait(lk, [&] { return phase != ph; });
    }
};

// ---- pageable host memory (what fft.FFT(x) on a plain Go slice hands us) ----
// cudaMemcpyAsync from pageable memory is staged by the driver on one thread; large calls go through a pinned ring of
// this library instead, filled by a few host threads, so the DMA engines see pinned memory and the copies overlap.
bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    const size_t min_part = (size_t)8 << 20;
    static const int max_parts = std::max(2, std::min(8, (int)std::thread::hardware_concurrency() / 2 - 1));
    int parts = (int)std::min<size_t>((size_t)max_parts, bytes / min_part);
    if (parts <= 1) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    const size_t per = ((bytes / (size_t)parts) + 4095) & ~(size_t)4095;
    for (int i = 1; i < parts; i++) {
        const size_t off = per * (size_t)i;
        if (off >= bytes) break;
        const size_t len = std::min(per, bytes - off);
        th.emplace_back([=] { memcpy((char*)dst + off, (const char*)src + off, len); });
    }
    memcpy(dst, src, std::min(per, bytes));
    for (auto& t : th) t.join();
}
struct PinnedRing {              // per device: two input and two output slots, grown on demand, kept for the life of the library
    char* in[2] = {nullptr, nullptr};
    char* out[2] = {nullptr, nullptr};
    size_t in_bytes = 0, out_bytes = 0;
};
PinnedRing g_ring[kMaxDev * kLanes];
Status ensure_ring(int slot, int dev, size_t in_bytes, size_t out_bytes) {
    PinnedRing& r = g_ring[slot];
    NearDevice near(dev);
    if (r.in_bytes < in_bytes) {
        for (int i = 0; i < 2; i++) { if (r.in[i]) cudaFreeHost(r.in[i]); r.in[i] = nullptr; }
        r.in_bytes = 0;
        for (int i = 0; i < 2; i++) GD_CUDA(cudaHostAlloc((void**)&r.in[i], in_bytes, cudaHostAllocPortable));
        r.in_bytes = in_bytes;
    }
    if (r.out_bytes < out_bytes) {
        for (int i = 0; i < 2; i++) { if (r.out[i]) cudaFreeHost(r.out[i]); r.out[i] = nullptr; }
        r.out_bytes = 0;
        for (int i = 0; i < 2; i++) GD_CUDA(cudaHostAlloc((void**)&r.out[i], out_bytes, cudaHostAllocPortable));
        r.out_bytes = out_bytes;
    }
    return ::gd::GD_OK;
}

}  // namespace

extern "C" {

int gd_init(int ndev) {
    Status s = ensure_device(0);
    if (s != ::gd::GD_OK) return (int)s;
    if (ndev <= 0 || ndev > g_ndev.load()) ndev = g_ndev.load();
    for (int i = 1; i < ndev; i++) {
        s = ensure_device(i);
        if (s != ::gd::GD_OK) return (int)s;
    }
    // the batched host-pointer calls now spread over devices 0 .. ndev-1 (one host thread per device); processes that
    // never call gd_init (one rank per GPU under torchrun: gd_use_device only) keep every call on their own device
    g_fanout.store(ndev);
    return ::gd::GD_OK;
}

int gd_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    g_fanout.store(1);
    for (int i = 0; i < g_ndev.load() * kLanes; i++) {
        PinnedRing& r = g_ring[i];
        for (int k = 0; k < 2; k++) { if (r.in[k]) cudaFreeHost(r.in[k]); if (r.out[k]) cudaFreeHost(r.out[k]); r.in[k] = r.out[k] = nullptr; }
        r.in_bytes = r.out_bytes = 0;
        if (!g_dev[i].ready) continue;
        g_dev[i].destroy();
        StageEvents& e = g_ev[i];
        if (e.ready) {
            for (int k = 0; k < 2; k++) { cudaEventDestroy(e.h2d[k]); cudaEventDestroy(e.comp[k]); cudaEventDestroy(e.d2h[k]); }
            e.ready = false;
        }
    }
    return ::gd::GD_OK;
}

const char* gd_last_error(void) { return last_error(); }

int gd_device_count(void) {
    int n = 0;
    for (int i = 0; i < g_ndev.load(); i++) n += lane_of(i, 0).ready ? 1 : 0;
    return n;
}

int gd_use_device(int dev) {
    Status s = ensure_device(dev);
    if (s != ::gd::GD_OK) return (int)s;
    t_dev = dev;
    return ::gd::GD_OK;
}

static int set_option_one(Device& d, const char* key, int64_t value);
int gd_set_option(const char* key, int64_t value) {
    if (!key) return (int)invalid_arg("gd_set_option: null key");
    for (int k = 0; k < kLanes; k++) {                      // every lane of the calling thread's device gets the value
        DevLock L__(k);
        if (L__.st != ::gd::GD_OK) return (int)L__.st;
        const int rc = set_option_one(*L__.d, key, value);
        if (rc) return rc;
    }
    return ::gd::GD_OK;
}
static int set_option_one(Device& d, const char* key, int64_t value) {
    if (!strcmp(key, "pass_scratch_mb")) { if (value < 1) return (int)invalid_arg("pass_scratch_mb < 1"); d.pass_scratch_budget = (size_t)value << 20; }
    else if (!strcmp(key, "l2_block_mb")) { if (value < 1) return (int)invalid_arg("l2_block_mb < 1"); d.l2_block_budget = (size_t)value << 20; }
    else if (!strcmp(key, "chunk_streams")) { if (value < 1 || value > 4) return (int)invalid_arg("chunk_streams out of range"); d.chunk_streams = (int)value; }
    else if (!strcmp(key, "l2_block_window")) d.l2_block_window = value != 0 && d.lane == 0;
    else if (!strcmp(key, "pwelch_bulk")) d.pwelch_bulk = value != 0;
    else if (!strcmp(key, "fanout_min_log2n")) { if (value < 12 || value > 40) return (int)invalid_arg("fanout_min_log2n out of range"); g_fanout_min_log2n.store((int)value); }
    else if (!strcmp(key, "fourstep_pipeline")) d.fourstep_pipeline = value != 0;
    else if (!strcmp(key, "fourstep_lines_sms")) { if (value < 0 || value > 1024) return (int)invalid_arg("fourstep_lines_sms out of range"); d.fourstep_lines_sms = (int)value; }
    else if (!strcmp(key, "fourstep_exchange_ctas")) d.fourstep_exchange_ctas = (int)value;
    else if (!strcmp(key, "fourstep_pipeline_mb")) { if (value < 1) return (int)invalid_arg("fourstep_pipeline_mb < 1"); d.fourstep_pipeline_mb = (int)value; }
    else if (!strcmp(key, "wide_tiles")) d.wide_tiles = value != 0;
    else if (!strcmp(key, "fused")) d.use_fused = value != 0;
    else if (!strcmp(key, "bluestein_fused")) d.bluestein_fused = value != 0;
    else if (!strcmp(key, "debug_alias")) d.debug_alias = value != 0;
    else if (!strcmp(key, "w32")) { if (value < 0 || value > 6) return (int)invalid_arg("w32 out of range"); d.w32 = (int)value; }
    else if (!strcmp(key, "tma")) d.use_tma = value != 0;
    else if (!strcmp(key, "tma_opt")) d.tma_opt = (int)value;
    else if (!strcmp(key, "tma14")) d.use_tma14 = value != 0;
    else if (!strcmp(key, "tma16")) d.use_tma16 = value != 0;
    else if (!strcmp(key, "tma19")) d.use_tma19 = value != 0;
    else if (!strcmp(key, "huge_min_log2n")) { if (value < 19 || value > 25) return (int)invalid_arg("huge_min_log2n out of range"); d.huge_min_log2n = (int)value; }
    else if (!strcmp(key, "real_widen")) d.real_widen = value != 0;
    else if (!strcmp(key, "bluestein_stream")) d.bluestein_stream = value != 0;
    else if (!strcmp(key, "bluestein_fuse_mul")) d.bluestein_fuse_mul = value != 0;
    else if (!strcmp(key, "bluestein_chunk_mb")) { if (value < 1 || value > 16384) return (int)invalid_arg("bluestein_chunk_mb out of range"); d.bluestein_chunk_bytes = (size_t)value << 20; }
    else if (!strcmp(key, "axis_single_max_log2")) { if (value < 4 || value > 12) return (int)invalid_arg("axis_single_max_log2 out of range"); d.axis_single_max_log2 = (int)value; }
    else if (!strcmp(key, "huge_l1")) { if (value != 0 && (value < 13 || value > 17)) return (int)invalid_arg("huge_l1 out of range"); d.huge_l1 = (int)value; }
    else if (!strcmp(key, "huge_sweeps")) { if (value != 0 && value != 3 && value != 4) return (int)invalid_arg("huge_sweeps out of range"); d.huge_sweeps = (int)value; }
    else if (!strcmp(key, "tma_grid_cap")) { if (value < 0 || value > 1024) return (int)invalid_arg("tma_grid_cap out of range"); d.tma_grid_cap = (int)value; }
    else if (!strcmp(key, "tma_prof")) d.tma_prof = value != 0;
    else if (!strcmp(key, "tma_delay")) { if (value < 0 || value > 4) return (int)invalid_arg("tma_delay out of range"); d.tma_delay = (int)value; }
    else if (!strcmp(key, "tma_slots")) { if (value < 2 || value > 6) return (int)invalid_arg("tma_slots out of range"); d.tma_slots = (int)value; }
    else if (!strcmp(key, "tiled_scratch")) d.tiled_scratch = value != 0;
    else if (!strcmp(key, "l2_window")) d.use_l2_window = value != 0 && d.lane == 0;
    else if (!strcmp(key, "fused_delay")) { if (value < 1 || value > 6) return (int)invalid_arg("fused_delay out of range"); d.fused_delay = (int)value; }
    else if (!strcmp(key, "fused_slot_mb")) { if (value < 1) return (int)invalid_arg("fused_slot_mb < 1"); d.fused_slot_budget = (size_t)value << 20; }
    else return (int)invalid_arg("gd_set_option: unknown key");
    return ::gd::GD_OK;
}

int64_t gd_kernel_launches(void) { return g_launches.load(); }

int gd_tma_profile_read(int64_t* out, int max_ctas) {
    if (!out || max_ctas < 1) return (int)invalid_arg("gd_tma_profile_read: bad arguments");
    GD_ENTER();
    if (!d.scratch[SCR_PROF]) return (int)invalid_arg("gd_tma_profile_read: no profiled launch yet (gd_set_option(\"tma_prof\", 1))");
    const int n = d.num_sms < max_ctas ? d.num_sms : max_ctas;
    GD_CUDA(cudaDeviceSynchronize());
    GD_CUDA(cudaMemcpy(out, d.scratch[SCR_PROF], (size_t)n * 32 * sizeof(long long), cudaMemcpyDeviceToHost));
    return n;
}

int64_t gd_bluestein_padded_len(int64_t n) {
    int64_t need = 2 * n - 1, la = 1;
    if ((n & (n - 1)) == 0 && n > 0 && need < 1) return 1;
    while (la < need) la <<= 1;
    return la;
}

// ---------------------------------------------------------------- host-pointer API

// ONE power-of-two transform of n = N1 * N2 points over `width` devices of this process (BASELINE config C5 without
// torchrun / NCCL): device g owns the columns [g W, (g + 1) W) of x viewed as [N1][N2]; lines over n1; ONE kernel that multiplies
// by w_N^(k1 n2), transposes and stores into the peers' receive buffers over NVLink (fourstep_exchange_kernel); lines over n2;
// device h then holds X[k1 + N1 k2] for k1 in [h K, (h + 1) K). Host <-> device traffic is two strided 2-D copies per device.
static int fft1d_fanout(const double* in, double* out, int lg, int dir, int width) {
    const int l1 = (lg + 1) / 2;
    const long long N1 = 1LL << l1, N2 = 1LL << (lg - l1), K = N1 / width, W = N2 / width, per = K * N2;
    std::vector<cpx*> A((size_t)width, nullptr), R((size_t)width, nullptr);
    HostBarrier bar(width);
    std::atomic<int> failed{0};
    return fan_out(width, [&](int i) -> int {
        int rc = 0;
        auto step = [&](Status s) { if (s != ::gd::GD_OK && !rc) { rc = (int)s; failed.store(1); } };
        auto cuda_step = [&](cudaError_t e, const char* what) { if (e != cudaSuccess) step(cuda_fail(e, what)); };
        DevLock L;
        if (L.st != ::gd::GD_OK) { failed.store(1); rc = (int)L.st; for (int k = 0; k < 2; k++) bar.wait(); return rc; }
        Device& d = *L.d;
        ScratchOrder order__(d, d.stream);
        for (int h = 0; h < width; h++)
            if (h != i) { cudaError_t e = cudaDeviceEnablePeerAccess(h, 0); if (e != cudaSuccess) cudaGetLastError(); }
        step(d.ensure_scratch(SCR_STAGE_IN, (size_t)per * sizeof(cpx), (void**)&A[(size_t)i]));
        step(d.ensure_scratch(SCR_STAGE_OUT, (size_t)per * sizeof(cpx), (void**)&R[(size_t)i]));
        cpx* a = A[(size_t)i];
        if (!rc) cuda_step(cudaMemcpy2DAsync(a, (size_t)W * sizeof(cpx), (const cpx*)in + (size_t)i * W, (size_t)N2 * sizeof(cpx),
                                             (size_t)W * sizeof(cpx), (size_t)N1, cudaMemcpyHostToDevice, d.stream), "cudaMemcpy2DAsync(H2D slab)");
        if (!rc) step(fft_strided(d, a, a, 1, N1, W, dir, d.stream));                       // lines over n1
        cudaStreamSynchronize(d.stream);
        bar.wait();                                                                         // every receive buffer exists and is idle
        if (!failed.load()) step(fourstep_exchange(a, R.data(), N1, W, i, width, lg, d.stream, 0, -1, 0, dir));
        cudaStreamSynchronize(d.stream);
        bar.wait();                                                                         // every device's stores have landed
        if (!failed.load()) step(fft_strided(d, R[(size_t)i], a, 1, N2, K, dir, d.stream));  // lines over n2: a[k2][k] = X[i K + k + N1 k2]
        if (!failed.load()) cuda_step(cudaMemcpy2DAsync((cpx*)out + (size_t)i * K, (size_t)N1 * sizeof(cpx), a, (size_t)K * sizeof(cpx),
                                                        (size_t)K * sizeof(cpx), (size_t)N2, cudaMemcpyDeviceToHost, d.stream), "cudaMemcpy2DAsync(D2H slab)");
        cudaError_t e = cudaStreamSynchronize(d.stream);
        if (e != cudaSuccess) step(cuda_fail(e, "fft1d_fanout"));
        if (!rc && failed.load()) { set_error("another device of the fan-out failed"); rc = (int)::gd::GD_ERR_CUDA; }
        return rc;
    });
}
static bool all_peers(int width) {
    for (int a = 0; a < width; a++)
        for (int b = 0; b < width; b++)
            if (a != b) { int ok = 0; if (cudaDeviceCanAccessPeer(&ok, a, b) != cudaSuccess || !ok) return false; }
    return true;
}

static int fft_host(const double* in, double* out, int64_t n, int64_t batch, bool real_in, int dir) {
    if (!in || !out || n < 1 || batch < 1 || (dir != 1 && dir != -1)) return (int)invalid_arg("fft: bad arguments");
    const int width = fanout_width();
    if (width > 1 && batch == 1 && !real_in && (n & (n - 1)) == 0 && (width & (width - 1)) == 0 && width <= 16) {
        // one large power-of-two transform: sharded four-step over the devices of this process
        int lg = 0;
        while ((1LL << lg) < n) lg++;
        const long long n2 = 1LL << (lg - (lg + 1) / 2);
        if (lg >= g_fanout_min_log2n.load() && lg <= 36 && n2 / width >= 32 && all_peers(width)) return fft1d_fanout(in, out, lg, dir, width);
    }
    if (width > 1 && batch >= 2 * (int64_t)width && (size_t)batch * (size_t)n >= ((size_t)1 << 22)) {
        // independent transforms: contiguous row ranges, one per device (no collective)
        const size_t in_el = real_in ? 1 : 2;
        return fan_out(width, [&](int i) {
            const long long r0 = share_begin(batch, width, i), r1 = share_begin(batch, width, i + 1);
            if (r1 <= r0) return 0;
            return fft_host(in + (size_t)r0 * n * in_el, out + (size_t)r0 * n * 2, n, r1 - r0, real_in, dir);
        });
    }
    GD_ENTER();
    ScratchOrder order__(d, d.stream);
    GD_TRY(ensure_events(d.slot()));
    StageEvents& ev = g_ev[d.slot()];
    const size_t in_el = real_in ? sizeof(double) : sizeof(cpx);
    // chunk the batch so H2D of chunk c+1, the kernels of chunk c and D2H of chunk c-1 overlap
    long long chunk = (long long)((128ull << 20) / ((size_t)n * sizeof(cpx)));
    if (chunk < 1) chunk = 1;
    if (chunk > batch) chunk = batch;
    char* din; cpx* dout;
    GD_TRY(d.ensure_scratch(SCR_STAGE_IN, 2 * (size_t)chunk * n * in_el, (void**)&din));
    GD_TRY(d.ensure_scratch(SCR_STAGE_OUT, 2 * (size_t)chunk * n * sizeof(cpx), (void**)&dout));
    if (chunk >= batch) {      // one chunk: nothing to overlap
        GD_TRY(up(d, din, in, (size_t)batch * n * in_el, d.stream));
        GD_TRY(fft1d(d, din, n, dout, n, n, batch, real_in, dir, d.stream));
        GD_TRY(down(d, out, dout, (size_t)batch * n * sizeof(cpx), d.stream));
        GD_CUDA(cudaStreamSynchronize(d.stream));
        return ::gd::GD_OK;
    }
    // plan (tables, Bluestein caches, scratch growth) before the pipeline starts
    GD_CUDA(cudaMemsetAsync(din, 0, (size_t)n * in_el, d.stream));
    GD_TRY(fft1d(d, din, n, dout, n, n, 1, real_in, dir, d.stream));
    GD_CUDA(cudaStreamSynchronize(d.stream));
    if ((!is_pinned(in) || !is_pinned(out)) && !getenv("GD_NO_STAGING_RING")) {
        // pageable memory: every chunk goes through the pinned ring, filled / drained by host threads while the copy
        // engines and the kernels work on the neighbouring chunks
        const size_t cin = (size_t)chunk * n * in_el, cout = (size_t)chunk * n * sizeof(cpx);
        GD_TRY(ensure_ring(d.slot(), d.dev, cin, cout));
        PinnedRing& ring = g_ring[d.slot()];
        const long long nchunks = (batch + chunk - 1) / chunk;
        std::thread drain;
        cudaError_t drain_err = cudaSuccess;
        struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{drain};      // early returns below
        for (long long c = 0; c <= nchunks; c++) {
            const int s = (int)(c & 1);
            if (c >= 1) {                           // drain chunk c-1 on a helper thread while chunk c is being filled
                const int sp = (int)((c - 1) & 1);
                const long long b0 = (c - 1) * chunk, nb = std::min<long long>(chunk, batch - b0);
                const int dev = d.dev;
                drain = std::thread([&, sp, b0, nb, dev] {
                    cudaSetDevice(dev);
                    drain_err = cudaEventSynchronize(ev.d2h[sp]);
                    if (drain_err == cudaSuccess) parallel_memcpy((cpx*)out + (size_t)b0 * n, ring.out[sp], (size_t)nb * n * sizeof(cpx));
                });
            }
            if (c < nchunks) {
                const long long b0 = c * chunk, nb = std::min<long long>(chunk, batch - b0);
                if (c >= 2) GD_CUDA(cudaEventSynchronize(ev.h2d[s]));                  // the slot's previous upload has left it
                parallel_memcpy(ring.in[s], (const char*)in + (size_t)b0 * n * in_el, (size_t)nb * n * in_el);
                char* di = din + (size_t)s * chunk * n * in_el;
                cpx* dob = dout + (size_t)s * chunk * n;
                if (c >= 2) GD_CUDA(cudaStreamWaitEvent(d.stream_in, ev.comp[s], 0));
                GD_CUDA(cudaMemcpyAsync(di, ring.in[s], (size_t)nb * n * in_el, cudaMemcpyHostToDevice, d.stream_in));
                GD_CUDA(cudaEventRecord(ev.h2d[s], d.stream_in));
                GD_CUDA(cudaStreamWaitEvent(d.stream, ev.h2d[s], 0));
                if (c >= 2) GD_CUDA(cudaStreamWaitEvent(d.stream, ev.d2h[s], 0));
                GD_TRY(fft1d(d, di, n, dob, n, n, nb, real_in, dir, d.stream));
                GD_CUDA(cudaEventRecord(ev.comp[s], d.stream));
                GD_CUDA(cudaStreamWaitEvent(d.stream_out, ev.comp[s], 0));
                // ring.out[s] was last read by the drain of chunk c-2, joined at the end of the previous iteration
                GD_CUDA(cudaMemcpyAsync(ring.out[s], dob, (size_t)nb * n * sizeof(cpx), cudaMemcpyDeviceToHost, d.stream_out));
                GD_CUDA(cudaEventRecord(ev.d2h[s], d.stream_out));
            }
            if (drain.joinable()) { drain.join(); if (drain_err != cudaSuccess) return (int)cuda_fail(drain_err, "cudaEventSynchronize(d2h)"); }
        }
        if (drain.joinable()) { drain.join(); if (drain_err != cudaSuccess) return (int)cuda_fail(drain_err, "cudaEventSynchronize(d2h)"); }
        GD_CUDA(cudaStreamSynchronize(d.stream));
        return ::gd::GD_OK;
    }
    long long c = 0;
    for (long long b0 = 0; b0 < batch; b0 += chunk, c++) {
        const int s = (int)(c & 1);
        const long long nb = std::min<long long>(chunk, batch - b0);
        char* di = din + (size_t)s * chunk * n * in_el;
        cpx* dob = dout + (size_t)s * chunk * n;
        if (c >= 2) GD_CUDA(cudaStreamWaitEvent(d.stream_in, ev.comp[s], 0));
        GD_CUDA(cudaMemcpyAsync(di, (const char*)in + (size_t)b0 * n * in_el, (size_t)nb * n * in_el, cudaMemcpyHostToDevice, d.stream_in));
        GD_CUDA(cudaEventRecord(ev.h2d[s], d.stream_in));
        GD_CUDA(cudaStreamWaitEvent(d.stream, ev.h2d[s], 0));
        if (c >= 2) GD_CUDA(cudaStreamWaitEvent(d.stream, ev.d2h[s], 0));
        GD_TRY(fft1d(d, di, n, dob, n, n, nb, real_in, dir, d.stream));
        GD_CUDA(cudaEventRecord(ev.comp[s], d.stream));
        GD_CUDA(cudaStreamWaitEvent(d.stream_out, ev.comp[s], 0));
        GD_CUDA(cudaMemcpyAsync((cpx*)out + (size_t)b0 * n, dob, (size_t)nb * n * sizeof(cpx), cudaMemcpyDeviceToHost, d.stream_out));
        GD_CUDA(cudaEventRecord(ev.d2h[s], d.stream_out));
    }
    GD_CUDA(cudaStreamSynchronize(d.stream_out));
    GD_CUDA(cudaStreamSynchronize(d.stream));
    return ::gd::GD_OK;
}

int gd_fft_c2c(const double* in, double* out, int64_t n, int dir) { return fft_host(in, out, n, 1, false, dir); }
int gd_fft_r2c_full(const double* in, double* out, int64_t n, int dir) { return fft_host(in, out, n, 1, true, dir); }
int gd_fft_batch_c2c(const double* in, double* out, int64_t n, int64_t batch, int dir) { return fft_host(in, out, n, batch, false, dir); }

int gd_convolve_c2c(const double* x, const double* y, double* out, int64_t n) {
    if (!x || !y || !out || n < 1) return (int)invalid_arg("convolve: bad arguments");
    GD_ENTER();
    ScratchOrder order__(d, d.stream);
    cpx *din, *dout;
    GD_TRY(d.ensure_scratch(SCR_STAGE_IN, 2 * (size_t)n * sizeof(cpx), (void**)&din));
    GD_TRY(d.ensure_scratch(SCR_STAGE_OUT, (size_t)n * sizeof(cpx), (void**)&dout));
    GD_TRY(up(d, din, x, (size_t)n * sizeof(cpx), d.stream));
    GD_TRY(up(d, din + n, y, (size_t)n * sizeof(cpx), d.stream));
    GD_TRY(convolve(d, din, din + n, dout, n, d.stream));
    GD_TRY(down(d, out, dout, (size_t)n * sizeof(cpx), d.stream));
    GD_CUDA(cudaStreamSynchronize(d.stream));
    return ::gd::GD_OK;
}

int gd_fftn_c2c(const double* in, double* out, const int64_t* dims, int nd, int dir) {
    if (!in || !out || !dims || nd < 1 || nd > 16 || (dir != 1 && dir != -1)) return (int)invalid_arg("fftn: bad arguments");
    long long ld[16], total = 1;
    for (int i = 0; i < nd; i++) { if (dims[i] < 1) return (int)invalid_arg("fftn: invalid dimensions"); ld[i] = dims[i]; total *= dims[i]; }
    GD_ENTER();
    ScratchOrder order__(d, d.stream);
    cpx* buf;
    GD_TRY(d.ensure_scratch(SCR_STAGE_OUT, (size_t)total * sizeof(cpx), (void**)&buf));
    GD_TRY(up(d, buf, in, (size_t)total * sizeof(cpx), d.stream));
    GD_TRY(fftn(d, buf, buf, ld, nd, dir, d.stream));
    GD_TRY(down(d, out, buf, (size_t)total * sizeof(cpx), d.stream));
    GD_CUDA(cudaStreamSynchronize(d.stream));
    return ::gd::GD_OK;
}

// fft.FFT2 over `width` devices of this process: row blocks in, all columns, all rows (fft/fft.go:138-151);
// both exchanges are this library's block-copy kernel storing into the peers' buffers over NVLink.
static int fft2_fanout(const double* in, double* out, int64_t rows, int64_t cols, int dir, int width) {
    const long long rg = rows / width, wc = cols / width;
    std::vector<cpx*> A((size_t)width, nullptr), B((size_t)width, nullptr);
    HostBarrier bar(width);
    std::atomic<int> failed{0};
    return fan_out(width, [&](int i) -> int {
        int rc = 0;
        // a device that fails must still meet the others at every barrier
        auto step = [&](Status s) { if (s != ::gd::GD_OK && !rc) { rc = (int)s; failed.store(1); } };
        DevLock L;
        if (L.st != ::gd::GD_OK) { failed.store(1); rc = (int)L.st; for (int k = 0; k < 4; k++) bar.wait(); return rc; }
        Device& d = *L.d;
        ScratchOrder order__(d, d.stream);
        for (int h = 0; h < width; h++)
            if (h != i) { cudaError_t e = cudaDeviceEnablePeerAccess(h, 0); if (e != cudaSuccess) cudaGetLastError(); }
        step(d.ensure_scratch(SCR_STAGE_IN, (size_t)rg * cols * sizeof(cpx), (void**)&A[(size_t)i]));
        step(d.ensure_scratch(SCR_STAGE_OUT, (size_t)rows * wc * sizeof(cpx), (void**)&B[(size_t)i]));
        if (!rc) step(up(d, A[(size_t)i], in + (size_t)i * rg * cols * 2, (size_t)rg * cols * sizeof(cpx), d.stream));
        cudaStreamSynchronize(d.stream);
        bar.wait();                                                   // every block uploaded, every buffer address known
        if (!failed.load())   // my columns [h*wc, (h+1)*wc) of my rows -> device h's column slab [rows][wc], rows [i*rg, (i+1)*rg)
            step(peer_block_copy(A[(size_t)i], B.data(), width, i, rg, wc, wc, cols, (long long)i * rg * wc, wc, d.stream));
        cudaStreamSynchronize(d.stream);
        bar.wait();
        if (!failed.load()) step(fft_strided(d, B[(size_t)i], B[(size_t)i], 1, rows, wc, dir, d.stream));          // every column
        if (!failed.load())   // rows [h*rg, (h+1)*rg) of my column slab -> device h's row block [rg][cols], columns [i*wc, (i+1)*wc)
            step(peer_block_copy(B[(size_t)i], A.data(), width, i, rg, wc, rg * wc, wc, (long long)i * wc, cols, d.stream));
        cudaStreamSynchronize(d.stream);
        bar.wait();
        if (!failed.load()) step(fft1d(d, A[(size_t)i], cols, A[(size_t)i], cols, cols, rg, false, dir, d.stream));   // every row
        if (!failed.load()) step(down(d, out + (size_t)i * rg * cols * 2, A[(size_t)i], (size_t)rg * cols * sizeof(cpx), d.stream));
        cudaStreamSynchronize(d.stream);
        bar.wait();
        if (!rc && failed.load()) { set_error("another device of the fan-out failed"); rc = (int)::gd::GD_ERR_CUDA; }
        return rc;
    });
}

int gd_fft2_c2c(const double* in, double* out, int64_t rows, int64_t cols, int dir) {
    const int width = fanout_width();
    if (in && out && width > 1 && (dir == 1 || dir == -1) && rows > 0 && cols > 0 && rows % width == 0 && cols % width == 0 &&
        (size_t)rows * (size_t)cols >= ((size_t)1 << 22) && cols / width < (1LL << 31) / 16) {
        bool peers = true;
        for (int a = 0; a < width && peers; a++)
            for (int b = 0; b < width && peers; b++)
                if (a != b) { int ok = 0; if (cudaDeviceCanAccessPeer(&ok, a, b) != cudaSuccess || !ok) peers = false; }
        if (peers) return fft2_fanout(in, out, rows, cols, dir, width);
    }
    int64_t dims[2] = {rows, cols};
    return gd_fftn_c2c(in, out, dims, 2, dir);
}

int gd_plan_warm(int64_t n) {
    if (n < 1) return (int)invalid_arg("plan_warm: n < 1");
    GD_ENTER();
    if (n == 1) return ::gd::GD_OK;
    if ((n & (n - 1)) == 0) {
        int lg = 0; while ((1LL << lg) < n) lg++;
        if (lg > 12 && lg <= 24) { TwiddleTable t; GD_TRY(d.twiddles(lg, &t)); }
    } else {
        const BluesteinPlan* pl;
        GD_TRY(d.bluestein(n, d.stream, &pl));
        GD_CUDA(cudaStreamSynchronize(d.stream));
    }
    return ::gd::GD_OK;
}

static int sample_size(int fmt) { return fmt == GD_SAMPLE_F64 ? 8 : fmt == GD_SAMPLE_F32 ? 4 : fmt == GD_SAMPLE_S16 ? 2 : fmt == GD_SAMPLE_U8 ? 1 : 0; }

// one device: segments 0 .. nsegs-1 of x. raw_host != NULL: the un-normalised per-bin sums go to the host (a share of a
// fan-out); otherwise pxx is finalised on the device with nsegs_total and norm.
static int pwelch_samples_one(const void* xv, int sample_fmt, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp,
                              int64_t nsegs, const double* win, double* raw_host, int64_t nsegs_total, double norm, double* pxx) {
    const size_t ssz = (size_t)sample_size(sample_fmt);
    const int64_t stride = nfft - noverlap;          // noverlap < 0 leaves gaps between segments, as spectral.Segment does
    const char* x = (const char*)xv;
    GD_ENTER();
    ScratchOrder order__(d, d.stream);
    GD_TRY(ensure_events(d.slot()));
    StageEvents& ev = g_ev[d.slot()];
    // stream the signal through two device buffers, a range of whole segments at a time; PCM formats travel as they are on
    // disk (1, 2 or 4 bytes per sample) and are decoded by the segment load of the kernel
    long long segs_per_chunk = std::max<long long>(1, (long long)(((256ull << 20) / ssz - (size_t)nfft) / (size_t)stride));
    if (segs_per_chunk > nsegs) segs_per_chunk = nsegs;
    // chunk starts stay 16-byte aligned for every format (the staged float64 path wants aligned ranges)
    while (segs_per_chunk > 1 && ((size_t)segs_per_chunk * stride * ssz) % 16) segs_per_chunk--;
    const size_t chunk_samples = ((size_t)(segs_per_chunk - 1) * stride + nfft + 15) & ~(size_t)15;
    char* dx;
    double* aux;
    GD_TRY(d.ensure_scratch(SCR_STAGE_IN, 2 * chunk_samples * ssz, (void**)&dx));
    GD_TRY(d.ensure_scratch(SCR_STAGE_OUT, ((size_t)fftlen + 3 * (size_t)lp) * sizeof(double), (void**)&aux));
    double *dwin = aux, *raw_tot = aux + fftlen, *raw_part = raw_tot + lp, *dpxx = raw_part + lp;
    GD_TRY(up(d, dwin, win, (size_t)fftlen * sizeof(double), d.stream));
    GD_CUDA(cudaMemsetAsync(raw_tot, 0, (size_t)lp * sizeof(double), d.stream));
    long long c = 0;
    for (long long s0 = 0; s0 < nsegs; s0 += segs_per_chunk, c++) {
        const int s = (int)(c & 1);
        const long long ns = std::min<long long>(segs_per_chunk, nsegs - s0);
        const size_t nsamp = (size_t)(ns - 1) * stride + nfft;
        char* dxi = dx + (size_t)s * chunk_samples * ssz;
        if (c >= 2) GD_CUDA(cudaStreamWaitEvent(d.stream_in, ev.comp[s], 0));
        GD_CUDA(cudaMemcpyAsync(dxi, x + (size_t)s0 * stride * ssz, nsamp * ssz, cudaMemcpyHostToDevice, d.stream_in));
        GD_CUDA(cudaEventRecord(ev.h2d[s], d.stream_in));
        GD_CUDA(cudaStreamWaitEvent(d.stream, ev.h2d[s], 0));
        GD_TRY(pwelch_partial(d, dxi, sample_fmt, nfft, stride, fftlen, lp, 0, ns, dwin, raw_part, d.stream));
        add_inplace_kernel<<<(unsigned)((lp + 127) / 128), 128, 0, d.stream>>>(raw_tot, raw_part, lp);
        g_launches++;
        GD_CUDA(cudaGetLastError());
        GD_CUDA(cudaEventRecord(ev.comp[s], d.stream));
    }
    if (raw_host) {
        GD_TRY(down(d, raw_host, raw_tot, (size_t)lp * sizeof(double), d.stream));
    } else {
        GD_TRY(pwelch_finalize(raw_tot, lp, nsegs_total, norm, dpxx, d.stream));
        GD_TRY(down(d, pxx, dpxx, (size_t)lp * sizeof(double), d.stream));
    }
    GD_CUDA(cudaStreamSynchronize(d.stream));
    return ::gd::GD_OK;
}

int gd_pwelch_samples(const void* xv, int sample_fmt, int64_t nx, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp,
                      int64_t nsegs, const double* win, double norm, double* pxx) {
    const size_t ssz = (size_t)sample_size(sample_fmt);
    if (!xv || !ssz || !win || !pxx || nfft < 1 || noverlap >= nfft || fftlen < nfft || lp < 1 || nsegs < 1)
        return (int)invalid_arg("pwelch: bad arguments");
    const int64_t stride = nfft - noverlap;
    if ((nsegs - 1) * stride + nfft > nx) return (int)invalid_arg("pwelch: x shorter than nsegs segments");
    const int width = fanout_width();
    if (width > 1 && nsegs >= 4 * (int64_t)width && (size_t)nsegs * (size_t)stride >= ((size_t)1 << 22)) {
        // independent segments: contiguous segment ranges, one per device; the partial sums (lp doubles each) are added in
        // device order on the host, which keeps runs reproducible, then scaled as pwelch.go:113-121,134-136 does
        std::vector<double> raw((size_t)width * (size_t)lp, 0.0);
        const int rc = fan_out(width, [&](int i) {
            const long long s0 = share_begin(nsegs, width, i), s1 = share_begin(nsegs, width, i + 1);
            if (s1 <= s0) return 0;
            return pwelch_samples_one((const char*)xv + (size_t)s0 * stride * ssz, sample_fmt, nfft, noverlap, fftlen, lp, s1 - s0, win,
                                      raw.data() + (size_t)i * lp, nsegs, norm, nullptr);
        });
        if (rc) return rc;
        for (int64_t j = 0; j < lp; j++) {
            double s = 0.0;
            for (int i = 0; i < width; i++) s += raw[(size_t)i * lp + j];
            double v = s / (double)nsegs;
            if (j > 0 && j < lp - 1) v *= 2.0;
            pxx[j] = v / norm;
        }
        return ::gd::GD_OK;
    }
    return pwelch_samples_one(xv, sample_fmt, nfft, noverlap, fftlen, lp, nsegs, win, nullptr, nsegs, norm, pxx);
}

int gd_pwelch_f64(const double* x, int64_t nx, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp,
                  int64_t nsegs, const double* win, double norm, double* pxx) {
    return gd_pwelch_samples(x, GD_SAMPLE_F64, nx, nfft, noverlap, fftlen, lp, nsegs, win, norm, pxx);
}

int gd_stft_f64(const double* x, int64_t nx, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp, int64_t nsegs,
                const double* win, double* out) {
    if (!x || !win || !out || nfft < 1 || noverlap >= nfft || fftlen < nfft || lp < 1 || lp > fftlen || nsegs < 1)
        return (int)invalid_arg("stft: bad arguments");
    const int64_t stride = nfft - noverlap;
    if ((nsegs - 1) * stride + nfft > nx) return (int)invalid_arg("stft: x shorter than nsegs segments");
    GD_ENTER();
    ScratchOrder order__(d, d.stream);
    // ranges of whole segments through one device buffer each way (the spectrogram is lp/stride times the signal)
    long long per = std::max<long long>(1, (long long)((128ull << 20) / ((size_t)std::max<int64_t>(stride, lp * 2) * sizeof(double))));
    if (per > nsegs) per = nsegs;
    const size_t in_samples = (size_t)(per - 1) * stride + nfft;
    double *dx, *dwin;
    cpx* dout;
    GD_TRY(d.ensure_scratch(SCR_STAGE_IN, (in_samples + (size_t)nfft) * sizeof(double), (void**)&dx));
    GD_TRY(d.ensure_scratch(SCR_STAGE_OUT, (size_t)per * lp * sizeof(cpx), (void**)&dout));
    dwin = dx + in_samples;
    GD_TRY(up(d, dwin, win, (size_t)nfft * sizeof(double), d.stream));
    for (long long s0 = 0; s0 < nsegs; s0 += per) {
        const long long ns = std::min<long long>(per, nsegs - s0);
        GD_TRY(up(d, dx, x + (size_t)s0 * stride, ((size_t)(ns - 1) * stride + nfft) * sizeof(double), d.stream));
        GD_TRY(stft(d, dx, nfft, stride, fftlen, lp, 0, ns, dwin, dout, d.stream));
        GD_TRY(down(d, out + (size_t)s0 * lp * 2, dout, (size_t)ns * lp * sizeof(cpx), d.stream));
    }
    GD_CUDA(cudaStreamSynchronize(d.stream));
    return ::gd::GD_OK;
}

int gd_fft_segments_c2c(const double* x, int64_t nx, int64_t seg_len, int64_t step, int64_t segs, int64_t fftlen, double* out) {
    if (!x || !out || seg_len < 1 || step < 1 || segs < 1 || fftlen < seg_len || (fftlen & (fftlen - 1)) || fftlen > (1LL << 24) ||
        (segs - 1) * step + seg_len > nx)
        return (int)invalid_arg("fft_segments: bad arguments (fftlen must be a power of two >= seg_len, segments inside x)");
    GD_ENTER();
    ScratchOrder order__(d, d.stream);
    const size_t used = (size_t)(segs - 1) * step + seg_len;
    cpx *dx, *dout;
    GD_TRY(d.ensure_scratch(SCR_STAGE_IN, used * sizeof(cpx), (void**)&dx));
    GD_TRY(d.ensure_scratch(SCR_STAGE_OUT, (size_t)segs * fftlen * sizeof(cpx), (void**)&dout));
    GD_TRY(up(d, dx, x, used * sizeof(cpx), d.stream));
    if (fftlen == 1) {
        GD_CUDA(cudaMemcpy2DAsync(dout, sizeof(cpx), dx, (size_t)step * sizeof(cpx), sizeof(cpx), (size_t)segs, cudaMemcpyDeviceToDevice, d.stream));
    } else {
        FusedOps ops;                     // dsputils.Segment slices (offset s*step, length seg_len) zero-padded as dsputils.ZeroPad2 does
        ops.ld_flags = LD_PAD; ops.n_valid_in = seg_len;
        int lg = 0;
        while ((1LL << lg) < fftlen) lg++;
        GD_TRY(fft_pow2(d, dx, step, dout, fftlen, lg, segs, ops, d.stream));
    }
    GD_TRY(down(d, out, dout, (size_t)segs * fftlen * sizeof(cpx), d.stream));
    GD_CUDA(cudaStreamSynchronize(d.stream));
    return ::gd::GD_OK;
}

int gd_convolve_linear_c2c(const double* x, int64_t nx, const double* h, int64_t nh, double* out) {
    if (!x || !h || !out || nx < 1 || nh < 1) return (int)invalid_arg("convolve_linear: bad arguments");
    GD_ENTER();
    ScratchOrder order__(d, d.stream);
    const size_t nout = (size_t)(nx + nh - 1);
    cpx *din, *dout;
    GD_TRY(d.ensure_scratch(SCR_STAGE_IN, (size_t)(nx + nh) * sizeof(cpx), (void**)&din));
    GD_TRY(d.ensure_scratch(SCR_STAGE_OUT, nout * sizeof(cpx), (void**)&dout));
    GD_TRY(up(d, din, x, (size_t)nx * sizeof(cpx), d.stream));
    GD_TRY(up(d, din + nx, h, (size_t)nh * sizeof(cpx), d.stream));
    GD_TRY(convolve_linear(d, din, nx, din + nx, nh, dout, d.stream));
    GD_TRY(down(d, out, dout, nout * sizeof(cpx), d.stream));
    GD_CUDA(cudaStreamSynchronize(d.stream));
    return ::gd::GD_OK;
}

// ---------------------------------------------------------------- streaming Pwelch
namespace {
struct PwStream {
    int dev = 0, fmt = 0;
    int64_t nfft = 0, stride = 0, fftlen = 0, lp = 0, nsegs = 0, carry = 0, skip = 0;
    size_t cap = 0;                 // samples the device buffer holds
    char* dx = nullptr;             // [carry | chunk]
    double *dwin = nullptr, *raw_tot = nullptr, *raw_part = nullptr;
};
}  // namespace

int gd_pwelch_stream_begin(void** handle, int sample_fmt, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp, const double* win) {
    if (!handle || !sample_size(sample_fmt) || !win || nfft < 1 || noverlap >= nfft || fftlen < nfft || lp < 1 || lp > fftlen / 2 + 1)
        return (int)invalid_arg("pwelch_stream_begin: bad arguments");
    GD_ENTER();
    std::unique_ptr<PwStream> s(new PwStream);
    s->dev = d.dev; s->fmt = sample_fmt; s->nfft = nfft; s->stride = nfft - noverlap; s->fftlen = fftlen; s->lp = lp;
    GD_CUDA(cudaMalloc((void**)&s->dwin, ((size_t)fftlen + 2 * (size_t)lp) * sizeof(double)));
    s->raw_tot = s->dwin + fftlen; s->raw_part = s->raw_tot + lp;
    GD_CUDA(cudaMemcpyAsync(s->dwin, win, (size_t)fftlen * sizeof(double), cudaMemcpyHostToDevice, d.stream));
    GD_CUDA(cudaMemsetAsync(s->raw_tot, 0, (size_t)lp * sizeof(double), d.stream));
    GD_CUDA(cudaStreamSynchronize(d.stream));
    *handle = s.release();
    return ::gd::GD_OK;
}

int gd_pwelch_stream_push(void* handle, const void* samples, int64_t n) {
    PwStream* s = (PwStream*)handle;
    if (!s || (!samples && n > 0) || n < 0) return (int)invalid_arg("pwelch_stream_push: bad arguments");
    if (n == 0) return ::gd::GD_OK;
    GD_ENTER();
    if (d.dev != s->dev) return (int)invalid_arg("pwelch_stream_push: the stream belongs to another device");
    ScratchOrder order__(d, d.stream);
    const size_t ssz = (size_t)sample_size(s->fmt);
    const char* src = (const char*)samples;
    // gapped segments (noverlap < 0): samples between two segments are dropped as they arrive
    if (s->skip > 0) { const int64_t k = std::min<int64_t>(s->skip, n); s->skip -= k; src += (size_t)k * ssz; n -= k; if (n == 0) return ::gd::GD_OK; }
    const size_t need = (size_t)s->carry + (size_t)n;
    if (need > s->cap) {
        char* nb = nullptr;
        const size_t cap = std::max<size_t>(need, (size_t)4 << 20);
        GD_CUDA(cudaMalloc((void**)&nb, cap * ssz));
        if (s->carry) GD_CUDA(cudaMemcpyAsync(nb, s->dx, (size_t)s->carry * ssz, cudaMemcpyDeviceToDevice, d.stream));
        GD_CUDA(cudaStreamSynchronize(d.stream));
        if (s->dx) cudaFree(s->dx);
        s->dx = nb; s->cap = cap;
    }
    GD_CUDA(cudaMemcpyAsync(s->dx + (size_t)s->carry * ssz, src, (size_t)n * ssz, cudaMemcpyHostToDevice, d.stream));
    const int64_t have = s->carry + n;
    int64_t ns = have >= s->nfft ? (have - s->nfft) / s->stride + 1 : 0;      // whole segments in [carry | chunk]
    if (ns > 0) {
        GD_TRY(pwelch_partial(d, s->dx, s->fmt, s->nfft, s->stride, s->fftlen, s->lp, 0, ns, s->dwin, s->raw_part, d.stream));
        add_inplace_kernel<<<(unsigned)((s->lp + 127) / 128), 128, 0, d.stream>>>(s->raw_tot, s->raw_part, s->lp);
        g_launches++;
        GD_CUDA(cudaGetLastError());
        s->nsegs += ns;
    }
    // keep what the next segment still needs: samples from ns * stride on
    const int64_t used = ns * s->stride;
    if (used >= have) { s->skip = used - have; s->carry = 0; }
    else {
        const int64_t keep = have - used;
        if (used > 0) {
            // the ranges may overlap: go through the partial-sum buffer's neighbour only when small, else a staged copy
            char* tmp = nullptr;
            GD_CUDA(cudaMalloc((void**)&tmp, (size_t)keep * ssz));
            GD_CUDA(cudaMemcpyAsync(tmp, s->dx + (size_t)used * ssz, (size_t)keep * ssz, cudaMemcpyDeviceToDevice, d.stream));
            GD_CUDA(cudaMemcpyAsync(s->dx, tmp, (size_t)keep * ssz, cudaMemcpyDeviceToDevice, d.stream));
            GD_CUDA(cudaStreamSynchronize(d.stream));
            cudaFree(tmp);
        }
        s->carry = keep;
    }
    GD_CUDA(cudaStreamSynchronize(d.stream));          // the caller may reuse `samples`
    return ::gd::GD_OK;
}

int gd_pwelch_stream_end(void* handle, double norm, double* pxx, int64_t* nsegs_out) {
    PwStream* s = (PwStream*)handle;
    if (!s) return (int)invalid_arg("pwelch_stream_end: null handle");
    GD_ENTER();
    Status rc = ::gd::GD_OK;
    if (nsegs_out) *nsegs_out = s->nsegs;
    if (pxx && s->nsegs > 0) {
        double* dp = s->raw_part;
        rc = pwelch_finalize(s->raw_tot, s->lp, s->nsegs, norm, dp, d.stream);
        if (rc == ::gd::GD_OK && cudaMemcpyAsync(pxx, dp, (size_t)s->lp * sizeof(double), cudaMemcpyDeviceToHost, d.stream) != cudaSuccess) rc = ::gd::GD_ERR_CUDA;
    }
    cudaStreamSynchronize(d.stream);
    if (s->dx) cudaFree(s->dx);
    if (s->dwin) cudaFree(s->dwin);
    delete s;
    return (int)rc;
}

void* gd_pinned_alloc(size_t bytes) {
    if (ensure_device(t_dev) != ::gd::GD_OK) return nullptr;
    void* p = nullptr;
    NearDevice near(t_dev);                  // page-locked pages are taken from the NUMA node of the calling thread's GPU
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { cuda_fail(e, "cudaHostAlloc"); return nullptr; }
    return p;
}
void gd_pinned_free(void* p) { if (p) cudaFreeHost(p); }

// ---------------------------------------------------------------- device-resident API

int gd_dev_alloc(void** p, size_t bytes) {
    if (!p) return (int)invalid_arg("gd_dev_alloc: null");
    GD_ENTER();
    (void)d;
    GD_CUDA(cudaMalloc(p, bytes));
    return ::gd::GD_OK;
}
int gd_dev_free(void* p) {
    GD_ENTER();
    (void)d;
    GD_CUDA(cudaFree(p));
    return ::gd::GD_OK;
}
int gd_memcpy_h2d(void* dst, const void* src, size_t bytes) {
    GD_ENTER();
    GD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, d.stream));
    GD_CUDA(cudaStreamSynchronize(d.stream));
    return ::gd::GD_OK;
}
int gd_memcpy_d2h(void* dst, const void* src, size_t bytes) {
    GD_ENTER();
    GD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, d.stream));
    GD_CUDA(cudaStreamSynchronize(d.stream));
    return ::gd::GD_OK;
}
int gd_stream_sync(void* stream) {
    GD_ENTER();
    GD_CUDA(cudaStreamSynchronize(pick(d, stream)));
    return ::gd::GD_OK;
}
int gd_fill_splitmix_dev(double* dst, int64_t n, uint64_t seed, uint64_t offset, void* stream) {
    GD_ENTER();
    return (int)fill_splitmix(dst, n, seed, offset, pick(d, stream));
}
int gd_fft_batch_c2c_dev(const double* in, double* out, int64_t n, int64_t batch, int dir, void* stream) {
    if (!in || !out || n < 1 || batch < 1 || (dir != 1 && dir != -1)) return (int)invalid_arg("fft_dev: bad arguments");
    GD_ENTER();
    ScratchOrder order__(d, pick(d, stream));
    return (int)fft1d(d, in, n, (cpx*)out, n, n, batch, false, dir, pick(d, stream));
}
int gd_fft_batch_r2c_full_dev(const double* in, double* out, int64_t n, int64_t batch, int dir, void* stream) {
    if (!in || !out || n < 1 || batch < 1 || (dir != 1 && dir != -1)) return (int)invalid_arg("fft_dev: bad arguments");
    GD_ENTER();
    ScratchOrder order__(d, pick(d, stream));
    return (int)fft1d(d, in, n, (cpx*)out, n, n, batch, true, dir, pick(d, stream));
}
int gd_convolve_c2c_dev(const double* x, const double* y, double* out, int64_t n, void* stream) {
    if (!x || !y || !out || n < 1) return (int)invalid_arg("convolve_dev: bad arguments");
    GD_ENTER();
    ScratchOrder order__(d, pick(d, stream));
    return (int)convolve(d, (const cpx*)x, (const cpx*)y, (cpx*)out, n, pick(d, stream));
}
int gd_fftn_c2c_dev(const double* in, double* out, const int64_t* dims, int nd, int dir, void* stream) {
    if (!in || !out || !dims || nd < 1 || nd > 16 || (dir != 1 && dir != -1)) return (int)invalid_arg("fftn_dev: bad arguments");
    long long ld[16];
    for (int i = 0; i < nd; i++) ld[i] = dims[i];
    GD_ENTER();
    ScratchOrder order__(d, pick(d, stream));
    return (int)fftn(d, (const cpx*)in, (cpx*)out, ld, nd, dir, pick(d, stream));
}
int gd_fourstep_twiddle_dev(double* blk, int64_t rows, int64_t cols, int64_t row0, int64_t col0, int log2n, void* stream) {
    if (!blk) return (int)invalid_arg("fourstep_twiddle_dev: null");
    GD_ENTER();
    return (int)fourstep_twiddle((cpx*)blk, rows, cols, row0, col0, log2n, pick(d, stream));
}
int gd_repack_gkw_dev(const double* in, double* out, int64_t g, int64_t k, int64_t w, void* stream) {
    if (!in || !out) return (int)invalid_arg("repack_gkw_dev: null");
    GD_ENTER();
    return (int)repack_gkw((const cpx*)in, (cpx*)out, g, k, w, pick(d, stream));
}
int gd_ipc_alloc(void** p, size_t bytes, unsigned char* handle64) {
    if (!p || !handle64 || bytes == 0) return (int)invalid_arg("gd_ipc_alloc: bad arguments");
    GD_ENTER();
    (void)d;
    GD_CUDA(cudaMalloc(p, bytes));
    cudaIpcMemHandle_t hnd;
    static_assert(sizeof(hnd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    GD_CUDA(cudaIpcGetMemHandle(&hnd, *p));
    memcpy(handle64, &hnd, 64);
    return ::gd::GD_OK;
}
int gd_ipc_open(const unsigned char* handle64, void** p) {
    if (!p || !handle64) return (int)invalid_arg("gd_ipc_open: bad arguments");
    GD_ENTER();
    (void)d;
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, handle64, 64);
    GD_CUDA(cudaIpcOpenMemHandle(p, hnd, cudaIpcMemLazyEnablePeerAccess));
    return ::gd::GD_OK;
}
int gd_ipc_close(void* p) {
    GD_ENTER();
    (void)d;
    GD_CUDA(cudaIpcCloseMemHandle(p));
    return ::gd::GD_OK;
}
int gd_fourstep_exchange_dev(const double* slab, void* const* peer_recv, int64_t n1, int64_t w, int rank, int world, int log2n,
                             void* stream) {
    if (!slab || !peer_recv) return (int)invalid_arg("fourstep_exchange_dev: null");
    GD_ENTER();
    return (int)fourstep_exchange((const cpx*)slab, (cpx* const*)peer_recv, n1, w, rank, world, log2n, pick(d, stream));
}
int gd_fourstep_lines_exchange_dev(const double* slab, double* tmp, void* const* peer_recv, int64_t n1, int64_t w, int rank, int world,
                                   int log2n, void* stream) {
    if (!slab || !tmp || !peer_recv) return (int)invalid_arg("fourstep_lines_exchange_dev: null");
    GD_ENTER();
    ScratchOrder order__(d, pick(d, stream));
    return (int)fourstep_lines_exchange(d, (const cpx*)slab, (cpx*)tmp, (cpx* const*)peer_recv, n1, w, rank, world, log2n, pick(d, stream));
}
int gd_fourstep_fused_supported(int64_t n1, int64_t n2, int world) {
    GD_ENTER();
    return fourstep_fused_supported(d, n1, n2, world) ? 1 : 0;
}
int gd_fourstep_lines_peer_dev(const double* slab, void* const* peer_recv, int64_t n1, int64_t w, int rank, int world, int log2n, int dir,
                               void* stream) {
    if (!slab || !peer_recv || (dir != 1 && dir != -1)) return (int)invalid_arg("fourstep_lines_peer_dev: null");
    GD_ENTER();
    ScratchOrder order__(d, pick(d, stream));
    return (int)fourstep_lines_peer(d, (const cpx*)slab, (cpx* const*)peer_recv, n1, w, rank, world, log2n, dir, pick(d, stream));
}
int gd_fourstep_rows_seg_dev(const double* recv, double* out, int64_t n2, int64_t k, int world, int dir, void* stream) {
    if (!recv || !out || (dir != 1 && dir != -1)) return (int)invalid_arg("fourstep_rows_seg_dev: null");
    GD_ENTER();
    ScratchOrder order__(d, pick(d, stream));
    return (int)fourstep_rows_seg(d, (const cpx*)recv, (cpx*)out, n2, k, world, dir, pick(d, stream));
}
int gd_peer_block_copy_dev(const double* src, void* const* peers, int world, int rank, int64_t rows, int64_t cols, int64_t src_step,
                           int64_t src_pitch, int64_t dst_off, int64_t dst_pitch, void* stream) {
    if (!src || !peers) return (int)invalid_arg("peer_block_copy_dev: null");
    GD_ENTER();
    return (int)peer_block_copy((const cpx*)src, (cpx* const*)peers, world, rank, rows, cols, src_step, src_pitch, dst_off, dst_pitch, pick(d, stream));
}
int gd_transpose_batched_dev(const double* in, double* out, int64_t batch, int64_t rows, int64_t cols, void* stream) {
    if (!in || !out) return (int)invalid_arg("transpose_batched_dev: null");
    GD_ENTER();
    return (int)transpose_batched((const cpx*)in, (cpx*)out, batch, rows, cols, pick(d, stream));
}
int gd_fft_strided_c2c_dev(const double* in, double* out, int64_t outer, int64_t len, int64_t stride, int dir, void* stream) {
    if (!in || !out || (dir != 1 && dir != -1)) return (int)invalid_arg("fft_strided_dev: bad arguments");
    GD_ENTER();
    ScratchOrder order__(d, pick(d, stream));
    return (int)fft_strided(d, (const cpx*)in, (cpx*)out, outer, len, stride, dir, pick(d, stream));
}
int gd_pwelch_partial_samples_dev(const void* x, int sample_fmt, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp, int64_t seg0,
                                  int64_t nseg, const double* win, double* raw, void* stream) {
    if (!x || !sample_size(sample_fmt) || !win || !raw || noverlap >= nfft) return (int)invalid_arg("pwelch_dev: bad arguments");
    GD_ENTER();
    ScratchOrder order__(d, pick(d, stream));
    return (int)pwelch_partial(d, x, sample_fmt, nfft, nfft - noverlap, fftlen, lp, seg0, nseg, win, raw, pick(d, stream));
}
int gd_pwelch_partial_dev(const double* x, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp, int64_t seg0,
                          int64_t nseg, const double* win, double* raw, void* stream) {
    return gd_pwelch_partial_samples_dev(x, GD_SAMPLE_F64, nfft, noverlap, fftlen, lp, seg0, nseg, win, raw, stream);
}
int gd_stft_f64_dev(const double* x, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp, int64_t seg0, int64_t nseg,
                    const double* win, double* out, void* stream) {
    if (!x || !win || !out || noverlap >= nfft) return (int)invalid_arg("stft_dev: bad arguments");
    GD_ENTER();
    ScratchOrder order__(d, pick(d, stream));
    return (int)stft(d, x, nfft, nfft - noverlap, fftlen, lp, seg0, nseg, win, (cpx*)out, pick(d, stream));
}
int gd_convolve_linear_c2c_dev(const double* x, int64_t nx, const double* h, int64_t nh, double* out, void* stream) {
    if (!x || !h || !out) return (int)invalid_arg("convolve_linear_dev: bad arguments");
    GD_ENTER();
    ScratchOrder order__(d, pick(d, stream));
    return (int)convolve_linear(d, (const cpx*)x, nx, (const cpx*)h, nh, (cpx*)out, pick(d, stream));
}
int gd_pwelch_finalize_dev(const double* raw, int64_t lp, int64_t nsegs, double norm, double* pxx, void* stream) {
    if (!raw || !pxx) return (int)invalid_arg("pwelch_finalize_dev: bad arguments");
    GD_ENTER();
    return (int)pwelch_finalize(raw, lp, nsegs, norm, pxx, pick(d, stream));
}

}  // extern "C"

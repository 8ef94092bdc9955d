// capi.cu -- the C ABI (include/godsp_b200.h): device contexts, staging, error convention.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <memory>

#include "../../include/godsp_b200.h"
#include "engine.h"

using namespace gd;

namespace {

constexpr int kMaxDev = 16;
Device g_dev[kMaxDev];
std::atomic<int> g_ndev{0};
std::mutex g_init_mu;
thread_local int t_dev = 0;

struct StageEvents {
    cudaEvent_t h2d[2] = {nullptr, nullptr}, comp[2] = {nullptr, nullptr}, d2h[2] = {nullptr, nullptr};
    bool ready = false;
};
StageEvents g_ev[kMaxDev];

Status ensure_events(int dev) {
    StageEvents& e = g_ev[dev];
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
    if (!g_dev[dev].ready) GD_TRY(g_dev[dev].init(dev));
    return ::gd::GD_OK;
}

// RAII: select + lock the calling thread's device
struct DevLock {
    Device* d = nullptr;
    Status st = ::gd::GD_OK;
    std::unique_lock<std::recursive_mutex> lk;
    DevLock() {
        st = ensure_device(t_dev);
        if (st != ::gd::GD_OK) return;
        d = &g_dev[t_dev];
        lk = std::unique_lock<std::recursive_mutex>(d->mu);
        cudaError_t e = cudaSetDevice(d->dev);
        if (e != cudaSuccess) st = cuda_fail(e, "cudaSetDevice");
    }
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
    return ::gd::GD_OK;
}

int gd_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    for (int i = 0; i < g_ndev.load(); i++) {
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
    for (int i = 0; i < g_ndev.load(); i++) n += g_dev[i].ready ? 1 : 0;
    return n;
}

int gd_use_device(int dev) {
    Status s = ensure_device(dev);
    if (s != ::gd::GD_OK) return (int)s;
    t_dev = dev;
    return ::gd::GD_OK;
}

int gd_set_option(const char* key, int64_t value) {
    GD_ENTER();
    if (!strcmp(key, "pass_scratch_mb")) { if (value < 1) return (int)invalid_arg("pass_scratch_mb < 1"); d.pass_scratch_budget = (size_t)value << 20; }
    else if (!strcmp(key, "l2_block_mb")) { if (value < 1) return (int)invalid_arg("l2_block_mb < 1"); d.l2_block_budget = (size_t)value << 20; }
    else if (!strcmp(key, "two_stream_chunks")) d.two_stream_chunks = value != 0;
    else if (!strcmp(key, "wide_tiles")) d.wide_tiles = value != 0;
    else if (!strcmp(key, "fused")) d.use_fused = value != 0;
    else if (!strcmp(key, "debug_alias")) d.debug_alias = value != 0;
    else if (!strcmp(key, "w32")) { if (value < 0 || value > 6) return (int)invalid_arg("w32 out of range"); d.w32 = (int)value; }
    else if (!strcmp(key, "tma")) d.use_tma = value != 0;
    else if (!strcmp(key, "tma_opt")) d.tma_opt = (int)value;
    else if (!strcmp(key, "tma_prof")) d.tma_prof = value != 0;
    else if (!strcmp(key, "tma_delay")) { if (value < 0 || value > 4) return (int)invalid_arg("tma_delay out of range"); d.tma_delay = (int)value; }
    else if (!strcmp(key, "tma_slots")) { if (value < 2 || value > 6) return (int)invalid_arg("tma_slots out of range"); d.tma_slots = (int)value; }
    else if (!strcmp(key, "tiled_scratch")) d.tiled_scratch = value != 0;
    else if (!strcmp(key, "l2_window")) d.use_l2_window = value != 0;
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

static int fft_host(const double* in, double* out, int64_t n, int64_t batch, bool real_in, int dir) {
    if (!in || !out || n < 1 || batch < 1 || (dir != 1 && dir != -1)) return (int)invalid_arg("fft: bad arguments");
    GD_ENTER();
    ScratchOrder order__(d, d.stream);
    GD_TRY(ensure_events(d.dev));
    StageEvents& ev = g_ev[d.dev];
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

int gd_fft2_c2c(const double* in, double* out, int64_t rows, int64_t cols, int dir) {
    int64_t dims[2] = {rows, cols};
    return gd_fftn_c2c(in, out, dims, 2, dir);
}

int gd_plan_warm(int64_t n) {
    if (n < 1) return (int)invalid_arg("plan_warm: n < 1");
    GD_ENTER();
    if (n == 1) return ::gd::GD_OK;
    if ((n & (n - 1)) == 0) {
        int lg = 0; while ((1LL << lg) < n) lg++;
        if (lg > 12) { TwiddleTable t; GD_TRY(d.twiddles(lg, &t)); }
        if (lg > 24) return (int)invalid_arg("plan_warm: n > 2^24");
    } else {
        const BluesteinPlan* pl;
        GD_TRY(d.bluestein(n, d.stream, &pl));
        GD_CUDA(cudaStreamSynchronize(d.stream));
    }
    return ::gd::GD_OK;
}

int gd_pwelch_f64(const double* x, int64_t nx, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp,
                  int64_t nsegs, const double* win, double norm, double* pxx) {
    if (!x || !win || !pxx || nfft < 1 || noverlap < 0 || noverlap >= nfft || fftlen < nfft || lp < 1 || nsegs < 1)
        return (int)invalid_arg("pwelch: bad arguments");
    const int64_t stride = nfft - noverlap;
    if ((nsegs - 1) * stride + nfft > nx) return (int)invalid_arg("pwelch: x shorter than nsegs segments");
    GD_ENTER();
    ScratchOrder order__(d, d.stream);
    GD_TRY(ensure_events(d.dev));
    StageEvents& ev = g_ev[d.dev];
    // stream the signal through two device buffers, a range of whole segments at a time
    long long segs_per_chunk = std::max<long long>(1, (long long)(((256ull << 20) / sizeof(double) - (size_t)nfft) / (size_t)stride));
    if (segs_per_chunk > nsegs) segs_per_chunk = nsegs;
    const size_t chunk_samples = (size_t)(segs_per_chunk - 1) * stride + nfft;
    double *dx, *aux;
    GD_TRY(d.ensure_scratch(SCR_STAGE_IN, 2 * chunk_samples * sizeof(double), (void**)&dx));
    GD_TRY(d.ensure_scratch(SCR_STAGE_OUT, ((size_t)fftlen + 3 * (size_t)lp) * sizeof(double), (void**)&aux));
    double *dwin = aux, *raw_tot = aux + fftlen, *raw_part = raw_tot + lp, *dpxx = raw_part + lp;
    GD_TRY(up(d, dwin, win, (size_t)fftlen * sizeof(double), d.stream));
    GD_CUDA(cudaMemsetAsync(raw_tot, 0, (size_t)lp * sizeof(double), d.stream));
    long long c = 0;
    for (long long s0 = 0; s0 < nsegs; s0 += segs_per_chunk, c++) {
        const int s = (int)(c & 1);
        const long long ns = std::min<long long>(segs_per_chunk, nsegs - s0);
        const size_t nsamp = (size_t)(ns - 1) * stride + nfft;
        double* dxi = dx + (size_t)s * chunk_samples;
        if (c >= 2) GD_CUDA(cudaStreamWaitEvent(d.stream_in, ev.comp[s], 0));
        GD_CUDA(cudaMemcpyAsync(dxi, x + (size_t)s0 * stride, nsamp * sizeof(double), cudaMemcpyHostToDevice, d.stream_in));
        GD_CUDA(cudaEventRecord(ev.h2d[s], d.stream_in));
        GD_CUDA(cudaStreamWaitEvent(d.stream, ev.h2d[s], 0));
        GD_TRY(pwelch_partial(d, dxi, nfft, stride, fftlen, lp, 0, ns, dwin, raw_part, d.stream));
        add_inplace_kernel<<<(unsigned)((lp + 127) / 128), 128, 0, d.stream>>>(raw_tot, raw_part, lp);
        g_launches++;
        GD_CUDA(cudaGetLastError());
        GD_CUDA(cudaEventRecord(ev.comp[s], d.stream));
    }
    GD_TRY(pwelch_finalize(raw_tot, lp, nsegs, norm, dpxx, d.stream));
    GD_TRY(down(d, pxx, dpxx, (size_t)lp * sizeof(double), d.stream));
    GD_CUDA(cudaStreamSynchronize(d.stream));
    return ::gd::GD_OK;
}

void* gd_pinned_alloc(size_t bytes) {
    if (ensure_device(t_dev) != ::gd::GD_OK) return nullptr;
    void* p = nullptr;
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
int gd_pwelch_partial_dev(const double* x, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp, int64_t seg0,
                          int64_t nseg, const double* win, double* raw, void* stream) {
    if (!x || !win || !raw || noverlap < 0 || noverlap >= nfft) return (int)invalid_arg("pwelch_dev: bad arguments");
    GD_ENTER();
    ScratchOrder order__(d, pick(d, stream));
    return (int)pwelch_partial(d, x, nfft, nfft - noverlap, fftlen, lp, seg0, nseg, win, raw, pick(d, stream));
}
int gd_pwelch_finalize_dev(const double* raw, int64_t lp, int64_t nsegs, double norm, double* pxx, void* stream) {
    if (!raw || !pxx) return (int)invalid_arg("pwelch_finalize_dev: bad arguments");
    GD_ENTER();
    return (int)pwelch_finalize(raw, lp, nsegs, norm, pxx, pick(d, stream));
}

}  // extern "C"

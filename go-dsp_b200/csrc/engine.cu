// engine.cu -- planner / executor: turns go-dsp's transforms into sequences of passes.
#include "engine.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fft_fused.cuh"
#include "pass_launch.cuh"
#include "bluestein_small.cuh"

namespace gd {

cudaError_t launch_fused(int log2l, bool wide, const FusedParams& a, long long total_items, int num_sms, cudaStream_t st);
int fused_tile_lines(int log2l, bool wide);
cudaError_t launch_pass32(int variant, const PassParams& a, int num_sms, cudaStream_t st);
cudaError_t launch_bluestein_small(int log2la, const BluesteinSmallParams& a, int num_sms, cudaStream_t st);
bool tma_fused_applicable(const void* in, long long in_dist, const cpx* out, long long out_dist, int ld_conj, int st_conj, double scale);
Status fft_tma_2p20(Device& d, const cpx* in, long long in_dist, cpx* out, long long out_dist, long long batch, int ld_conj,
                    int st_conj, double scale, cudaStream_t st);
int pass32_tile_lines(int variant);
// fused TMA four-step for 2^13 .. 2^18 points (fft_tma14.cuh; one translation unit per size)
#define GD_TMA2D_DECL(LG)                                                                                                                \
    bool tma##LG##_rows_applicable(const void* in, long long in_dist, const cpx* out, long long out_dist, long long batch, int ld_conj,  \
                                   int st_conj, double scale);                                                                           \
    bool tma##LG##_cols_applicable(const cpx* src, const cpx* dst, long long len, long long ncols, long long pitch);                     \
    Status fft_tma_2p##LG(Device& d, int mode, const cpx* in, long long in_dist, cpx* out, long long out_dist, long long count, bool inv, \
                          double scale, cudaStream_t st, const Tma2dExtra& ex);
GD_TMA2D_DECL(13) GD_TMA2D_DECL(14) GD_TMA2D_DECL(15) GD_TMA2D_DECL(16) GD_TMA2D_DECL(17) GD_TMA2D_DECL(18) GD_TMA2D_DECL(19)
struct Tma2dEntry {
    bool (*rows_ok)(const void*, long long, const cpx*, long long, long long, int, int, double);
    bool (*cols_ok)(const cpx*, const cpx*, long long, long long, long long);
    Status (*run)(Device&, int, const cpx*, long long, cpx*, long long, long long, bool, double, cudaStream_t, const Tma2dExtra&);
    int unit;                                            // transforms / columns per phase: 2^20 / N
};
static const Tma2dEntry* tma2d_entry(const Device& d, int log2n) {
    static const Tma2dEntry tab[7] = {
        {tma13_rows_applicable, tma13_cols_applicable, fft_tma_2p13, 128}, {tma14_rows_applicable, tma14_cols_applicable, fft_tma_2p14, 64},
        {tma15_rows_applicable, tma15_cols_applicable, fft_tma_2p15, 32},  {tma16_rows_applicable, tma16_cols_applicable, fft_tma_2p16, 16},
        {tma17_rows_applicable, tma17_cols_applicable, fft_tma_2p17, 8},   {tma18_rows_applicable, tma18_cols_applicable, fft_tma_2p18, 4},
        {tma19_rows_applicable, tma19_cols_applicable, fft_tma_2p19, 2}};
    if (!d.use_tma || log2n < 13 || log2n > 19 || (log2n == 19 && !d.use_tma19)) return nullptr;
    if (log2n == 14 ? !d.use_tma14 : !d.use_tma16) return nullptr;       // "tma14": the 2^14 kernel; "tma16": every other size of the family
    return &tab[log2n - 13];
}

// ------------------------------------------------------------------ errors
std::atomic<long long> g_launches{0};
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error() { return g_err.c_str(); }
Status cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return e == cudaErrorMemoryAllocation ? GD_ERR_NOMEM : GD_ERR_CUDA;
}
static Status invalid(const char* msg) { g_err = msg; return GD_ERR_INVALID; }

static inline int ilog2ll(long long v) { int r = 0; while (v > 1) { v >>= 1; r++; } return r; }
static inline bool is_pow2(long long x) { return (x & (x - 1)) == 0; }

// ------------------------------------------------------------------ tables
// exp(-2 pi i e / M) for e = e0, e0+step, ... (count entries), long-double accurate, exact on the axes
static void host_twiddles(std::vector<cpx>& out, long long M, long long step, long long count) {
    out.resize((size_t)count);
    for (long long j = 0; j < count; j++) {
        long long e = (j * step) % M;
        long double c, s;
        if (e == 0) { c = 1; s = 0; }
        else if (4 * e == M) { c = 0; s = 1; }
        else if (2 * e == M) { c = -1; s = 0; }
        else if (4 * e == 3 * M) { c = 0; s = -1; }
        else {
            long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)e / (long double)M;
            c = cosl(a); s = sinl(a);
        }
        out[(size_t)j] = make_double2((double)c, (double)(-s));
    }
}

static Status upload(const std::vector<cpx>& h, cpx** dptr) {
    GD_CUDA(cudaMalloc((void**)dptr, h.size() * sizeof(cpx)));
    GD_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(cpx), cudaMemcpyHostToDevice));
    return GD_OK;
}

Status Device::init(int device, int lane_index) {
    std::lock_guard<std::recursive_mutex> lk(mu);
    if (ready) return GD_OK;
    dev = device;
    lane = lane_index;
    GD_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    GD_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) {
        set_error("go-dsp_b200 is built for sm_100a (B200) only; device '" + std::string(prop.name) + "' is sm_" +
                  std::to_string(prop.major) + std::to_string(prop.minor));
        return GD_ERR_UNSUPPORTED;
    }
    num_sms = prop.multiProcessorCount;
    l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
    l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    GD_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    GD_CUDA(cudaStreamCreateWithFlags(&stream_in, cudaStreamNonBlocking));
    GD_CUDA(cudaStreamCreateWithFlags(&stream_out, cudaStreamNonBlocking));
    for (int i = 0; i < AUX_STREAMS; i++) {
        GD_CUDA(cudaStreamCreateWithFlags(&stream_aux[i], cudaStreamNonBlocking));
        GD_CUDA(cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming));
    }
    GD_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    GD_CUDA(cudaEventCreateWithFlags(&ev_pipe, cudaEventDisableTiming));
    for (int k = 5; k <= 13; k++) {
        std::vector<cpx> h;
        host_twiddles(h, 1LL << k, 1, 1LL << k);
        GD_TRY(upload(h, &wl[k]));
    }
    if (const char* s = getenv("GD_PASS_SCRATCH_MB")) {
        long v = atol(s);
        if (v > 0) pass_scratch_budget = (size_t)v << 20;
    }
    if (const char* s = getenv("GD_L2_BLOCK_MB")) { long v = atol(s); if (v > 0) l2_block_budget = (size_t)v << 20; }
    if (const char* s = getenv("GD_WIDE_TILES")) wide_tiles = atoi(s) != 0;
    if (const char* s = getenv("GD_FUSED")) use_fused = atoi(s) != 0;
    if (const char* s = getenv("GD_TILED")) tiled_scratch = atoi(s) != 0;
    if (const char* s = getenv("GD_W32")) w32 = atoi(s);
    if (const char* s = getenv("GD_TMA")) use_tma = atoi(s) != 0;
    if (const char* s = getenv("GD_TMA_OPT")) tma_opt = atoi(s);
    if (const char* s = getenv("GD_L2_WINDOW")) use_l2_window = atoi(s) != 0;
    if (const char* s = getenv("GD_FUSED_DELAY")) { int v = atoi(s); if (v >= 1 && v <= 6) fused_delay = v; }
    // the persisting L2 set-aside is one per GPU: only lane 0 carves and resets it (the other lanes' fused kernels run
    // without the window, with the evict-last hints alone)
    if (lane != 0) { use_l2_window = false; l2_block_window = false; }
    if (getenv("GD_VERBOSE")) fprintf(stderr, "[godsp] dev %d: %d SMs, L2 %d MiB, persisting max %zu MiB, window max %zu MiB\n", dev, num_sms, prop.l2CacheSize >> 20, l2_persist_max >> 20, l2_window_max >> 20);
    ready = true;
    return GD_OK;
}

void Device::destroy() {
    std::lock_guard<std::recursive_mutex> lk(mu);
    if (!ready) return;
    cudaSetDevice(dev);
    cudaDeviceSynchronize();
    if (scratch_evt) { cudaEventDestroy(scratch_evt); scratch_evt = nullptr; scratch_used = false; }
    for (auto& w : wl) { if (w) cudaFree(w); w = nullptr; }
    for (auto& kv : tw) { cudaFree(kv.second.lo); if (kv.second.hi) cudaFree(kv.second.hi); }
    tw.clear();
    for (auto& kv : blue) { cudaFree(kv.second.chirp_inv); cudaFree(kv.second.bhat); }
    blue.clear();
    for (int i = 0; i < SCR_NSLOTS; i++) { if (scratch[i]) cudaFree(scratch[i]); scratch[i] = nullptr; scratch_bytes[i] = 0; }
    cudaStreamDestroy(stream); cudaStreamDestroy(stream_in); cudaStreamDestroy(stream_out);
    for (int i = 0; i < AUX_STREAMS; i++) { cudaStreamDestroy(stream_aux[i]); cudaEventDestroy(ev_join[i]); stream_aux[i] = nullptr; ev_join[i] = nullptr; }
    cudaEventDestroy(ev_fork); cudaEventDestroy(ev_pipe);
    stream = stream_in = stream_out = nullptr;
    ev_fork = ev_pipe = nullptr;
    ready = false;
}

Status Device::ensure_scratch(ScratchSlot s, size_t bytes, void** out) {
    if (scratch_bytes[s] < bytes) {
        if (scratch[s]) {
            GD_CUDA(cudaDeviceSynchronize());     // nobody may still be using the old block
            GD_CUDA(cudaFree(scratch[s]));
            scratch[s] = nullptr; scratch_bytes[s] = 0;
        }
        size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        GD_CUDA(cudaMalloc(&scratch[s], want));
        scratch_bytes[s] = want;
    }
    *out = scratch[s];
    return GD_OK;
}

Status Device::l2_release() {
    if (l2_dirty && !l2_hold) {
        GD_CUDA(cudaCtxResetPersistingL2Cache());
        GD_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
        l2_carved = 0;
        l2_dirty = false;
    }
    return GD_OK;
}

Status Device::twiddles(int log2m, TwiddleTable* out) {
    auto it = tw.find(log2m);
    if (it == tw.end()) {
        if (log2m > 24) return invalid("twiddle table: M > 2^24 not supported");
        long long M = 1LL << log2m;
        TwiddleTable t;
        std::vector<cpx> h;
        host_twiddles(h, M, 1, M < 4096 ? M : 4096);
        GD_TRY(upload(h, &t.lo));
        if (log2m > 12) {
            host_twiddles(h, M, 4096, M >> 12);
            GD_TRY(upload(h, &t.hi));
        }
        it = tw.emplace(log2m, t).first;
    }
    *out = it->second;
    return GD_OK;
}

// keeps the persisting L2 set-aside carved for its lifetime: handing it back (cudaCtxResetPersistingL2Cache + limit 0) and carving
// it again between the launches of a chunk loop costs tens of microseconds of host time each way
struct L2Hold {
    Device& d;
    explicit L2Hold(Device& dv) : d(dv) { d.l2_hold++; }
    ~L2Hold() { d.l2_hold--; }
};

// ------------------------------------------------------------------ pass dispatch
static Status launch_pass(Device& d, int log2l, const PassParams& p, cudaStream_t st) {
    bool generic = (p.ld_flags & (LD_REAL | LD_PAD | LD_MULAUX | LD_REVERSE)) ||
                   (p.st_flags & (ST_MULAUX | ST_DIV | ST_TRUNC));
    GD_TRY(d.l2_release());
    cudaError_t e;
    if (log2l == 10 && !generic && d.w32 >= 1 && d.w32 <= 6) e = launch_pass32(d.w32, p, d.num_sms, st);   // 32 points per thread
    else if (log2l >= 1 && log2l <= 8) e = launch_pass_small(log2l, p, generic, d.num_sms, st);
    else if (log2l <= 10) e = launch_pass_mid(log2l, d.wide_tiles, p, generic, d.num_sms, st);
    else if (log2l <= 12) e = launch_pass_big(log2l, d.wide_tiles, p, generic, d.num_sms, st);
    else return invalid("launch_pass: line length out of range");
    if (e != cudaSuccess) return cuda_fail(e, "fft_pass_kernel launch");
    g_launches++;
    return GD_OK;
}

static PassParams base_params(Device& d, int log2l) {
    PassParams p;
    memset(&p, 0, sizeof(p));
    p.inner = 1;
    p.scale = 1.0; p.div = 1.0;
    p.wl = log2l >= 5 ? d.wl[log2l] : nullptr;
    return p;
}

// Spread the independent chunks of one call over `ways` streams: the caller's and the device's auxiliary ones. Each way
// has its own scratch block; short launches of different chunks fill each other's ramps and tails.
struct ForkJoin {
    Device& d;
    cudaStream_t st;
    int ways;
    bool window = false;
    ForkJoin(Device& dev, cudaStream_t s, int want) : d(dev), st(s), ways(want < 1 ? 1 : want) {
        if (ways > 1 + Device::AUX_STREAMS) ways = 1 + Device::AUX_STREAMS;
        for (int i = 0; i < Device::AUX_STREAMS; i++) if (s == d.stream_aux[i]) ways = 1;
        if (ways > 1) {
            cudaEventRecord(d.ev_fork, st);
            for (int i = 0; i + 1 < ways; i++) cudaStreamWaitEvent(d.stream_aux[i], d.ev_fork, 0);
        }
    }
    int way(long long i) const { return (int)(i % ways); }
    cudaStream_t stream(long long i) const { const int w = way(i); return w == 0 ? st : d.stream_aux[w - 1]; }
    // keep the inter-pass blocks in L2: persisting access-policy window over the scratch on the streams of this call
    void persist(void* base, size_t bytes) {
        if (!d.l2_block_window || !d.use_l2_window || d.l2_persist_max == 0 || d.l2_window_max == 0) return;
        const size_t want = bytes < d.l2_persist_max ? bytes : d.l2_persist_max;
        if (d.l2_carved != want) {
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) { cudaGetLastError(); return; }
            d.l2_carved = want;
        }
        set(base, bytes < d.l2_window_max ? bytes : d.l2_window_max, (double)want);
        window = true;
        d.l2_hold++;                         // launch_pass must not hand the set-aside back between the chunks
    }
    void set(void* base, size_t bytes, double carved) {
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr = base;
        attr.accessPolicyWindow.num_bytes = bytes;
        if (bytes) {
            const double ratio = carved / (double)bytes;
            attr.accessPolicyWindow.hitRatio = (float)(ratio > 1.0 ? 1.0 : ratio);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        }
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
        for (int i = 0; i + 1 < ways; i++) cudaStreamSetAttribute(d.stream_aux[i], cudaStreamAttributeAccessPolicyWindow, &attr);
    }
    ~ForkJoin() {
        if (window) { set(nullptr, 0, 0.0); d.l2_hold--; d.l2_dirty = true; }
        for (int i = 0; i + 1 < ways; i++) { cudaEventRecord(d.ev_join[i], d.stream_aux[i]); cudaStreamWaitEvent(st, d.ev_join[i], 0); }
    }
};

// ------------------------------------------------------------------ power-of-two transforms
static Status fft_axis(Device& d, const cpx* src, cpx* dst, long long outer, long long len, long long s, int dir, cudaStream_t st,
                       long long col0 = 0, long long ncols = -1);
Status transpose_batched(const cpx* in, cpx* out, long long batch, long long rows, long long cols, cudaStream_t st);

// Large single transforms (2^22 .. 2^34 points; the reference has no length limit, fft/fft.go:72-87) as an outer four-step over
// the fused size family. x is a row-major [N1][N2] matrix.
//  two sweeps (N2 <= 4096, N1 = 2^13 .. 2^17): (1) every column through the fused kernel in column mode, whose stores carry
//      w_N^(n2 k1) (TW2 in fft_tma14.cuh); (2) every row (length N2, contiguous) through one pass kernel with the transposed store
//      X[k1 + N1 k2]: tiles of T adjacent rows write T x 16 contiguous bytes per k2.
//  three sweeps (beyond): (1) as above, (2) one transpose to [N2][N1], (3) columns of length N2 in place: out[k2][k1] = X[k1 + N1 k2].
//  four sweeps (huge_sweeps = 4, the first formulation: columns, twiddle kernel, rows, transpose) stay as the cross-check.
// Forward and inverse (every sub-step inverted, conjugate twiddle) only; needs a second N-element buffer.
static Tma2dExtra huge_extra(int log2n, long long nmat, long long in_mdist, long long out_mdist) {
    Tma2dExtra ex;
    ex.tw2_log2m = log2n; ex.nmat = nmat; ex.in_mdist = in_mdist; ex.out_mdist = out_mdist;
    return ex;
}
static Status fft_pow2_huge(Device& d, const void* in, long long in_dist, cpx* out, long long out_dist, int log2n, long long batch,
                            const FusedOps& ops, cudaStream_t st) {
    const bool fwd = ops.ld_flags == 0 && ops.st_flags == 0;
    const bool inv = ops.ld_flags == LD_CONJ && ops.st_flags == (ST_CONJ | ST_SCALE);
    if (!fwd && !inv) return invalid("transforms above 2^24 points support plain forward / inverse only (no Bluestein above a padded length of 2^24)");
    if (log2n > 34) return invalid("fft_pow2: N > 2^34");
    const int dir = inv ? -1 : +1;
    const long long N = 1LL << log2n;
    // the split: N1 through the fused family in column mode (2^13 .. 2^17 rows), N2 = N / N1
    int l1 = d.huge_l1 ? d.huge_l1 : (log2n <= 24 ? 16 : log2n <= 29 ? 17 : (log2n + 1) / 2);
    if (l1 > log2n - 4) l1 = log2n - 4;
    int l2 = log2n - l1;
    const Tma2dEntry* te = (d.huge_sweeps != 4 && l1 >= 13 && l1 <= 17) ? tma2d_entry(d, l1) : nullptr;
    const bool three = te && (l2 > 12 || d.huge_sweeps == 3);
    if (te && !three && l2 > 12) te = nullptr;
    if (!te) { l1 = (log2n + 1) / 2; l2 = log2n - l1; }
    const long long N1 = 1LL << l1, N2 = 1LL << l2;
    // two sweeps: the transforms of a batch go through each sweep together, as many as fit the scratch budget (one launch per sweep
    // instead of one per transform: at 2^21 .. 2^23 points a launch per transform is mostly ramp and tail)
    long long per = 1;
    if (te && !three && batch > 1 && N2 / te->unit <= 512 && in_dist < (1LL << 35)) {
        per = (long long)(d.pass_scratch_budget / ((size_t)N * sizeof(cpx)));
        if (per < 1) per = 1;
        if (per > batch) per = batch;
    }
    cpx* tmp;
    GD_TRY(d.ensure_scratch(SCR_HUGE, (size_t)per * N * sizeof(cpx), (void**)&tmp));
    // the fused kernel's persisting L2 set-aside stays carved across the sweeps and the transforms of the batch (handing it back and
    // carving it again around every pass kernel costs more host time than a 2^21-point transform takes)
    L2Hold hold(d);
    if (per > 1 && te->cols_ok((const cpx*)in, tmp, N1, N2, N2)) {
        for (long long b0 = 0; b0 < batch; b0 += per) {
            const long long nb = batch - b0 < per ? batch - b0 : per;
            GD_TRY(te->run(d, 1, (const cpx*)in + b0 * in_dist, N2, tmp, N2, N2, inv, inv ? 1.0 / (double)N1 : 1.0, st, huge_extra(log2n, nb, in_dist, N)));
            PassParams r = base_params(d, l2);
            r.in = tmp; r.out = out + b0 * out_dist;
            r.nlines = nb * N1; r.inner = N1;
            r.in_qs = N; r.in_is = N2; r.in_es = 1;
            r.out_qs = out_dist; r.out_is = 1; r.out_es = (int)N1;
            r.in_mode = MODE_ROW; r.out_mode = MODE_COL;
            if (inv) { r.ld_flags = LD_CONJ; r.st_flags = ST_CONJ | ST_SCALE; r.scale = 1.0 / (double)N2; }
            GD_TRY(launch_pass(d, l2, r, st));
        }
        return GD_OK;
    }
    for (long long b = 0; b < batch; b++) {
        const cpx* src = (const cpx*)in + b * in_dist;
        cpx* dst = out + b * out_dist;
        if (te && te->cols_ok(src, tmp, N1, N2, N2)) {
            cpx* mid = three ? dst : tmp;                                        // three sweeps: columns -> dst, transpose -> tmp, columns -> dst
            GD_TRY(te->run(d, 1, src, N2, mid, N2, N2, inv, inv ? 1.0 / (double)N1 : 1.0, st, huge_extra(log2n, 1, 0, 0)));   // (1) columns n2, twiddle on store
            if (three) {
                GD_TRY(transpose_batched(mid, tmp, 1, N1, N2, st));              // (2) [N1][N2] -> [N2][N1]
                GD_TRY(fft_axis(d, tmp, dst, 1, N2, N1, dir, st));               // (3) lines over n2 at stride N1
                continue;
            }
            PassParams r = base_params(d, l2);                                   // (2) rows k1 of tmp -> X[k1 + N1 k2]
            r.in = tmp; r.out = dst;
            r.nlines = N1; r.inner = N1;
            r.in_qs = N; r.in_is = N2; r.in_es = 1;
            r.out_qs = N; r.out_is = 1; r.out_es = (int)N1;
            r.in_mode = MODE_ROW; r.out_mode = MODE_COL;
            if (inv) { r.ld_flags = LD_CONJ; r.st_flags = ST_CONJ | ST_SCALE; r.scale = 1.0 / (double)N2; }
            GD_TRY(launch_pass(d, l2, r, st));
            continue;
        }
        GD_TRY(fft_axis(d, src, tmp, 1, N1, N2, dir, st));                       // (1) columns: lines over n1, stride N2
        GD_TRY(fourstep_twiddle(tmp, N1, N2, 0, 0, log2n, st, dir));             // (2)
        GD_TRY(fft1d(d, tmp, N2, tmp, N2, N2, N1, false, dir, st));              // (3) rows, in place
        GD_TRY(transpose_batched(tmp, dst, 1, N1, N2, st));                      // (4)
    }
    return GD_OK;
}

Status fft_pow2(Device& d, const void* in, long long in_dist, cpx* out, long long out_dist, int log2n,
                long long batch, const FusedOps& ops, cudaStream_t st) {
    if (log2n < 1 || batch < 1) return invalid("fft_pow2: bad size");
    const long long N = 1LL << log2n;
    if (log2n <= 12) {
        PassParams p = base_params(d, log2n);
        p.in = in; p.out = out; p.nlines = batch; p.inner = 1;
        p.in_qs = in_dist; p.in_is = 0; p.in_es = 1;
        p.out_qs = out_dist; p.out_is = 0; p.out_es = 1;
        p.in_mode = p.out_mode = MODE_ROW;
        p.ld_flags = ops.ld_flags; p.st_flags = ops.st_flags;
        p.aux_in = ops.aux_in; p.aux_out = ops.aux_out;
        p.n_valid_in = ops.n_valid_in; p.n_valid_out = ops.n_valid_out;
        p.scale = ops.scale; p.div = ops.div;
        return launch_pass(d, log2n, p, st);
    }
    // plain forward / inverse transforms from 2^22 points (2^21 in a batch, where every sweep is one launch over the batch) take the outer four-step
    const int hmin = batch > 1 ? d.huge_min_log2n - 1 : d.huge_min_log2n;
    if (log2n > 24 || (log2n >= hmin && ops.ld_flags == 0 && ops.st_flags == 0) ||
        (log2n >= hmin && ops.ld_flags == LD_CONJ && ops.st_flags == (ST_CONJ | ST_SCALE)))
        return fft_pow2_huge(d, in, in_dist, out, out_dist, log2n, batch, ops, st);
    const bool lean = !(ops.ld_flags & ~LD_CONJ) && !(ops.st_flags & ~(ST_CONJ | ST_SCALE));
    if (d.use_tma && lean && log2n == 20 && !d.debug_alias) {
        const int lc = (ops.ld_flags & LD_CONJ) ? 1 : 0, sc = (ops.st_flags & ST_CONJ) ? 1 : 0;
        const double scl = (ops.st_flags & ST_SCALE) ? ops.scale : 1.0;
        if (tma_fused_applicable(in, in_dist, out, out_dist, lc, sc, scl))
            return fft_tma_2p20(d, (const cpx*)in, in_dist, out, out_dist, batch, lc, sc, scl, st);
    }
    if (const Tma2dEntry* te = (lean && !d.debug_alias) ? tma2d_entry(d, log2n) : nullptr) {
        // 2^13 .. 2^18: whole phases of the batch through the fused kernel, a remainder of less than one phase through the chunks below
        const int lc = (ops.ld_flags & LD_CONJ) ? 1 : 0, sc = (ops.st_flags & ST_CONJ) ? 1 : 0;
        const double scl = (ops.st_flags & ST_SCALE) ? ops.scale : 1.0;
        const long long main = batch - batch % te->unit;
        if (main > 0 && te->rows_ok(in, in_dist, out, out_dist, main, lc, sc, scl)) {
            GD_TRY(te->run(d, 0, (const cpx*)in, in_dist, out, out_dist, main, lc != 0, scl, st, Tma2dExtra()));
            if (main == batch) return GD_OK;
            return fft_pow2(d, (const cpx*)in + main * in_dist, in_dist, out + main * out_dist, out_dist, log2n, batch - main, ops, st);
        }
    }
    if (d.use_fused && lean && (log2n % 2) == 0 && log2n >= 16 && batch * (2LL << (log2n / 2)) < (1LL << 30)) {
        // both passes in one persistent kernel, intermediate resident in L2 (fft_fused.cuh)
        const int l = log2n / 2, T = fused_tile_lines(l, d.wide_tiles), tpt = (1 << l) / T;
        TwiddleTable tw;
        GD_TRY(d.twiddles(log2n, &tw));
        long long J = (2LL * d.num_sms * 2 + tpt - 1) / tpt;               // ~2 rounds of resident CTAs per phase
        long long jcap = (long long)(d.fused_slot_budget / ((size_t)N * sizeof(cpx)));
        if (jcap < 1) jcap = 1;
        if (J > jcap) J = jcap;
        if (J > batch) J = batch;
        FusedParams f;
        memset(&f, 0, sizeof(f));
        f.in = (const cpx*)in; f.out = out; f.in_dist = in_dist; f.out_dist = out_dist;
        f.batch = (int)batch; f.group_tf = (int)J; f.ngroups = (int)((batch + J - 1) / J);
        f.log2n = log2n; f.ld_conj = (ops.ld_flags & LD_CONJ) ? 1 : 0; f.st_flags = ops.st_flags; f.scale = ops.scale;
        f.tw_lo = tw.lo; f.tw_hi = tw.hi; f.wl = d.wl[l];
        if (const char* dbg = getenv("GD_FUSED_DEBUG")) f.debug = atoi(dbg);
        f.delay = d.fused_delay; f.nslots = d.fused_delay + 2;
        GD_TRY(d.ensure_scratch(SCR_PASS, (size_t)(d.fused_delay + 2) * J * N * sizeof(cpx), (void**)&f.scratch));
        int* cnt;
        GD_TRY(d.ensure_scratch(SCR_CNT, (2 * (size_t)f.ngroups + 1) * sizeof(int), (void**)&cnt));
        GD_CUDA(cudaMemsetAsync(cnt, 0, (2 * (size_t)f.ngroups + 1) * sizeof(int), st));
        f.done1 = cnt; f.done2 = cnt + f.ngroups; f.next_item = cnt + 2 * f.ngroups;
        long long total_items = 2LL * f.ngroups * J * tpt;
        // keep the scratch slots resident in L2: persisting access-policy window for this launch
        const size_t scr_bytes = (size_t)(d.fused_delay + 2) * J * N * sizeof(cpx);
        const bool window = d.use_l2_window && d.l2_persist_max > 0 && d.l2_window_max > 0;
        if (window) {      // carve a persisting set-aside just large enough for the scratch slots (the rest of L2
                           // keeps merging the partial-line stores of the streaming side)
            size_t want = scr_bytes < d.l2_persist_max ? scr_bytes : d.l2_persist_max;
            if (d.l2_carved != want) {
                GD_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
                d.l2_carved = want;
            }
        }
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        if (window) {
            attr.accessPolicyWindow.base_ptr = f.scratch;
            attr.accessPolicyWindow.num_bytes = scr_bytes < d.l2_window_max ? scr_bytes : d.l2_window_max;
            double ratio = (double)d.l2_carved / (double)attr.accessPolicyWindow.num_bytes;
            attr.accessPolicyWindow.hitRatio = (float)(ratio > 1.0 ? 1.0 : ratio);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            GD_CUDA(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
        }
        d.l2_dirty = window;
        cudaError_t e = launch_fused(l, d.wide_tiles, f, total_items, d.num_sms, st);
        if (window) {
            attr.accessPolicyWindow.num_bytes = 0;
            cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
        }
        if (e != cudaSuccess) return cuda_fail(e, "fft_fused_kernel launch");
        g_launches++;
        return GD_OK;
    }
    // four-step: N = N1 * N2, n = n1*N2 + n2, k = k1 + N1*k2
    if (d.debug_alias) { in_dist = 0; out_dist = 0; }   // experiment: every transform on the same L2-resident buffers
    const int l1 = (log2n + 1) / 2, l2 = log2n - l1;
    const long long N1 = 1LL << l1, N2 = 1LL << l2;
    TwiddleTable tw;
    GD_TRY(d.twiddles(log2n, &tw));
    // The inter-pass block of a chunk is written by pass 1 and read back by pass 2 right away: chunks are sized to stay
    // in L2 (l2_block_budget), so HBM sees 32 B per point, not 64. Chunks rotate over chunk_streams streams (one scratch
    // block each), which hides the launch gaps and tails of these short launches behind the other chunks' kernels.
    const size_t budget = d.pass_scratch_budget < d.l2_block_budget ? d.pass_scratch_budget : d.l2_block_budget;
    long long chunk = (long long)(budget / ((size_t)N * sizeof(cpx)));
    if (chunk < 1) chunk = 1;
    if (chunk > batch) chunk = batch;
    const long long nchunks = (batch + chunk - 1) / chunk;
    ForkJoin fj(d, st, nchunks >= 2 * d.chunk_streams ? d.chunk_streams : 1);
    cpx* scr0;
    GD_TRY(d.ensure_scratch(SCR_PASS, (size_t)fj.ways * chunk * N * sizeof(cpx), (void**)&scr0));
    const bool real_in = ops.ld_flags & LD_REAL;
    if (nchunks > 1) fj.persist(scr0, (size_t)fj.ways * chunk * N * sizeof(cpx));
    long long ci = 0;
    for (long long b0 = 0; b0 < batch; b0 += chunk, ci++) {
        long long nb = batch - b0 < chunk ? batch - b0 : chunk;
        cudaStream_t st = fj.stream(ci);
        cpx* scr = scr0 + (size_t)fj.way(ci) * chunk * N;
        // pass 1: columns n2 of every transform, length N1, twiddle w_N^(n2*k1) on store; the intermediate
        // is tile-major (T adjacent columns = one contiguous block) so these stores are fully coalesced
        const bool lean32 = d.w32 >= 1 && d.w32 <= 6 && !(ops.ld_flags & ~LD_CONJ) && !(ops.st_flags & ~(ST_CONJ | ST_SCALE));
        const int T1 = (l1 == 10 && lean32) ? pass32_tile_lines(d.w32) : pass_tile_lines(l1, d.wide_tiles);
        const int T2 = (l2 == 10 && lean32) ? pass32_tile_lines(d.w32) : pass_tile_lines(l2, d.wide_tiles);
        const bool tiled = d.tiled_scratch && T1 == T2 && ((N2 / 32) % T1) == 0 && !(ops.st_flags & ~(ST_CONJ | ST_SCALE)) && !(ops.ld_flags & ~LD_CONJ);
        PassParams p = base_params(d, l1);
        p.in = real_in ? (const void*)((const double*)in + b0 * in_dist) : (const void*)((const cpx*)in + b0 * in_dist);
        p.out = scr;
        p.nlines = nb * N2; p.inner = N2;
        p.in_qs = in_dist; p.in_is = 1; p.in_es = (int)N2;
        p.out_qs = d.debug_alias ? 0 : N; p.out_is = 1; p.out_es = (int)N2;
        p.in_mode = p.out_mode = MODE_COL;
        p.ld_flags = ops.ld_flags; p.aux_in = ops.aux_in; p.n_valid_in = ops.n_valid_in;
        p.st_flags = ST_TWIDDLE; p.tw_sel = 0; p.tw_log2m = log2n; p.tw_lo = tw.lo; p.tw_hi = tw.hi;
        if (tiled) { p.out_tiled = 1; p.tiled_len = (int)N1; }
        GD_TRY(launch_pass(d, l1, p, st));
        // pass 2: rows k1, length N2, transposed store X[k1 + N1*k2]
        PassParams r = base_params(d, l2);
        r.in = scr; r.out = out + b0 * out_dist;
        r.nlines = nb * N1; r.inner = N1;
        r.in_qs = d.debug_alias ? 0 : N; r.in_is = N2; r.in_es = 1;
        r.out_qs = out_dist; r.out_is = 1; r.out_es = (int)N1;
        r.in_mode = MODE_ROW; r.out_mode = MODE_COL;
        if (tiled) { r.in_tiled = 1; r.tiled_len = (int)N1; r.in_mode = MODE_COL; }
        r.st_flags = ops.st_flags; r.aux_out = ops.aux_out; r.n_valid_out = ops.n_valid_out;
        r.scale = ops.scale; r.div = ops.div;
        GD_TRY(launch_pass(d, l2, r, st));
    }
    return GD_OK;
}

// ------------------------------------------------------------------ small helper kernels
__global__ void copy1_kernel(const void* in, long long in_dist, cpx* out, long long out_dist, long long batch, int real_in) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= batch) return;
    cpx v;
    if (real_in) v = make_double2(((const double*)in)[b * in_dist], 0.0);
    else v = ((const cpx*)in)[b * in_dist];
    out[b * out_dist] = v;
}

// chirp tables with the reference's exact argument arithmetic (fft/bluestein.go:48-57):
// arg = fl(fl(pi / N) * fl(i*i)); i*i is exact in int64 and in double for N < 9.4e7.
__global__ void chirp_kernel(long long n, long long la, cpx* chirp_inv, cpx* b) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s, c;
    if (i == 0) { s = 0.0; c = 1.0; }
    else {
        double coef = __ddiv_rn(3.14159265358979323846, (double)n);
        double arg = __dmul_rn(coef, (double)(i * i));
        sincos(arg, &s, &c);
    }
    chirp_inv[i] = make_double2(c, -s);
    b[i] = make_double2(c, s);
    if (i != 0) b[la - i] = make_double2(c, s);
}

// Bluestein above a padded length of 2^24 (the power-of-two transforms there are the outer four-step, which takes no fused
// load / store operators): the same element operations as LD_PAD | LD_MULAUX | LD_REAL | LD_REVERSE and
// ST_MULAUX | ST_TRUNC | ST_DIV of the pass kernel, as two streaming kernels.
__global__ void bluestein_prep_kernel(const void* __restrict__ in, long long in_dist, int real_in, int reverse, long long n, long long la,
                                      const cpx* __restrict__ chirp_inv, cpx* __restrict__ a) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x, t = blockIdx.y;       // blockIdx.y: transform of the chunk
    a += t * la;
    for (; i < la; i += stride) {
        cpx v = make_double2(0.0, 0.0);
        if (i < n) {
            const long long src = t * in_dist + ((reverse && i != 0) ? n - i : i);   // fft/fft.go:39-43
            if (real_in) v.x = __ldg(reinterpret_cast<const double*>(in) + src);
            else v = __ldg(reinterpret_cast<const cpx*>(in) + src);
            v = cmul(v, __ldg(chirp_inv + i));                              // fft/bluestein.go:70-73
        }
        a[i] = v;
    }
}
__global__ void bluestein_post_kernel(const cpx* __restrict__ r, long long la, const cpx* __restrict__ chirp_inv, long long n, int inverse,
                                      double div, cpx* __restrict__ out, long long out_dist) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x, t = blockIdx.y;
    r += t * la; out += t * out_dist;
    for (; i < n; i += stride) {
        cpx v = cmul(r[i], __ldg(chirp_inv + i));                           // fft/bluestein.go:89-91
        if (inverse) { v.x /= div; v.y /= div; }                            // fft/fft.go:47-50
        out[i] = v;
    }
}
// a[t][i] *= b[i] for every transform t of the chunk (fft/fft.go:63-66 with the cached FFT(b))
__global__ void bluestein_mul_kernel(cpx* __restrict__ a, const cpx* __restrict__ b, long long la) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    a += (long long)blockIdx.y * la;
    for (; i < la; i += stride) a[i] = cmul(a[i], __ldg(b + i));
}

__global__ void pointwise_mul_kernel(cpx* a, const cpx* b, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) a[i] = cmul(a[i], b[i]);
}

// lines (o, i) of length len and element stride s  <->  dense [line][len], through a padded 32 x 32 shared-memory tile: adjacent lines are adjacent in the source, adjacent points adjacent in the
// dense copy, so both sides move 512 contiguous bytes per warp (the plain kernels read / write 16 bytes per stride s)
template <bool SCATTER>
__global__ void __launch_bounds__(256) lines_tiled_kernel(const cpx* __restrict__ src, cpx* __restrict__ dst, long long line0, long long nlines,
                                                          long long len, long long s) {
    __shared__ cpx tile[32][33];
    const long long l_base = (long long)blockIdx.x * 32, j_base = (long long)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (!SCATTER) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const long long j = j_base + ty + 8 * k, l = l_base + tx;
            if (l < nlines && j < len) { const long long line = line0 + l, o = line / s, i = line - o * s; tile[ty + 8 * k][tx] = src[o * len * s + i + j * s]; }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const long long l = l_base + ty + 8 * k, j = j_base + tx;
            if (l < nlines && j < len) dst[l * len + j] = tile[tx][ty + 8 * k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const long long l = l_base + ty + 8 * k, j = j_base + tx;
            if (l < nlines && j < len) tile[tx][ty + 8 * k] = src[l * len + j];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const long long j = j_base + ty + 8 * k, l = l_base + tx;
            if (l < nlines && j < len) { const long long line = line0 + l, o = line / s, i = line - o * s; dst[o * len * s + i + j * s] = tile[ty + 8 * k][tx]; }
        }
    }
}
// the same one element per thread (lines longer than 2^21 points)
__global__ void gather_lines_kernel(const cpx* src, cpx* dst, long long line0, long long nlines, long long len, long long s) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nlines * len) return;
    long long l = t / len, j = t - l * len;
    long long line = line0 + l, o = line / s, i = line - o * s;
    dst[t] = src[o * len * s + i + j * s];
}
__global__ void scatter_lines_kernel(const cpx* src, cpx* dst, long long line0, long long nlines, long long len, long long s) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nlines * len) return;
    long long l = t / len, j = t - l * len;
    long long line = line0 + l, o = line / s, i = line - o * s;
    dst[o * len * s + i + j * s] = src[t];
}

__global__ void splitmix_kernel(double* out, long long n, unsigned long long seed, unsigned long long offset) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = splitmix_unit(seed, offset + (unsigned long long)i);
}

// ---- distributed four-step helpers (config C5: one transform of N = N1*N2 points over G GPUs) ----
// block[r][c] *= w_N^((row0 + r) * (col0 + c)), N = 2^log2n <= 2^40. One thread owns 16 consecutive c of a row:
// base and step come from sincospi of exactly reduced exponents, the run by a 15-term product chain.
__global__ void fourstep_twiddle_kernel(cpx* __restrict__ blk, long long rows, long long cols, long long row0,
                                        long long col0, int log2n, double sgn) {
    const long long runs_per_row = (cols + 15) / 16;
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= rows * runs_per_row) return;
    const long long r = t / runs_per_row, c = (t - r * runs_per_row) * 16;
    const unsigned long long mask = (1ULL << log2n) - 1ULL, gr = (unsigned long long)(row0 + r);
    const double invn = 1.0 / (double)(1ULL << log2n);
    double s, co;
    sincospi(sgn * 2.0 * (double)((gr * (unsigned long long)(col0 + c)) & mask) * invn, &s, &co);
    cpx w = make_double2(co, s);
    sincospi(sgn * 2.0 * (double)(gr & mask) * invn, &s, &co);
    const cpx step = make_double2(co, s);
    cpx* p = blk + r * cols + c;
    const int n = (int)(cols - c < 16 ? cols - c : 16);
    for (int i = 0; i < n; i++) { p[i] = cmul(p[i], w); w = cmul(w, step); }
}
// all-to-all receive buffer [G][K][W] (source rank, local row, source's column) -> rows [K][G*W]
__global__ void repack_gkw_kernel(const cpx* __restrict__ in, cpx* __restrict__ out, long long G, long long K, long long W) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= G * K * W) return;
    const long long j = t % W, k = (t / W) % K, g = t / (W * K);
    out[k * (G * W) + g * W + j] = in[t];
}

// batched transpose in[b][r][c] -> out[b][c][r] through a padded 32 x 32 shared-memory tile (both sides coalesced):
// turns each source rank's [K][W] all-to-all block into [W][K], so the received slab is [N2][K] and the second
// four-step pass runs as strided lines with its output already in natural order
__global__ void __launch_bounds__(256) transpose_batched_kernel(const cpx* __restrict__ in, cpx* __restrict__ out,
                                                                long long rows, long long cols) {
    __shared__ cpx tile[32][33];
    const long long b = blockIdx.z;
    const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const cpx* src = in + b * rows * cols;
    cpx* dst = out + b * rows * cols;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const long long r = r0 + ty + 8 * i, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + 8 * i][tx] = src[r * cols + c];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const long long c = c0 + ty + 8 * i, r = r0 + tx;
        if (r < rows && c < cols) dst[c * rows + r] = tile[tx][ty + 8 * i];
    }
}
Status transpose_batched(const cpx* in, cpx* out, long long batch, long long rows, long long cols, cudaStream_t st) {
    if (batch < 1 || rows < 1 || cols < 1 || in == out) return invalid("transpose_batched: bad arguments");
    const long long gx = (cols + 31) / 32, gy = (rows + 31) / 32;
    if (gy > 65535 || batch > 65535) return invalid("transpose_batched: grid too large");      // rows <= 2^21
    transpose_batched_kernel<<<dim3((unsigned)gx, (unsigned)gy, (unsigned)batch), 256, 0, st>>>(in, out, rows, cols);
    g_launches++;
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

// The exchange step of the sharded four-step as ONE kernel over peer memory (NVLink P2P stores) instead of
// twiddle kernel + NCCL all-to-all + transpose kernel: block (h, k-tile, c-tile) reads slab[h*K + k][c] (this rank's
// [N1][W] slab after the length-N1 lines), multiplies by w_N^((h*K + k) * (g*W + c)), transposes the 64 x 32 tile
// through shared memory and stores it straight into rank h's receive buffer at [g*W + c][k] (rows of K elements,
// 1 KiB contiguous per row segment). One HBM read + one NVLink write per element; the receive buffer is then
// already [N2][K], ready for the length-N2 lines.
struct PeerPtrs { cpx* p[16]; };      // passed as __grid_constant__: indexed in the parameter bank, no local-memory copy
__global__ void __launch_bounds__(256) fourstep_exchange_kernel(const cpx* __restrict__ slab, const __grid_constant__ PeerPtrs peers, long long K,
                                                                long long W, int g, int log2n, int world_, long long cbeg, long long cend,
                                                                long long nbx, long long nby, double sgn) {
    __shared__ cpx tile[64][33];
    // consecutive blocks go to different peers, and rank g starts with peer g + 1: with one peer per grid slice every rank
    // wrote to the same destination at the same time (8 GPUs: 56 ms against 34 ms for NCCL's all-to-all)
    const int world = world_;
    // grid-stride over the (x, y) tiles: a pipelined call launches only a few CTAs so that the exchange, which is bound by
    // NVLink, leaves the SMs to the line kernels it overlaps with
    const long long ntx = nbx, nty = nby;
    for (long long t = blockIdx.x; t < ntx * nty; t += gridDim.x) {
    const long long bxi = t % ntx, byi = t / ntx;
    __syncthreads();
    const int h = (int)((bxi % world + g + 1) % world);
    const long long k0 = byi * 64, c0 = cbeg + (bxi / world) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long c = c0 + tx;
    if (c < cend) {
        const unsigned long long mask = (1ULL << log2n) - 1ULL, n2 = (unsigned long long)g * W + c;
        const double invn = 1.0 / (double)(1ULL << log2n);
        const unsigned long long k1a = (unsigned long long)h * K + k0 + ty;
        double sn, cs;
        sincospi(sgn * 2.0 * (double)((k1a * n2) & mask) * invn, &sn, &cs);       // sgn = -1: forward; +1: inverse (conjugate twiddle)
        cpx w = make_double2(cs, sn);
        sincospi(sgn * 2.0 * (double)((8ULL * n2) & mask) * invn, &sn, &cs);
        const cpx step = make_double2(cs, sn);
        const cpx* src = slab + ((long long)h * K + k0 + ty) * W + c;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (k0 + ty + 8 * i < K) tile[ty + 8 * i][tx] = cmul(src[(long long)(8 * i) * W], w);
            w = cmul(w, step);
        }
    }
    __syncthreads();
    cpx* dst = peers.p[h] + ((long long)g * W + c0) * K + k0;
    for (int rr = ty; rr < 32; rr += 8) {
        if (c0 + rr >= cend) break;
#pragma unroll
        for (int kk = tx; kk < 64; kk += 32)
            if (k0 + kk < K) dst[(long long)rr * K + kk] = tile[kk][rr];
    }
    }
}
Status fourstep_exchange(const cpx* slab, cpx* const* peer_recv, long long n1, long long w, int rank, int world, int log2n,
                         cudaStream_t st, long long cbeg, long long ccount, int max_ctas, int dir) {
    if (!slab || !peer_recv || world < 1 || world > 16 || rank < 0 || rank >= world || n1 % world || w < 1 || log2n < 1 || log2n > 40)
        return invalid("fourstep_exchange: bad arguments");
    if (ccount < 0) { cbeg = 0; ccount = w; }
    if (cbeg < 0 || cbeg + ccount > w || ccount < 1) return invalid("fourstep_exchange: bad column range");
    const long long K = n1 / world;
    PeerPtrs pp;
    for (int i = 0; i < 16; i++) pp.p[i] = i < world ? peer_recv[i] : nullptr;
    const long long gx = (ccount + 31) / 32, gy = (K + 63) / 64;
    long long grid = gx * world * gy;
    if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
    if (grid > 2147483647LL) grid = 2147483647LL;
    fourstep_exchange_kernel<<<(unsigned)grid, 256, 0, st>>>(slab, pp, K, w, rank, log2n, world, cbeg, cbeg + ccount, gx * world, gy,
                                                          dir < 0 ? 1.0 : -1.0);
    g_launches++;
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

// First half of the sharded four-step, pipelined: the length-n1 lines of a block of columns (slab -> tmp) run on the
// caller's stream while the exchange kernel of the previous block (tmp -> peers over NVLink) runs on another one, so the
// NVLink stores overlap the butterflies instead of following them. The caller fences the ranks before and after.
Status fourstep_lines_exchange(Device& d, const cpx* slab, cpx* tmp, cpx* const* peer_recv, long long n1, long long w, int rank, int world,
                               int log2n, cudaStream_t st) {
    if (!slab || !tmp || slab == tmp) return invalid("fourstep_lines_exchange: bad arguments");
    const bool blockable = is_pow2(n1) && n1 > 4096 && n1 <= (1LL << 24) && (double)n1 * (double)w < 2147483648.0;
    long long pb = (long long)(((size_t)d.fourstep_pipeline_mb << 20) / ((size_t)n1 * sizeof(cpx)));   // columns per pipeline block
    pb = (pb / 32) * 32;
    if (!d.fourstep_pipeline || !blockable || pb < 32 || w < 2 * pb) {
        GD_TRY(fft_axis(d, slab, tmp, 1, n1, w, +1, st));
        return fourstep_exchange(tmp, peer_recv, n1, w, rank, world, log2n, st, 0, -1, 0);
    }
    cudaStream_t sx = d.stream_aux[Device::AUX_STREAMS - 1];                          // not one of the chunk streams fft_axis rotates over
    if (st == sx) return invalid("fourstep_lines_exchange: stream clash");
    cudaEventRecord(d.ev_fork, st);
    cudaStreamWaitEvent(sx, d.ev_fork, 0);
    // the fused line kernel is persistent with one CTA per SM: while an exchange runs beside it, it leaves some SMs free
    const int saved_cap = d.tma_grid_cap;
    if (d.fourstep_lines_sms > 0) d.tma_grid_cap = d.fourstep_lines_sms;
    Status rc = GD_OK;
    for (long long c0 = 0; c0 < w && rc == GD_OK; c0 += pb) {
        const long long nc = w - c0 < pb ? w - c0 : pb;
        rc = fft_axis(d, slab, tmp, 1, n1, w, +1, st, c0, nc);
        if (rc != GD_OK) break;
        cudaEventRecord(d.ev_pipe, st);
        cudaStreamWaitEvent(sx, d.ev_pipe, 0);
        rc = fourstep_exchange(tmp, peer_recv, n1, w, rank, world, log2n, sx, c0, nc, d.fourstep_exchange_ctas);
    }
    d.tma_grid_cap = saved_cap;
    GD_TRY(rc);
    GD_CUDA(cudaEventRecord(d.ev_pipe, sx));
    GD_CUDA(cudaStreamWaitEvent(st, d.ev_pipe, 0));
    return GD_OK;
}

// The sharded four-step with the exchange fused into the first line pass (TW2 = 2 in fft_tma14.cuh): the length-n1 lines of this
// rank's [n1][w] slab run through the fused family in column mode, the outputs leave multiplied by w_N^(k1 n2), and the TMA stores
// of pass 2 go straight into the ranks' receive buffers over NVLink: row k1 of the slab is row k1 % K of block `rank` ([K][w],
// K = n1 / world) of rank k1 / K. No intermediate slab, no exchange kernel, NVLink busy while the butterflies run.
bool fourstep_fused_supported(Device& d, long long n1, long long n2, int world) {
    if (world < 1 || world > 8 || !is_pow2(n1) || !is_pow2(n2) || n1 % world || n2 % world) return false;
    const int l1 = ilog2ll(n1), l2 = ilog2ll(n2);
    if (l1 < 13 || l1 > 17 || l2 < 13 || l2 > 18 || !tma2d_entry(d, l1) || !tma2d_entry(d, l2)) return false;
    const long long w = n2 / world, k = n1 / world;
    const int lb1 = 1 << (l1 / 2), la2 = 1 << ((l2 + 1) / 2);                 // LB of the n1 lines, LA of the n2 lines
    return w % tma2d_entry(d, l1)->unit == 0 && k % tma2d_entry(d, l2)->unit == 0 && lb1 % world == 0 &&
           (world == 1 || (la2 % world == 0 && world % 2 == 0)) && w < (1LL << 30);
}
Status fourstep_lines_peer(Device& d, const cpx* slab, cpx* const* peer_recv, long long n1, long long w, int rank, int world, int log2n,
                           int dir, cudaStream_t st) {
    // log2n = 0: no twiddle (the column pass of FFT2 on row blocks: rows [h*K, (h+1)*K) of the transformed slab go to rank h)
    const long long n2 = log2n ? (1LL << log2n) / n1 : w * world;
    if (!slab || !peer_recv || rank < 0 || rank >= world || !fourstep_fused_supported(d, n1, n2, world) || (log2n && w * world * n1 != (1LL << log2n)))
        return invalid("fourstep_lines_peer: unsupported shape (see gd_fourstep_fused_supported)");
    const Tma2dEntry* te = tma2d_entry(d, ilog2ll(n1));
    if (!te->cols_ok(slab, slab, n1, w, w)) return invalid("fourstep_lines_peer: slab alignment");
    Tma2dExtra ex;
    ex.tw2_log2m = log2n; ex.tw2_col0 = (long long)rank * w;
    ex.npeer = world; ex.peer = peer_recv; ex.peer_off = (long long)rank * (n1 / world) * w; ex.rank = rank;
    return te->run(d, 1, slab, w, const_cast<cpx*>(slab), w, w, dir < 0, dir < 0 ? 1.0 / (double)n1 : 1.0, st, ex);
}
// Second half: this rank's receive buffer [world][K][w] -- row k1 of the spectrum's [n1][n2] view in `world` segments of w = n2 / world
// points -- through the fused family in row mode on segmented lines: out[k1 local][k2] = X[k1 + n1 k2], K transforms of n2 points.
Status fourstep_rows_seg(Device& d, const cpx* recv, cpx* out, long long n2, long long k, int world, int dir, cudaStream_t st) {
    if (!recv || !out || recv == out || world < 1 || n2 % world) return invalid("fourstep_rows_seg: bad arguments");
    const long long w = n2 / world;
    if (world == 1) return fft1d(d, recv, n2, out, n2, n2, k, false, dir, st);
    const Tma2dEntry* te = is_pow2(n2) ? tma2d_entry(d, ilog2ll(n2)) : nullptr;
    if (!te || k % te->unit || ((uintptr_t)recv % 16) || ((uintptr_t)out % 16)) return invalid("fourstep_rows_seg: unsupported shape (see gd_fourstep_fused_supported)");
    Tma2dExtra ex;
    ex.seg = world; ex.seg_dist = k * w;
    return te->run(d, 0, recv, w, out, n2, k, dir < 0, dir < 0 ? 1.0 / (double)n2 : 1.0, st, ex);
}

// FFT2 on row blocks: both exchanges are strided block copies into peer memory (no repack kernels, no NCCL data
// movement). For every peer h: dst_h[dst_off + r*dst_pitch + c] = src[h*src_step + r*src_pitch + c], r < rows, c < cols.
__global__ void __launch_bounds__(256) peer_block_copy_kernel(const cpx* __restrict__ src, const __grid_constant__ PeerPtrs peers, long long rows, long long cols,
                                                              long long src_step, long long src_pitch, long long dst_off, long long dst_pitch,
                                                              int world, int rank) {
    const int h = (int)((blockIdx.x % world + rank + 1) % world);      // consecutive blocks -> different peers, rotated by rank
    const long long bx = blockIdx.x / world, nbx = gridDim.x / world;
    const cpx* s = src + (long long)h * src_step;
    cpx* d = peers.p[h] + dst_off;
    for (long long r = blockIdx.y; r < rows; r += gridDim.y)
        for (long long c = bx * blockDim.x + threadIdx.x; c < cols; c += nbx * blockDim.x)
            d[r * dst_pitch + c] = s[r * src_pitch + c];
}
Status peer_block_copy(const cpx* src, cpx* const* peers, int world, int rank, long long rows, long long cols, long long src_step,
                       long long src_pitch, long long dst_off, long long dst_pitch, cudaStream_t st) {
    if (!src || !peers || world < 1 || world > 16 || rows < 1 || cols < 1) return invalid("peer_block_copy: bad arguments");
    PeerPtrs pp;
    for (int i = 0; i < 16; i++) pp.p[i] = i < world ? peers[i] : nullptr;
    long long gx = (cols + 255) / 256;
    if (gx > 64) gx = 64;
    long long gy = rows < 4096 ? rows : 4096;
    peer_block_copy_kernel<<<dim3((unsigned)(gx * world), (unsigned)gy, 1), 256, 0, st>>>(src, pp, rows, cols, src_step, src_pitch, dst_off, dst_pitch, world, rank);
    g_launches++;
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

Status fourstep_twiddle(cpx* blk, long long rows, long long cols, long long row0, long long col0, int log2n, cudaStream_t st, int dir) {
    if (rows < 1 || cols < 1 || log2n < 1 || log2n > 40) return invalid("fourstep_twiddle: bad arguments");
    long long threads = rows * ((cols + 15) / 16);
    fourstep_twiddle_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(blk, rows, cols, row0, col0, log2n, dir < 0 ? 1.0 : -1.0);
    g_launches++;
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}
Status repack_gkw(const cpx* in, cpx* out, long long G, long long K, long long W, cudaStream_t st) {
    if (G < 1 || K < 1 || W < 1 || in == out) return invalid("repack_gkw: bad arguments");
    long long n = G * K * W;
    repack_gkw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, G, K, W);
    g_launches++;
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

static inline unsigned grid_for(long long n, int block) {
    long long g = (n + block - 1) / block;
    return (unsigned)(g < 1 ? 1 : g);
}

Status fill_splitmix(double* out, long long n, unsigned long long seed, unsigned long long offset, cudaStream_t st) {
    if (n <= 0) return GD_OK;
    long long g = (n + 255) / 256;
    if (g > 148 * 32) g = 148 * 32;
    splitmix_kernel<<<(unsigned)g, 256, 0, st>>>(out, n, seed, offset);
    g_launches++;
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

// ------------------------------------------------------------------ Bluestein
Status Device::bluestein(long long n, cudaStream_t st, const BluesteinPlan** out) {
    auto it = blue.find(n);
    if (it == blue.end()) {
        BluesteinPlan pl;
        pl.n = n;
        long long need = 2 * n - 1;                 // dsputils.NextPowerOf2(2N-1), dsputils/dsputils.go:39-45
        pl.la = 1; pl.log2la = 0;
        while (pl.la < need) { pl.la <<= 1; pl.log2la++; }
        if (pl.log2la > 30) return invalid("bluestein: padded length > 2^30 not supported");
        GD_CUDA(cudaMalloc((void**)&pl.chirp_inv, (size_t)n * sizeof(cpx)));
        GD_CUDA(cudaMalloc((void**)&pl.bhat, (size_t)pl.la * sizeof(cpx)));
        cpx* b;
        GD_TRY(ensure_scratch(SCR_AUX, (size_t)pl.la * sizeof(cpx), (void**)&b));
        GD_CUDA(cudaMemsetAsync(b, 0, (size_t)pl.la * sizeof(cpx), st));
        chirp_kernel<<<grid_for(n, 256), 256, 0, st>>>(n, pl.la, pl.chirp_inv, b);
        g_launches++;
        GD_CUDA(cudaGetLastError());
        FusedOps none;
        GD_TRY(fft_pow2(*this, b, pl.la, pl.bhat, pl.la, pl.log2la, 1, none, st));   // FFT(b), fft/fft.go:61
        it = blue.emplace(n, pl).first;
    }
    *out = &it->second;
    return GD_OK;
}

static Status bluestein_fft(Device& d, const void* in, long long in_dist, cpx* out, long long out_dist, long long n,
                            long long batch, bool real_in, int dir, cudaStream_t st) {
    const BluesteinPlan* pl;
    GD_TRY(d.bluestein(n, st, &pl));
    const long long la = pl->la;
    if (pl->log2la <= 13 && d.bluestein_fused) {
        // the whole transform of a line in one kernel, the padded sequence stays on the SM (bluestein_small.cuh)
        BluesteinSmallParams b;
        b.in = in; b.out = out; b.in_dist = in_dist; b.out_dist = out_dist; b.n = n; b.batch = batch;
        b.chirp = pl->chirp_inv; b.bhat = pl->bhat; b.wl = pl->log2la >= 5 ? d.wl[pl->log2la] : nullptr;
        b.scale = 1.0 / (double)la; b.div = (double)n;
        b.real_in = real_in ? 1 : 0; b.inverse = dir < 0 ? 1 : 0;
        GD_TRY(d.l2_release());
        cudaError_t e = launch_bluestein_small(pl->log2la, b, d.num_sms, st);
        if (e != cudaSuccess) return cuda_fail(e, "bluestein_small_kernel launch");
        g_launches++;
        return GD_OK;
    }
    if (pl->log2la > 24 || (d.bluestein_stream && pl->log2la >= 14 && batch * la >= (1LL << 21))) {
        // The power-of-two transforms as plain forward / inverse ones -- the fused size family, the 2^20 kernel and the outer four-step
        // take no fused load / store operators -- with the chirp products, the padding, the product with FFT(b) and the truncation as
        // streaming kernels around them: 2-3x the rate of the two-launch passes with fused operators (n = 30000: 12.7 GS/s), and the
        // only formulation above a padded length of 2^24. A chunk of transforms at a time.
        long long per = (long long)(d.bluestein_chunk_bytes / ((size_t)la * sizeof(cpx)));
        if (per < 1) per = 1;
        if (per > batch) per = batch;
        if (per > 32768) per = 32768;                                          // grid.y
        cpx* A;
        GD_TRY(d.ensure_scratch(SCR_A, (size_t)per * la * sizeof(cpx), (void**)&A));
        FusedOps fwd, inv;
        inv.ld_flags = LD_CONJ; inv.st_flags = ST_CONJ | ST_SCALE; inv.scale = 1.0 / (double)la;
        L2Hold hold(d);
        for (long long b0 = 0; b0 < batch; b0 += per) {
            const long long nb = batch - b0 < per ? batch - b0 : per;
            long long gx = (la + 1023) / 1024;                                 // 4 elements per thread
            if (gx * nb > (long long)d.num_sms * 64) gx = ((long long)d.num_sms * 64 + nb - 1) / nb;
            if (gx < 1) gx = 1;
            const dim3 g((unsigned)gx, (unsigned)nb, 1);
            const void* src = real_in ? (const void*)((const double*)in + b0 * in_dist) : (const void*)((const cpx*)in + b0 * in_dist);
            bluestein_prep_kernel<<<g, 256, 0, st>>>(src, in_dist, real_in ? 1 : 0, dir < 0 ? 1 : 0, n, la, pl->chirp_inv, A);
            GD_CUDA(cudaGetLastError());
            // forward transform with the product with FFT(b) on its stores where the fused family takes the whole chunk (padded length
            // 2^14 .. 2^19, whole phases), else the transform and a product sweep
            const Tma2dEntry* te = d.bluestein_fuse_mul ? tma2d_entry(d, pl->log2la) : nullptr;
            if (te && nb % te->unit == 0 && te->rows_ok(A, la, A, la, nb, 0, 0, 1.0)) {
                Tma2dExtra ex;
                ex.aux = pl->bhat;
                GD_TRY(te->run(d, 0, A, la, A, la, nb, false, 1.0, st, ex));
            } else {
                GD_TRY(fft_pow2(d, A, la, A, la, pl->log2la, nb, fwd, st));
                bluestein_mul_kernel<<<g, 256, 0, st>>>(A, pl->bhat, la);
                GD_CUDA(cudaGetLastError());
                g_launches++;
            }
            GD_TRY(fft_pow2(d, A, la, A, la, pl->log2la, nb, inv, st));
            bluestein_post_kernel<<<g, 256, 0, st>>>(A, la, pl->chirp_inv, n, dir < 0 ? 1 : 0, (double)n, out + b0 * out_dist, out_dist);
            GD_CUDA(cudaGetLastError());
            g_launches += 2;
        }
        return GD_OK;
    }
    // chunks large enough for the two transforms' inner L2-sized blocks to rotate over the chunk streams: with 64 MiB chunks every
    // launch carried ~10 us of HBM time and the GPU idled between them (n = 4095: 8.7 GS/s)
    long long chunk = (long long)(d.bluestein_chunk_bytes / ((size_t)la * sizeof(cpx)));
    if (chunk < 1) chunk = 1;
    if (chunk > batch) chunk = batch;
    cpx* A;
    GD_TRY(d.ensure_scratch(SCR_A, (size_t)chunk * la * sizeof(cpx), (void**)&A));
    const bool inv = dir < 0;
    L2Hold hold(d);                                        // one set-aside for every chunk's two transforms
    for (long long b0 = 0; b0 < batch; b0 += chunk) {
        long long nb = batch - b0 < chunk ? batch - b0 : chunk;
        const void* src = real_in ? (const void*)((const double*)in + b0 * in_dist) : (const void*)((const cpx*)in + b0 * in_dist);
        // a = x * conj(chirp), zero-padded to la; A = FFT(a) * FFT(b)   (bluestein.go:70-76, fft.go:60-66)
        // inverse: the reference transforms the index-reversed input and divides by N (fft.go:39-50); the
        // chirp table is only accurate to ~N*eps rad, so the reversal is reproduced literally, not by conjugation
        FusedOps f1;
        f1.ld_flags = LD_PAD | LD_MULAUX | (real_in ? LD_REAL : 0) | (inv ? LD_REVERSE : 0);
        f1.aux_in = pl->chirp_inv; f1.n_valid_in = n;
        f1.st_flags = ST_MULAUX; f1.aux_out = pl->bhat;
        GD_TRY(fft_pow2(d, src, in_dist, A, la, pl->log2la, nb, f1, st));
        // r = IFFT(A) (conj . FFT . conj, / la); r[k] * conj(chirp)[k], k < N   (fft.go:35-52, bluestein.go:89-93)
        FusedOps f2;
        f2.ld_flags = LD_CONJ;
        f2.st_flags = ST_CONJ | ST_SCALE | ST_MULAUX | ST_TRUNC | (inv ? ST_DIV : 0);
        f2.scale = 1.0 / (double)la; f2.div = (double)n;
        f2.aux_out = pl->chirp_inv; f2.n_valid_out = n;
        GD_TRY(fft_pow2(d, A, la, out + b0 * out_dist, out_dist, pl->log2la, nb, f2, st));
    }
    return GD_OK;
}

// out[t][i] = (in[t][i], 0): batched real -> complex widening (transforms t = blockIdx.y, blockIdx.y + gridDim.y, ...)
__global__ void widen_real_kernel(const double* __restrict__ in, long long in_dist, cpx* __restrict__ out, long long out_dist, long long n,
                                  long long batch) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = blockIdx.y; t < batch; t += gridDim.y) {
        const double* src = in + t * in_dist;
        cpx* dst = out + t * out_dist;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = make_double2(__ldg(src + i), 0.0);
    }
}

Status fft1d(Device& d, const void* in, long long in_dist, cpx* out, long long out_dist, long long n,
             long long batch, bool real_in, int dir, cudaStream_t st) {
    if (n < 1 || batch < 1) return invalid("fft1d: bad size");
    if (n == 1) {                                          // fft/fft.go:76-80 (IFFT: x/1)
        copy1_kernel<<<grid_for(batch, 256), 256, 0, st>>>(in, in_dist, out, out_dist, batch, real_in ? 1 : 0);
        g_launches++;
        GD_CUDA(cudaGetLastError());
        return GD_OK;
    }
    if (is_pow2(n) && real_in && (const void*)in != (const void*)out) {
        // dsputils.ToComplex (dsputils/dsputils.go:25-31) as its own sweep where the transform that follows has no fused load operator:
        // the outer four-step (from 2^22 points; above 2^24 the only path) and whole phases of the fused kernels (2^13 .. 2^20)
        const int lg = ilog2ll(n);
        const int hmin = batch > 1 ? d.huge_min_log2n - 1 : d.huge_min_log2n;
        if (lg > 24 || lg >= hmin || (d.real_widen && lg >= 13 && (double)batch * (double)n >= (double)(1LL << 21))) {
            long long gx = (n + 1023) / 1024;
            const long long by = batch < 32768 ? batch : 32768;
            if (gx * by > (long long)d.num_sms * 64) gx = ((long long)d.num_sms * 64 + by - 1) / by;
            if (gx < 1) gx = 1;
            widen_real_kernel<<<dim3((unsigned)gx, (unsigned)by, 1), 256, 0, st>>>((const double*)in, in_dist, out, out_dist, n, batch);
            g_launches++;
            GD_CUDA(cudaGetLastError());
            return fft1d(d, out, out_dist, out, out_dist, n, batch, false, dir, st);
        }
    }
    if (is_pow2(n)) {
        FusedOps ops;
        ops.ld_flags = real_in ? LD_REAL : 0;
        if (dir < 0) {
            ops.ld_flags |= LD_CONJ;
            ops.st_flags = ST_CONJ | ST_SCALE;
            ops.scale = 1.0 / (double)n;                   // exact for power-of-two n: same bits as x / N
        }
        return fft_pow2(d, in, in_dist, out, out_dist, ilog2ll(n), batch, ops, st);
    }
    return bluestein_fft(d, in, in_dist, out, out_dist, n, batch, real_in, dir, st);
}

Status convolve(Device& d, const cpx* x, const cpx* y, cpx* out, long long n, cudaStream_t st) {
    if (n < 1) return invalid("convolve: bad size");
    cpx *X, *Y;
    GD_TRY(d.ensure_scratch(SCR_B, (size_t)n * sizeof(cpx), (void**)&X));
    GD_TRY(d.ensure_scratch(SCR_C, (size_t)n * sizeof(cpx), (void**)&Y));
    GD_TRY(fft1d(d, y, n, Y, n, n, 1, false, +1, st));
    if (is_pow2(n) && n >= 2 && n < (1LL << d.huge_min_log2n)) {
        // FFT(x) * FFT(y) (fft.go:63-66) fused into the store of FFT(x): ST_MULAUX with aux = FFT(y); from 2^22 points the plain
        // transform (outer four-step) and a product sweep are faster than the two passes that take the operator
        FusedOps f;
        f.st_flags = ST_MULAUX; f.aux_out = Y;
        GD_TRY(fft_pow2(d, x, n, X, n, ilog2ll(n), 1, f, st));
    } else {
        GD_TRY(fft1d(d, x, n, X, n, n, 1, false, +1, st));
        pointwise_mul_kernel<<<grid_for(n, 256), 256, 0, st>>>(X, Y, n);
        g_launches++;
        GD_CUDA(cudaGetLastError());
    }
    return fft1d(d, X, n, out, n, n, 1, false, -1, st);
}

// ------------------------------------------------------------------ linear convolution (overlap-save)
// xe[i] = x[i - lead] for lead <= i < lead + nx, else 0
__global__ void ols_extend_kernel(const cpx* __restrict__ x, long long nx, long long lead, long long total, cpx* __restrict__ xe) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long j = i - lead;
    xe[i] = (j >= 0 && j < nx) ? x[j] : make_double2(0.0, 0.0);
}
// out[b*step + j] = blocks[b*lb + lead + j], j < step: the part of every circular block result that is free of wrap-around
__global__ void ols_take_kernel(const cpx* __restrict__ blocks, long long b0, long long nb, long long lb, long long lead, long long step,
                                long long nout, cpx* __restrict__ out) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nb * step) return;
    const long long b = t / step, j = t - b * step, o = (b0 + b) * step + j;
    if (o < nout) out[o] = blocks[b * lb + lead + j];
}

// Linear convolution y[n] = sum_m h[m] x[n - m], n < nx + nh - 1, as circular convolutions of length lb on blocks of x that
// overlap by nh - 1 samples (overlap-save). Each block is what fft.Convolve computes (fft/fft.go:55-69: IFFT(FFT(a) * FFT(b)))
// with FFT(h) taken once; the overlapping blocks are read in place (rows `step` apart), the product with FFT(h) is fused
// into the forward transform's store and conj / scale into the inverse.
Status convolve_linear(Device& d, const cpx* x, long long nx, const cpx* h, long long nh, cpx* out, cudaStream_t st) {
    if (!x || !h || !out || nx < 1 || nh < 1) return invalid("convolve_linear: bad arguments");
    if (nh > nx) { const cpx* t = x; x = h; h = t; long long tn = nx; nx = nh; nh = tn; }      // convolution commutes: h is the short one
    const long long nout = nx + nh - 1, lead = nh - 1;
    if (nh > (1LL << 21)) return invalid("convolve_linear: the shorter operand must not exceed 2^21 points");
    long long lb = 4096;
    while (lb < 4 * nh) lb <<= 1;
    if (nout <= lb) { lb = 2; while (lb < nout) lb <<= 1; }
    int lg = 0;
    while ((1LL << lg) < lb) lg++;
    const long long step = lb - lead, nblocks = (nout + step - 1) / step, total = (nblocks - 1) * step + lb;
    cpx *xe, *H, *A;
    GD_TRY(d.ensure_scratch(SCR_B, (size_t)total * sizeof(cpx), (void**)&xe));
    GD_TRY(d.ensure_scratch(SCR_C, (size_t)lb * sizeof(cpx), (void**)&H));
    long long chunk = (long long)((256ull << 20) / ((size_t)lb * sizeof(cpx)));
    if (chunk < 1) chunk = 1;
    if (chunk > nblocks) chunk = nblocks;
    GD_TRY(d.ensure_scratch(SCR_A, (size_t)chunk * lb * sizeof(cpx), (void**)&A));
    ols_extend_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, nx, lead, total, xe);
    ols_extend_kernel<<<grid_for(lb, 256), 256, 0, st>>>(h, nh, 0, lb, H);
    g_launches += 2;
    GD_CUDA(cudaGetLastError());
    FusedOps none;
    GD_TRY(fft_pow2(d, H, lb, H, lb, lg, 1, none, st));                      // FFT(h), once
    for (long long b0 = 0; b0 < nblocks; b0 += chunk) {
        const long long nb = nblocks - b0 < chunk ? nblocks - b0 : chunk;
        FusedOps f1;                                                         // A = FFT(block) * FFT(h)            (fft.go:60-66)
        f1.st_flags = ST_MULAUX; f1.aux_out = H;
        GD_TRY(fft_pow2(d, xe + b0 * step, step, A, lb, lg, nb, f1, st));
        FusedOps f2;                                                         // IFFT: conj . FFT . conj, exact 1/lb (fft.go:35-52)
        f2.ld_flags = LD_CONJ; f2.st_flags = ST_CONJ | ST_SCALE; f2.scale = 1.0 / (double)lb;
        GD_TRY(fft_pow2(d, A, lb, A, lb, lg, nb, f2, st));
        ols_take_kernel<<<grid_for(nb * step, 256), 256, 0, st>>>(A, b0, nb, lb, lead, step, nout, out);
        g_launches++;
        GD_CUDA(cudaGetLastError());
    }
    return GD_OK;
}

// ------------------------------------------------------------------ N-d transforms
// one axis: lines (o, i), o < outer, i < s, element stride s, length len. src may equal dst.
// col0 / ncols: only columns [col0, col0 + ncols) of every block (power-of-two lengths above 4096 only)
static Status fft_axis(Device& d, const cpx* src, cpx* dst, long long outer, long long len, long long s, int dir, cudaStream_t st,
                       long long col0, long long ncols) {
    const long long nlines = outer * s;
    bool sub = ncols >= 0 && !(col0 == 0 && ncols == s);
    if (len == 1) {
        if (src != dst) GD_CUDA(cudaMemcpyAsync(dst, src, (size_t)nlines * sizeof(cpx), cudaMemcpyDeviceToDevice, st));
        return GD_OK;
    }
    if (s == 1) return fft1d(d, src, len, dst, len, len, outer, false, dir, st);
    const bool p2 = is_pow2(len);
    const bool fits31 = (double)len * (double)s < 2147483648.0;   // in-line offsets are 32-bit in the pass kernel
    // one strided pass: T adjacent columns per tile, T x 16 contiguous bytes per row. 4096-point lines leave 32 bytes per row, so wide
    // matrices take the blocked four-step below instead (FFT2 4096 x 16384: 2.19 -> 1.22 ms; 2048-point lines measured equal either way)
    if (p2 && fits31 && !sub && (len <= (1LL << d.axis_single_max_log2) || (len <= 4096 && s < 64))) {
        int l = ilog2ll(len);
        PassParams p = base_params(d, l);
        p.in = src; p.out = dst; p.nlines = nlines; p.inner = s;
        p.in_qs = p.out_qs = len * s; p.in_is = p.out_is = 1; p.in_es = p.out_es = (int)s;
        p.in_mode = p.out_mode = MODE_COL;
        if (dir < 0) { p.ld_flags = LD_CONJ; p.st_flags = ST_CONJ | ST_SCALE; p.scale = 1.0 / (double)len; }
        return launch_pass(d, l, p, st);
    }
    if (sub && !(p2 && len >= 4 && len <= (1LL << 24) && fits31))
        return invalid("fft_axis: a column range needs a power-of-two length in [4, 2^24]");
    if (const Tma2dEntry* te = p2 ? tma2d_entry(d, ilog2ll(len)) : nullptr) {
        // columns of a matrix with 2^13 .. 2^17 rows (2^14: fft.FFT2 on 16384 x 16384; 2^16: the line passes of the sharded
        // 2^32-point transform): whole phases of columns in one fused launch, intermediate resident in L2 (fft_tma14.cuh);
        // a remainder of less than one phase of columns goes through the column-range path below
        const long long cfirst = sub ? col0 : 0, ctotal = sub ? ncols : s;
        const long long main = ctotal - ctotal % te->unit;
        if (main > 0 && te->cols_ok(src + cfirst, dst + cfirst, len, main, s)) {
            for (long long o = 0; o < outer; o++)
                GD_TRY(te->run(d, 1, src + o * len * s + cfirst, s, dst + o * len * s + cfirst, s, main, dir < 0, dir < 0 ? 1.0 / (double)len : 1.0, st, Tma2dExtra()));
            if (main == ctotal) return GD_OK;
            col0 = cfirst + main; ncols = ctotal - main;            // the remainder: fewer columns than a phase, two-launch path below
        }
    }
    sub = ncols >= 0 && !(col0 == 0 && ncols == s);
    if (p2 && len <= (1LL << 24) && fits31) {
        // strided four-step on blocks of cb adjacent columns; the inter-pass block [len][cb] stays in L2
        const int lg = ilog2ll(len), l1 = (lg + 1) / 2, l2 = lg - l1;
        const long long R1 = 1LL << l1, R2 = 1LL << l2;
        TwiddleTable tw;
        GD_TRY(d.twiddles(lg, &tw));
        // blocks sized to stay in L2 between the passes, rotating over several streams (see fft_pow2)
        const size_t budget = d.pass_scratch_budget < d.l2_block_budget ? d.pass_scratch_budget : d.l2_block_budget;
        long long cb = (long long)(budget / ((size_t)len * sizeof(cpx)));
        if (cb < 8) cb = 8;
        if (cb > s) cb = s;
        const long long cfirst = sub ? col0 : 0, clast = sub ? col0 + ncols : s;
        const long long nblocks = outer * ((clast - cfirst + cb - 1) / cb);
        ForkJoin fj(d, st, nblocks >= 2 * d.chunk_streams ? d.chunk_streams : 1);
        cpx* scr0;
        GD_TRY(d.ensure_scratch(SCR_PASS, (size_t)fj.ways * len * cb * sizeof(cpx), (void**)&scr0));
        if (nblocks > 1) fj.persist(scr0, (size_t)fj.ways * len * cb * sizeof(cpx));
        long long ci = 0;
        for (long long o = 0; o < outer; o++)
            for (long long c0 = cfirst; c0 < clast; c0 += cb, ci++) {
                long long nc = clast - c0 < cb ? clast - c0 : cb;
                cudaStream_t st = fj.stream(ci);
                cpx* scr = scr0 + (size_t)fj.way(ci) * len * cb;
                const cpx* sp = src + o * len * s + c0;
                cpx* dp = dst + o * len * s + c0;
                // pass 1: lines (n2, c): length R1 over n1 (stride R2*s); out block[(k1*R2 + n2)][c]
                PassParams p = base_params(d, l1);
                p.in = sp; p.out = scr; p.nlines = R2 * nc; p.inner = nc;
                p.in_qs = s; p.in_is = 1; p.in_es = (int)(R2 * s);
                p.out_qs = nc; p.out_is = 1; p.out_es = (int)(R2 * nc);
                p.in_mode = p.out_mode = MODE_COL;
                if (dir < 0) p.ld_flags = LD_CONJ;
                p.st_flags = ST_TWIDDLE; p.tw_sel = 1; p.tw_log2m = lg; p.tw_lo = tw.lo; p.tw_hi = tw.hi;
                GD_TRY(launch_pass(d, l1, p, st));
                // pass 2: lines (k1, c): length R2 over n2; out row (k1 + R1*k2)
                PassParams r = base_params(d, l2);
                r.in = scr; r.out = dp; r.nlines = R1 * nc; r.inner = nc;
                r.in_qs = R2 * nc; r.in_is = 1; r.in_es = (int)nc;
                r.out_qs = s; r.out_is = 1; r.out_es = (int)(R1 * s);
                r.in_mode = r.out_mode = MODE_COL;
                if (dir < 0) { r.st_flags = ST_CONJ | ST_SCALE; r.scale = 1.0 / (double)len; }
                GD_TRY(launch_pass(d, l2, r, st));
            }
        return GD_OK;
    }
    // any other length (Bluestein lines): gather to dense lines, transform, scatter back
    long long chunk = (long long)(d.bluestein_chunk_bytes / 2 / ((size_t)len * sizeof(cpx)));
    if (chunk < 1) chunk = 1;
    if (chunk > nlines) chunk = nlines;
    cpx* buf;
    GD_TRY(d.ensure_scratch(SCR_B, (size_t)chunk * len * sizeof(cpx), (void**)&buf));
    for (long long l0 = 0; l0 < nlines; l0 += chunk) {
        long long nl = nlines - l0 < chunk ? nlines - l0 : chunk;
        const bool tiled = (len + 31) / 32 <= 65535;
        const dim3 tg((unsigned)((nl + 31) / 32), (unsigned)((len + 31) / 32), 1);
        if (tiled) lines_tiled_kernel<false><<<tg, 256, 0, st>>>(src, buf, l0, nl, len, s);
        else gather_lines_kernel<<<grid_for(nl * len, 256), 256, 0, st>>>(src, buf, l0, nl, len, s);
        g_launches++;
        GD_CUDA(cudaGetLastError());
        GD_TRY(fft1d(d, buf, len, buf, len, len, nl, false, dir, st));
        if (tiled) lines_tiled_kernel<true><<<tg, 256, 0, st>>>(buf, dst, l0, nl, len, s);
        else scatter_lines_kernel<<<grid_for(nl * len, 256), 256, 0, st>>>(buf, dst, l0, nl, len, s);
        g_launches++;
        GD_CUDA(cudaGetLastError());
    }
    return GD_OK;
}

Status fft_strided(Device& d, const cpx* src, cpx* dst, long long outer, long long len, long long s, int dir, cudaStream_t st) {
    if (outer < 1 || len < 1 || s < 1) return invalid("fft_strided: bad arguments");
    return fft_axis(d, src, dst, outer, len, s, dir, st);
}

Status fftn(Device& d, const cpx* in, cpx* out, const long long* dims, int nd, int dir, cudaStream_t st) {
    if (nd < 1 || nd > 16) return invalid("fftn: bad rank");
    long long total = 1;
    for (int i = 0; i < nd; i++) { if (dims[i] < 1) return invalid("fftn: invalid dimensions"); total *= dims[i]; }
    // the reference sweeps axis 0 first (fft/fft.go:175-189); FFT2 does columns (axis 0) then rows (fft.go:138-151)
    const cpx* src = in;
    long long outer = 1;
    for (int a = 0; a < nd; a++) {
        long long s = total / (outer * dims[a]);
        GD_TRY(fft_axis(d, src, out, outer, dims[a], s, dir, st));
        src = out;
        outer *= dims[a];
    }
    return GD_OK;
}

}  // namespace gd

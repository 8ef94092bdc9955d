// Stand-alone harness for the TMA-fed four-step kernels (go-dsp_b200/csrc/fft_tma.cuh): batched 2^20-point
// complex128 FFT = pass 1 + pass 2, CUDA-event timing, check of two transforms against a host radix-2 FFT.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I../../go-dsp_b200/csrc tma_pass.cu -o tma_pass
// usage: tma_pass [batch=64] [alias_mask=-1] [iters=5] [grid=148] [fused=0] [delay=2] [persist=1]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>
#include <algorithm>
#include "fft_tma_legacy.cuh"
using namespace gd;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(1); }
    return (EncodeFn)fn;
}
// [batch][1024 rows][1024 complex] viewed as doubles: dims {2048, 1024, batch}; box {8, 256, 1}
static CUtensorMap make_map(EncodeFn enc, void* base, long long batch, long long dist_elems, int promo = 0) {
    CUtensorMap m;
    cuuint64_t dims[3] = {2048, 1024, (cuuint64_t)batch};
    cuuint64_t strides[2] = {16384, (cuuint64_t)dist_elems * 16};
    cuuint32_t box[3] = {2 * TMA_T, TMA_BOX_ROWS, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}
__global__ void fill(double* x, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long z = (unsigned long long)(i + 1) * 0x9E3779B97F4A7C15ULL;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
        x[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
    }
}
static void host_fft(std::vector<std::complex<double>>& a) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; i++) { size_t bit = n >> 1; for (; j & bit; bit >>= 1) j ^= bit; j ^= bit; if (i < j) std::swap(a[i], a[j]); }
    for (size_t len = 2; len <= n; len <<= 1) {
        std::vector<std::complex<double>> w(len / 2);
        for (size_t k = 0; k < len / 2; k++) { long double ang = -2.0L * M_PIl * k / len; w[k] = {(double)cosl(ang), (double)sinl(ang)}; }
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; k++) { auto u = a[i + k], v = a[i + k + len / 2] * w[k]; a[i + k] = u + v; a[i + k + len / 2] = u - v; }
    }
}

int main(int argc, char** argv) {
    long long batch = argc > 1 ? atoll(argv[1]) : 64;
    int mask = argc > 2 ? atoi(argv[2]) : -1;
    int iters = argc > 3 ? atoi(argv[3]) : 5;
    int grid = argc > 4 ? atoi(argv[4]) : 148;
    int fused = argc > 5 ? atoi(argv[5]) : 0;
    int delay = argc > 6 ? atoi(argv[6]) : 2;
    int persist = argc > 7 ? atoi(argv[7]) : 1;
    const long long N = 1 << 20;
    long long nbuf = mask >= 0 ? (mask + 1) : batch;
    cpx *x, *mid, *out;
    CK(cudaMalloc(&x, nbuf * N * 16)); CK(cudaMalloc(&mid, nbuf * N * 16)); CK(cudaMalloc(&out, nbuf * N * 16));
    fill<<<1184, 256>>>((double*)x, nbuf * N * 2);
    CK(cudaMemset(out, 0, nbuf * N * 16));
    // tables
    std::vector<cpx> wl(1024), lo(4096), hi(256);
    for (int p = 0; p < 1024; p++) { long double a = -2.0L * M_PIl * p / 1024; wl[p] = make_double2((double)cosl(a), (double)sinl(a)); }
    for (int e = 0; e < 4096; e++) { long double a = -2.0L * M_PIl * e / N; lo[e] = make_double2((double)cosl(a), (double)sinl(a)); }
    for (int e = 0; e < 256; e++) { long double a = -2.0L * M_PIl * (e * 4096.0L) / N; hi[e] = make_double2((double)cosl(a), (double)sinl(a)); }
    cpx *dwl, *dlo, *dhi;
    CK(cudaMalloc(&dwl, 1024 * 16)); CK(cudaMalloc(&dlo, 4096 * 16)); CK(cudaMalloc(&dhi, 256 * 16));
    CK(cudaMemcpy(dwl, wl.data(), 1024 * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dlo, lo.data(), 4096 * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dhi, hi.data(), 256 * 16, cudaMemcpyHostToDevice));
    EncodeFn enc = get_encode();
    CUtensorMap mx = make_map(enc, x, nbuf, N), mm = make_map(enc, mid, nbuf, N), mo = make_map(enc, out, nbuf, N);
    CK(cudaFuncSetAttribute(fft_tma_kernel<TMA_OUT_ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM));
    CK(cudaFuncSetAttribute(fft_tma_kernel<TMA_OUT_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM));
    TmaPassParams p1{}, p2{};
    p1.ntiles = batch * 256; p1.out = mid; p1.out_dist = N; p1.wl = dwl; p1.tw_lo = dlo; p1.tw_hi = dhi; p1.tw_log2m = 20;
    p1.scale = 1.0; p1.tf_mask = mask;
    p1.dbg_noload = getenv("TMA_NOLOAD") ? 1 : 0; p1.dbg_nostore = getenv("TMA_NOSTORE") ? 1 : 0;
    p2 = p1; p2.out = nullptr; p2.tw_lo = nullptr; p2.tw_hi = nullptr;
    auto run = [&]() {
        fft_tma_kernel<TMA_OUT_ROWS><<<grid, TMA_THREADS, TMA_SMEM>>>(mx, mx, p1);
        fft_tma_kernel<TMA_OUT_TILE><<<grid, TMA_THREADS, TMA_SMEM>>>(mm, mo, p2);
    };
    run();
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    float best = 1e9, b1 = 1e9, b2 = 1e9;
    for (int i = 0; i < iters; i++) {
        cudaEventRecord(e0);
        fft_tma_kernel<TMA_OUT_ROWS><<<grid, TMA_THREADS, TMA_SMEM>>>(mx, mx, p1);
        cudaEventRecord(e1);
        fft_tma_kernel<TMA_OUT_TILE><<<grid, TMA_THREADS, TMA_SMEM>>>(mm, mo, p2);
        cudaEventRecord(e2);
        CK(cudaDeviceSynchronize());
        float m1, m2; cudaEventElapsedTime(&m1, e0, e1); cudaEventElapsedTime(&m2, e1, e2);
        if (m1 + m2 < best) best = m1 + m2;
        if (m1 < b1) b1 = m1;
        if (m2 < b2) b2 = m2;
    }
    printf("batch %lld mask %d grid %d: pass1 %.3f ms (%.1f Gpt/s)  pass2 %.3f ms (%.1f Gpt/s)  total %.3f ms = %.1f GS/s\n", batch, mask, grid,
           b1, batch * N / b1 / 1e6, b2, batch * N / b2 / 1e6, best, batch * N / best / 1e6);
    // check (only meaningful without aliasing)
    if (mask < 0) {
        for (long long tf : {0LL, batch - 1}) {
            std::vector<std::complex<double>> h(N), g(N);
            CK(cudaMemcpy(h.data(), x + tf * N, N * 16, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(g.data(), out + tf * N, N * 16, cudaMemcpyDeviceToHost));
            host_fft(h);
            long double num = 0, den = 0;
            for (long long i = 0; i < N; i++) { num += std::norm(g[i] - h[i]); den += std::norm(h[i]); }
            printf("transform %lld: rel L2 err %.3e\n", tf, (double)sqrtl(num / den));
        }
    }
    if (fused) {
        const int S = delay + 2;
        cpx* scratch; int* cnt;
        CK(cudaMalloc(&scratch, (size_t)S * N * 16));
        CK(cudaMalloc(&cnt, (2 * batch + 2) * sizeof(int)));
        const int promo = getenv("TMA_PROMO") ? atoi(getenv("TMA_PROMO")) : 0;
        CUtensorMap ms = make_map(enc, scratch, S, N, getenv("TMA_PROMO_INT") ? atoi(getenv("TMA_PROMO_INT")) : 0);
        mx = make_map(enc, x, nbuf, N, promo);
        CK(cudaFuncSetAttribute(fft_tma_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM));
        CK(cudaFuncSetAttribute(fft_tma_fused2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA2_SMEM));
        CK(cudaFuncSetAttribute(fft_tma_fused2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA2_SMEM));
        cudaStream_t st; CK(cudaStreamCreate(&st));
        if (persist) {
            int maxp = 0, maxw = 0;
            cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, 0);
            cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, 0);
            size_t want = (size_t)S * N * 16;
            printf("persisting L2 max %d MiB, window max %d MiB, want %zu MiB\n", maxp >> 20, maxw >> 20, want >> 20);
            CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want < (size_t)maxp ? want : (size_t)maxp));
            cudaStreamAttrValue v{};
            v.accessPolicyWindow.base_ptr = scratch;
            v.accessPolicyWindow.num_bytes = want < (size_t)maxw ? want : (size_t)maxw;
            v.accessPolicyWindow.hitRatio = 1.0f;
            v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v));
        }
        TmaFusedParams f{};
        f.batch = (int)batch; f.delay = delay; f.nslots = S; f.scratch = scratch; f.done1 = cnt; f.done2 = cnt + batch; f.queue = cnt + 2 * batch;
        f.wl = dwl; f.tw_lo = dlo; f.tw_hi = dhi; f.tw_log2m = 20; f.scale = 1.0;
        long long* dstats; CK(cudaMalloc(&dstats, grid * 12 * sizeof(long long))); CK(cudaMemset(dstats, 0, grid * 12 * sizeof(long long)));
        if (getenv("TMA_STATS")) f.stats = dstats;
        f.hints = getenv("TMA_HINTS") ? atoi(getenv("TMA_HINTS")) : 0;
        f.p2_stg = getenv("TMA_P2STG") ? atoi(getenv("TMA_P2STG")) : 0;
        f.out = out; f.out_dist = N;
        f.dbg_nodeps = getenv("TMA_NODEPS") ? atoi(getenv("TMA_NODEPS")) : 0;
        f.dbg_nop1st = getenv("TMA_NOP1ST") ? 1 : 0;
        f.dbg_nop2st = getenv("TMA_NOP2ST") ? 1 : 0;
        f.two_queues = getenv("TMA_2Q") ? atoi(getenv("TMA_2Q")) : 0;
        f.dbg_acqload = getenv("TMA_ACQLOAD") ? 1 : 0;
        f.dbg_nopubfence = getenv("TMA_NOPUBFENCE") ? 1 : 0;
        f.dbg_out_alias = getenv("TMA_OUTALIAS") ? ~1 : 0;
        f.dbg_in_alias = getenv("TMA_INALIAS") ? ~1 : 0;
        f.dbg_noload = getenv("TMA_NOLOAD") ? 1 : 0;
        CK(cudaMemsetAsync(out, 0, nbuf * N * 16, st));
        float fb = 1e9;
        for (int i = 0; i < iters + 1; i++) {
            CK(cudaMemsetAsync(cnt, 0, (2 * batch + 2) * sizeof(int), st));
            cudaEventRecord(e0, st);
            if (fused == 3) fft_tma_fused2_kernel<true><<<grid, TMA_THREADS, TMA2_SMEM, st>>>(mx, ms, mo, f);
            else if (fused == 2) fft_tma_fused2_kernel<false><<<grid, TMA_THREADS, TMA2_SMEM, st>>>(mx, ms, mo, f);
            else fft_tma_fused_kernel<<<grid, TMA_THREADS, TMA_SMEM, st>>>(mx, ms, mo, f);
            cudaEventRecord(e1, st);
            CK(cudaStreamSynchronize(st));
            float m; cudaEventElapsedTime(&m, e0, e1);
            if (i > 0 && m < fb) fb = m;
        }
        printf("FUSED%d batch %lld delay %d grid %d persist %d: %.3f ms = %.1f GS/s\n", fused, batch, delay, grid, persist, fb, batch * N / fb / 1e6);
        if (f.stats) {
            std::vector<long long> hs(grid * 12);
            CK(cudaMemcpy(hs.data(), dstats, grid * 12 * sizeof(long long), cudaMemcpyDeviceToHost));
            {
                double l1 = 0, n1 = 0, l2 = 0, n2 = 0;
                for (int c = 0; c < grid; c++) { l1 += hs[grid * 8 + c * 4]; n1 += hs[grid * 8 + c * 4 + 1]; l2 += hs[grid * 8 + c * 4 + 2]; n2 += hs[grid * 8 + c * 4 + 3]; }
                printf("  load latency issue->landed: P1 tiles (x from HBM) %.0f cycles avg, P2 tiles (Int from L2) %.0f cycles avg\n", l1 / n1, l2 / n2);
            }
            double avg[8] = {0};
            for (int c = 0; c < grid; c++) for (int k = 0; k < 8; k++) avg[k] += (double)hs[c * 8 + k] / grid;
            printf("  loader cycles: total %.0f  wait-buffer %.0f  (storer: store+drain %.0f)  dep-P1done %.0f  dep-slot %.0f | consumer wait-full g0 %.0f g1 %.0f\n",
                   avg[0], avg[1], avg[2], avg[3], avg[4], avg[5], avg[6]);
            printf("  tiles taken per CTA: avg %.1f\n", avg[7]);
            for (int k : {3, 5, 7}) {
                std::vector<long long> v(grid);
                for (int c = 0; c < grid; c++) v[c] = hs[c * 8 + k];
                std::vector<long long> so = v; std::sort(so.begin(), so.end());
                printf("  stat %d: min %lld  p10 %lld  median %lld  p90 %lld  max %lld | first CTAs:", k, so[0], so[grid / 10], so[grid / 2], so[grid * 9 / 10], so[grid - 1]);
                for (int c = 0; c < 12 && c < grid; c++) printf(" %lld", v[c] / 1000);
                printf(" (k cycles)\n");
            }
        }
        for (long long tf : {0LL, batch / 2, batch - 1}) {
            std::vector<std::complex<double>> h(N), g(N);
            CK(cudaMemcpy(h.data(), x + tf * N, N * 16, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(g.data(), out + tf * N, N * 16, cudaMemcpyDeviceToHost));
            host_fft(h);
            long double num = 0, den = 0;
            for (long long i = 0; i < N; i++) { num += std::norm(g[i] - h[i]); den += std::norm(h[i]); }
            printf("fused transform %lld: rel L2 err %.3e\n", tf, (double)sqrtl(num / den));
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

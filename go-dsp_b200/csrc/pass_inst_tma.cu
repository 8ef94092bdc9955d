// Host side of the TMA-fed fused four-step (fft_tma.cuh): tensor maps, scratch slots under a persisting L2
// window, dependency counters, one persistent launch per chunk of at most 128 transforms of 2^20 points.
#include <string.h>
#include "engine.h"
#include "fft_tma.cuh"

namespace gd {

static Status invalid(const char* msg) { set_error(msg); return GD_ERR_INVALID; }

typedef CUresult (*TmaEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup: no link-time dependency on libcuda
static Status tma_encoder(TmaEncodeFn* out) {
    static TmaEncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        GD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) return invalid("cuTensorMapEncodeTiled is not available in this driver");
        fn = (TmaEncodeFn)p;
    }
    *out = fn;
    return GD_OK;
}

// [count][1024 rows][1024 complex] with transform pitch `dist` elements, as a rank-3 tensor of doubles;
// box = one quarter of a tile: 8 doubles (4 adjacent complex columns) x 256 rows
static Status tma_map(TmaEncodeFn enc, const void* base, long long count, long long dist, CUtensorMap* m) {
    cuuint64_t dims[3] = {2 * TMA_L, TMA_L, (cuuint64_t)count};
    cuuint64_t strides[2] = {TMA_L * 16, (cuuint64_t)dist * 16};
    cuuint32_t box[3] = {2 * TMA_T, TMA_BOX_ROWS, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return invalid("cuTensorMapEncodeTiled failed (pointer alignment or pitch?)");
    return GD_OK;
}

bool tma_fused_applicable(const void* in, long long in_dist, const cpx* out, long long out_dist, int ld_conj, int st_conj, double scale) {
    // forward (no hooks) or inverse (conj . forward . conj, scale folded into the twiddle); TMA needs 16-byte aligned
    // bases and pitches (always true for complex128) and pitches below 2^40 bytes
    const bool fwd = !ld_conj && !st_conj && scale == 1.0, inv = ld_conj && st_conj;
    return (fwd || inv) && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0 && in_dist >= (long long)TMA_L * TMA_L &&
           out_dist >= (long long)TMA_L * TMA_L && in_dist < (1LL << 36) && out_dist < (1LL << 36);
}

template <bool INV, bool PROF>
static cudaError_t launch_one(int grid, const CUtensorMap& mx, const CUtensorMap& mi, const CUtensorMap& mo, const TmaFusedParams& f,
                              cudaStream_t st) {
    // the attribute is per device: set it on every call (cheap), a process may drive several GPUs
    cudaError_t e = cudaFuncSetAttribute(fft_tma_fused_kernel<INV, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM);
    if (e != cudaSuccess) return e;
    fft_tma_fused_kernel<INV, PROF><<<grid, TMA_THREADS, TMA_SMEM, st>>>(mx, mi, mo, f);
    return cudaGetLastError();
}

// batched butterflies of N = 2^20 points; inverse = (ld_conj, st_conj, scale = 1/N)
Status fft_tma_2p20(Device& d, const cpx* in, long long in_dist, cpx* out, long long out_dist, long long batch, int ld_conj,
                    int st_conj, double scale, cudaStream_t st) {
    const bool inv = ld_conj && st_conj;
    TmaEncodeFn enc;
    GD_TRY(tma_encoder(&enc));
    const long long N = (long long)TMA_L * TMA_L;
    const int S = d.tma_slots;
    const int D = d.tma_delay < S - 1 ? d.tma_delay : S - 1;      // P2(g - S) must precede P1(g) in the sequence: D <= S - 1
    TwiddleTable tw;
    GD_TRY(d.twiddles(20, &tw));
    cpx* scratch;
    const size_t scr_bytes = (size_t)S * N * sizeof(cpx);
    GD_TRY(d.ensure_scratch(SCR_TMA, scr_bytes, (void**)&scratch));
    const long long CH = 512;               // transforms per launch (one tensor map per array; every launch ramps up and drains 3 phases)
    int* cnt;
    GD_TRY(d.ensure_scratch(SCR_CNT, (2 * (size_t)CH + 2) * sizeof(int), (void**)&cnt));
    long long* prof = nullptr;
    if (d.tma_prof) {
        GD_TRY(d.ensure_scratch(SCR_PROF, (size_t)d.num_sms * TMA_PROF_SLOTS * sizeof(long long), (void**)&prof));
        GD_CUDA(cudaMemsetAsync(prof, 0, (size_t)d.num_sms * TMA_PROF_SLOTS * sizeof(long long), st));
    }
    CUtensorMap m_int;
    GD_TRY(tma_map(enc, scratch, S, N, &m_int));
    // keep the scratch slots resident in L2: persisting access-policy window while the launches run
    const bool window = d.use_l2_window && d.l2_persist_max > 0 && d.l2_window_max > 0;
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    if (window) {
        size_t want = scr_bytes < d.l2_persist_max ? scr_bytes : d.l2_persist_max;
        if (d.l2_carved != want) {
            GD_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
            d.l2_carved = want;
        }
        attr.accessPolicyWindow.base_ptr = scratch;
        attr.accessPolicyWindow.num_bytes = scr_bytes < d.l2_window_max ? scr_bytes : d.l2_window_max;
        double ratio = (double)d.l2_carved / (double)attr.accessPolicyWindow.num_bytes;
        attr.accessPolicyWindow.hitRatio = (float)(ratio > 1.0 ? 1.0 : ratio);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        GD_CUDA(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
        d.l2_dirty = true;
    }
    Status rc = GD_OK;
    for (long long b0 = 0; b0 < batch && rc == GD_OK; b0 += CH) {
        const long long nb = batch - b0 < CH ? batch - b0 : CH;
        CUtensorMap m_x, m_out;
        if ((rc = tma_map(enc, in + b0 * in_dist, nb, in_dist, &m_x)) != GD_OK) break;
        if ((rc = tma_map(enc, out + b0 * out_dist, nb, out_dist, &m_out)) != GD_OK) break;
        TmaFusedParams f;
        memset(&f, 0, sizeof(f));
        f.batch = (int)nb; f.delay = D; f.nslots = S; f.scratch = scratch;
        f.done1 = cnt; f.done2 = cnt + CH; f.queue = cnt + 2 * CH;
        f.wl = d.wl[10]; f.tw_lo = tw.lo; f.tw_hi = tw.hi; f.tw_log2m = 20;
        f.scale = scale;
        f.opt = d.tma_opt;
        f.prof = prof;
        cudaError_t e = cudaMemsetAsync(cnt, 0, (2 * (size_t)CH + 2) * sizeof(int), st);
        if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemsetAsync(counters)"); break; }
        const long long nitems = 2 * nb * (TMA_L / TMA_T);
        const int sms = d.tma_grid_cap > 0 && d.tma_grid_cap < d.num_sms ? d.tma_grid_cap : d.num_sms;
        const int grid = (int)(nitems < sms ? nitems : sms);
        if (prof) e = inv ? launch_one<true, true>(grid, m_x, m_int, m_out, f, st) : launch_one<false, true>(grid, m_x, m_int, m_out, f, st);
        else e = inv ? launch_one<true, false>(grid, m_x, m_int, m_out, f, st) : launch_one<false, false>(grid, m_x, m_int, m_out, f, st);
        if (e != cudaSuccess) { rc = cuda_fail(e, "fft_tma_fused_kernel launch"); break; }
        g_launches++;
    }
    if (window) {
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    }
    return rc;
}

}  // namespace gd

// Host side of the TMA-fed fused four-step for 2^14- and 2^16-point transforms (fft_tma14.cuh, N = LEN x LEN with LEN = 128
// or 256): rows of a batch or columns of a row-major matrix; tensor maps, scratch slots under a persisting L2 window,
// dependency counters, one persistent launch per up to 512 phases. Included by one translation unit per LEN.
#pragma once
#include <math.h>
#include <string.h>
#include "engine.h"
#include "fft_tma14.cuh"

namespace gd {

static Status invalid14(const char* msg) { set_error(msg); return GD_ERR_INVALID; }

typedef CUresult (*TmaEncodeFn14)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static Status encoder14(TmaEncodeFn14* out) {
    static TmaEncodeFn14 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        GD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) return invalid14("cuTensorMapEncodeTiled is not available in this driver");
        fn = (TmaEncodeFn14)p;
    }
    *out = fn;
    return GD_OK;
}

// rank-4 map over doubles: dims d0..d3 (d0 in doubles), byte strides s1..s3 of dims 1..3, box b0..b3
static Status map4(TmaEncodeFn14 enc, const void* base, const cuuint64_t (&dims)[4], const cuuint64_t (&strides)[3], const cuuint32_t (&box)[4],
                   CUtensorMap* m) {
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return invalid14("cuTensorMapEncodeTiled failed (2^14 fused kernel: pointer alignment or pitch?)");
    return GD_OK;
}

template <int LEN>
static bool tma2d_rows_applicable(const void* in, long long in_dist, const cpx* out, long long out_dist, long long batch, int ld_conj, int st_conj,
                                  double scale) {
    using SH = T14Shape<LEN>;
    const long long N = (long long)LEN * LEN;
    const bool fwd = !ld_conj && !st_conj && scale == 1.0, inv = ld_conj && st_conj;
    return (fwd || inv) && batch >= 2 * SH::TPP && batch % SH::TPP == 0 && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0 && in_dist >= N &&
           out_dist >= N && in_dist < (1LL << 35) && out_dist < (1LL << 35);
}
template <int LEN>
static bool tma2d_cols_applicable(const cpx* src, const cpx* dst, long long len, long long s) {
    using SH = T14Shape<LEN>;
    return len == (long long)LEN * LEN && s >= 2 * SH::CPP && s % SH::CPP == 0 && s < (1LL << 30) && ((uintptr_t)src % 16) == 0 &&
           ((uintptr_t)dst % 16) == 0;
}

template <int LEN, int MODE, bool INV, bool PROF = false>
static cudaError_t launch14(int grid, const CUtensorMap& mx, const CUtensorMap& mi, const CUtensorMap& mo, const Tma14Params& f, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(fft_tma14_kernel<LEN, MODE, INV, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, T14_SMEM);   // per device
    if (e != cudaSuccess) return e;
    fft_tma14_kernel<LEN, MODE, INV, PROF><<<grid, TMA_THREADS, T14_SMEM, st>>>(mx, mi, mo, f);
    return cudaGetLastError();
}

// mode ROWS: `count` transforms of N = LEN^2 points, transform t at in + t * in_dist / out + t * out_dist (count % TPP == 0).
// mode COLS: the N rows of a row-major matrix with `count` columns (count % CPP == 0): every column is a transform;
//            in_dist / out_dist are ignored (row pitch = count).
template <int LEN>
static Status fft_tma_2d(Device& d, int mode, const cpx* in, long long in_dist, cpx* out, long long out_dist, long long count, bool inv, double scale,
                         cudaStream_t st) {
    using SH = T14Shape<LEN>;
    constexpr cuuint64_t L = LEN, LINES = SH::LINES, N = (cuuint64_t)LEN * LEN;
    TmaEncodeFn14 enc;
    GD_TRY(encoder14(&enc));
    const int S = d.tma_slots;
    const int D = d.tma_delay < S - 1 ? d.tma_delay : S - 1;
    TwiddleTable tw;
    GD_TRY(d.twiddles(SH::LOG2N, &tw));
    cpx* scratch;
    const size_t slot_elems = (size_t)1 << 20, scr_bytes = (size_t)S * slot_elems * sizeof(cpx);
    GD_TRY(d.ensure_scratch(SCR_TMA, scr_bytes, (void**)&scratch));
    const long long CH = 512;
    int* cnt;
    GD_TRY(d.ensure_scratch(SCR_CNT, (2 * (size_t)CH + 2) * sizeof(int), (void**)&cnt));
    CUtensorMap m_int;
    if (mode == T14_ROWS) {
        // Int[t][n2][k1], t < TPP * S: dims (k1 in doubles... the tile box is LINES adjacent elements of a row x LEN / 2 rows
        const cuuint64_t dims[4] = {2 * L, L, (cuuint64_t)(SH::TPP * S), 1};
        const cuuint64_t str[3] = {L * 16, N * 16, N * 16 * (cuuint64_t)(SH::TPP * S)};
        const cuuint32_t box[4] = {(cuuint32_t)(2 * LINES), (cuuint32_t)(L / 2), 1, 1};
        GD_TRY(map4(enc, scratch, dims, str, box, &m_int));
    } else {
        // Int[tb][n2][k1][LINES t], tb < TBP * S
        const cuuint64_t dims[4] = {2 * LINES, L, L, (cuuint64_t)(SH::TBP * S)};
        const cuuint64_t str[3] = {LINES * 16, LINES * 16 * L, LINES * 16 * L * L};
        const cuuint32_t box[4] = {(cuuint32_t)(2 * LINES), 1, (cuuint32_t)(L / 2), 1};
        GD_TRY(map4(enc, scratch, dims, str, box, &m_int));
    }
    Tma14Params f;
    memset(&f, 0, sizeof(f));
    f.wl = LEN == 256 ? d.wl[8] : nullptr;
    for (int j = 1; j < 4 && LEN == 128; j++)
        for (int k = 0; k < 32; k++) {
            const int e = (j * k) % 128;
            long double c = 1, s = 0;
            if (e == 0) { c = 1; s = 0; } else if (e == 32) { c = 0; s = 1; } else if (e == 64) { c = -1; s = 0; } else if (e == 96) { c = 0; s = -1; }
            else { const long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)e / 128.0L; c = cosl(ang); s = sinl(ang); }
            f.w128[j - 1][k] = make_double2((double)c, (double)(-s));
        }
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
    const long long UNIT = mode == T14_ROWS ? SH::TPP : SH::CPP;      // transforms / columns per phase
    const long long groups = count / UNIT;
    for (long long g0 = 0; g0 < groups && rc == GD_OK; g0 += CH) {
        const long long ng = groups - g0 < CH ? groups - g0 : CH;
        CUtensorMap m_x, m_out;
        if (mode == T14_ROWS) {
            const cuuint64_t dims[4] = {2 * L, L, (cuuint64_t)(ng * UNIT), 1};
            const cuuint32_t box[4] = {(cuuint32_t)(2 * LINES), (cuuint32_t)(L / 2), 1, 1};
            const cuuint64_t sx[3] = {L * 16, (cuuint64_t)in_dist * 16, (cuuint64_t)in_dist * 16 * (cuuint64_t)(ng * UNIT)};
            const cuuint64_t so[3] = {L * 16, (cuuint64_t)out_dist * 16, (cuuint64_t)out_dist * 16 * (cuuint64_t)(ng * UNIT)};
            if ((rc = map4(enc, in + g0 * UNIT * in_dist, dims, sx, box, &m_x)) != GD_OK) break;
            if ((rc = map4(enc, out + g0 * UNIT * out_dist, dims, so, box, &m_out)) != GD_OK) break;
        } else {
            const cuuint64_t dims[4] = {(cuuint64_t)(2 * ng * UNIT), L, L, 1};
            const cuuint32_t box[4] = {(cuuint32_t)(2 * LINES), 1, (cuuint32_t)(L / 2), 1};
            const cuuint64_t str[3] = {(cuuint64_t)count * 16, (cuuint64_t)count * 16 * L, (cuuint64_t)count * 16 * L * L};
            if ((rc = map4(enc, in + g0 * UNIT, dims, str, box, &m_x)) != GD_OK) break;
            if ((rc = map4(enc, out + g0 * UNIT, dims, str, box, &m_out)) != GD_OK) break;
        }
        f.batch = (int)ng; f.delay = D; f.nslots = S; f.scratch = scratch;
        f.done1 = cnt; f.done2 = cnt + CH; f.queue = cnt + 2 * CH;
        f.tw_lo = tw.lo; f.tw_hi = tw.hi; f.scale = scale;
        f.out = mode == T14_ROWS ? out + g0 * UNIT * out_dist : out + g0 * UNIT;
        f.out_dist = mode == T14_ROWS ? out_dist : count;
        f.opt = d.tma_opt >> 4;
        f.prof = nullptr;
        if (d.tma_prof) {
            long long* pr;
            GD_TRY(d.ensure_scratch(SCR_PROF, (size_t)d.num_sms * TMA_PROF_SLOTS * sizeof(long long), (void**)&pr));
            GD_CUDA(cudaMemsetAsync(pr, 0, (size_t)d.num_sms * TMA_PROF_SLOTS * sizeof(long long), st));
            f.prof = pr;
        }
        cudaError_t e = cudaMemsetAsync(cnt, 0, (2 * (size_t)CH + 2) * sizeof(int), st);
        if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemsetAsync(counters)"); break; }
        const long long nitems = 2 * ng * 256;
        const int grid = (int)(nitems < d.num_sms ? nitems : d.num_sms);
        if (f.prof && !inv) e = mode == T14_ROWS ? launch14<LEN, T14_ROWS, false, true>(grid, m_x, m_int, m_out, f, st) : launch14<LEN, T14_COLS, false, true>(grid, m_x, m_int, m_out, f, st);
        else if (mode == T14_ROWS) e = inv ? launch14<LEN, T14_ROWS, true>(grid, m_x, m_int, m_out, f, st) : launch14<LEN, T14_ROWS, false>(grid, m_x, m_int, m_out, f, st);
        else e = inv ? launch14<LEN, T14_COLS, true>(grid, m_x, m_int, m_out, f, st) : launch14<LEN, T14_COLS, false>(grid, m_x, m_int, m_out, f, st);
        if (e != cudaSuccess) { rc = cuda_fail(e, "fft_tma14_kernel launch"); break; }
        g_launches++;
    }
    if (window) {
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    }
    return rc;
}

}  // namespace gd

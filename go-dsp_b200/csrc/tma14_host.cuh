// Host side of the TMA-fed fused four-step for 2^13- .. 2^18-point transforms (fft_tma14.cuh, N = LA x LB): rows of a batch
// or columns of a row-major matrix; tensor maps, scratch slots under a persisting L2 window, dependency counters, one
// persistent launch per up to 512 phases. Included by one translation unit per size (pass_inst_tma<log2 N>.cu).
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

template <int LA, int LB>
static bool tma2d_rows_applicable(const void* in, long long in_dist, const cpx* out, long long out_dist, long long batch, int ld_conj, int st_conj,
                                  double scale) {
    using SH = T14Shape<LA, LB>;
    const long long N = SH::N;
    const bool fwd = !ld_conj && !st_conj && scale == 1.0, inv = ld_conj && st_conj;
    return (fwd || inv) && batch >= SH::UNIT && batch % SH::UNIT == 0 && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0 && in_dist >= N &&
           out_dist >= N && in_dist < (1LL << 35) && out_dist < (1LL << 35);
}
// columns [0, ncols) of a row-major matrix with N rows and row pitch `pitch`
template <int LA, int LB>
static bool tma2d_cols_applicable(const cpx* src, const cpx* dst, long long len, long long ncols, long long pitch) {
    using SH = T14Shape<LA, LB>;
    return LB <= 256 && len == (long long)SH::N && ncols >= SH::UNIT && ncols % SH::UNIT == 0 && pitch >= ncols && pitch < (1LL << 30) &&
           ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0;
}

template <int LA, int LB, int MODE, bool INV, bool PROF = false, int TW2 = 0>
static cudaError_t launch14(int grid, const CUtensorMap& mx, const CUtensorMap& mi, const CUtensorMap& mo, const Tma14Params& f, cudaStream_t st,
                            const T14Peers* peers = nullptr) {
    cudaError_t e = cudaFuncSetAttribute(fft_tma14_kernel<LA, LB, MODE, INV, PROF, TW2>, cudaFuncAttributeMaxDynamicSharedMemorySize, T14_SMEM);   // per device
    if (e != cudaSuccess) return e;
    if constexpr (TW2 == 2) fft_tma14_kernel<LA, LB, MODE, INV, PROF, TW2><<<grid, TMA_THREADS, T14_SMEM, st>>>(mx, mi, mo, f, *peers);
    else fft_tma14_kernel<LA, LB, MODE, INV, PROF, TW2><<<grid, TMA_THREADS, T14_SMEM, st>>>(mx, mi, mo, f, 0);
    return cudaGetLastError();
}

static inline cuuint64_t clamp_stride(unsigned long long v) { return v < (1ULL << 39) ? v : (1ULL << 39); }   // stride of a dimension of extent 1

// mode ROWS: `count` transforms of N = LA * LB points, transform t at in + t * in_dist / out + t * out_dist (count % UNIT == 0).
// mode COLS: the first `count` columns of a row-major matrix with N rows (count % UNIT == 0): every column is a transform;
//            in_dist = out_dist = the row pitch of the matrix in elements. tw2_log2m > 0: output k of column c leaves multiplied by
//            w_M^((tw2_col0 + c) k), M = 2^tw2_log2m (conjugated for inv): the twiddle of an outer four-step over these lines.
//            nmat > 1: the same columns of nmat matrices (matrix m at in + m * in_mdist / out + m * out_mdist) in the same launches;
//            count / UNIT must be a power of two <= 512 (the matrices ride on dimension 3 of the tensor maps).
//            ex.npeer = G > 0 (with tw2_log2m): output row k of the slab goes to rank h = k / (N / G), row k % (N / G) of the matrix with
//            row pitch out_dist at ex.peer[h] + ex.peer_off: the stores of pass 2 are the exchange of the sharded four-step.
// mode ROWS, ex.seg = G > 1: transform t is G segments of N / G contiguous points, segment g at in + g * ex.seg_dist + t * in_dist.
template <int LA, int LB>
static Status fft_tma_2d(Device& d, int mode, const cpx* in, long long in_dist, cpx* out, long long out_dist, long long count, bool inv, double scale,
                         cudaStream_t st, const Tma2dExtra& ex) {
    const int tw2_log2m = ex.tw2_log2m;
    const long long tw2_col0 = ex.tw2_col0, in_mdist = ex.in_mdist, out_mdist = ex.out_mdist;
    long long nmat = ex.nmat;
    if (ex.npeer && (mode != T14_COLS || ex.npeer > 8 || (int)LB % ex.npeer || !ex.peer || nmat > 1)) return invalid14("fused kernel: bad peer-store launch");
    if (ex.aux && (mode != T14_ROWS || inv || tw2_log2m || ex.npeer)) return invalid14("fused kernel: the aux product needs forward row mode");
    if (ex.seg > 1 && (mode != T14_ROWS || LA > 512 || (int)LA % ex.seg || (ex.seg & 1))) return invalid14("fused kernel: bad segmented-row launch");
    using SH = T14Shape<LA, LB>;
    constexpr cuuint64_t A = LA, Bq = LB, N = SH::N, LNA = SH::LINES_A, LNB = SH::LINES_B, UNIT = SH::UNIT;
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
    if (tw2_log2m && (mode != T14_COLS || LB > 256 || tw2_log2m < 0 || tw2_log2m > 40)) return invalid14("fused kernel: the outer twiddle needs column mode");
    if (mode == T14_ROWS) {
        // Int[t][n2][k1], t < UNIT * S: a pass-2 tile is LINES_B adjacent k1 x LB / 2 rows n2 per half
        const cuuint64_t dims[4] = {2 * A, Bq, UNIT * (cuuint64_t)S, 1};
        const cuuint64_t str[3] = {A * 16, N * 16, clamp_stride(N * 16 * UNIT * (cuuint64_t)S)};
        const cuuint32_t box[4] = {(cuuint32_t)(2 * LNB), (cuuint32_t)(Bq / 2), 1, 1};
        GD_TRY(map4(enc, scratch, dims, str, box, &m_int));
    } else {
        // Int[tb][n2][k1][LINES_A t], tb < TBP * S: a pass-2 tile is RA adjacent k1 x LINES_A columns x LB / 2 rows n2 per half
        const cuuint64_t dims[4] = {2 * LNA, A, Bq, (cuuint64_t)(SH::TBP * S)};
        const cuuint64_t str[3] = {LNA * 16, LNA * 16 * A, LNA * 16 * A * Bq};
        const cuuint32_t box[4] = {(cuuint32_t)(2 * LNA), (cuuint32_t)SH::RA, (cuuint32_t)(Bq / 2), 1};
        GD_TRY(map4(enc, scratch, dims, str, box, &m_int));
    }
    Tma14Params f;
    memset(&f, 0, sizeof(f));
    f.wla = d.wl[T14Len<LA>::LOG2]; f.wlb = d.wl[T14Len<LB>::LOG2];
    for (int j = 1; j < 4; j++)
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
    const long long gpm = count / (long long)UNIT;                     // phases per matrix
    int bshift = 31;
    if (nmat > 1) {
        if (mode != T14_COLS || gpm > CH || (gpm & (gpm - 1)) || in_mdist >= (1LL << 35) || out_mdist >= (1LL << 35)) return invalid14("fused kernel: bad multi-matrix launch");
        for (bshift = 0; (1LL << bshift) < gpm; bshift++) {}
    } else nmat = 1;
    const long long groups = gpm * nmat;
    for (long long g0 = 0; g0 < groups && rc == GD_OK; g0 += CH) {
        const long long ng = groups - g0 < CH ? groups - g0 : CH;
        const bool multi = nmat > 1;
        const long long m0 = multi ? g0 / gpm : 0;                     // first matrix of this launch
        const cuuint64_t nm = multi ? (cuuint64_t)(ng / gpm) : 1;      // matrices of this launch (CH is a multiple of gpm)
        const cuuint64_t nt = multi ? (cuuint64_t)count : (cuuint64_t)ng * UNIT;   // transforms / columns (per matrix) of this launch
        CUtensorMap m_x, m_out;
        T14Peers peer_maps;
        if (mode == T14_ROWS) {
            const cuuint64_t dout[4] = {2 * A, Bq, nt, 1};
            const cuuint32_t bo[4] = {(cuuint32_t)(2 * LNB), (cuuint32_t)(Bq / 2), 1, 1};
            const cuuint64_t so[3] = {A * 16, (cuuint64_t)out_dist * 16, clamp_stride((cuuint64_t)out_dist * 16 * nt)};
            if (ex.seg > 1) {
                // rows n1 = g * (LA / G) + r of transform t: (r, g, t) are dimensions 1, 2, 3; half a tile is G / 2 segments
                const cuuint64_t sg = (cuuint64_t)ex.seg;
                const cuuint64_t dx[4] = {2 * Bq, A / sg, sg, nt};
                const cuuint32_t bx[4] = {(cuuint32_t)(2 * LNA), (cuuint32_t)(A / sg), (cuuint32_t)(sg / 2), 1};
                const cuuint64_t sx[3] = {Bq * 16, (cuuint64_t)ex.seg_dist * 16, (cuuint64_t)in_dist * 16};
                if ((rc = map4(enc, in + g0 * (long long)UNIT * in_dist, dx, sx, bx, &m_x)) != GD_OK) break;
            } else {
                const cuuint64_t dx[4] = {2 * Bq, A, nt, 1};
                const cuuint32_t bx[4] = {(cuuint32_t)(2 * LNA), (cuuint32_t)(A / 2 > 256 ? 256 : A / 2), 1, 1};
                const cuuint64_t sx[3] = {Bq * 16, (cuuint64_t)in_dist * 16, clamp_stride((cuuint64_t)in_dist * 16 * nt)};
                if ((rc = map4(enc, in + g0 * (long long)UNIT * in_dist, dx, sx, bx, &m_x)) != GD_OK) break;
            }
            if ((rc = map4(enc, out + g0 * (long long)UNIT * out_dist, dout, so, bo, &m_out)) != GD_OK) break;
        } else {
            const cuuint64_t pi_ = (cuuint64_t)in_dist * 16, po = (cuuint64_t)out_dist * 16;
            const cuuint64_t dx[4] = {2 * nt, Bq, A, nm}, dout[4] = {2 * nt, A, Bq, nm};
            const cuuint32_t bx[4] = {(cuuint32_t)(2 * LNA), 1, (cuuint32_t)(A / 2), 1}, bo[4] = {(cuuint32_t)(2 * LNA), (cuuint32_t)SH::RA, (cuuint32_t)(Bq / 2), 1};
            const cuuint64_t sx[3] = {pi_, pi_ * Bq, multi ? (cuuint64_t)in_mdist * 16 : clamp_stride(pi_ * N)};
            const cuuint64_t so[3] = {po, po * A, multi ? (cuuint64_t)out_mdist * 16 : clamp_stride(po * N)};
            if ((rc = map4(enc, multi ? in + m0 * in_mdist : in + g0 * (long long)UNIT, dx, sx, bx, &m_x)) != GD_OK) break;
            if (ex.npeer) {
                // one output map per rank: its LB / G rows of k2 (N / G rows of the slab), the same columns
                const cuuint64_t per = Bq / (cuuint64_t)ex.npeer;
                const cuuint64_t dp[4] = {2 * nt, A, per, 1};
                const cuuint32_t bp[4] = {(cuuint32_t)(2 * LNA), (cuuint32_t)SH::RA, (cuuint32_t)(per < Bq / 2 ? per : Bq / 2), 1};
                const cuuint64_t sp[3] = {po, po * A, clamp_stride(po * A * per)};
                for (int h = 0; h < ex.npeer && rc == GD_OK; h++) rc = map4(enc, ex.peer[h] + ex.peer_off + g0 * (long long)UNIT, dp, sp, bp, &peer_maps.m[h]);
                if (rc != GD_OK) break;
                m_out = peer_maps.m[0];
            } else if ((rc = map4(enc, multi ? out + m0 * out_mdist : out + g0 * (long long)UNIT, dout, so, bo, &m_out)) != GD_OK) break;
        }
        f.batch = (int)ng; f.delay = D; f.nslots = S; f.scratch = scratch;
        f.done1 = cnt; f.done2 = cnt + CH; f.queue = cnt + 2 * CH;
        f.tw_lo = tw.lo; f.tw_hi = tw.hi; f.scale = scale;
        f.tw2_log2m = tw2_log2m; f.tw2_col0 = multi ? tw2_col0 : tw2_col0 + g0 * (long long)UNIT;
        f.bshift = bshift; f.npeer = ex.npeer; f.prot = ex.rank; f.aux = ex.aux; f.seg = ex.seg > 1 ? ex.seg : 0;
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
        const int sms = d.tma_grid_cap > 0 && d.tma_grid_cap < d.num_sms ? d.tma_grid_cap : d.num_sms;
        const int grid = (int)(nitems < sms ? nitems : sms);
        if constexpr (LB <= 256) {
            if (ex.npeer) e = inv ? launch14<LA, LB, T14_COLS, true, false, 2>(grid, m_x, m_int, m_out, f, st, &peer_maps) : launch14<LA, LB, T14_COLS, false, false, 2>(grid, m_x, m_int, m_out, f, st, &peer_maps);
            else if (tw2_log2m) e = inv ? launch14<LA, LB, T14_COLS, true, false, 1>(grid, m_x, m_int, m_out, f, st) : launch14<LA, LB, T14_COLS, false, false, 1>(grid, m_x, m_int, m_out, f, st);
            else if (f.prof && !inv && LA == LB) e = mode == T14_ROWS ? launch14<LA, LB, T14_ROWS, false, LA == LB>(grid, m_x, m_int, m_out, f, st) : launch14<LA, LB, T14_COLS, false, LA == LB>(grid, m_x, m_int, m_out, f, st);
            else if (mode == T14_ROWS && ex.aux) e = launch14<LA, LB, T14_ROWS, false, false, 3>(grid, m_x, m_int, m_out, f, st);
            else if (mode == T14_ROWS) e = inv ? launch14<LA, LB, T14_ROWS, true>(grid, m_x, m_int, m_out, f, st) : launch14<LA, LB, T14_ROWS, false>(grid, m_x, m_int, m_out, f, st);
            else e = inv ? launch14<LA, LB, T14_COLS, true>(grid, m_x, m_int, m_out, f, st) : launch14<LA, LB, T14_COLS, false>(grid, m_x, m_int, m_out, f, st);
        } else {
            if (mode != T14_ROWS) { rc = invalid14("fused kernel: columns need LB <= 256"); break; }
            if (ex.aux) e = launch14<LA, LB, T14_ROWS, false, false, 3>(grid, m_x, m_int, m_out, f, st);
            else e = inv ? launch14<LA, LB, T14_ROWS, true>(grid, m_x, m_int, m_out, f, st) : launch14<LA, LB, T14_ROWS, false>(grid, m_x, m_int, m_out, f, st);
        }
        if (e != cudaSuccess) { rc = cuda_fail(e, "fft_tma14_kernel launch"); break; }
        g_launches++;
    }
    if (window) {
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    }
    return rc;
}

// entry points of one size (defined in pass_inst_tma<log2 N>.cu)
#define GD_TMA2D_ENTRY(LG, LA, LB)                                                                                                      \
    bool tma##LG##_rows_applicable(const void* in, long long in_dist, const cpx* out, long long out_dist, long long batch, int ld_conj, \
                                   int st_conj, double scale) {                                                                         \
        return tma2d_rows_applicable<LA, LB>(in, in_dist, out, out_dist, batch, ld_conj, st_conj, scale);                               \
    }                                                                                                                                   \
    bool tma##LG##_cols_applicable(const cpx* src, const cpx* dst, long long len, long long ncols, long long pitch) {                   \
        return tma2d_cols_applicable<LA, LB>(src, dst, len, ncols, pitch);                                                              \
    }                                                                                                                                   \
    Status fft_tma_2p##LG(Device& d, int mode, const cpx* in, long long in_dist, cpx* out, long long out_dist, long long count, bool inv, \
                          double scale, cudaStream_t st, const Tma2dExtra& ex) {                                                        \
        return fft_tma_2d<LA, LB>(d, mode, in, in_dist, out, out_dist, count, inv, scale, st, ex);                                      \
    }

}  // namespace gd

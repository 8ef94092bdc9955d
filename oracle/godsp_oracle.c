/*
 * godsp_oracle.c -- CPU restatement of go-dsp's FFT / Welch-PSD path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: a plain-C
 * restatement (same operation order, no FMA contraction: build with
 * -O2 -ffp-contract=off) of the reference Go sources.  It is never linked,
 * imported or called by the product library (go-dsp_b200/): only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it, and only as the checker or the timed CPU baseline.
 *
 * The reference cannot run here (pure Go, no Go toolchain in the image or on
 * the GPU box), so parity is pinned by the reference's OWN golden vectors
 * (tests/golden/reference_vectors.json, transcribed with file:line from
 * fft/fft_test.go, spectral/{pwelch,spectral}_test.go, window/window_test.go,
 * dsputils/{matrix,dsputils}_test.go) -- tests/test_oracle_golden.py checks every one.
 * What those vectors do not pin (N > 256, Convolve, IFFTReal, multi-segment
 * Pwelch, Bluestein beyond N=5) is pinned only by the cited source lines.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the go-dsp repository root).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } c128;

/* Go complex128 multiply on amd64: (ac-bd) + (ad+bc)i, no fused multiply-add. */
static inline c128 cmul(c128 a, c128 b) {
    c128 r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}
static inline c128 cadd(c128 a, c128 b) { c128 r = { a.re + b.re, a.im + b.im }; return r; }
static inline c128 csub(c128 a, c128 b) { c128 r = { a.re - b.re, a.im - b.im }; return r; }

/* ------------------------------------------------------------------ dsputils */

/* dsputils/dsputils.go:34-36  IsPowerOf2: x&(x-1)==0 (true for 0). */
int gdo_is_pow2(int64_t x) { return (x & (x - 1)) == 0; }

/* Go math.Log2 (src/math/log10.go): Frexp, exact for powers of two, else
 * Log(frac)*(1/Ln2)+exp. */
static double go_log2(double x) {
    int e;
    double frac = frexp(x, &e);
    if (frac == 0.5) return (double)(e - 1);
    return log(frac) * (1.0 / M_LN2) + (double)e;
}

/* dsputils/dsputils.go:39-45  NextPowerOf2: int(Pow(2, Ceil(Log2(float64(x))))). */
int64_t gdo_next_pow2(int64_t x) {
    if (gdo_is_pow2(x)) return x;
    return (int64_t)ldexp(1.0, (int)ceil(go_log2((double)x)));
}

/* ------------------------------------------------------------------ radix2.go */

/* fft/radix2.go:172-180  log2 (integer). */
static unsigned ilog2(uint64_t v) { unsigned r = 0; for (v >>= 1; v != 0; v >>= 1) r++; return r; }

/* fft/radix2.go:184-199  reverseBits(v, s): reverse the low s bits of v. */
uint64_t gdo_reverse_bits(uint64_t v, uint64_t s) {
    uint64_t r = v & 1;
    s--;
    for (v >>= 1; v != 0; v >>= 1) { r <<= 1; r |= v & 1; s--; }
    return r << s;
}

/* fft/radix2.go:25-73  grow-only twiddle cache: table[N][k]; N=4 is seeded with the
 * exact {1,-i,-1,i}; even entries are copied from table[N/2], odd entries are
 * Sincos(-2*Pi/N*k). Index by log2(N). */
#define MAXLOG 40
static c128 *g_factors[MAXLOG];

static void ensure_factors(int64_t n) {
    #pragma omp critical(gdo_factors)
    {
        if (g_factors[2] == NULL) {
            c128 *f = (c128 *)malloc(4 * sizeof(c128));
            f[0].re = 1;  f[0].im = 0;  f[1].re = 0;  f[1].im = -1;
            f[2].re = -1; f[2].im = 0;  f[3].re = 0;  f[3].im = 1;
            g_factors[2] = f;
        }
        /* radix2.go:52-64 */
        for (int64_t i = 8, p = 4; i <= n; p = i, i <<= 1) {
            unsigned li = ilog2((uint64_t)i), lp = ilog2((uint64_t)p);
            if (g_factors[li] == NULL) {
                c128 *f = (c128 *)malloc((size_t)i * sizeof(c128));
                for (int64_t nn = 0, j = 0; nn < i; nn += 2, j++) f[nn] = g_factors[lp][j];
                for (int64_t nn = 1; nn < i; nn += 2) {
                    double s, c;
                    sincos(-2 * M_PI / (double)i * (double)nn, &s, &c);
                    f[nn].re = c; f[nn].im = s;
                }
                g_factors[li] = f;
            }
        }
    }
}

void gdo_ensure_radix2_factors(int64_t n) { if (n >= 4) ensure_factors(n); }

/* Read access for tests: copies table[N] (N power of two >= 4). */
void gdo_radix2_factors(int64_t n, double *out) {
    ensure_factors(n);
    memcpy(out, g_factors[ilog2((uint64_t)n)], (size_t)n * sizeof(c128));
}

/* fft/radix2.go:157-168  reorderData: r[reverseBits(n, log2 N)] = x[n]. */
static void reorder_data(const c128 *x, c128 *r, int64_t lx) {
    uint64_t s = ilog2((uint64_t)lx);
    for (uint64_t n = 0; n < (uint64_t)lx; n++) r[gdo_reverse_bits(n, s)] = x[n];
}

/* fft/radix2.go:80-154  radix2FFT.  The goroutine pool only partitions the
 * independent butterflies of one stage; a sequential sweep gives bit-identical
 * results.  out must not alias x. */
static void radix2_fft(const c128 *x, c128 *out, int64_t lx) {
    if (lx == 2) {            /* stage==2 only; no factor table is touched */
        out[0] = cadd(x[0], x[1]); out[1] = csub(x[0], x[1]); return;
    }
    ensure_factors(lx);
    const c128 *factors = g_factors[ilog2((uint64_t)lx)];
    c128 *t = (c128 *)malloc((size_t)lx * sizeof(c128));
    c128 *r = (c128 *)malloc((size_t)lx * sizeof(c128));
    reorder_data(x, r, lx);
    for (int64_t stage = 2; stage <= lx; stage <<= 1) {
        int64_t blocks = lx / stage, s_2 = stage / 2;
        for (int64_t nb = 0; nb < lx; nb += stage) {
            if (stage != 2) {
                for (int64_t j = 0; j < s_2; j++) {          /* radix2.go:106-113 */
                    int64_t idx = j + nb, idx2 = idx + s_2;
                    c128 ridx = r[idx];
                    c128 w_n = cmul(r[idx2], factors[blocks * j]);
                    t[idx] = cadd(ridx, w_n);
                    t[idx2] = csub(ridx, w_n);
                }
            } else {                                          /* radix2.go:114-120 */
                c128 rn = r[nb], rn1 = r[nb + 1];
                t[nb] = cadd(rn, rn1);
                t[nb + 1] = csub(rn, rn1);
            }
        }
        c128 *tmp = r; r = t; t = tmp;                       /* radix2.go:150 */
    }
    memcpy(out, r, (size_t)lx * sizeof(c128));
    free(t); free(r);
}

/* ------------------------------------------------------------------ fft.go / bluestein.go */

static void fft_any(const c128 *x, c128 *out, int64_t lx);
static void ifft_any(const c128 *x, c128 *out, int64_t lx);

/* fft/fft.go:55-69  Convolve (circular): IFFT(FFT(x)*FFT(y)). */
static void convolve(const c128 *x, const c128 *y, c128 *out, int64_t n) {
    c128 *fx = (c128 *)malloc((size_t)n * sizeof(c128));
    c128 *fy = (c128 *)malloc((size_t)n * sizeof(c128));
    fft_any(x, fx, n);
    fft_any(y, fy, n);
    for (int64_t i = 0; i < n; i++) fx[i] = cmul(fx[i], fy[i]);
    ifft_any(fx, out, n);
    free(fx); free(fy);
}

/* fft/bluestein.go:48-57  chirp tables: f[i]=(cos,sin)(Pi/N * float64(i*i)),
 * inv = conj(f), i=0 forced to (1,0).  The angle is formed exactly as the
 * reference forms it (division first, then one rounded multiply). */
static void bluestein_factors(int64_t n, c128 *f, c128 *inv) {
    for (int64_t i = 0; i < n; i++) {
        double s, c;
        if (i == 0) { s = 0; c = 1; }
        else sincos(M_PI / (double)n * (double)(i * i), &s, &c);
        f[i].re = c;   f[i].im = s;
        inv[i].re = c; inv[i].im = -s;
    }
}

void gdo_bluestein_factors(int64_t n, double *f, double *inv) {
    bluestein_factors(n, (c128 *)f, (c128 *)inv);
}

/* fft/bluestein.go:68-94  bluesteinFFT. */
static void bluestein_fft(const c128 *x, c128 *out, int64_t lx) {
    int64_t la = gdo_next_pow2(lx * 2 - 1);
    c128 *a = (c128 *)calloc((size_t)la, sizeof(c128));
    c128 *b = (c128 *)calloc((size_t)la, sizeof(c128));
    c128 *r = (c128 *)malloc((size_t)la * sizeof(c128));
    c128 *f = (c128 *)malloc((size_t)lx * sizeof(c128));
    c128 *inv = (c128 *)malloc((size_t)lx * sizeof(c128));
    bluestein_factors(lx, f, inv);
    for (int64_t n = 0; n < lx; n++) a[n] = cmul(x[n], inv[n]);    /* :74-76 */
    for (int64_t i = 0; i < lx; i++) {                              /* :78-85 */
        b[i] = f[i];
        if (i != 0) b[la - i] = f[i];
    }
    convolve(a, b, r, la);                                          /* :87 */
    for (int64_t i = 0; i < lx; i++) out[i] = cmul(r[i], inv[i]);   /* :89-93 */
    free(a); free(b); free(r); free(f); free(inv);
}

int64_t gdo_bluestein_padded_len(int64_t lx) { return gdo_next_pow2(lx * 2 - 1); }

/* fft/fft.go:72-87  FFT dispatch. */
static void fft_any(const c128 *x, c128 *out, int64_t lx) {
    if (lx <= 1) { if (lx == 1) out[0] = x[0]; return; }
    if (gdo_is_pow2(lx)) { radix2_fft(x, out, lx); return; }
    bluestein_fft(x, out, lx);
}

/* fft/fft.go:35-52  IFFT: index reversal, FFT, divide by complex(N,0)
 * (Go's complex division with a zero imaginary divisor is re/N, im/N). */
static void ifft_any(const c128 *x, c128 *out, int64_t lx) {
    c128 *r = (c128 *)malloc((size_t)lx * sizeof(c128));
    r[0] = x[0];
    for (int64_t i = 1; i < lx; i++) r[i] = x[lx - i];
    fft_any(r, out, lx);
    double N = (double)lx;
    for (int64_t n = 0; n < lx; n++) { out[n].re /= N; out[n].im /= N; }
    free(r);
}

void gdo_fft(const double *in, double *out, int64_t n) { fft_any((const c128 *)in, (c128 *)out, n); }
void gdo_ifft(const double *in, double *out, int64_t n) { if (n > 0) ifft_any((const c128 *)in, (c128 *)out, n); }

/* dsputils/dsputils.go:25-31 ToComplex + fft/fft.go:25-32 FFTReal / IFFTReal. */
static c128 *to_complex(const double *x, int64_t n) {
    c128 *y = (c128 *)malloc((size_t)(n > 0 ? n : 1) * sizeof(c128));
    for (int64_t i = 0; i < n; i++) { y[i].re = x[i]; y[i].im = 0; }
    return y;
}
void gdo_fft_real(const double *in, double *out, int64_t n) {
    c128 *y = to_complex(in, n); fft_any(y, (c128 *)out, n); free(y);
}
void gdo_ifft_real(const double *in, double *out, int64_t n) {
    c128 *y = to_complex(in, n); if (n > 0) ifft_any(y, (c128 *)out, n); free(y);
}
void gdo_convolve(const double *x, const double *y, double *out, int64_t n) {
    convolve((const c128 *)x, (const c128 *)y, (c128 *)out, n);
}

/* fft/fft.go:123-154  computeFFT2 on a contiguous row-major rows x cols array:
 * all column transforms first, then all row transforms. */
void gdo_fft2(const double *in_, double *out_, int64_t rows, int64_t cols, int inverse) {
    const c128 *in = (const c128 *)in_; c128 *out = (c128 *)out_;
    void (*f)(const c128 *, c128 *, int64_t) = inverse ? ifft_any : fft_any;
    c128 *t = (c128 *)malloc((size_t)(rows > cols ? rows : cols) * sizeof(c128));
    c128 *u = (c128 *)malloc((size_t)(rows > cols ? rows : cols) * sizeof(c128));
    for (int64_t i = 0; i < cols; i++) {
        for (int64_t j = 0; j < rows; j++) t[j] = in[j * cols + i];
        f(t, u, rows);
        for (int64_t n = 0; n < rows; n++) out[n * cols + i] = u[n];
    }
    for (int64_t n = 0; n < rows; n++) {
        memcpy(t, out + n * cols, (size_t)cols * sizeof(c128));
        f(t, out + n * cols, cols);
    }
    free(t); free(u);
}

/* fft/fft.go:166-224 computeFFTN + dsputils/matrix.go:37-57,110-175: flat
 * row-major storage (last dim fastest, offsets[i] = prod dims[i+1:]); for each
 * axis, transform every line along it (Dim gather, FFT, SetDim scatter),
 * ping-ponging between two matrices. */
void gdo_fftn(const double *in_, double *out_, const int64_t *dims, int nd, int inverse) {
    void (*f)(const c128 *, c128 *, int64_t) = inverse ? ifft_any : fft_any;
    int64_t total = 1, offsets[32];
    for (int i = nd - 1; i >= 0; i--) { offsets[i] = total; total *= dims[i]; }
    c128 *t = (c128 *)malloc((size_t)total * sizeof(c128));
    c128 *r = (c128 *)calloc((size_t)total, sizeof(c128));
    memcpy(t, in_, (size_t)total * sizeof(c128));
    int64_t maxd = 1;
    for (int i = 0; i < nd; i++) if (dims[i] > maxd) maxd = dims[i];
    c128 *line = (c128 *)malloc((size_t)maxd * sizeof(c128));
    c128 *res = (c128 *)malloc((size_t)maxd * sizeof(c128));
    for (int ax = 0; ax < nd; ax++) {
        int64_t len = dims[ax], stride = offsets[ax];
        int64_t outer = total / (len * stride);
        for (int64_t o = 0; o < outer; o++)
            for (int64_t i = 0; i < stride; i++) {
                int64_t base = o * len * stride + i;
                for (int64_t j = 0; j < len; j++) line[j] = t[base + j * stride];
                f(line, res, len);
                for (int64_t j = 0; j < len; j++) r[base + j * stride] = res[j];
            }
        c128 *tmp = r; r = t; t = tmp;                           /* fft.go:188 */
    }
    memcpy(out_, t, (size_t)total * sizeof(c128));
    free(t); free(r); free(line); free(res);
}

/* ------------------------------------------------------------------ window.go */

enum { GDO_WIN_RECT = 0, GDO_WIN_HAMMING, GDO_WIN_HANN, GDO_WIN_BARTLETT, GDO_WIN_FLATTOP, GDO_WIN_BLACKMAN };

/* window/window.go:32-152; every generator returns [1] for L == 1. */
int gdo_window(int id, int64_t L, double *r) {
    if (L <= 0) return 0;
    if (id == GDO_WIN_RECT) { for (int64_t i = 0; i < L; i++) r[i] = 1; return 0; }   /* :32-40 */
    if (L == 1) { r[0] = 1; return 0; }
    int64_t N = L - 1;
    switch (id) {
    case GDO_WIN_HAMMING: {                                        /* :44-58 */
        double coef = M_PI * 2 / (double)N;
        for (int64_t n = 0; n <= N; n++) r[n] = 0.54 - 0.46 * cos(coef * (double)n);
        return 0; }
    case GDO_WIN_HANN: {                                           /* :62-76 */
        double coef = 2 * M_PI / (double)N;
        for (int64_t n = 0; n <= N; n++) r[n] = 0.5 * (1 - cos(coef * (double)n));
        return 0; }
    case GDO_WIN_BARTLETT: {                                       /* :80-99 */
        double coef = 2 / (double)N;
        int64_t n = 0;
        for (; n <= N / 2; n++) r[n] = coef * (double)n;
        for (; n <= N; n++) r[n] = 2 - coef * (double)n;
        return 0; }
    case GDO_WIN_FLATTOP: {                                        /* :103-135 */
        const double a0 = 0.21557895, a1 = 0.41663158, a2 = 0.277263158, a3 = 0.083578947, a4 = 0.006947368;
        double coef = 2 * M_PI / (double)N;
        for (int64_t n = 0; n <= N; n++) {
            double factor = (double)n * coef;
            double t0 = a0, t1 = a1 * cos(factor), t2 = a2 * cos(2 * factor);
            double t3 = a3 * cos(3 * factor), t4 = a4 * cos(4 * factor);
            r[n] = t0 - t1 + t2 - t3 + t4;
        }
        return 0; }
    case GDO_WIN_BLACKMAN: {                                       /* :138-152 */
        for (int64_t n = 0; n <= N; n++) {
            double t1 = -0.5 * cos(2 * M_PI * (double)n / (double)N);
            double t2 = 0.08 * cos(4 * M_PI * (double)n / (double)N);
            r[n] = 0.42 + t1 + t2;
        }
        return 0; }
    }
    return -1;
}

/* ------------------------------------------------------------------ spectral */

/* spectral/spectral.go:22-33  Segment count. */
int64_t gdo_segment_count(int64_t lx, int64_t size, int64_t noverlap) {
    int64_t stride = size - noverlap;
    if (lx == size) return 1;
    if (lx > size) return (lx - size) / stride + 1;
    return 0;
}

/* spectral/spectral.go:35-44  Segment copy into a dense [segments][size] array. */
void gdo_segment(const double *x, int64_t lx, int64_t size, int64_t noverlap, double *out) {
    int64_t segs = gdo_segment_count(lx, size, noverlap), stride = size - noverlap;
    for (int64_t i = 0, off = 0; i < segs; i++, off += stride)
        memcpy(out + i * size, x + off, (size_t)size * sizeof(double));
}

/* One segment of the Pwelch loop, spectral/pwelch.go:107-121: zero-pad to pad
 * (no-op when pad <= nfft), window of length len(segment), FFTReal, and
 * d = |X[j]|^2 / nsegs (x2 for 0<j<lp-1) accumulated into pxx. */
static void pwelch_one(const double *seg, int64_t nfft, int64_t fftlen, int64_t lp, int64_t nsegs,
                       const double *win_apply, double *pxx, c128 *buf, c128 *spec) {
    for (int64_t i = 0; i < fftlen; i++) {
        double v = i < nfft ? seg[i] : 0.0;
        buf[i].re = v * win_apply[i]; buf[i].im = 0;
    }
    fft_any(buf, spec, fftlen);
    for (int64_t j = 0; j < lp; j++) {
        double d = (spec[j].re * spec[j].re + spec[j].im * spec[j].im) / (double)nsegs;
        if (j > 0 && j < lp - 1) d *= 2.0;
        pxx[j] += d;
    }
}

/* spectral/pwelch.go:74-145  Pwelch.  nfft/pad/noverlap are the raw option values
 * (0 = default); win_apply = wf(max(pad,nfft)) and win_norm = wf(nfft) are the two
 * evaluations of PwelchOptions.Window the reference performs (:109 via
 * window.go:26, and :124).  Returns lp (= len(Pxx) = len(freqs)), 0 for empty x.
 * threads<=1 follows the reference's sequential accumulation order exactly;
 * threads>1 (bench baseline only) gives each thread a contiguous range of
 * segments and sums the per-thread partials in thread order. */
int64_t gdo_pwelch(const double *x, int64_t lx, double Fs, int64_t nfft, int64_t pad, int64_t noverlap,
                   const double *win_apply, const double *win_norm, int scale_off,
                   double *pxx, double *freqs, int threads) {
    if (lx == 0) return 0;                                          /* :75-77 */
    if (nfft == 0) nfft = 256;                                      /* :85-87 */
    if (pad == 0) pad = nfft;                                       /* :93-95 */
    double *xp = NULL;
    if (lx < nfft) {                                                /* :97-99 */
        xp = (double *)calloc((size_t)nfft, sizeof(double));
        memcpy(xp, x, (size_t)lx * sizeof(double));
        x = xp; lx = nfft;
    }
    int64_t lp = pad / 2 + 1;                                       /* :101 */
    int64_t fftlen = pad > nfft ? pad : nfft;                       /* :108 ZeroPadF */
    int64_t nsegs = gdo_segment_count(lx, nfft, noverlap);          /* :104 */
    int64_t stride = nfft - noverlap;
    for (int64_t j = 0; j < lp; j++) pxx[j] = 0;
    if (gdo_is_pow2(fftlen) && fftlen >= 4) ensure_factors(fftlen);
    if (threads <= 1) {
        c128 *buf = (c128 *)malloc((size_t)fftlen * sizeof(c128));
        c128 *spec = (c128 *)malloc((size_t)fftlen * sizeof(c128));
        for (int64_t s = 0; s < nsegs; s++)
            pwelch_one(x + s * stride, nfft, fftlen, lp, nsegs, win_apply, pxx, buf, spec);
        free(buf); free(spec);
    } else {
#ifdef _OPENMP
        double *part = (double *)calloc((size_t)threads * (size_t)lp, sizeof(double));
        #pragma omp parallel num_threads(threads)
        {
            int t = omp_get_thread_num(), nt = omp_get_num_threads();
            int64_t s0 = nsegs * t / nt, s1 = nsegs * (t + 1) / nt;
            c128 *buf = (c128 *)malloc((size_t)fftlen * sizeof(c128));
            c128 *spec = (c128 *)malloc((size_t)fftlen * sizeof(c128));
            for (int64_t s = s0; s < s1; s++)
                pwelch_one(x + s * stride, nfft, fftlen, lp, nsegs, win_apply, part + (size_t)t * lp, buf, spec);
            free(buf); free(spec);
        }
        for (int t = 0; t < threads; t++)
            for (int64_t j = 0; j < lp; j++) pxx[j] += part[(size_t)t * lp + j];
        free(part);
#else
        return -1;
#endif
    }
    double norm = 0;                                                /* :124-128 */
    for (int64_t i = 0; i < nfft; i++) norm += win_norm[i] * win_norm[i];   /* math.Pow(x,2) == x*x */
    if (!scale_off) norm *= Fs;                                     /* :130-132 */
    for (int64_t i = 0; i < lp; i++) pxx[i] /= norm;                /* :134-136 */
    double coef = Fs / (double)pad;                                 /* :139 */
    for (int64_t i = 0; i < lp; i++) freqs[i] = (double)i * coef;   /* :140-142 */
    free(xp);
    return lp;
}

/* ------------------------------------------------------------------ callers / formats either side of the path */

/* spectral/pwelch.go:104-113 without the accumulate (STFT / spectrogram): for every segment of
 * spectral.Segment(x, nfft, noverlap): ZeroPadF to fftlen = max(pad, nfft), window.Apply, FFTReal; the
 * first lp bins of each spectrum go to out[s*lp + j]. Returns the segment count. */
int64_t gdo_stft(const double *x, int64_t lx, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp,
                 const double *win_apply, double *out) {
    int64_t nsegs = gdo_segment_count(lx, nfft, noverlap), stride = nfft - noverlap;
    if (gdo_is_pow2(fftlen) && fftlen >= 4) ensure_factors(fftlen);
    c128 *buf = (c128 *)malloc((size_t)fftlen * sizeof(c128));
    c128 *spec = (c128 *)malloc((size_t)fftlen * sizeof(c128));
    for (int64_t s = 0; s < nsegs; s++) {
        const double *seg = x + s * stride;
        for (int64_t i = 0; i < fftlen; i++) {
            double v = i < nfft ? seg[i] : 0.0;
            buf[i].re = v * win_apply[i]; buf[i].im = 0;
        }
        fft_any(buf, spec, fftlen);
        for (int64_t j = 0; j < lp; j++) { out[2 * (s * lp + j)] = spec[j].re; out[2 * (s * lp + j) + 1] = spec[j].im; }
    }
    free(buf); free(spec);
    return nsegs;
}

/* dsputils/dsputils.go:89-115  Segment: the (length, step) of the segs aliasing slices; returns 0, or -1 for the
 * reference's panic("too many segments"). */
int gdo_dsputils_segment(int64_t lx, int64_t segs, double noverlap, int64_t *length_out, int64_t *step_out) {
    int64_t overlap = 0, length, step = 0, tot;
    for (length = lx; length > 0; length--) {                        /* :94-101 */
        overlap = (int64_t)((double)length * noverlap);
        tot = segs * (length - overlap) + overlap;
        if (tot <= lx) { step = length - overlap; break; }
    }
    if (length == 0) return -1;                                      /* :103-105 */
    *length_out = length; *step_out = step;
    return 0;
}

/* wav/wav.go:59-110  New: RIFF/WAVE header walk. hdr[0..5] = AudioFormat, NumChannels, SampleRate, ByteRate,
 * BlockAlign, BitsPerSample; hdr[6] = Samples, hdr[7] = Duration in ns, hdr[8] = offset of the data bytes,
 * hdr[9] = data chunk size. Returns 0, or: 1 short read, 2 missing RIFF, 3 missing WAVE, 4 bad fmt size,
 * 5 unknown audio format, 6 data chunk before fmt. */
int gdo_wav_new(const unsigned char *b, int64_t n, int64_t *hdr) {
    if (n < 12) return 1;                                            /* :62-64 */
    if (memcmp(b, "RIFF", 4) != 0) return 2;                         /* :65-67 */
    if (memcmp(b + 8, "WAVE", 4) != 0) return 3;                     /* :68-70 */
    int64_t pos = 12;
    int has_fmt = 0;
    for (;;) {
        if (pos + 8 > n) return 1;                                   /* :73-75 */
        uint32_t sz = (uint32_t)b[pos + 4] | ((uint32_t)b[pos + 5] << 8) | ((uint32_t)b[pos + 6] << 16) | ((uint32_t)b[pos + 7] << 24);
        const unsigned char *typ = b + pos;
        pos += 8;
        if (memcmp(typ, "fmt ", 4) == 0) {                           /* :78-96 */
            if (sz < 16) return 4;
            if (pos + (int64_t)sz > n) return 1;
            const unsigned char *f = b + pos;
            hdr[0] = f[0] | (f[1] << 8);
            hdr[1] = f[2] | (f[3] << 8);
            hdr[2] = (int64_t)((uint32_t)f[4] | ((uint32_t)f[5] << 8) | ((uint32_t)f[6] << 16) | ((uint32_t)f[7] << 24));
            hdr[3] = (int64_t)((uint32_t)f[8] | ((uint32_t)f[9] << 8) | ((uint32_t)f[10] << 16) | ((uint32_t)f[11] << 24));
            hdr[4] = f[12] | (f[13] << 8);
            hdr[5] = f[14] | (f[15] << 8);
            if (hdr[0] != 1 && hdr[0] != 3) return 5;
            has_fmt = 1;
            pos += sz;
        } else if (memcmp(typ, "data", 4) == 0) {                    /* :97-104 */
            if (!has_fmt) return 6;
            hdr[6] = (int64_t)sz / hdr[5] * 8;                       /* int(sz) / int(BitsPerSample) * 8, left to right */
            hdr[7] = hdr[6] * 1000000000LL / hdr[2] / hdr[1];        /* Duration(Samples) * Second / SampleRate / NumChannels */
            hdr[8] = pos; hdr[9] = sz;
            return 0;
        } else {
            pos += sz;                                               /* :105-106 io.CopyN(ioutil.Discard, r, sz) */
        }
    }
}

/* wav/wav.go:138-161  ReadFloats on n raw little-endian samples: fmt 1 = float32 as is, 2 = int16 ->
 * (float32(v) - MinInt16) / (MaxInt16 - MinInt16), 3 = uint8 -> float32(v) / MaxUint8, all in float32. */
int gdo_wav_read_floats(const unsigned char *data, int fmt, int64_t n, float *out) {
    for (int64_t i = 0; i < n; i++) {
        if (fmt == 3) {
            volatile float f = (float)data[i];
            out[i] = f / 255.0f;                                     /* :145-148 */
        } else if (fmt == 2) {
            int16_t v = (int16_t)((uint16_t)data[2 * i] | ((uint16_t)data[2 * i + 1] << 8));
            volatile float f = (float)v - (-32768.0f);
            out[i] = f / (32767.0f - (-32768.0f));                   /* :150-153 */
        } else if (fmt == 1) {
            memcpy(out + i, data + 4 * i, 4);                        /* :154-155 */
        } else return -1;
    }
    return 0;
}

/* ------------------------------------------------------------------ bench helpers */

/* Independent transforms of one batch, one per OpenMP thread at a time (bench
 * baseline only; every transform is the sequential radix2_fft above). */
void gdo_fft_batch(const double *in, double *out, int64_t n, int64_t batch, int threads) {
    if (gdo_is_pow2(n) && n >= 4) ensure_factors(n);
#ifdef _OPENMP
    #pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 1)
#endif
    for (int64_t b = 0; b < batch; b++)
        fft_any((const c128 *)in + b * n, (c128 *)out + b * n, n);
}

/* SplitMix64 counter-based generator shared by oracle, CUDA and numpy
 * (SURVEY.md 8d): v(i) in [-1,1). */
static inline double splitmix_unit(uint64_t seed, uint64_t i) {
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}
void gdo_fill_splitmix(double *out, int64_t n, uint64_t seed, uint64_t offset) {
    for (int64_t i = 0; i < n; i++) out[i] = splitmix_unit(seed, offset + (uint64_t)i);
}

int gdo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// godsp.cpp -- implementation of the host-side mirror (godsp.hpp) over include/godsp_b200.h.
// Argument checks, option defaults, window tables, norms and frequency vectors are formed
// with the reference's integer / floating expressions; every transform is a C-ABI call.
#include "godsp.hpp"

#include <cmath>
#include <cstring>

#include "../../include/godsp_b200.h"

namespace godsp {

static void check(int status, const char* what) {
    if (status != 0) throw Panic(std::string(what) + ": go-dsp_b200 status " + std::to_string(status) + ": " + gd_last_error());
}

// ============================================================================ dsputils
namespace dsputils {

cvec ToComplex(const rvec& x) {
    cvec y(x.size());
    for (size_t n = 0; n < x.size(); n++) y[n] = cplx(x[n], 0);
    return y;
}
bool IsPowerOf2(int64_t x) { return (x & (x - 1)) == 0; }

// math.Log2 as Go defines it (Frexp; exact for powers of two; else Log(frac)*(1/Ln2)+exp)
static double go_log2(double x) {
    int e;
    double frac = std::frexp(x, &e);
    if (frac == 0.5) return (double)(e - 1);
    return std::log(frac) * (1.0 / M_LN2) + (double)e;
}
int64_t NextPowerOf2(int64_t x) {
    if (IsPowerOf2(x)) return x;
    return (int64_t)std::ldexp(1.0, (int)std::ceil(go_log2((double)x)));
}
cvec ZeroPad(const cvec& x, int64_t length) {
    if ((int64_t)x.size() >= length) return x;
    cvec r((size_t)length);
    std::copy(x.begin(), x.end(), r.begin());
    return r;
}
rvec ZeroPadF(const rvec& x, int64_t length) {
    if ((int64_t)x.size() >= length) return x;
    rvec r((size_t)length, 0.0);
    std::copy(x.begin(), x.end(), r.begin());
    return r;
}
cvec ZeroPad2(const cvec& x) { return ZeroPad(x, NextPowerOf2((int64_t)x.size())); }
std::vector<cvec> ToComplex2(const std::vector<rvec>& x) {
    std::vector<cvec> y(x.size());
    for (size_t n = 0; n < x.size(); n++) y[n] = ToComplex(x[n]);
    return y;
}
std::vector<std::pair<int64_t, int64_t>> Segment(int64_t lx, int64_t segs, double noverlap) {
    int64_t overlap = 0, length = 0, step = 0, tot = 0;
    for (length = lx; length > 0; length--) {
        overlap = (int64_t)((double)length * noverlap);
        tot = segs * (length - overlap) + overlap;
        if (tot <= lx) { step = length - overlap; break; }
    }
    if (length == 0) throw Panic("too many segments");
    std::vector<std::pair<int64_t, int64_t>> r((size_t)segs);
    int64_t s = 0;
    for (auto& pr : r) { pr = {s, length}; s += step; }
    return r;
}
bool Float64Equal(double a, double b) { return std::fabs(a - b) <= 1e-8 || std::fabs(1 - a / b) <= 1e-8; }
bool ComplexEqual(cplx a, cplx b) { return Float64Equal(a.real(), b.real()) && Float64Equal(a.imag(), b.imag()); }
bool PrettyClose(const rvec& a, const rvec& b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++) if (!Float64Equal(a[i], b[i])) return false;
    return true;
}
bool PrettyCloseC(const cvec& a, const cvec& b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++) if (!ComplexEqual(a[i], b[i])) return false;
    return true;
}

Matrix Matrix::MakeMatrix(const cvec& x, const std::vector<int64_t>& dims) {
    Matrix m;
    int64_t length = 1;
    m.offsets_.assign(dims.size(), 0);
    for (int i = (int)dims.size() - 1; i >= 0; i--) {
        if (dims[i] < 1) throw Panic("invalid dimensions");
        m.offsets_[i] = length;
        length *= dims[i];
    }
    if ((int64_t)x.size() != length) throw Panic("incorrect dimensions");
    m.list_ = x;
    m.dims_ = dims;
    return m;
}
Matrix Matrix::MakeMatrix2(const std::vector<cvec>& x) {
    if (x.empty()) throw Panic("runtime error: index out of range [0] with length 0");
    std::vector<int64_t> dims = {(int64_t)x.size(), (int64_t)x[0].size()};
    cvec r((size_t)(dims[0] * dims[1]));
    for (size_t n = 0; n < x.size(); n++) {
        if ((int64_t)x[n].size() != dims[1]) throw Panic("ragged array");
        std::copy(x[n].begin(), x[n].end(), r.begin() + n * dims[1]);
    }
    return MakeMatrix(r, dims);
}
Matrix Matrix::MakeEmptyMatrix(const std::vector<int64_t>& dims) {
    int64_t x = 1;
    for (auto v : dims) x *= v;
    return MakeMatrix(cvec((size_t)(x > 0 ? x : 0)), dims);
}
int64_t Matrix::offset(const std::vector<int64_t>& idx) const {
    if (idx.size() != dims_.size()) throw Panic("incorrect dimensions");
    int64_t i = 0;
    for (size_t n = 0; n < idx.size(); n++) {
        if (idx[n] > dims_[n]) throw Panic("incorrect dimensions");
        i += idx[n] * offsets_[n];
    }
    if (i < 0 || i >= (int64_t)list_.size()) throw Panic("runtime error: index out of range");
    return i;
}
std::vector<int64_t> Matrix::indexes(const std::vector<int64_t>& idx) const {
    if (idx.size() != dims_.size()) throw Panic("runtime error: index out of range");
    int i = -1;
    for (size_t n = 0; n < idx.size(); n++) {
        if (idx[n] == -1) {
            if (i >= 0) throw Panic("only one dimension index allowed");
            i = (int)n;
        } else if (idx[n] >= dims_[n]) throw Panic("dimension out of bounds");
    }
    if (i == -1) throw Panic("must specify one dimension index");
    int64_t x = 0;
    for (size_t n = 0; n < idx.size(); n++) if (idx[n] >= 0) x += offsets_[n] * idx[n];
    std::vector<int64_t> r((size_t)dims_[i]);
    for (size_t j = 0; j < r.size(); j++) r[j] = x + offsets_[i] * (int64_t)j;
    return r;
}
cvec Matrix::Dim(const std::vector<int64_t>& idx) const {
    auto inds = indexes(idx);
    cvec r(inds.size());
    for (size_t n = 0; n < inds.size(); n++) r[n] = list_[(size_t)inds[n]];
    return r;
}
void Matrix::SetDim(const cvec& x, const std::vector<int64_t>& idx) {
    auto inds = indexes(idx);
    if (x.size() != inds.size()) throw Panic("incorrect array length");
    for (size_t n = 0; n < inds.size(); n++) list_[(size_t)inds[n]] = x[n];
}
std::vector<cvec> Matrix::To2D() const {
    if (dims_.size() != 2) throw Panic("can only convert 2-D Matrixes");
    std::vector<cvec> r((size_t)dims_[0]);
    for (int64_t i = 0; i < dims_[0]; i++) r[(size_t)i].assign(list_.begin() + i * dims_[1], list_.begin() + (i + 1) * dims_[1]);
    return r;
}
bool Matrix::PrettyClose(const Matrix& n) const {
    for (size_t i = 0; i < dims_.size(); i++) if (i >= n.dims_.size() || dims_[i] != n.dims_[i]) return false;
    return PrettyCloseC(list_, n.list_);
}
}  // namespace dsputils

// ============================================================================ window
namespace window {
void Apply(rvec& x, const Func& wf) {
    rvec w = wf((int64_t)x.size());
    if (w.size() < x.size()) throw Panic("runtime error: index out of range");
    for (size_t i = 0; i < x.size(); i++) x[i] *= w[i];
}
rvec Rectangular(int64_t L) { return rvec((size_t)(L > 0 ? L : 0), 1.0); }

// every generator below returns {1} for L == 1
template <class F>
static rvec generate(int64_t L, F f) {
    if (L < 0) throw Panic("makeslice: len out of range");
    rvec r((size_t)L);
    if (L == 1) { r[0] = 1; return r; }
    for (int64_t n = 0; n < L; n++) r[(size_t)n] = f(n, L - 1);
    return r;
}
rvec Hamming(int64_t L) {
    return generate(L, [](int64_t n, int64_t N) { double coef = M_PI * 2 / (double)N; return 0.54 - 0.46 * std::cos(coef * (double)n); });
}
rvec Hann(int64_t L) {
    return generate(L, [](int64_t n, int64_t N) { double coef = 2 * M_PI / (double)N; return 0.5 * (1 - std::cos(coef * (double)n)); });
}
rvec Bartlett(int64_t L) {
    return generate(L, [](int64_t n, int64_t N) { double coef = 2 / (double)N; return n <= N / 2 ? coef * (double)n : 2 - coef * (double)n; });
}
rvec FlatTop(int64_t L) {
    return generate(L, [](int64_t n, int64_t N) {
        const double a0 = 0.21557895, a1 = 0.41663158, a2 = 0.277263158, a3 = 0.083578947, a4 = 0.006947368;
        double factor = (double)n * (2 * M_PI / (double)N);
        double t1 = a1 * std::cos(factor), t2 = a2 * std::cos(2 * factor), t3 = a3 * std::cos(3 * factor), t4 = a4 * std::cos(4 * factor);
        return a0 - t1 + t2 - t3 + t4;
    });
}
rvec Blackman(int64_t L) {
    return generate(L, [](int64_t n, int64_t N) {
        double t1 = -0.5 * std::cos(2 * M_PI * (double)n / (double)N);
        double t2 = 0.08 * std::cos(4 * M_PI * (double)n / (double)N);
        return 0.42 + t1 + t2;
    });
}
}  // namespace window

// ============================================================================ fft
namespace fft {
static int g_worker_pool_size = 0;
void SetWorkerPoolSize(int n) { g_worker_pool_size = n < 0 ? 0 : n; }   // the GPU path has no worker pool
int WorkerPoolSize() { return g_worker_pool_size; }

static cvec run1d(const void* in, int64_t n, bool real_in, int dir) {
    cvec r((size_t)n);
    if (n == 0) return r;
    int st = real_in ? gd_fft_r2c_full((const double*)in, (double*)r.data(), n, dir)
                     : gd_fft_c2c((const double*)in, (double*)r.data(), n, dir);
    check(st, real_in ? "gd_fft_r2c_full" : "gd_fft_c2c");
    return r;
}
cvec FFT(const cvec& x) { return run1d(x.data(), (int64_t)x.size(), false, +1); }
cvec IFFT(const cvec& x) {
    if (x.empty()) throw Panic("runtime error: index out of range [0] with length 0");   // fft.go:40
    return run1d(x.data(), (int64_t)x.size(), false, -1);
}
cvec FFTReal(const rvec& x) { return run1d(x.data(), (int64_t)x.size(), true, +1); }
cvec IFFTReal(const rvec& x) {
    if (x.empty()) throw Panic("runtime error: index out of range [0] with length 0");
    return run1d(x.data(), (int64_t)x.size(), true, -1);
}
cvec Convolve(const cvec& x, const cvec& y) {
    if (x.size() != y.size()) throw Panic("arrays not of equal size");
    if (x.empty()) throw Panic("runtime error: index out of range [0] with length 0");   // IFFT of an empty product
    cvec r(x.size());
    check(gd_convolve_c2c((const double*)x.data(), (const double*)y.data(), (double*)r.data(), (int64_t)x.size()), "gd_convolve_c2c");
    return r;
}

static std::vector<cvec> fft2(const std::vector<cvec>& x, int dir) {
    const int64_t rows = (int64_t)x.size();
    if (rows == 0) throw Panic("empty input array");
    const int64_t cols = (int64_t)x[0].size();
    for (auto& row : x) if ((int64_t)row.size() != cols) throw Panic("ragged input array");
    std::vector<cvec> r((size_t)rows, cvec((size_t)cols));
    if (cols == 0) return r;
    // [][]complex128 rows are separate allocations: stage them into one pinned block
    const size_t bytes = (size_t)rows * cols * sizeof(cplx);
    cplx* in = (cplx*)gd_pinned_alloc(2 * bytes);
    if (!in) throw Panic(std::string("gd_pinned_alloc: ") + gd_last_error());
    cplx* out = in + (size_t)rows * cols;
    for (int64_t i = 0; i < rows; i++) std::memcpy(in + i * cols, x[(size_t)i].data(), (size_t)cols * sizeof(cplx));
    int st = gd_fft2_c2c((const double*)in, (double*)out, rows, cols, dir);
    if (st == 0) for (int64_t i = 0; i < rows; i++) std::memcpy(r[(size_t)i].data(), out + i * cols, (size_t)cols * sizeof(cplx));
    gd_pinned_free(in);
    check(st, "gd_fft2_c2c");
    return r;
}
std::vector<cvec> FFT2(const std::vector<cvec>& x) { return fft2(x, +1); }
std::vector<cvec> IFFT2(const std::vector<cvec>& x) { return fft2(x, -1); }
std::vector<cvec> FFT2Real(const std::vector<rvec>& x) { return fft2(dsputils::ToComplex2(x), +1); }
std::vector<cvec> IFFT2Real(const std::vector<rvec>& x) { return fft2(dsputils::ToComplex2(x), -1); }

static dsputils::Matrix fftn(const dsputils::Matrix& m, int dir) {
    auto dims = m.Dimensions();
    dsputils::Matrix r = dsputils::Matrix::MakeEmptyMatrix(dims);
    check(gd_fftn_c2c((const double*)m.list().data(), (double*)r.list().data(), dims.data(), (int)dims.size(), dir), "gd_fftn_c2c");
    return r;
}
dsputils::Matrix FFTN(const dsputils::Matrix& m) { return fftn(m, +1); }
dsputils::Matrix IFFTN(const dsputils::Matrix& m) { return fftn(m, -1); }

void EnsureRadix2Factors(int64_t input_len) {
    if (input_len >= 1) check(gd_plan_warm(input_len), "gd_plan_warm");
}
uint64_t reverseBits(uint64_t v, uint64_t s) {
    uint64_t r = 0;
    for (uint64_t b = 0; b < s; b++) r |= ((v >> b) & 1ULL) << (s - 1 - b);   // low s bits of v, reversed
    return r;
}
}  // namespace fft

// ============================================================================ spectral
namespace spectral {
int64_t SegmentCount(int64_t lx, int64_t size, int64_t noverlap) {
    const int64_t stride = size - noverlap;
    if (lx == size) return 1;
    if (lx > size) {
        if (stride == 0) throw Panic("runtime error: integer divide by zero");
        return (lx - size) / stride + 1;
    }
    return 0;
}
std::vector<rvec> Segment(const rvec& x, int64_t size, int64_t noverlap) {
    const int64_t segs = SegmentCount((int64_t)x.size(), size, noverlap), stride = size - noverlap;
    if (segs < 0) throw Panic("runtime error: makeslice: len out of range");
    std::vector<rvec> r((size_t)segs);
    int64_t off = 0;
    for (auto& seg : r) {
        if (off < 0 || off + size > (int64_t)x.size()) throw Panic("runtime error: index out of range");
        seg.assign(x.begin() + off, x.begin() + off + size);
        off += stride;
    }
    return r;
}

std::pair<rvec, rvec> Pwelch(const rvec& xin, double Fs, const PwelchOptions* o) {
    if (xin.empty()) return {};
    if (!o) throw Panic("runtime error: invalid memory address or nil pointer dereference");
    int64_t nfft = o->NFFT, pad = o->Pad;
    const int64_t noverlap = o->Noverlap;
    window::Func wf = o->Window;
    if (nfft == 0) nfft = 256;
    if (!wf) wf = window::Hann;
    if (pad == 0) pad = nfft;
    if (nfft < 1 || pad < 0) throw Panic("runtime error: makeslice: len out of range");
    rvec padded;
    const rvec* x = &xin;
    if ((int64_t)xin.size() < nfft) { padded = dsputils::ZeroPadF(xin, nfft); x = &padded; }
    const int64_t lp = pad / 2 + 1;
    const int64_t nsegs = SegmentCount((int64_t)x->size(), nfft, noverlap);
    // Noverlap < 0 is valid in the reference: stride = size - noverlap > size leaves gaps between the segments
    if (nsegs < 1 || noverlap >= nfft) throw Panic("runtime error: makeslice: len out of range");
    const int64_t fftlen = pad > nfft ? pad : nfft;          // len(ZeroPadF(segment, pad))
    rvec win = wf(fftlen);                                    // window.Apply(x, wf) -> wf(len(x))
    if ((int64_t)win.size() < fftlen) throw Panic("runtime error: index out of range");
    rvec w = wf(nfft);
    double norm = 0;
    for (double v : w) norm += v * v;                         // math.Pow(x, 2), summed in order
    if (!o->Scale_off) norm *= Fs;
    std::pair<rvec, rvec> out;
    out.first.assign((size_t)lp, 0.0);
    check(gd_pwelch_f64(x->data(), (int64_t)x->size(), nfft, noverlap, fftlen, lp, nsegs, win.data(), norm, out.first.data()),
          "gd_pwelch_f64");
    out.second.resize((size_t)lp);
    const double coef = Fs / (double)pad;
    for (int64_t i = 0; i < lp; i++) out.second[(size_t)i] = (double)i * coef;
    return out;
}
}  // namespace spectral
}  // namespace godsp

// ============================================================================ flat C surface
// What tests/ and the ctypes binding (go-dsp_b200/godsp) call: one function per Go API entry,
// returning 0 or -1 with the panic message in gdh_last_panic().
using namespace godsp;
static thread_local std::string g_panic;
#define GDH_API extern "C" __attribute__((visibility("default")))
template <class F>
static int guarded(F f) {
    try { f(); return 0; }
    catch (const Panic& p) { g_panic = p.what(); return -1; }
    catch (const std::exception& e) { g_panic = std::string("exception: ") + e.what(); return -2; }
}
typedef void (*gdh_window_cb)(int64_t L, double* out, void* ctx);

GDH_API const char* gdh_last_panic() { return g_panic.c_str(); }
GDH_API int gdh_fft(const double* in, int64_t n, double* out, int dir, int real_in) {
    return guarded([&] {
        cvec r;
        if (real_in) { rvec x(in, in + n); r = dir > 0 ? fft::FFTReal(x) : fft::IFFTReal(x); }
        else { cvec x((const cplx*)in, (const cplx*)in + n); r = dir > 0 ? fft::FFT(x) : fft::IFFT(x); }
        std::memcpy(out, r.data(), r.size() * sizeof(cplx));
    });
}
GDH_API int gdh_convolve(const double* x, int64_t nx, const double* y, int64_t ny, double* out) {
    return guarded([&] {
        cvec r = fft::Convolve(cvec((const cplx*)x, (const cplx*)x + nx), cvec((const cplx*)y, (const cplx*)y + ny));
        std::memcpy(out, r.data(), r.size() * sizeof(cplx));
    });
}
// rows given as separate pointers + lengths, like [][]complex128
GDH_API int gdh_fft2(const double* const* rows, const int64_t* lens, int64_t nrows, double* const* out_rows, int dir, int real_in) {
    return guarded([&] {
        std::vector<cvec> x((size_t)nrows);
        for (int64_t i = 0; i < nrows; i++) {
            if (real_in) x[(size_t)i] = dsputils::ToComplex(rvec(rows[i], rows[i] + lens[i]));
            else x[(size_t)i].assign((const cplx*)rows[i], (const cplx*)rows[i] + lens[i]);
        }
        auto r = dir > 0 ? fft::FFT2(x) : fft::IFFT2(x);
        for (int64_t i = 0; i < nrows; i++) std::memcpy(out_rows[i], r[(size_t)i].data(), r[(size_t)i].size() * sizeof(cplx));
    });
}
GDH_API int gdh_fftn(const double* in, const int64_t* dims, int nd, double* out, int dir) {
    return guarded([&] {
        std::vector<int64_t> d(dims, dims + nd);
        int64_t total = 1;
        for (auto v : d) total *= v;
        auto m = dsputils::Matrix::MakeMatrix(cvec((const cplx*)in, (const cplx*)in + (total > 0 ? total : 0)), d);
        auto r = dir > 0 ? fft::FFTN(m) : fft::IFFTN(m);
        std::memcpy(out, r.list().data(), r.list().size() * sizeof(cplx));
    });
}
GDH_API int gdh_ensure_radix2_factors(int64_t n) { return guarded([&] { fft::EnsureRadix2Factors(n); }); }
GDH_API void gdh_set_worker_pool_size(int n) { fft::SetWorkerPoolSize(n); }
GDH_API int gdh_worker_pool_size() { return fft::WorkerPoolSize(); }
GDH_API uint64_t gdh_reverse_bits(uint64_t v, uint64_t s) { return fft::reverseBits(v, s); }
GDH_API int64_t gdh_next_power_of2(int64_t x) { return dsputils::NextPowerOf2(x); }
GDH_API int gdh_is_power_of2(int64_t x) { return dsputils::IsPowerOf2(x) ? 1 : 0; }
GDH_API int gdh_window(int id, int64_t L, double* out) {
    return guarded([&] {
        rvec r;
        switch (id) {
            case 0: r = window::Rectangular(L); break;
            case 1: r = window::Hamming(L); break;
            case 2: r = window::Hann(L); break;
            case 3: r = window::Bartlett(L); break;
            case 4: r = window::FlatTop(L); break;
            case 5: r = window::Blackman(L); break;
            default: throw Panic("unknown window");
        }
        std::memcpy(out, r.data(), r.size() * sizeof(double));
    });
}
GDH_API int64_t gdh_segment_count(int64_t lx, int64_t size, int64_t noverlap) {
    int64_t n = -1;
    guarded([&] { n = spectral::SegmentCount(lx, size, noverlap); });
    return n;
}
GDH_API int gdh_dsputils_segment(int64_t lx, int64_t segs, double noverlap, int64_t* offsets, int64_t* length) {
    return guarded([&] {
        auto r = dsputils::Segment(lx, segs, noverlap);
        for (size_t i = 0; i < r.size(); i++) offsets[i] = r[i].first;
        *length = r.empty() ? 0 : r[0].second;
    });
}
// spectral.Pwelch: window_id >= 0 picks a package window, -1 uses the callback (PwelchOptions.Window),
// -2 means nil (default Hann). has_opts == 0 passes a nil *PwelchOptions. Returns len(Pxx) via *lp_out.
GDH_API int gdh_pwelch(const double* x, int64_t nx, double Fs, int has_opts, int64_t nfft, int64_t pad, int64_t noverlap,
                       int scale_off, int window_id, gdh_window_cb cb, void* ctx, double* pxx, double* freqs, int64_t cap,
                       int64_t* lp_out) {
    return guarded([&] {
        spectral::PwelchOptions o;
        o.NFFT = nfft; o.Pad = pad; o.Noverlap = noverlap; o.Scale_off = scale_off != 0;
        if (window_id == -1 && cb) {
            o.Window = [cb, ctx](int64_t L) { rvec w((size_t)(L > 0 ? L : 0)); cb(L, w.data(), ctx); return w; };
        } else if (window_id >= 0) {
            o.Window = [window_id](int64_t L) {
                rvec w((size_t)(L > 0 ? L : 0));
                if (gdh_window(window_id, L, w.data()) != 0) throw Panic(g_panic);
                return w;
            };
        }
        auto r = spectral::Pwelch(rvec(x, x + nx), Fs, has_opts ? &o : nullptr);
        *lp_out = (int64_t)r.first.size();
        if ((int64_t)r.first.size() > cap) throw Panic("gdh_pwelch: output buffer too small");
        std::memcpy(pxx, r.first.data(), r.first.size() * sizeof(double));
        std::memcpy(freqs, r.second.data(), r.second.size() * sizeof(double));
    });
}

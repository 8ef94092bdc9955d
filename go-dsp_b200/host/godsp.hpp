// godsp.hpp -- host-side mirror of go-dsp's exported API over the B200 C ABI.
//
// The reference is a Go package; no Go toolchain exists in this image or on the GPU box, so
// the host side that a cgo shim would hold is written here in C++ with the reference's names,
// argument meaning and error behaviour (Go panics become godsp::Panic carrying the reference's
// message). Everything numerical below the option handling goes through include/godsp_b200.h;
// nothing here computes a transform on the CPU.
//
//   godsp::fft       <- fft/fft.go, fft/radix2.go (EnsureRadix2Factors, reverseBits)
//   godsp::spectral  <- spectral/pwelch.go, spectral/spectral.go
//   godsp::window    <- window/window.go        (host-side O(L) tables, as in the Go drop-in)
//   godsp::dsputils  <- dsputils/dsputils.go, dsputils/matrix.go, dsputils/compare.go
#pragma once
#include <complex>
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace godsp {

using cplx = std::complex<double>;
using cvec = std::vector<cplx>;
using rvec = std::vector<double>;

// Go's panic(string) / runtime panics
struct Panic : std::runtime_error {
    explicit Panic(const std::string& m) : std::runtime_error(m) {}
};

namespace dsputils {
cvec ToComplex(const rvec& x);                                   // dsputils.go:25-31
bool IsPowerOf2(int64_t x);                                      // dsputils.go:34-36 (true for 0)
int64_t NextPowerOf2(int64_t x);                                 // dsputils.go:39-45
cvec ZeroPad(const cvec& x, int64_t length);                     // dsputils.go:49-58
rvec ZeroPadF(const rvec& x, int64_t length);                    // dsputils.go:61-69
cvec ZeroPad2(const cvec& x);                                    // dsputils.go:72-74
std::vector<cvec> ToComplex2(const std::vector<rvec>& x);        // dsputils.go:77-83
// dsputils.go:89-115: segs equal-length windows with noverlap (fraction) overlap; returns
// (offset, length) pairs into x -- the Go version returns aliasing sub-slices.
std::vector<std::pair<int64_t, int64_t>> Segment(int64_t lx, int64_t segs, double noverlap);
bool Float64Equal(double a, double b);                           // compare.go:94-96
bool ComplexEqual(cplx a, cplx b);                               // compare.go:84-91
bool PrettyClose(const rvec& a, const rvec& b);                  // compare.go:27-38
bool PrettyCloseC(const cvec& a, const cvec& b);                 // compare.go:41-52

// dsputils/matrix.go:21-216: flat row-major N-d container, last dimension fastest
class Matrix {
   public:
    static Matrix MakeMatrix(const cvec& x, const std::vector<int64_t>& dims);      // matrix.go:37-57
    static Matrix MakeMatrix2(const std::vector<cvec>& x);                          // matrix.go:60-72
    static Matrix MakeEmptyMatrix(const std::vector<int64_t>& dims);                // matrix.go:83-90
    Matrix Copy() const { return *this; }                                           // matrix.go:75-80
    std::vector<int64_t> Dimensions() const { return dims_; }                       // matrix.go:144-148
    cvec Dim(const std::vector<int64_t>& idx) const;                                // matrix.go:156-164
    void SetDim(const cvec& x, const std::vector<int64_t>& idx);                    // matrix.go:166-175
    cplx Value(const std::vector<int64_t>& idx) const { return list_[offset(idx)]; }    // matrix.go:179-181
    void SetValue(cplx x, const std::vector<int64_t>& idx) { list_[offset(idx)] = x; } // matrix.go:185-187
    std::vector<cvec> To2D() const;                                                 // matrix.go:191-203
    bool PrettyClose(const Matrix& n) const;                                        // matrix.go:207-216
    const cvec& list() const { return list_; }     // same-module accessor used by fft.FFTN
    cvec& list() { return list_; }

   private:
    cvec list_;
    std::vector<int64_t> dims_, offsets_;
    int64_t offset(const std::vector<int64_t>& idx) const;                          // matrix.go:93-108
    std::vector<int64_t> indexes(const std::vector<int64_t>& idx) const;            // matrix.go:110-141
};
}  // namespace dsputils

namespace window {
using Func = std::function<rvec(int64_t)>;
void Apply(rvec& x, const Func& windowFunction);                 // window.go:25-29
rvec Rectangular(int64_t L);                                     // window.go:32-40
rvec Hamming(int64_t L);                                         // window.go:44-58
rvec Hann(int64_t L);                                            // window.go:62-76
rvec Bartlett(int64_t L);                                        // window.go:80-99
rvec FlatTop(int64_t L);                                         // window.go:103-135
rvec Blackman(int64_t L);                                        // window.go:138-152
}  // namespace window

namespace fft {
cvec FFTReal(const rvec& x);                                     // fft.go:25-27
cvec IFFTReal(const rvec& x);                                    // fft.go:30-32
cvec IFFT(const cvec& x);                                        // fft.go:35-52
cvec Convolve(const cvec& x, const cvec& y);                     // fft.go:55-69  panics "arrays not of equal size"
cvec FFT(const cvec& x);                                         // fft.go:72-87
void SetWorkerPoolSize(int n);                                   // fft.go:95-101 (kept; no effect on the GPU path)
int WorkerPoolSize();
std::vector<cvec> FFT2Real(const std::vector<rvec>& x);          // fft.go:104-106
std::vector<cvec> FFT2(const std::vector<cvec>& x);              // fft.go:109-111 panics "empty input array" / "ragged input array"
std::vector<cvec> IFFT2Real(const std::vector<rvec>& x);         // fft.go:114-116
std::vector<cvec> IFFT2(const std::vector<cvec>& x);             // fft.go:119-121
dsputils::Matrix FFTN(const dsputils::Matrix& m);                // fft.go:157-159
dsputils::Matrix IFFTN(const dsputils::Matrix& m);               // fft.go:162-164
void EnsureRadix2Factors(int64_t input_len);                     // radix2.go:35-37 -> plan / table warm-up
uint64_t reverseBits(uint64_t v, uint64_t s);                    // radix2.go:184-199 (kept: fft_test.go:241-249 calls it)
}  // namespace fft

namespace spectral {
struct PwelchOptions {                                           // pwelch.go:28-65 (zero values = defaults)
    int64_t NFFT = 0;            // 0 -> 256
    window::Func Window;         // empty -> window::Hann
    int64_t Pad = 0;             // 0 -> NFFT
    int64_t Noverlap = 0;
    bool Scale_off = false;
};
// pwelch.go:74-145; returns (Pxx, freqs)
std::pair<rvec, rvec> Pwelch(const rvec& x, double Fs, const PwelchOptions* o);
std::vector<rvec> Segment(const rvec& x, int64_t size, int64_t noverlap);   // spectral.go:22-47
int64_t SegmentCount(int64_t lx, int64_t size, int64_t noverlap);          // spectral.go:27-33
}  // namespace spectral

}  // namespace godsp

// Package fft provides forward and inverse fast Fourier transform functions.
//
// B200 build: every transform runs on the GPU through the C ABI of libgodsp_b200
// (include/godsp_b200.h). The exported API is go-dsp's; there is no CPU fallback --
// a failed device call panics.
package fft

/*
#cgo LDFLAGS: -lgodsp_b200
#include <stdlib.h>
#include "godsp_b200.h"
*/
import "C"

import (
	"unsafe"

	"github.com/mjibson/go-dsp/dsputils"
)

func check(st C.int, what string) {
	if st != 0 {
		panic(what + ": " + C.GoString(C.gd_last_error()))
	}
}

func cptr(x []complex128) *C.double { return (*C.double)(unsafe.Pointer(&x[0])) }
func fptr(x []float64) *C.double    { return (*C.double)(unsafe.Pointer(&x[0])) }

// FFT returns the forward FFT of the complex-valued slice.
func FFT(x []complex128) []complex128 {
	r := make([]complex128, len(x))
	if len(x) == 0 {
		return r
	}
	check(C.gd_fft_c2c(cptr(x), cptr(r), C.int64_t(len(x)), 1), "gd_fft_c2c")
	return r
}

// IFFT returns the inverse FFT of the complex-valued slice.
func IFFT(x []complex128) []complex128 {
	_ = x[0] // the reference indexes x[0] first: an empty slice panics the same way
	r := make([]complex128, len(x))
	check(C.gd_fft_c2c(cptr(x), cptr(r), C.int64_t(len(x)), -1), "gd_fft_c2c")
	return r
}

// FFTReal returns the forward FFT of the real-valued slice (widening fused on the device).
func FFTReal(x []float64) []complex128 {
	r := make([]complex128, len(x))
	if len(x) == 0 {
		return r
	}
	check(C.gd_fft_r2c_full(fptr(x), cptr(r), C.int64_t(len(x)), 1), "gd_fft_r2c_full")
	return r
}

// IFFTReal returns the inverse FFT of the real-valued slice.
func IFFTReal(x []float64) []complex128 {
	_ = x[0]
	r := make([]complex128, len(x))
	check(C.gd_fft_r2c_full(fptr(x), cptr(r), C.int64_t(len(x)), -1), "gd_fft_r2c_full")
	return r
}

// Convolve returns the convolution of x * y.
func Convolve(x, y []complex128) []complex128 {
	if len(x) != len(y) {
		panic("arrays not of equal size")
	}
	_ = x[0]
	r := make([]complex128, len(x))
	check(C.gd_convolve_c2c(cptr(x), cptr(y), cptr(r), C.int64_t(len(x))), "gd_convolve_c2c")
	return r
}

// FFTBatch is additive (go-dsp has no batch call, SURVEY.md 8f rank 1): x holds len(x)/n transforms of n points
// back to back; the result has the same layout. One cgo call, chunked H2D / kernels / D2H overlap inside the library;
// slices wrapped around gd_pinned_alloc memory (PinnedComplex) transfer asynchronously. dir: +1 forward, -1 inverse.
func FFTBatch(x []complex128, n int, dir int) []complex128 {
	if n <= 0 || len(x)%n != 0 {
		panic("FFTBatch: len(x) must be a multiple of n")
	}
	r := make([]complex128, len(x))
	if len(x) == 0 {
		return r
	}
	check(C.gd_fft_batch_c2c(cptr(x), cptr(r), C.int64_t(n), C.int64_t(len(x)/n), C.int(dir)), "gd_fft_batch_c2c")
	return r
}

// PinnedComplex returns a slice of n complex128 in page-locked host memory (free it with FreePinned).
func PinnedComplex(n int) []complex128 {
	p := C.gd_pinned_alloc(C.size_t(n) * 16)
	if p == nil {
		panic("gd_pinned_alloc: " + C.GoString(C.gd_last_error()))
	}
	return unsafe.Slice((*complex128)(p), n)
}

// FreePinned releases a slice obtained from PinnedComplex.
func FreePinned(x []complex128) {
	if len(x) > 0 {
		C.gd_pinned_free(unsafe.Pointer(&x[0]))
	}
}

var worker_pool_size = 0

// SetWorkerPoolSize is kept for API compatibility; the GPU path has no worker pool.
func SetWorkerPoolSize(n int) {
	if n < 0 {
		n = 0
	}
	worker_pool_size = n
}

// EnsureRadix2Factors warms the plan / twiddle (or Bluestein) caches for input_len.
func EnsureRadix2Factors(input_len int) {
	if input_len >= 1 {
		check(C.gd_plan_warm(C.int64_t(input_len)), "gd_plan_warm")
	}
}

// reverseBits returns the first s bits of v in reverse order (used by fft_test.go).
func reverseBits(v, s uint) uint {
	var r uint
	for b := uint(0); b < s; b++ {
		r |= ((v >> b) & 1) << (s - 1 - b)
	}
	return r
}

// fft2 stages the separately allocated rows into one pinned block, runs the 2-D transform
// (columns, then rows, as the reference orders them) and copies the rows back.
func fft2(x [][]complex128, dir C.int) [][]complex128 {
	rows := len(x)
	if rows == 0 {
		panic("empty input array")
	}
	cols := len(x[0])
	r := make([][]complex128, rows)
	for i := 0; i < rows; i++ {
		if len(x[i]) != cols {
			panic("ragged input array")
		}
		r[i] = make([]complex128, cols)
	}
	if cols == 0 {
		return r
	}
	n := rows * cols
	p := C.gd_pinned_alloc(C.size_t(2 * n * 16))
	if p == nil {
		panic("gd_pinned_alloc: " + C.GoString(C.gd_last_error()))
	}
	defer C.gd_pinned_free(p)
	buf := unsafe.Slice((*complex128)(p), 2*n)
	for i := 0; i < rows; i++ {
		copy(buf[i*cols:(i+1)*cols], x[i])
	}
	check(C.gd_fft2_c2c((*C.double)(p), (*C.double)(unsafe.Pointer(&buf[n])), C.int64_t(rows), C.int64_t(cols), dir), "gd_fft2_c2c")
	for i := 0; i < rows; i++ {
		copy(r[i], buf[n+i*cols:n+(i+1)*cols])
	}
	return r
}

// FFT2 returns the 2-dimensional, forward FFT of the complex-valued matrix.
func FFT2(x [][]complex128) [][]complex128 { return fft2(x, 1) }

// IFFT2 returns the 2-dimensional, inverse FFT of the complex-valued matrix.
func IFFT2(x [][]complex128) [][]complex128 { return fft2(x, -1) }

// FFT2Real returns the 2-dimensional, forward FFT of the real-valued matrix.
func FFT2Real(x [][]float64) [][]complex128 { return fft2(dsputils.ToComplex2(x), 1) }

// IFFT2Real returns the 2-dimensional, inverse FFT of the real-valued matrix.
func IFFT2Real(x [][]float64) [][]complex128 { return fft2(dsputils.ToComplex2(x), -1) }

// lastAxisLines walks every line along the last dimension in row-major order; for the
// Matrix layout (last dimension fastest) their concatenation is the flat backing list.
func lastAxisLines(dims []int, f func(idx []int)) {
	nd := len(dims)
	idx := make([]int, nd)
	idx[nd-1] = -1
	for {
		f(idx)
		k := nd - 2
		for ; k >= 0; k-- {
			idx[k]++
			if idx[k] < dims[k] {
				break
			}
			idx[k] = 0
		}
		if k < 0 {
			return
		}
	}
}

func fftn(m *dsputils.Matrix, dir C.int) *dsputils.Matrix {
	dims := m.Dimensions()
	total, last := 1, dims[len(dims)-1]
	for _, d := range dims {
		total *= d
	}
	in := make([]complex128, 0, total)
	lastAxisLines(dims, func(idx []int) { in = append(in, m.Dim(idx)...) })
	out := make([]complex128, total)
	cd := make([]C.int64_t, len(dims))
	for i, d := range dims {
		cd[i] = C.int64_t(d)
	}
	check(C.gd_fftn_c2c(cptr(in), cptr(out), &cd[0], C.int(len(dims)), dir), "gd_fftn_c2c")
	r := dsputils.MakeEmptyMatrix(dims)
	off := 0
	lastAxisLines(dims, func(idx []int) { r.SetDim(out[off:off+last], idx); off += last })
	return r
}

// FFTN returns the forward FFT of the matrix m, computed in all N dimensions.
func FFTN(m *dsputils.Matrix) *dsputils.Matrix { return fftn(m, 1) }

// IFFTN returns the inverse FFT of the matrix m, computed in all N dimensions.
func IFFTN(m *dsputils.Matrix) *dsputils.Matrix { return fftn(m, -1) }

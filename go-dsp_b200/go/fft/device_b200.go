package fft

/*
#include <stdlib.h>
#include "godsp_b200.h"
*/
import "C"

import "unsafe"

// Additive API of the B200 build (SURVEY.md 8f): go-dsp itself has none of these. They exist so that a chain of
// transforms pays the PCIe copy once and so that a whole box can be used from one process.

// Init makes the batched calls (FFTBatch, FFT2, spectral.Pwelch) spread over the first ndev GPUs of the box, one host
// thread per device inside the library (ndev <= 0: every visible GPU). Without it every call runs on one device.
func Init(ndev int) { check(C.gd_init(C.int(ndev)), "gd_init") }

// DeviceBuffer is n complex128 elements resident in GPU memory: an opaque handle and a length.
type DeviceBuffer struct {
	ptr unsafe.Pointer
	n   int
}

// NewDeviceBuffer allocates n elements on the calling thread's device.
func NewDeviceBuffer(n int) *DeviceBuffer {
	b := &DeviceBuffer{n: n}
	sz := n
	if sz < 1 {
		sz = 1
	}
	check(C.gd_dev_alloc(&b.ptr, C.size_t(16*sz)), "gd_dev_alloc")
	return b
}

// Upload copies x (len(x) == b.Len()) to the device; Download returns a fresh slice with the contents.
func (b *DeviceBuffer) Upload(x []complex128) {
	if len(x) != b.n {
		panic("DeviceBuffer.Upload: length mismatch")
	}
	if b.n > 0 {
		check(C.gd_memcpy_h2d(b.ptr, unsafe.Pointer(&x[0]), C.size_t(16*b.n)), "gd_memcpy_h2d")
	}
}
func (b *DeviceBuffer) Download() []complex128 {
	r := make([]complex128, b.n)
	if b.n > 0 {
		check(C.gd_memcpy_d2h(unsafe.Pointer(&r[0]), b.ptr, C.size_t(16*b.n)), "gd_memcpy_d2h")
	}
	return r
}
func (b *DeviceBuffer) Len() int { return b.n }
func (b *DeviceBuffer) Free() {
	if b.ptr != nil {
		check(C.gd_dev_free(b.ptr), "gd_dev_free")
		b.ptr = nil
	}
}

// FFT runs Len()/n transforms of n points back to back on resident data (dir +1 = FFT, -1 = IFFT).
func (b *DeviceBuffer) FFT(n, dir int) *DeviceBuffer {
	if n <= 0 || b.n%n != 0 {
		panic("DeviceBuffer.FFT: length must be a multiple of n")
	}
	r := NewDeviceBuffer(b.n)
	check(C.gd_fft_batch_c2c_dev((*C.double)(b.ptr), (*C.double)(r.ptr), C.int64_t(n), C.int64_t(b.n/n), C.int(dir), nil), "gd_fft_batch_c2c_dev")
	check(C.gd_stream_sync(nil), "gd_stream_sync")
	return r
}

// Convolve is fft.Convolve on resident data.
func (b *DeviceBuffer) Convolve(y *DeviceBuffer) *DeviceBuffer {
	if b.n != y.n {
		panic("arrays not of equal size")
	}
	r := NewDeviceBuffer(b.n)
	check(C.gd_convolve_c2c_dev((*C.double)(b.ptr), (*C.double)(y.ptr), (*C.double)(r.ptr), C.int64_t(b.n), nil), "gd_convolve_c2c_dev")
	check(C.gd_stream_sync(nil), "gd_stream_sync")
	return r
}

// ConvolveLinear returns the linear convolution of x and h (len(x)+len(h)-1 values) by overlap-save on the circular
// Convolve: what Convolve(ZeroPad(x, m), ZeroPad(h, m))[:n] gives for m = NextPowerOf2(n), n = len(x)+len(h)-1.
func ConvolveLinear(x, h []complex128) []complex128 {
	if len(x) == 0 || len(h) == 0 {
		return []complex128{}
	}
	r := make([]complex128, len(x)+len(h)-1)
	check(C.gd_convolve_linear_c2c(cptr(x), C.int64_t(len(x)), cptr(h), C.int64_t(len(h)), cptr(r)), "gd_convolve_linear_c2c")
	return r
}

// FFTSegments returns FFT(dsputils.ZeroPad2(s)) for every slice s = x[i*step : i*step+length] that dsputils.Segment
// returns (pass its length and step); the slices are described to the GPU, not copied. fftlen = NextPowerOf2(length).
func FFTSegments(x []complex128, length, step, segs, fftlen int) [][]complex128 {
	flat := make([]complex128, segs*fftlen)
	check(C.gd_fft_segments_c2c(cptr(x), C.int64_t(len(x)), C.int64_t(length), C.int64_t(step), C.int64_t(segs), C.int64_t(fftlen), cptr(flat)), "gd_fft_segments_c2c")
	r := make([][]complex128, segs)
	for i := range r {
		r[i] = flat[i*fftlen : (i+1)*fftlen]
	}
	return r
}

package spectral

/*
#include "godsp_b200.h"
*/
import "C"

import (
	"unsafe"

	"github.com/mjibson/go-dsp/window"
)

// Additive API of the B200 build (SURVEY.md 8f ranks 1-3): wav ingest, streaming Pwelch, spectrogram.

type resolved struct {
	nfft, pad, fftlen, lp int
	win                   []float64
	norm                  float64
}

// option defaults, the two window evaluations and the norm exactly as Pwelch forms them (spectral/pwelch.go:79-102,124-128)
func resolve(o *PwelchOptions) resolved {
	r := resolved{nfft: o.NFFT, pad: o.Pad}
	wf := o.Window
	if r.nfft == 0 {
		r.nfft = 256
	}
	if wf == nil {
		wf = window.Hann
	}
	if r.pad == 0 {
		r.pad = r.nfft
	}
	r.fftlen = r.nfft
	if r.pad > r.fftlen {
		r.fftlen = r.pad
	}
	r.lp = r.pad/2 + 1
	r.win = wf(r.fftlen)
	for _, v := range wf(r.nfft) {
		r.norm += v * v
	}
	return r
}

func freqsOf(Fs float64, pad, lp int) []float64 {
	f := make([]float64, lp)
	coef := Fs / float64(pad)
	for i := range f {
		f[i] = float64(i) * coef
	}
	return f
}

// PwelchStream is spectral.Pwelch over a signal that arrives in chunks: what wav.ReadSamples returns ([]uint8, []int16,
// []float32) or []float64 goes to the GPU as it is; the wav.ReadFloats conversion (wav/wav.go:138-161) happens in the
// segment load of the kernel, so the signal crosses PCIe at its on-disk width.
type PwelchStream struct {
	h   unsafe.Pointer
	o   *PwelchOptions
	r   resolved
	fmt C.int
}

func NewPwelchStream(o *PwelchOptions, sampleFmt int) *PwelchStream {
	s := &PwelchStream{o: o, r: resolve(o), fmt: C.int(sampleFmt)}
	st := C.gd_pwelch_stream_begin(&s.h, s.fmt, C.int64_t(s.r.nfft), C.int64_t(o.Noverlap), C.int64_t(s.r.fftlen), C.int64_t(s.r.lp),
		(*C.double)(unsafe.Pointer(&s.r.win[0])))
	if st != 0 {
		panic("gd_pwelch_stream_begin: " + C.GoString(C.gd_last_error()))
	}
	return s
}

// Push accepts []uint8, []int16, []float32 or []float64 matching the stream's sample format.
func (s *PwelchStream) Push(samples interface{}) {
	var p unsafe.Pointer
	var n int
	switch d := samples.(type) {
	case []uint8:
		p, n = unsafe.Pointer(&d[0]), len(d)
	case []int16:
		p, n = unsafe.Pointer(&d[0]), len(d)
	case []float32:
		p, n = unsafe.Pointer(&d[0]), len(d)
	case []float64:
		p, n = unsafe.Pointer(&d[0]), len(d)
	default:
		panic("PwelchStream.Push: unsupported sample type")
	}
	if st := C.gd_pwelch_stream_push(s.h, p, C.int64_t(n)); st != 0 {
		panic("gd_pwelch_stream_push: " + C.GoString(C.gd_last_error()))
	}
}

func (s *PwelchStream) Finish(Fs float64) (Pxx, freqs []float64) {
	norm := s.r.norm
	if !s.o.Scale_off {
		norm *= Fs
	}
	Pxx = make([]float64, s.r.lp)
	var nsegs C.int64_t
	if st := C.gd_pwelch_stream_end(s.h, C.double(norm), (*C.double)(unsafe.Pointer(&Pxx[0])), &nsegs); st != 0 {
		panic("gd_pwelch_stream_end: " + C.GoString(C.gd_last_error()))
	}
	return Pxx, freqsOf(Fs, s.r.pad, s.r.lp)
}

// Spectrogram is the segment loop of Pwelch without the accumulate (spectral/pwelch.go:104-113):
// S[c][j] = FFT(window * segment c, zero-padded)[j], j < pad/2+1.
func Spectrogram(x []float64, Fs float64, o *PwelchOptions) (S [][]complex128, freqs []float64) {
	r := resolve(o)
	stride := r.nfft - o.Noverlap
	nsegs := 0
	if len(x) == r.nfft {
		nsegs = 1
	} else if len(x) > r.nfft {
		nsegs = (len(x)-r.nfft)/stride + 1
	}
	flat := make([]complex128, nsegs*r.lp)
	if nsegs > 0 {
		st := C.gd_stft_f64((*C.double)(unsafe.Pointer(&x[0])), C.int64_t(len(x)), C.int64_t(r.nfft), C.int64_t(o.Noverlap), C.int64_t(r.fftlen),
			C.int64_t(r.lp), C.int64_t(nsegs), (*C.double)(unsafe.Pointer(&r.win[0])), (*C.double)(unsafe.Pointer(&flat[0])))
		if st != 0 {
			panic("gd_stft_f64: " + C.GoString(C.gd_last_error()))
		}
	}
	S = make([][]complex128, nsegs)
	for c := range S {
		S[c] = flat[c*r.lp : (c+1)*r.lp]
	}
	return S, freqsOf(Fs, r.pad, r.lp)
}

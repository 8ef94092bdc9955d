package spectral

/*
#cgo LDFLAGS: -lgodsp_b200
#include "godsp_b200.h"
*/
import "C"

import (
	"unsafe"

	"github.com/mjibson/go-dsp/dsputils"
	"github.com/mjibson/go-dsp/window"
)

// PwelchOptions keeps go-dsp's fields and zero-value defaults.
type PwelchOptions struct {
	// NFFT is the number of data points used in each block for the FFT (default 256).
	NFFT int
	// Window returns the window values for a given length (default window.Hann).
	Window func(int) []float64
	// Pad is the length each segment is zero-padded to before the FFT (default NFFT).
	Pad int
	// Noverlap is the number of points of overlap between blocks (default 0).
	Noverlap int
	// Scale_off disables scaling of the density by the sampling frequency.
	Scale_off bool
}

// Pwelch estimates the power spectral density of x using Welch's method. The option
// defaults, window tables, norm and frequency vector are formed here exactly as go-dsp
// forms them; the segment loop (gather, window, transform, |X|^2 accumulation) is one
// fused kernel on the device. Returns the PSD Pxx and the frequencies freqs.
func Pwelch(x []float64, Fs float64, o *PwelchOptions) (Pxx, freqs []float64) {
	if len(x) == 0 {
		return []float64{}, []float64{}
	}
	nfft, pad, noverlap, wf := o.NFFT, o.Pad, o.Noverlap, o.Window
	if nfft == 0 {
		nfft = 256
	}
	if wf == nil {
		wf = window.Hann
	}
	if pad == 0 {
		pad = nfft
	}
	if len(x) < nfft {
		x = dsputils.ZeroPadF(x, nfft)
	}
	lp := pad/2 + 1
	stride := nfft - noverlap
	var nsegs int // len(Segment(x, nfft, noverlap)) without materialising the copies
	if len(x) == nfft {
		nsegs = 1
	} else {
		nsegs = (len(x)-nfft)/stride + 1
	}
	if nsegs < 1 { // Noverlap < 0 is valid: stride > nfft leaves gaps between segments, as spectral.Segment does
		panic("runtime error: makeslice: len out of range")
	}
	fftlen := nfft
	if pad > fftlen {
		fftlen = pad
	}
	win := wf(fftlen) // window.Apply(x, wf) evaluates wf(len(x)) on the padded segment
	_ = win[fftlen-1]
	var norm float64
	for _, v := range wf(nfft) {
		norm += v * v
	}
	if !o.Scale_off {
		norm *= Fs
	}
	Pxx = make([]float64, lp)
	st := C.gd_pwelch_f64((*C.double)(unsafe.Pointer(&x[0])), C.int64_t(len(x)), C.int64_t(nfft), C.int64_t(noverlap),
		C.int64_t(fftlen), C.int64_t(lp), C.int64_t(nsegs), (*C.double)(unsafe.Pointer(&win[0])), C.double(norm),
		(*C.double)(unsafe.Pointer(&Pxx[0])))
	if st != 0 {
		panic("gd_pwelch_f64: " + C.GoString(C.gd_last_error()))
	}
	freqs = make([]float64, lp)
	coef := Fs / float64(pad)
	for i := range freqs {
		freqs[i] = float64(i) * coef
	}
	return
}

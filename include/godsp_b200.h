/*
 * godsp_b200.h -- C ABI of libgodsp_b200.so, the B200 (sm_100a) engine behind go-dsp's
 * exported Go API for the FFT / Welch-PSD path.
 *
 * The reference (github.com/mjibson/go-dsp) is pure Go with no FFI; these are the entry
 * points a cgo shim inside its fft/ and spectral/ packages binds (INTEGRATION.md shows the
 * shim). Each function names the reference interface it replaces (file:line relative to the
 * go-dsp repository root).
 *
 * Conventions
 *   - complex128 slices are passed as `double*` to interleaved (re, im) pairs -- the memory
 *     layout of Go's []complex128 and of CUDA's double2; []float64 as `double*`.
 *   - every function returns 0 on success or a negative gd_status; gd_last_error() gives the
 *     message (thread-local). There is NO CPU fallback: without a B200 every compute entry
 *     point fails with GD_ERR_CUDA / GD_ERR_UNSUPPORTED and the Go shim panics.
 *   - host-pointer entry points are synchronous: the caller owns all buffers, `out` must not
 *     alias `in`, nothing is retained after return (cgo pointer rules). Pageable and pinned
 *     (gd_pinned_alloc) host memory are both accepted; pinned memory is copied asynchronously.
 *   - entry points may be called from any OS thread (goroutines migrate): each call selects
 *     its CUDA device explicitly and serialises on that device's mutex.
 *   - dir: +1 forward (fft.FFT), -1 inverse including the 1/N (fft.IFFT).
 */
#ifndef GODSP_B200_H
#define GODSP_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define GD_API __attribute__((visibility("default")))
#else
#define GD_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

enum gd_status {
    GD_OK = 0,
    GD_ERR_INVALID = -1,      /* bad argument (the Go shim checks and panics with the reference's strings first) */
    GD_ERR_CUDA = -2,         /* CUDA runtime failure */
    GD_ERR_NOMEM = -3,        /* device or pinned allocation failed */
    GD_ERR_UNSUPPORTED = -4,  /* size outside what this build plans (e.g. one transform > 2^24 points on one GPU) */
    GD_ERR_NOT_INIT = -5
};

/* ---- lifecycle ------------------------------------------------------------------------- */
/* Initialise devices 0..ndev-1 (ndev <= 0: every visible device). Idempotent. Called lazily
 * (with ndev = 1) by every other entry point, so a Go program never has to call it. */
GD_API int gd_init(int ndev);
GD_API int gd_shutdown(void);
GD_API const char* gd_last_error(void);
GD_API int gd_device_count(void);            /* devices initialised so far */
GD_API int gd_use_device(int dev);           /* device used by subsequent calls from this thread (default 0) */
/* Tuning knobs: "pass_scratch_mb" (inter-pass scratch kept L2-resident), "wide_tiles" (0/1), "fanout_min_log2n". */
GD_API int gd_set_option(const char* key, int64_t value);

/* ---- fft package ----------------------------------------------------------------------- */
/* fft.FFT (fft/fft.go:72-87) / fft.IFFT (fft/fft.go:35-52): any n >= 1; power-of-two n runs the
 * Stockham passes that replace radix2FFT (fft/radix2.go:80-154), other n the fused Bluestein path
 * that replaces bluesteinFFT (fft/bluestein.go:68-94). After gd_init(ndev > 1), a power-of-two n of at least
 * 2^26 points (option "fanout_min_log2n") is ONE transform sharded over the devices of the process: four-step
 * with the twiddle, the transpose and the peer-memory stores over NVLink in one kernel (BASELINE config C5). */
GD_API int gd_fft_c2c(const double* in, double* out, int64_t n, int dir);
/* fft.FFTReal / fft.IFFTReal (fft/fft.go:25-32): float64 in, full n-bin complex out; the
 * dsputils.ToComplex widening (dsputils/dsputils.go:25-31) is fused into the first load. */
GD_API int gd_fft_r2c_full(const double* in_real, double* out, int64_t n, int dir);
/* `batch` independent transforms stored back to back (additive API: the reference has no batch
 * call; this is what a loop of fft.FFT over rows lowers to). */
GD_API int gd_fft_batch_c2c(const double* in, double* out, int64_t n, int64_t batch, int dir);
/* fft.Convolve (fft/fft.go:55-69), equal lengths (the shim panics "arrays not of equal size"). */
GD_API int gd_convolve_c2c(const double* x, const double* y, double* out, int64_t n);
/* fft.FFT2 / fft.IFFT2 (fft/fft.go:109-154) on a dense row-major rows x cols array (the shim
 * stages [][]complex128 rows into one pinned block). */
GD_API int gd_fft2_c2c(const double* in, double* out, int64_t rows, int64_t cols, int dir);
/* fft.FFTN / fft.IFFTN (fft/fft.go:157-224) on dsputils.Matrix's flat row-major list
 * (dsputils/matrix.go:21-57: last dimension fastest). */
GD_API int gd_fftn_c2c(const double* in, double* out, const int64_t* dims, int nd, int dir);
/* fft.EnsureRadix2Factors (fft/radix2.go:35-37): build and cache the tables / Bluestein plan for n. */
GD_API int gd_plan_warm(int64_t n);
/* dsputils.NextPowerOf2(2n-1) (dsputils/dsputils.go:39-45) as the engine computes it: the
 * Bluestein padded length for n (bit-exact requirement). */
GD_API int64_t gd_bluestein_padded_len(int64_t n);

/* ---- spectral package ------------------------------------------------------------------ */
/* The segment loop of spectral.Pwelch (spectral/pwelch.go:104-122,134-136). The Go side keeps
 * the option defaults, window evaluation, norm and freqs (pwelch.go:85-102,124-132,138-142):
 *   nfft, noverlap : resolved option values;  fftlen = max(pad, nfft);  lp = pad/2 + 1
 *   nsegs          : len(spectral.Segment(x, nfft, noverlap)) (spectral/spectral.go:22-33)
 *   win            : Window(fftlen) (window.Apply, window/window.go:25-29)
 *   norm           : sum(Window(nfft)^2) [* Fs]
 * x must hold at least (nsegs-1)*(nfft-noverlap) + nfft samples. pxx receives lp values. */
GD_API int gd_pwelch_f64(const double* x, int64_t nx, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp,
                  int64_t nsegs, const double* win, double norm, double* pxx);


/* ---- the callers and data formats either side of the path ---------------- */
/* Sample formats: float64, or the formats wav.ReadSamples yields (wav/wav.go:113-136), decoded on the device exactly as
 * wav.ReadFloats does (wav/wav.go:138-161, float32 arithmetic: uint8 -> v/255, int16 -> (v+32768)/65535, float32 as is)
 * and widened to float64. The decode is part of the Pwelch kernel's segment load, so the signal crosses PCIe at its
 * on-disk width (1, 2 or 4 bytes per sample instead of 8). */
enum { GD_SAMPLE_F64 = 0, GD_SAMPLE_F32 = 1, GD_SAMPLE_S16 = 2, GD_SAMPLE_U8 = 3 };
/* gd_pwelch_f64 for any sample format. */
GD_API int gd_pwelch_samples(const void* x, int sample_fmt, int64_t nx, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp,
                      int64_t nsegs, const double* win, double norm, double* pxx);
/* Streaming spectral.Pwelch: the signal arrives in chunks of any size (a wav file read block by block); segments that
 * straddle two chunks are handled by the library. begin -> push* -> end; end writes Pxx (lp values, same scaling as
 * gd_pwelch_f64 with the final segment count) and the number of segments, and frees the handle. Pass pxx = NULL to abandon. */
GD_API int gd_pwelch_stream_begin(void** handle, int sample_fmt, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp, const double* win);
GD_API int gd_pwelch_stream_push(void* handle, const void* samples, int64_t n);
GD_API int gd_pwelch_stream_end(void* handle, double norm, double* pxx, int64_t* nsegs_out);
/* STFT / spectrogram: the segment loop of spectral.Pwelch without the accumulate (spectral/pwelch.go:104-113):
 * out[c*lp + j] = FFT(win * segment c, zero-padded to fftlen)[j], j < lp (lp <= fftlen), c < nsegs; complex128. */
GD_API int gd_stft_f64(const double* x, int64_t nx, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp, int64_t nsegs,
                const double* win, double* out);
/* fft.FFT of every slice dsputils.Segment returns (dsputils/dsputils.go:89-115: segs slices of seg_len elements, `step`
 * apart, aliasing x), each zero-padded to fftlen as dsputils.ZeroPad2 does (dsputils.go:72-75; fftlen a power of two
 * >= seg_len). The slices are described, not copied: the transform reads them in place. out: segs x fftlen complex128. */
GD_API int gd_fft_segments_c2c(const double* x, int64_t nx, int64_t seg_len, int64_t step, int64_t segs, int64_t fftlen, double* out);
/* Linear (non-circular) convolution, nx + nh - 1 outputs, by overlap-save on top of fft.Convolve's circular product
 * (fft/fft.go:55-69); equal to Convolve of both operands zero-padded (dsputils.ZeroPad) to a power of two >= nx + nh - 1. */
GD_API int gd_convolve_linear_c2c(const double* x, int64_t nx, const double* h, int64_t nh, double* out);

/* ---- staging memory for the shim -------------------------------------------------------- */
GD_API void* gd_pinned_alloc(size_t bytes);
GD_API void gd_pinned_free(void* p);

/* ---- device-resident API (benchmarks, verification, the additive batched Go API) -------- */
/* Device pointers; asynchronous on `stream` (a cudaStream_t; NULL = the library's stream). */
GD_API int gd_dev_alloc(void** p, size_t bytes);
GD_API int gd_dev_free(void* p);
GD_API int gd_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes);
GD_API int gd_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes);
GD_API int gd_stream_sync(void* stream);
GD_API int gd_fill_splitmix_dev(double* dst_dev, int64_t n, uint64_t seed, uint64_t offset, void* stream);
GD_API int gd_fft_batch_c2c_dev(const double* in_dev, double* out_dev, int64_t n, int64_t batch, int dir, void* stream);
GD_API int gd_fft_batch_r2c_full_dev(const double* in_dev, double* out_dev, int64_t n, int64_t batch, int dir, void* stream);
GD_API int gd_convolve_c2c_dev(const double* x_dev, const double* y_dev, double* out_dev, int64_t n, void* stream);
GD_API int gd_fftn_c2c_dev(const double* in_dev, double* out_dev, const int64_t* dims, int nd, int dir, void* stream);
/* Lines of length len and element stride `stride` (stride adjacent lines per block, `outer` blocks): the FFTN axis
 * primitive (fft/fft.go:175-189), exposed for the distributed four-step. in may equal out. */
GD_API int gd_fft_strided_c2c_dev(const double* in_dev, double* out_dev, int64_t outer, int64_t len, int64_t stride, int dir, void* stream);
/* Distributed four-step building blocks (BASELINE config 5: one 2^32-point transform over 8 GPUs):
 * blk[r][c] *= w_N^((row0+r)*(col0+c)), N = 2^log2n;  and the all-to-all receive layout [G][K][W] -> rows [K][G*W]. */
GD_API int gd_fourstep_twiddle_dev(double* blk_dev, int64_t rows, int64_t cols, int64_t row0, int64_t col0, int log2n, void* stream);
GD_API int gd_repack_gkw_dev(const double* in_dev, double* out_dev, int64_t g, int64_t k, int64_t w, void* stream);
/* in[batch][rows][cols] -> out[batch][cols][rows] (complex128): per-source-rank transpose of the all-to-all receive buffer */
GD_API int gd_transpose_batched_dev(const double* in_dev, double* out_dev, int64_t batch, int64_t rows, int64_t cols, void* stream);
/* The exchange step as one kernel over peer memory (NVLink P2P stores) instead of twiddle + all-to-all + transpose:
 * slab = this rank's [n1][w] block after the length-n1 lines; element (k1, c) times w_N^(k1 * (rank*w + c)) is stored
 * into peer_recv[k1 / (n1/world)] at [rank*w + c][k1 % (n1/world)] (rows of n1/world elements). peer_recv: host array of
 * `world` device pointers (own buffer at index `rank`, the others opened with gd_ipc_open). The caller orders the
 * ranks around it (a stream-ordered collective before and after). */
GD_API int gd_fourstep_exchange_dev(const double* slab_dev, void* const* peer_recv, int64_t n1, int64_t w, int rank, int world,
                                    int log2n, void* stream);
/* The length-n1 lines of the slab (slab -> tmp, both [n1][w]) and the exchange above, pipelined over blocks of columns:
 * the NVLink stores of one block overlap the butterflies of the next. Same rendezvous rules as gd_fourstep_exchange_dev. */
GD_API int gd_fourstep_lines_exchange_dev(const double* slab_dev, double* tmp_dev, void* const* peer_recv, int64_t n1, int64_t w, int rank,
                                          int world, int log2n, void* stream);
/* The same transform with the exchange FUSED INTO THE FIRST LINE PASS: the length-n1 lines of the slab run through the fused
 * TMA kernel, the outputs leave multiplied by w_N^(k1 n2), and the kernel's TMA stores land directly in the ranks' receive buffers
 * over NVLink: row k1 of the slab becomes row k1 % K of block `rank` ([K][w], K = n1/world) of peer_recv[k1 / K], so every receive
 * buffer is [world][K][w] (no transpose, no intermediate slab, no separate exchange kernel). gd_fourstep_rows_seg_dev then
 * transforms the K rows of a receive buffer -- each n2 = world*w points in `world` segments -- into out[k1 local][k2] =
 * X[k1 + n1*k2] ([K][n2]). gd_fourstep_fused_supported: 1 when both line lengths are in the fused kernel's range for this
 * world size (n1 = 2^13..2^17, n2 = 2^13..2^18, world 1, 2, 4 or 8). Same rendezvous rules as gd_fourstep_exchange_dev.
 * dir = -1: inverse lines (1/length each) and the conjugate twiddle. log2n = 0: no twiddle -- the column pass of fft.FFT2 on row
 * blocks (fft/fft.go:138-144) whose stores land in the ranks' row blocks, followed by the row pass (fft.go:146-151) on segmented
 * rows. Replaces the log2(N) sweeps of fft/radix2.go:131-151 for one transform spread over the GPUs of a node. */
GD_API int gd_fourstep_fused_supported(int64_t n1, int64_t n2, int world);
GD_API int gd_fourstep_lines_peer_dev(const double* slab_dev, void* const* peer_recv, int64_t n1, int64_t w, int rank, int world, int log2n,
                                      int dir, void* stream);
GD_API int gd_fourstep_rows_seg_dev(const double* recv_dev, double* out_dev, int64_t n2, int64_t k, int world, int dir, void* stream);
/* FFT2 on row blocks: an exchange as strided block copies into peer memory; for every peer h (complex128 elements):
 * peers[h][dst_off + r*dst_pitch + c] = src[h*src_step + r*src_pitch + c], r < rows, c < cols */
GD_API int gd_peer_block_copy_dev(const double* src_dev, void* const* peers, int world, int rank, int64_t rows, int64_t cols, int64_t src_step,
                                  int64_t src_pitch, int64_t dst_off, int64_t dst_pitch, void* stream);
/* cudaMalloc'ed buffer + its 64-byte CUDA IPC handle; open / close a peer's handle in this process */
GD_API int gd_ipc_alloc(void** p, size_t bytes, unsigned char* handle64);
GD_API int gd_ipc_open(const unsigned char* handle64, void** p);
GD_API int gd_ipc_close(void* p);
/* raw[j] = sum over segments seg0..seg0+nseg-1 of |FFT(win * segment)[j]|^2, j < lp (one GPU's share) */
GD_API int gd_pwelch_partial_dev(const double* x_dev, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp,
                          int64_t seg0, int64_t nseg, const double* win_dev, double* raw_dev, void* stream);
GD_API int gd_pwelch_partial_samples_dev(const void* x_dev, int sample_fmt, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp,
                                  int64_t seg0, int64_t nseg, const double* win_dev, double* raw_dev, void* stream);
GD_API int gd_stft_f64_dev(const double* x_dev, int64_t nfft, int64_t noverlap, int64_t fftlen, int64_t lp, int64_t seg0, int64_t nseg,
                    const double* win_dev, double* out_dev, void* stream);
GD_API int gd_convolve_linear_c2c_dev(const double* x_dev, int64_t nx, const double* h_dev, int64_t nh, double* out_dev, void* stream);
/* pxx[j] = raw[j] / nsegs (x2 for 0 < j < lp-1) / norm */
GD_API int gd_pwelch_finalize_dev(const double* raw_dev, int64_t lp, int64_t nsegs, double norm, double* pxx_dev, void* stream);
/* number of kernels this library has launched on the calling thread's device since gd_init */
GD_API int64_t gd_kernel_launches(void);
/* Measurement only: after gd_set_option("tma_prof", 1) the fused 2^20 kernel keeps 32 cycle counters per CTA (where its
 * loader, storers and consumer groups waited); copies up to max_ctas * 32 values of the last launch to `out` and returns
 * the number of CTAs, or a negative status. */
GD_API int gd_tma_profile_read(int64_t* out, int max_ctas);

#ifdef __cplusplus
}
#endif
#endif /* GODSP_B200_H */

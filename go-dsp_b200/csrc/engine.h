// engine.h -- device-side planner / executor behind the C ABI (include/godsp_b200.h).
// Everything here works on DEVICE pointers and is asynchronous on the given stream.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "fft_pass.cuh"

namespace gd {

enum Status : int {
    GD_OK = 0,
    GD_ERR_INVALID = -1,    // bad argument
    GD_ERR_CUDA = -2,       // CUDA runtime failure (message in gd_last_error)
    GD_ERR_NOMEM = -3,      // device / pinned allocation failed
    GD_ERR_UNSUPPORTED = -4,
    GD_ERR_NOT_INIT = -5,
};

extern std::atomic<long long> g_launches;   // kernels launched by this library

void set_error(const std::string& msg);
const char* last_error();
Status cuda_fail(cudaError_t e, const char* what);

#define GD_CUDA(call)                                              \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return ::gd::cuda_fail(e__, #call); \
    } while (0)
#define GD_TRY(call)                          \
    do {                                      \
        ::gd::Status s__ = (call);            \
        if (s__ != ::gd::GD_OK) return s__;   \
    } while (0)

// first-pass load / last-pass store hooks of a (possibly multi-pass) power-of-two transform
struct FusedOps {
    int ld_flags = 0;
    const cpx* aux_in = nullptr;
    long long n_valid_in = 0;
    int st_flags = 0;
    const cpx* aux_out = nullptr;
    long long n_valid_out = 0;
    double scale = 1.0;
    double div = 1.0;
};

struct TwiddleTable {       // w_M^e two-level table
    cpx* lo = nullptr;
    cpx* hi = nullptr;
};

struct BluesteinPlan {      // fft/bluestein.go:26-65 cache, plus the cached FFT(b) (bluestein.go:78-87)
    long long n = 0, la = 0;
    int log2la = 0;
    cpx* chirp_inv = nullptr;   // conj chirp, n entries
    cpx* bhat = nullptr;        // FFT_la(b), la entries
};

enum ScratchSlot { SCR_PASS = 0, SCR_A, SCR_B, SCR_C, SCR_STAGE_IN, SCR_STAGE_OUT, SCR_PWELCH, SCR_AUX, SCR_CNT, SCR_TMA, SCR_PROF, SCR_HUGE, SCR_NSLOTS };

struct Device {
    int dev = -1;
    int lane = 0;                        // execution lane of the device this object is (capi.cu); lane 0 owns the persisting L2 window
    int slot() const { return dev * 2 + lane; }
    int num_sms = 0;
    bool ready = false;
    cudaStream_t stream = nullptr;       // compute stream of the host-pointer entry points
    cudaStream_t stream_in = nullptr;    // H2D
    cudaStream_t stream_out = nullptr;   // D2H
    static constexpr int AUX_STREAMS = 3;
    cudaStream_t stream_aux[AUX_STREAMS] = {nullptr};   // the other L2-sized chunks of a multi-pass transform (ForkJoin, engine.cu)
    cudaEvent_t ev_fork = nullptr, ev_pipe = nullptr, ev_join[AUX_STREAMS] = {nullptr};
    cpx* wl[14] = {nullptr};             // intra-line tables exp(-2 pi i e/L), L = 2^k
    std::map<int, TwiddleTable> tw;      // keyed by log2 M
    std::map<long long, BluesteinPlan> blue;
    void* scratch[SCR_NSLOTS] = {nullptr};
    size_t scratch_bytes[SCR_NSLOTS] = {0};
    size_t pass_scratch_budget = 1ull << 30;    // upper bound of the inter-pass scratch of the two-launch four-step path
    bool real_widen = true;                     // real input, whole phases of a fused kernel: widen to complex in its own sweep, then the plain transform
    bool bluestein_fuse_mul = true;             // ... with the product with FFT(b) on the stores of the forward transform (fused family sizes)
    bool bluestein_stream = true;               // padded length >= 2^14, at least 2^21 padded points per call: plain transforms + streaming kernels
    size_t bluestein_chunk_bytes = 1ull << 30;  // padded sequences of one Bluestein chunk (between its two transforms)
    size_t l2_block_budget = 24ull << 20;       // inter-pass block of one chunk: small enough to stay in L2 between the passes
    bool fourstep_pipeline = false;             // sharded four-step: exchange of a column block behind the lines of the next (measured: no gain)
    int fourstep_pipeline_mb = 256;             // slab bytes per pipeline block
    int fourstep_lines_sms = 0;                 // pipelined sharded four-step: CTAs of the fused line kernel while an exchange runs beside it (0 = all SMs)
    int fourstep_exchange_ctas = 0;             // cap on the CTAs of a pipelined exchange launch (0 = one per tile)
    bool pwelch_bulk = true;                    // L = 4096 float64: bulk-copy fed kernel (pwelch.cu)
    int chunk_streams = 2;                      // streams the chunks of one call rotate over (1 .. 1 + AUX_STREAMS)
    bool l2_block_window = true;                // persisting L2 window over the inter-pass blocks of a chunked call
    int l2_hold = 0;                            // depth of chunked calls in flight: keep the set-aside between their launches (L2Hold)
    bool wide_tiles = false;             // 512-thread tiles for L >= 1024
    int w32 = 2;                         // 1024-point lean passes: 32 points per thread (fft_w32.cuh); 0 = 16-point kernel
    bool debug_alias = false;            // timing experiment only: all transforms of a batch alias one buffer (results are garbage)
    bool tiled_scratch = false;          // tile-major four-step intermediate (measured slower: pass 2 loses its contiguous row reads)
    size_t l2_persist_max = 0;           // cudaDevAttrMaxPersistingL2CacheSize
    size_t l2_window_max = 0;            // cudaDevAttrMaxAccessPolicyWindowSize
    bool use_l2_window = true;
    size_t l2_carved = 0;                // current cudaLimitPersistingL2CacheSize on this device
    bool l2_dirty = false;               // persisting lines / set-aside left behind by a fused launch
    int fused_delay = 2;                 // phases between P1(g) and P2(g) in the fused schedule
    bool bluestein_fused = true;         // padded length <= 4096: one kernel per Bluestein transform (bluestein_small.cuh)
    bool use_fused = false;              // one persistent kernel for both four-step passes (N = L*L), L2-resident scratch
    size_t fused_slot_budget = 16ull << 20;   // bytes of L2-resident intermediate per scratch slot (fused_delay + 2 slots)
    bool use_tma = true;                 // N = 2^20 lean transforms: TMA-fed fused four-step, intermediate resident in L2 (fft_tma.cuh)
    bool use_tma14 = true;               // 2^14-point lines (batched rows, columns of a 2^14-row matrix): fused kernel of fft_tma14.cuh
    int huge_min_log2n = 22;             // plain forward / inverse transforms of at least 2^this points (one less in a batch) take the outer
                                         // four-step over the fused family (fft_pow2_huge): two sweeps at 0.30-0.37 of HBM, where the 4096-point
                                         // strided line passes of the two-launch schedule fall to 6-23 GS/s
    int axis_single_max_log2 = 11;       // strided lines of up to 2^this points (any length on matrices narrower than 64 columns) are one pass
    int huge_l1 = 0;                     // measurement: log2 of the column length of that outer four-step (0 = the rule in fft_pow2_huge)
    int huge_sweeps = 0;                 // measurement / cross-check: 3 or 4 forces the three- / four-sweep formulation
    int tma_grid_cap = 0;                // fused size family: at most this many CTAs (0 = one per SM); leaves SMs to a concurrent kernel
    bool use_tma19 = true;               // 2^19-point transforms: the same kernel with 1024-point pass-1 sub-lines (rows only)
    bool use_tma16 = true;               // 2^16-point lines: the same kernel with 256-point sub-lines
    int tma_opt = 0;                     // measurement switches of the fused kernel (TmaFusedParams::opt)
    int tma_prof = 0;                    // measurement: cycle counters of the fused kernel (gd_tma_profile_read)
    int tma_delay = 2;                   // P1 phases the schedule runs ahead of P2 (fft_tma.cuh)
    int tma_slots = 3;                   // scratch slots of 16 MiB (one transform each) kept in L2; delay <= slots - 1
    // The scratch blocks, dependency counters and the persisting L2 window are per device, not per stream: a call on
    // another stream first waits (on the device) for the previous user, see ScratchOrder.
    cudaEvent_t scratch_evt = nullptr;
    cudaStream_t scratch_last = nullptr;
    bool scratch_used = false;
    std::recursive_mutex mu;             // every public entry point locks its device

    Status init(int device, int lane_index = 0);
    void destroy();
    Status ensure_scratch(ScratchSlot s, size_t bytes, void** out);
    // give the whole L2 back to kernels that do not use the persisting window (called lazily by their launchers)
    Status l2_release();
    Status twiddles(int log2m, TwiddleTable* out);
    Status bluestein(long long n, cudaStream_t st, const BluesteinPlan** out);
};

// Stream-orders the users of the device's shared scratch state: construct after locking the device, before any work is
// enqueued on `st`; the destructor records the hand-over point.
struct ScratchOrder {
    Device& d;
    cudaStream_t st;
    ScratchOrder(Device& dev, cudaStream_t s) : d(dev), st(s) {
        if (d.scratch_used && d.scratch_last != st && d.scratch_evt) cudaStreamWaitEvent(st, d.scratch_evt, 0);
    }
    ~ScratchOrder() {
        if (!d.scratch_evt) cudaEventCreateWithFlags(&d.scratch_evt, cudaEventDisableTiming);
        if (d.scratch_evt) { cudaEventRecord(d.scratch_evt, st); d.scratch_last = st; d.scratch_used = true; }
    }
};

// ---- transforms (device pointers) ----------------------------------------------------------
// batched power-of-two transform, forward butterflies only; direction is expressed through ops
Status fft_pow2(Device& d, const void* in, long long in_dist, cpx* out, long long out_dist, int log2n,
                long long batch, const FusedOps& ops, cudaStream_t st);
// fft.FFT / fft.IFFT semantics for any n >= 1 (fft/fft.go:35-52,72-87); real_in: input is float64
Status fft1d(Device& d, const void* in, long long in_dist, cpx* out, long long out_dist, long long n,
             long long batch, bool real_in, int dir, cudaStream_t st);
// fft.Convolve (fft/fft.go:55-69)
Status convolve(Device& d, const cpx* x, const cpx* y, cpx* out, long long n, cudaStream_t st);
// fft.FFTN / IFFTN on a contiguous row-major array (fft/fft.go:166-224, dsputils/matrix.go:37-57);
// nd == 2 is fft.FFT2 / IFFT2 (fft/fft.go:123-154). in may equal out.
Status fftn(Device& d, const cpx* in, cpx* out, const long long* dims, int nd, int dir, cudaStream_t st);

// ---- Welch PSD -----------------------------------------------------------------------------
// raw[k] (+)= sum over local segments of |FFT(w * seg)[k]|^2 folded to k < lp; segments seg0..seg0+nseg-1
// fmt: 0 float64, 1 float32, 2 int16, 3 uint8 -- the PCM formats are decoded as wav.ReadFloats does (wav/wav.go:138-161)
Status pwelch_partial(Device& d, const void* x, int fmt, long long nfft, long long stride, long long fftlen,
                      long long lp, long long seg0, long long nseg, const double* win, double* raw,
                      cudaStream_t st);
// STFT / spectrogram: out[c][j] = FFT(win * segment c, zero-padded to fftlen)[j], j < lp  (spectral/pwelch.go:104-113 without the sum)
Status stft(Device& d, const double* x, long long nfft, long long stride, long long fftlen, long long lp, long long seg0,
            long long nseg, const double* win, cpx* out, cudaStream_t st);
// linear (non-circular) convolution by overlap-save on top of the circular Convolve (fft/fft.go:55-69): out has nx + nh - 1 elements
Status convolve_linear(Device& d, const cpx* x, long long nx, const cpx* h, long long nh, cpx* out, cudaStream_t st);
// pxx[j] = raw[j] / nsegs (x2 for 0<j<lp-1) / norm      (spectral/pwelch.go:113-121,134-136)
Status pwelch_finalize(const double* raw, long long lp, long long nsegs, double norm, double* pxx, cudaStream_t st);

// ---- distributed four-step (C5) building blocks -------------------------------------------
// optional operators of a fused-family launch (tma14_host.cuh)
struct Tma2dExtra {
    int tw2_log2m = 0;               // COLS: outputs leave multiplied by w_M^((tw2_col0 + column) k), M = 2^tw2_log2m
    long long tw2_col0 = 0;
    long long nmat = 1, in_mdist = 0, out_mdist = 0;   // COLS: the same columns of nmat matrices in the same launches
    int npeer = 0;                   // COLS + tw2: rows of the output spread over npeer ranks' buffers (peer[h] + peer_off, row pitch out_dist)
    cpx* const* peer = nullptr;
    long long peer_off = 0;
    int rank = 0;
    const cpx* aux = nullptr;        // ROWS, forward: output k of every transform leaves multiplied by aux[k]
    int seg = 0;                     // ROWS: a transform is seg segments, seg_dist elements apart
    long long seg_dist = 0;
};
bool fourstep_fused_supported(Device& d, long long n1, long long n2, int world);
Status fourstep_lines_peer(Device& d, const cpx* slab, cpx* const* peer_recv, long long n1, long long w, int rank, int world, int log2n, int dir, cudaStream_t st);
Status fourstep_rows_seg(Device& d, const cpx* recv, cpx* out, long long n2, long long k, int world, int dir, cudaStream_t st);
Status fourstep_twiddle(cpx* blk, long long rows, long long cols, long long row0, long long col0, int log2n, cudaStream_t st, int dir = 1);
Status repack_gkw(const cpx* in, cpx* out, long long G, long long K, long long W, cudaStream_t st);
Status fourstep_exchange(const cpx* slab, cpx* const* peer_recv, long long n1, long long w, int rank, int world, int log2n,
                         cudaStream_t st, long long cbeg = 0, long long ccount = -1, int max_ctas = 0, int dir = +1);
Status fourstep_lines_exchange(Device& d, const cpx* slab, cpx* tmp, cpx* const* peer_recv, long long n1, long long w, int rank, int world,
                               int log2n, cudaStream_t st);
Status peer_block_copy(const cpx* src, cpx* const* peers, int world, int rank, long long rows, long long cols, long long src_step,
                       long long src_pitch, long long dst_off, long long dst_pitch, cudaStream_t st);
Status transpose_batched(const cpx* in, cpx* out, long long batch, long long rows, long long cols, cudaStream_t st);
// strided lines: `outer` blocks of `len` lines-elements with element stride s (fft_axis)
Status fft_strided(Device& d, const cpx* src, cpx* dst, long long outer, long long len, long long s, int dir, cudaStream_t st);

// ---- utilities -----------------------------------------------------------------------------
Status fill_splitmix(double* out, long long n, unsigned long long seed, unsigned long long offset, cudaStream_t st);

}  // namespace gd

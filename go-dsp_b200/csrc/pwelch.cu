// pwelch.cu -- Welch PSD hot loop (spectral/pwelch.go:104-122) as one fused kernel:
// gather overlapping segments straight from the signal (no spectral.Segment copies,
// spectral/spectral.go:35-44), apply the window (window/window.go:25-29, evaluated once
// instead of once per segment), transform, and accumulate |X|^2 per bin in registers.
//
// Two real segments ride in one complex transform: z = w*(a + i*b), and because only
// the SUM of periodograms is needed, |A[k]|^2 + |B[k]|^2 = (|Z[k]|^2 + |Z[L-k]|^2)/2 --
// no per-segment separation. Each thread keeps 16 running sums (bins p + P*i) for its
// whole share of segments; a deterministic fold adds the per-group partials.
#include <stdint.h>

#include "engine.h"
#include "fft_pass.cuh"

namespace gd {

// Sample formats of the signal: float64, or the on-disk formats of wav/wav.go decoded exactly as
// Wav.ReadFloats does (wav/wav.go:138-161, float32 arithmetic) and then widened, as a caller's float64(f) would:
//   uint8:  float32(v) / 255            int16:  (float32(v) + 32768) / 65535            float32: as is
enum : int { SMP_F64 = 0, SMP_F32 = 1, SMP_S16 = 2, SMP_U8 = 3 };
__host__ __device__ __forceinline__ int sample_bytes(int fmt) { return fmt == SMP_F64 ? 8 : fmt == SMP_F32 ? 4 : fmt == SMP_S16 ? 2 : 1; }
template <int FMT>
__device__ __forceinline__ double load_sample(const void* __restrict__ x, long long i) {
    if constexpr (FMT == SMP_F64) return __ldg(reinterpret_cast<const double*>(x) + i);
    else if constexpr (FMT == SMP_F32) return (double)__ldg(reinterpret_cast<const float*>(x) + i);
    else if constexpr (FMT == SMP_S16) return (double)__fdiv_rn((float)__ldg(reinterpret_cast<const short*>(x) + i) + 32768.0f, 65535.0f);
    else return (double)__fdiv_rn((float)__ldg(reinterpret_cast<const unsigned char*>(x) + i), 255.0f);
}
__device__ __forceinline__ double load_sample_rt(const void* __restrict__ x, long long i, int fmt) {
    switch (fmt) {
        case SMP_F32: return load_sample<SMP_F32>(x, i);
        case SMP_S16: return load_sample<SMP_S16>(x, i);
        case SMP_U8: return load_sample<SMP_U8>(x, i);
    }
    return load_sample<SMP_F64>(x, i);
}

template <int LOG2L>
struct PwShape {
    static constexpr int L = 1 << LOG2L;
    static constexpr int P = L / 16;
    static constexpr int T = P >= 256 ? 1 : 256 / P;
    static constexpr int LS = line_stride(L, T);
};

// PREFETCH: the samples of the next segment pair (one contiguous range of stride + nfft doubles: the two
// segments overlap) are copied with cp.async into the group's idle exchange buffer while the last
// butterfly step and the |Z|^2 accumulation of the current pair run; needs 16-byte aligned ranges.
// FMT: sample format; the decode is part of the segment load (PREFETCH stages raw float64 only).
template <int LOG2L, bool PREFETCH, int FMT>
__global__ void __launch_bounds__(PwShape<LOG2L>::T * PwShape<LOG2L>::P, 2)
pwelch_fused_kernel(const void* __restrict__ xv, long long nfft, long long stride, long long seg0, long long nseg,
                    const double* __restrict__ win, double* __restrict__ partial, const cpx* __restrict__ wl) {
    static_assert(!PREFETCH || FMT == SMP_F64, "the staged path copies float64 ranges");
    const double* __restrict__ x = reinterpret_cast<const double*>(xv);
    using SH = PwShape<LOG2L>;
    constexpr int L = SH::L, P = SH::P, T = SH::T, LS = SH::LS;
    constexpr int NSTEP = PassShape<LOG2L>::NSTEP, LASTR = PassShape<LOG2L>::LASTR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* sm = reinterpret_cast<cpx*>(smem_raw);

    const int tid = threadIdx.x, p = tid % P, ell = tid / P;
    cpx* sl = sm + ell * LS;
    const double* stage = reinterpret_cast<const double*>(sl);
    const long long npairs = (nseg + 1) / 2;
    const long long GG = (long long)gridDim.x * T;
    const long long gg = (long long)blockIdx.x * T + ell;
    const long long u0 = gg * npairs / GG, u1 = (gg + 1) * npairs / GG;
    // uniform trip count for the whole CTA (barriers inside the loop)
    const long long g_first = (long long)blockIdx.x * T, g_last = g_first + T - 1;
    long long iters = 0;
    for (long long g = g_first; g <= g_last; g++) {
        long long c = (g + 1) * npairs / GG - g * npairs / GG;
        iters = c > iters ? c : iters;
    }

    auto prefetch = [&](long long u) {             // this group's threads copy [xa, xa + count) into sl
        if (u < u1) {
            const double* xa = x + (seg0 + 2 * u) * stride;
            const long long count = (2 * u + 1 < nseg) ? stride + nfft : nfft;
            const long long n16 = count >> 1;
            for (long long c = p; c < n16; c += P) cp_async16(sl + c, xa + 2 * c, 16);
            if ((count & 1) && p == 0) {
                unsigned s = (unsigned)__cvta_generic_to_shared(reinterpret_cast<double*>(sl) + count - 1);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(xa + count - 1) : "memory");
            }
        }
        cp_async_commit();
    };

    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = 0.0;
    if constexpr (PREFETCH) prefetch(u0);

    for (long long it = 0; it < iters; it++) {
        const long long u = u0 + it;
        const bool act = u < u1;
        const bool has_b = act && (2 * u + 1 < nseg);
        cpx z[16];
        if constexpr (PREFETCH) {
            cp_async_wait_all();
            __syncthreads();                      // the pair's samples (copied by all threads of the group) are visible
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int n = p + P * i;
                double a = 0.0, b = 0.0, w = 0.0;
                if (act && n < nfft) {
                    w = __ldg(win + n);
                    a = stage[n];
                    if (has_b) b = stage[stride + n];
                }
                z[i] = make_double2(w * a, w * b);
            }
        } else {
            const long long ia = (seg0 + 2 * u) * stride, ib = ia + stride;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int n = p + P * i;
                double a = 0.0, b = 0.0, w = 0.0;
                if (act && n < nfft) {
                    w = __ldg(win + n);
                    a = load_sample<FMT>(xv, ia + n);
                    if (has_b) b = load_sample<FMT>(xv, ib + n);
                }
                z[i] = make_double2(w * a, w * b);
            }
        }
        butterfly_step<L, 16, 1>(z, p, wl);
        if constexpr (PREFETCH) __syncthreads();  // all staged samples consumed before the scatter overwrites them
        scatter_step<L, 16, 1>(z, p, sl);
        __syncthreads();
        gather_step<L>(z, p, sl);
        if constexpr (NSTEP == 3) {
            __syncthreads();
            butterfly_step<L, 16, 16>(z, p, wl);
            scatter_step<L, 16, 16>(z, p, sl);
            __syncthreads();
            gather_step<L>(z, p, sl);
        }
        __syncthreads();                          // exchange buffer idle again
        if constexpr (PREFETCH) prefetch(u + 1);
        butterfly_step<L, LASTR, (NSTEP == 2 ? 16 : 256)>(z, p, wl);
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fma(z[i].x, z[i].x, fma(z[i].y, z[i].y, acc[i]));
    }
#pragma unroll
    for (int i = 0; i < 16; i++) partial[gg * L + p + P * i] = acc[i];
}

// ---- bulk-fed kernel for L = 4096 (config C4) --------------------------------------------------------------------------
// Same arithmetic as pwelch_fused_kernel<12>, different plumbing. The load/store unit is this kernel's co-limiting
// pipe (ncu: 59 % of its wavefront rate against 52 % of the FP64 pipe), so the sample ranges no longer pass through
// it on the way in: one thread issues ONE bulk copy per segment pair (cp.async.bulk, the TMA engine, completion on an
// mbarrier) into a staging buffer of its own, as soon as the previous pair's samples have been taken -- a whole
// iteration ahead, instead of 24 LDGSTS per thread issued late in the iteration. The exchange buffer is unpadded
// (element i lives at i ^ ((i >> 4) & 7): radix-16 scatters and gathers stay conflict-free), which is what makes
// 64 KiB + 48 KiB fit twice per SM; four CTA-wide barriers per pair instead of six; with 50 % overlap the second
// segment's first half IS the first segment's second half, so 8 of the 32 staged loads per thread are not repeated.
constexpr int PWB_L = 4096, PWB_P = 256;
constexpr int PWB_STAGE_DOUBLES = 2 * PWB_L;                                    // stride + nfft <= 2 L samples per pair
constexpr int PWB_SMEM = PWB_L * 16 + PWB_STAGE_DOUBLES * 8 + 64;               // exchange + staging + mbarrier

__device__ __forceinline__ unsigned pw_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(PWB_P, 2)
pwelch_bulk_kernel(const double* __restrict__ x, long long nfft, long long stride, long long seg0, long long nseg,
                   const double* __restrict__ win, double* __restrict__ partial, const cpx* __restrict__ wl) {
    constexpr int L = PWB_L, P = PWB_P;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cpx* sl = reinterpret_cast<cpx*>(smem_raw);
    double* stage = reinterpret_cast<double*>(smem_raw + L * 16);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + L * 16 + PWB_STAGE_DOUBLES * 8);
    const int p = threadIdx.x;
    const long long npairs = (nseg + 1) / 2;
    const long long u0 = (long long)blockIdx.x * npairs / gridDim.x, u1 = ((long long)blockIdx.x + 1) * npairs / gridDim.x;
    const bool half_overlap = (2 * stride == nfft) && nfft == L;

    if (p == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(pw_smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    auto fetch = [&](long long u) {               // thread 0: one bulk copy of the pair's sample range
        const long long count = (2 * u + 1 < nseg) ? stride + nfft : nfft;
        const unsigned bytes = (unsigned)(count * 8);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(pw_smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     ::"r"(pw_smem_u32(stage)), "l"(x + (seg0 + 2 * u) * stride), "r"(bytes), "r"(pw_smem_u32(bar)) : "memory");
    };
    if (p == 0 && u0 < u1) fetch(u0);

    // exchange addressing: element i at i ^ ((i >> 4) & 7)
    const int m1 = p & 7;                         // scatter 1: i = 16 p + r            -> 16 p + (r ^ m1)
    cpx* sc1 = sl + 16 * p;
    const int k2 = p & 15;                        // scatter 2: i = 256 (p >> 4) + k2 + 16 r -> low bits k2 ^ (r & 7)
    cpx* sc2 = sl + 16 * (p - k2);
    const cpx* ga = sl + (p ^ ((p >> 4) & 7));    // gathers: i = p + 256 q                -> (p ^ c) + 256 q

    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = 0.0;
    unsigned phase = 0;
    for (long long u = u0; u < u1; u++) {
        const bool has_b = 2 * u + 1 < nseg;
        {   // the pair's samples have landed
            asm volatile(
                "{\n.reg .pred q;\nPW_WAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n@q bra PW_DONE_%=;\nbra PW_WAIT_%=;\nPW_DONE_%=:\n}\n"
                ::"r"(pw_smem_u32(bar)), "r"(phase) : "memory");
            phase ^= 1u;
        }
        cpx z[16];
        if (half_overlap && has_b) {
            double a[24];
#pragma unroll
            for (int i = 0; i < 24; i++) a[i] = stage[p + P * i];          // samples p + 256 i of the 6144-sample range
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const double w = __ldg(win + p + P * i);
                z[i] = make_double2(w * a[i], w * a[i + 8]);              // b[n] = range[n + 2048]
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int n = p + P * i;
                double a = 0.0, b = 0.0, w = 0.0;
                if (n < nfft) {
                    w = __ldg(win + n);
                    a = stage[n];
                    if (has_b) b = stage[stride + n];
                }
                z[i] = make_double2(w * a, w * b);
            }
        }
        __syncthreads();                          // staging consumed; the previous pair's gathers are done as well
        if (p == 0 && u + 1 < u1) fetch(u + 1);
        butterfly_step<L, 16, 1>(z, p, wl);
#pragma unroll
        for (int r = 0; r < 16; r++) sc1[r ^ m1] = z[r];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; q++) z[q] = ga[P * q];
        butterfly_step<L, 16, 16>(z, p, wl);
        __syncthreads();                          // every gather of the first exchange is done
#pragma unroll
        for (int r = 0; r < 16; r++) sc2[16 * r + (k2 ^ (r & 7))] = z[r];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; q++) z[q] = ga[P * q];
        butterfly_step<L, 16, 256>(z, p, wl);
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fma(z[i].x, z[i].x, fma(z[i].y, z[i].y, acc[i]));
    }
#pragma unroll
    for (int i = 0; i < 16; i++) partial[(long long)blockIdx.x * L + p + P * i] = acc[i];
}

// raw[j] = 0.5 * (S[j] + S[(L-j) mod L]),  S[k] = sum over groups (fixed order) of partial[g][k]
__global__ void pwelch_fold_kernel(const double* __restrict__ partial, long long groups, int L, long long lp,
                                   double* __restrict__ raw) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= lp) return;
    long long m = (L - j) % L;
    double s0 = 0.0, s1 = 0.0;
    for (long long g = 0; g < groups; g++) { s0 += partial[g * L + j]; s1 += partial[g * L + m]; }
    raw[j] = 0.5 * (s0 + s1);
}

// power-of-two fftlen above 4096: two real segments ride in one complex transform here as well (see the top of the file).
// z[u][n] = win[n] * (a[n] + i b[n]), a = segment seg0 + 2 (u0 + u), b = the next one (0 past the last segment), zero-padded
__global__ void pwelch_pack_pairs_kernel(const void* __restrict__ x, int fmt, long long nfft, long long stride, long long seg0,
                                         long long nseg, long long u0, long long npairs, long long fftlen,
                                         const double* __restrict__ win, cpx* __restrict__ buf) {
    const long long tot = npairs * fftlen, step = (long long)gridDim.x * blockDim.x;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < tot; t += step) {
        const long long u = t / fftlen, n = t - u * fftlen, sa = 2 * (u0 + u);
        double a = 0.0, b = 0.0;
        if (n < nfft) {
            const double w = __ldg(win + n);
            a = w * load_sample_rt(x, (seg0 + sa) * stride + n, fmt);
            if (sa + 1 < nseg) b = w * load_sample_rt(x, (seg0 + sa + 1) * stride + n, fmt);
        }
        buf[t] = make_double2(a, b);
    }
}
// partial[g][k] (+)= sum over the pairs u = g, g + G, ... of this chunk (ascending) of |Z_u[k]|^2: a fixed summation order
__global__ void pwelch_accum_pairs_kernel(const cpx* __restrict__ buf, long long npairs, long long fftlen, int first,
                                          double* __restrict__ partial) {
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= fftlen) return;
    const long long g = blockIdx.y, G = gridDim.y;
    double s = first ? 0.0 : partial[g * fftlen + k];
    for (long long u = g; u < npairs; u += G) {
        const cpx v = buf[u * fftlen + k];
        s = fma(v.x, v.x, fma(v.y, v.y, s));
    }
    partial[g * fftlen + k] = s;
}

// general path (any fftlen): dense windowed, zero-padded complex segments
__global__ void pwelch_gather_kernel(const void* __restrict__ x, int fmt, long long nfft, long long stride, long long seg0,
                                     long long nseg, long long fftlen, const double* __restrict__ win, cpx* __restrict__ buf) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nseg * fftlen) return;
    long long c = t / fftlen, n = t - c * fftlen;
    double v = 0.0;
    if (n < nfft) v = load_sample_rt(x, (seg0 + c) * stride + n, fmt) * win[n];
    buf[t] = make_double2(v, 0.0);
}
// window as a complex array (w[n], 0): the aux operand of the pass kernel's fused load multiply (STFT)
__global__ void window_to_complex_kernel(const double* __restrict__ win, long long n, cpx* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_double2(win[i], 0.0);
}
// dense spectra [nseg][fftlen] -> the first lp bins of each, [nseg][lp]
__global__ void take_bins_kernel(const cpx* __restrict__ in, long long nseg, long long fftlen, long long lp, cpx* __restrict__ out) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nseg * lp) return;
    long long c = t / lp, j = t - c * lp;
    out[t] = in[c * fftlen + j];
}
__global__ void pwelch_accum_kernel(const cpx* __restrict__ buf, long long nseg, long long fftlen, long long lp,
                                    double* __restrict__ raw) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= lp) return;
    double s = raw[j];
    for (long long c = 0; c < nseg; c++) {
        cpx v = buf[c * fftlen + j];
        s += v.x * v.x + v.y * v.y;
    }
    raw[j] = s;
}
__global__ void pwelch_finalize_kernel(const double* __restrict__ raw, long long lp, double nsegs, double norm,
                                       double* __restrict__ pxx) {
    long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= lp) return;
    double d = raw[j] / nsegs;                 // spectral/pwelch.go:114
    if (j > 0 && j < lp - 1) d *= 2.0;         // :116-118
    pxx[j] = d / norm;                         // :134-136
}

typedef void (*PwKernel)(const void*, long long, long long, long long, long long, const double*, double*, const cpx*);
template <int LOG2L>
static Status launch_fused(Device& d, const void* x, int fmt, long long nfft, long long stride, long long lp, long long seg0,
                           long long nseg, const double* win, double* raw, cudaStream_t st) {
    using SH = PwShape<LOG2L>;
    // the staged path needs float64 samples and every pair's sample range 16-byte aligned
    // (pair u starts 2*u*stride doubles = 16*u*stride bytes after the first one)
    // and stride <= nfft (a pair's range of stride + nfft samples must fit the 2 L-sample buffer; noverlap < 0 leaves gaps)
    const bool aligned = fmt == SMP_F64 && stride <= nfft && (reinterpret_cast<uintptr_t>((const double*)x + seg0 * stride) & 15) == 0;
    PwKernel kern = aligned ? (PwKernel)pwelch_fused_kernel<LOG2L, true, SMP_F64>
                  : fmt == SMP_F64 ? (PwKernel)pwelch_fused_kernel<LOG2L, false, SMP_F64>
                  : fmt == SMP_F32 ? (PwKernel)pwelch_fused_kernel<LOG2L, false, SMP_F32>
                  : fmt == SMP_S16 ? (PwKernel)pwelch_fused_kernel<LOG2L, false, SMP_S16>
                                   : (PwKernel)pwelch_fused_kernel<LOG2L, false, SMP_U8>;
    if constexpr (LOG2L == 12) {
        // config C4's shape: float64, nfft = 4096, aligned ranges of an even number of samples -> the bulk-fed kernel
        if (aligned && d.pwelch_bulk && nfft == 4096 && (stride % 2) == 0) {
            static int bulk_bps[16] = {0};
            int& b = bulk_bps[d.dev & 15];
            if (!b) {
                GD_CUDA(cudaFuncSetAttribute(pwelch_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PWB_SMEM));
                GD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, pwelch_bulk_kernel, PWB_P, PWB_SMEM));
                if (b < 1) { set_error("pwelch_bulk_kernel does not fit on an SM"); return GD_ERR_CUDA; }
            }
            const long long npairs = (nseg + 1) / 2;
            long long grid = (long long)d.num_sms * b;
            if (grid > npairs) grid = npairs;
            double* partial;
            GD_TRY(d.ensure_scratch(SCR_PWELCH, (size_t)grid * PWB_L * sizeof(double), (void**)&partial));
            pwelch_bulk_kernel<<<(unsigned)grid, PWB_P, PWB_SMEM, st>>>((const double*)x, nfft, stride, seg0, nseg, win, partial, d.wl[12]);
            g_launches++;
            GD_CUDA(cudaGetLastError());
            pwelch_fold_kernel<<<(unsigned)((lp + 127) / 128), 128, 0, st>>>(partial, grid, PWB_L, lp, raw);
            g_launches++;
            GD_CUDA(cudaGetLastError());
            return GD_OK;
        }
    }
    const int threads = SH::T * SH::P, smem = SH::T * SH::LS * (int)sizeof(cpx);
    static int bps[5][16] = {{0}};                       // per (variant, device): the opt-in below is a per-device attribute
    int& blocks_per_sm = bps[aligned ? 4 : fmt][d.dev & 15];
    if (!blocks_per_sm) {
        GD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int b = 0;
        GD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, threads, smem));
        if (b < 1) { set_error("pwelch_fused_kernel does not fit on an SM"); return GD_ERR_CUDA; }
        blocks_per_sm = b;
    }
    const long long npairs = (nseg + 1) / 2;
    long long grid = (long long)d.num_sms * blocks_per_sm;
    long long need = (npairs + SH::T - 1) / SH::T;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    const long long groups = grid * SH::T;
    double* partial;
    GD_TRY(d.ensure_scratch(SCR_PWELCH, (size_t)groups * SH::L * sizeof(double), (void**)&partial));
    kern<<<(unsigned)grid, threads, smem, st>>>(x, nfft, stride, seg0, nseg, win, partial, d.wl[LOG2L]);
    g_launches++;
    GD_CUDA(cudaGetLastError());
    pwelch_fold_kernel<<<(unsigned)((lp + 127) / 128), 128, 0, st>>>(partial, groups, SH::L, lp, raw);
    g_launches++;
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

Status pwelch_partial(Device& d, const void* x, int fmt, long long nfft, long long stride, long long fftlen, long long lp,
                      long long seg0, long long nseg, const double* win, double* raw, cudaStream_t st) {
    if (nfft < 1 || stride < 1 || fftlen < nfft || lp < 1 || lp > fftlen / 2 + 1 || nseg < 0 || fmt < SMP_F64 || fmt > SMP_U8) {
        set_error("pwelch: bad arguments");
        return GD_ERR_INVALID;
    }
    GD_TRY(d.l2_release());
    if (nseg == 0) {
        GD_CUDA(cudaMemsetAsync(raw, 0, (size_t)lp * sizeof(double), st));
        return GD_OK;
    }
    const bool p2 = (fftlen & (fftlen - 1)) == 0;
    if (p2 && fftlen >= 32 && fftlen <= 4096) {
        switch (fftlen) {
            case 32: return launch_fused<5>(d, x, fmt, nfft, stride, lp, seg0, nseg, win, raw, st);
            case 64: return launch_fused<6>(d, x, fmt, nfft, stride, lp, seg0, nseg, win, raw, st);
            case 128: return launch_fused<7>(d, x, fmt, nfft, stride, lp, seg0, nseg, win, raw, st);
            case 256: return launch_fused<8>(d, x, fmt, nfft, stride, lp, seg0, nseg, win, raw, st);
            case 512: return launch_fused<9>(d, x, fmt, nfft, stride, lp, seg0, nseg, win, raw, st);
            case 1024: return launch_fused<10>(d, x, fmt, nfft, stride, lp, seg0, nseg, win, raw, st);
            case 2048: return launch_fused<11>(d, x, fmt, nfft, stride, lp, seg0, nseg, win, raw, st);
            case 4096: return launch_fused<12>(d, x, fmt, nfft, stride, lp, seg0, nseg, win, raw, st);
        }
    }
    if (p2 && fftlen > 4096) {
        // pairs of segments packed into complex transforms -> batched power-of-two transforms (the fused size family up to
        // 2^18 points) -> per-bin sums over groups of pairs in a fixed order -> the fold of the fused path
        const long long npairs = (nseg + 1) / 2;
        long long cp = (long long)((1ull << 30) / ((size_t)fftlen * sizeof(cpx)));          // pairs per chunk: 1 GiB of spectra
        if (cp < 1) cp = 1;
        if (cp > npairs) cp = npairs;
        const long long unit = fftlen <= (1LL << 18) ? (1LL << 20) / fftlen : 1;             // whole phases of the fused kernel
        if (cp > unit) cp -= cp % unit;
        long long G = (1LL << 23) / fftlen;                                                 // groups: partial sums of 64 MiB at most
        if (G > 128) G = 128;
        if (G > cp) G = cp;
        if (G < 1) G = 1;
        cpx* buf;
        double* partial;
        GD_TRY(d.ensure_scratch(SCR_PWELCH, (size_t)cp * fftlen * sizeof(cpx), (void**)&buf));
        GD_TRY(d.ensure_scratch(SCR_AUX, (size_t)G * fftlen * sizeof(double), (void**)&partial));
        for (long long u0 = 0; u0 < npairs; u0 += cp) {
            const long long np_ = npairs - u0 < cp ? npairs - u0 : cp;
            long long pg = (np_ * fftlen + 255) / 256;
            if (pg > (long long)d.num_sms * 32) pg = (long long)d.num_sms * 32;
            pwelch_pack_pairs_kernel<<<(unsigned)pg, 256, 0, st>>>(x, fmt, nfft, stride, seg0, nseg, u0, np_, fftlen, win, buf);
            GD_CUDA(cudaGetLastError());
            GD_TRY(fft1d(d, buf, fftlen, buf, fftlen, fftlen, np_, false, +1, st));
            pwelch_accum_pairs_kernel<<<dim3((unsigned)((fftlen + 255) / 256), (unsigned)G, 1), 256, 0, st>>>(buf, np_, fftlen, u0 == 0 ? 1 : 0, partial);
            GD_CUDA(cudaGetLastError());
            g_launches += 2;
        }
        pwelch_fold_kernel<<<(unsigned)((lp + 127) / 128), 128, 0, st>>>(partial, G, (int)fftlen, lp, raw);
        g_launches++;
        GD_CUDA(cudaGetLastError());
        return GD_OK;
    }
    // general path: dense windowed segments -> batched transform (any length) -> per-bin sums
    long long chunk = (long long)((64ull << 20) / ((size_t)fftlen * sizeof(cpx)));
    if (chunk < 1) chunk = 1;
    if (chunk > nseg) chunk = nseg;
    cpx* buf;
    GD_TRY(d.ensure_scratch(SCR_PWELCH, (size_t)chunk * fftlen * sizeof(cpx), (void**)&buf));
    GD_CUDA(cudaMemsetAsync(raw, 0, (size_t)lp * sizeof(double), st));
    for (long long c0 = 0; c0 < nseg; c0 += chunk) {
        long long nc = nseg - c0 < chunk ? nseg - c0 : chunk;
        long long tot = nc * fftlen;
        pwelch_gather_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(x, fmt, nfft, stride, seg0 + c0, nc, fftlen, win, buf);
        g_launches++;
        GD_CUDA(cudaGetLastError());
        GD_TRY(fft1d(d, buf, fftlen, buf, fftlen, fftlen, nc, false, +1, st));
        pwelch_accum_kernel<<<(unsigned)((lp + 127) / 128), 128, 0, st>>>(buf, nc, fftlen, lp, raw);
        g_launches++;
        GD_CUDA(cudaGetLastError());
    }
    return GD_OK;
}

// STFT / spectrogram: the Welch loop without the accumulate (spectral/pwelch.go:104-113): out[c][j] = FFT(win * segment c)[j],
// j < lp, segments seg0 .. seg0 + nseg - 1 of stride `stride`. Power-of-two lengths are ONE fused launch per pass: the pass
// kernel gathers the overlapping segments straight from the signal (rows `stride` samples apart), multiplies by the window,
// zero-pads, transforms and stores only the first lp bins.
Status stft(Device& d, const double* x, long long nfft, long long stride, long long fftlen, long long lp, long long seg0,
            long long nseg, const double* win, cpx* out, cudaStream_t st) {
    if (nfft < 1 || stride < 1 || fftlen < nfft || lp < 1 || lp > fftlen || nseg < 0) { set_error("stft: bad arguments"); return GD_ERR_INVALID; }
    if (nseg == 0) return GD_OK;
    GD_TRY(d.l2_release());
    const bool p2 = (fftlen & (fftlen - 1)) == 0;
    if (p2 && fftlen >= 2 && fftlen <= (1LL << 24)) {
        cpx* wc;
        GD_TRY(d.ensure_scratch(SCR_AUX, (size_t)nfft * sizeof(cpx), (void**)&wc));
        window_to_complex_kernel<<<(unsigned)((nfft + 255) / 256), 256, 0, st>>>(win, nfft, wc);
        g_launches++;
        GD_CUDA(cudaGetLastError());
        FusedOps ops;
        ops.ld_flags = LD_REAL | LD_PAD | LD_MULAUX;
        ops.aux_in = wc; ops.n_valid_in = nfft;
        ops.st_flags = ST_TRUNC; ops.n_valid_out = lp;
        int lg = 0;
        while ((1LL << lg) < fftlen) lg++;
        return fft_pow2(d, x + seg0 * stride, stride, out, lp, lg, nseg, ops, st);
    }
    // any other length: dense windowed segments -> batched transform -> first lp bins
    long long chunk = (long long)((64ull << 20) / ((size_t)fftlen * sizeof(cpx)));
    if (chunk < 1) chunk = 1;
    if (chunk > nseg) chunk = nseg;
    cpx* buf;
    GD_TRY(d.ensure_scratch(SCR_PWELCH, (size_t)chunk * fftlen * sizeof(cpx), (void**)&buf));
    for (long long c0 = 0; c0 < nseg; c0 += chunk) {
        const long long nc = nseg - c0 < chunk ? nseg - c0 : chunk, tot = nc * fftlen;
        pwelch_gather_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(x, SMP_F64, nfft, stride, seg0 + c0, nc, fftlen, win, buf);
        g_launches++;
        GD_CUDA(cudaGetLastError());
        GD_TRY(fft1d(d, buf, fftlen, buf, fftlen, fftlen, nc, false, +1, st));
        take_bins_kernel<<<(unsigned)((nc * lp + 255) / 256), 256, 0, st>>>(buf, nc, fftlen, lp, out + c0 * lp);
        g_launches++;
        GD_CUDA(cudaGetLastError());
    }
    return GD_OK;
}

Status pwelch_finalize(const double* raw, long long lp, long long nsegs, double norm, double* pxx, cudaStream_t st) {
    if (lp < 1 || nsegs < 1) { set_error("pwelch_finalize: bad arguments"); return GD_ERR_INVALID; }
    pwelch_finalize_kernel<<<(unsigned)((lp + 127) / 128), 128, 0, st>>>(raw, lp, (double)nsegs, norm, pxx);
    g_launches++;
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

}  // namespace gd

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
template <int LOG2L, bool PREFETCH>
__global__ void __launch_bounds__(PwShape<LOG2L>::T * PwShape<LOG2L>::P, 2)
pwelch_fused_kernel(const double* __restrict__ x, long long nfft, long long stride, long long seg0, long long nseg,
                    const double* __restrict__ win, double* __restrict__ partial, const cpx* __restrict__ wl) {
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
            const double* xa = x + (seg0 + 2 * u) * stride;
            const double* xb = xa + stride;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int n = p + P * i;
                double a = 0.0, b = 0.0, w = 0.0;
                if (act && n < nfft) {
                    w = __ldg(win + n);
                    a = __ldg(xa + n);
                    if (has_b) b = __ldg(xb + n);
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

// general path (any fftlen): dense windowed, zero-padded complex segments
__global__ void pwelch_gather_kernel(const double* __restrict__ x, long long nfft, long long stride, long long seg0,
                                     long long nseg, long long fftlen, const double* __restrict__ win, cpx* __restrict__ buf) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nseg * fftlen) return;
    long long c = t / fftlen, n = t - c * fftlen;
    double v = 0.0;
    if (n < nfft) v = x[(seg0 + c) * stride + n] * win[n];
    buf[t] = make_double2(v, 0.0);
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

template <int LOG2L>
static Status launch_fused(Device& d, const double* x, long long nfft, long long stride, long long lp, long long seg0,
                           long long nseg, const double* win, double* raw, cudaStream_t st) {
    using SH = PwShape<LOG2L>;
    // the staged path needs every pair's sample range 16-byte aligned
    // (pair u starts 2*u*stride doubles = 16*u*stride bytes after the first one)
    const bool aligned = (reinterpret_cast<uintptr_t>(x + seg0 * stride) & 15) == 0;
    auto kern = aligned ? pwelch_fused_kernel<LOG2L, true> : pwelch_fused_kernel<LOG2L, false>;
    const int threads = SH::T * SH::P, smem = SH::T * SH::LS * (int)sizeof(cpx);
    static int bps[2][16] = {{0}};                       // per (variant, device): the opt-in below is a per-device attribute
    int& blocks_per_sm = bps[aligned ? 1 : 0][d.dev & 15];
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

Status pwelch_partial(Device& d, const double* x, long long nfft, long long stride, long long fftlen, long long lp,
                      long long seg0, long long nseg, const double* win, double* raw, cudaStream_t st) {
    if (nfft < 1 || stride < 1 || fftlen < nfft || lp < 1 || lp > fftlen / 2 + 1 || nseg < 0) {
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
            case 32: return launch_fused<5>(d, x, nfft, stride, lp, seg0, nseg, win, raw, st);
            case 64: return launch_fused<6>(d, x, nfft, stride, lp, seg0, nseg, win, raw, st);
            case 128: return launch_fused<7>(d, x, nfft, stride, lp, seg0, nseg, win, raw, st);
            case 256: return launch_fused<8>(d, x, nfft, stride, lp, seg0, nseg, win, raw, st);
            case 512: return launch_fused<9>(d, x, nfft, stride, lp, seg0, nseg, win, raw, st);
            case 1024: return launch_fused<10>(d, x, nfft, stride, lp, seg0, nseg, win, raw, st);
            case 2048: return launch_fused<11>(d, x, nfft, stride, lp, seg0, nseg, win, raw, st);
            case 4096: return launch_fused<12>(d, x, nfft, stride, lp, seg0, nseg, win, raw, st);
        }
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
        pwelch_gather_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(x, nfft, stride, seg0 + c0, nc, fftlen, win, buf);
        g_launches++;
        GD_CUDA(cudaGetLastError());
        GD_TRY(fft1d(d, buf, fftlen, buf, fftlen, fftlen, nc, false, +1, st));
        pwelch_accum_kernel<<<(unsigned)((lp + 127) / 128), 128, 0, st>>>(buf, nc, fftlen, lp, raw);
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

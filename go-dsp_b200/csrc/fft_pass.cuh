// fft_pass.cuh -- one batched "pass": many independent length-L (L = 2^m <= 4096)
// transforms, T lines per CTA, each line held 16 points per thread in registers.
//
// A pass is the unit every transform of the library is planned from:
//   * N <= 4096:  one pass (row mode in, row mode out);
//   * N  > 4096:  four-step N = N1*N2 -- pass 1 = N2 column transforms of length N1
//                 with the twiddle w_N^(n2*k1) fused into its stores, pass 2 = N1 row
//                 transforms of length N2 with the transposed (stride-N1) store fused
//                 (replaces the log2(N) sweeps of fft/radix2.go:131-151);
//   * FFT2/FFTN: passes over strided lines (fft/fft.go:138-151,175-189 gather/scatter
//                 loops become the pass's address arithmetic);
//   * Bluestein:  chirp multiply / zero-pad fused into loads, pointwise product and
//                 post-multiply / truncation fused into stores (fft/bluestein.go:70-93).
#pragma once
#include "fft_core.cuh"

namespace gd {

enum : int { MODE_ROW = 0, MODE_COL = 1 };
enum : int {
    LD_REAL = 1,     // input is float64, imag = 0                     (dsputils.go:25-31 ToComplex)
    LD_PAD = 2,      // elements with local offset >= n_valid_in read as 0 (dsputils.go:49-58 ZeroPad)
    LD_MULAUX = 4,   // multiply by aux_in[local offset]                (bluestein.go:74-76)
    LD_SWAP = 8,     // swap re/im after the other load ops (inverse transform = swap . forward . swap)
    LD_REVERSE = 16, // read x[(n_valid_in - n) mod n_valid_in] instead of x[n]  (fft.go:39-43 IFFT index reversal)
};
enum : int {
    ST_TWIDDLE = 1,  // multiply output k of the line by w_M^(mult*k)   (four-step twiddle)
    ST_MULAUX = 2,   // multiply by aux_out[local offset]               (fft.go:63-66, bluestein.go:89-91)
    ST_SCALE = 4,    // multiply by scale (exact 1/N for power-of-two N; fft.go:47-50)
    ST_DIV = 8,      // divide by div (non power-of-two N keeps the reference's true division)
    ST_TRUNC = 16,   // skip outputs with local offset >= n_valid_out   (bluestein.go:93)
    ST_SWAP = 32,    // swap re/im before the other store ops
};

struct PassParams {
    const void* in;
    void* out;
    long long nlines;          // number of lines in this launch
    long long inner;           // line id = q*inner + i
    long long in_qs, in_is, in_es;     // element strides: base = q*qs + i*is, element j at + j*es
    long long out_qs, out_is, out_es;
    int in_mode, out_mode;     // MODE_ROW: lanes run along the line; MODE_COL: lanes run across adjacent lines
    int ld_flags, st_flags;
    int tw_sel;                // twiddle multiplier: 0 -> i, 1 -> q
    int tw_log2m;              // M = 2^tw_log2m
    const cpx* tw_lo;          // exp(-2 pi i e / M), e < min(M, 4096)
    const cpx* tw_hi;          // exp(-2 pi i h * 4096 / M), h < M / 4096
    const cpx* wl;             // exp(-2 pi i e / L), e < L (intra-line step twiddles)
    const cpx* aux_in;
    const cpx* aux_out;
    long long n_valid_in, n_valid_out;
    double scale, div;
};

template <int LOG2L>
struct PassShape {
    static constexpr int L = 1 << LOG2L;
    static constexpr int PPT = L < 16 ? L : 16;      // points per thread
    static constexpr int P = L / PPT;                // threads per line
    static constexpr int NSTEP = LOG2L <= 4 ? 1 : (LOG2L + 3) / 4;
    static constexpr int LASTR = LOG2L <= 4 ? L : ((LOG2L % 4) ? (1 << (LOG2L % 4)) : 16);
};

__device__ __forceinline__ cpx tw_lookup(const PassParams& a, unsigned long long e) {
    if (a.tw_log2m <= 12) return __ldg(a.tw_lo + e);
    cpx lo = __ldg(a.tw_lo + (e & 4095ULL)), hi = __ldg(a.tw_hi + (e >> 12));
    return cmul(hi, lo);
}

// One Stockham step of radix R on the 16 registers of this thread: butterfly q takes
// x[q + NB*r], its index in the line is j = p + P*q, k = j mod NS.
template <int L, int R, int NS>
__device__ __forceinline__ void butterfly_step(cpx (&x)[16], int p, const cpx* __restrict__ wl) {
    constexpr int P = L / 16, NB = 16 / R;
#pragma unroll
    for (int q = 0; q < NB; q++) {
        if constexpr (NS > 1) {
            int j = p + P * q;
            int k = j & (NS - 1);
            cpx w = __ldg(wl + k * (L / (NS * R)));
            twiddle_powers<R, NB>(&x[q], w);
        }
        dft<R, NB>(&x[q]);
    }
}

// scatter the step's outputs into the line's shared-memory region (Stockham auto-sort)
template <int L, int R, int NS>
__device__ __forceinline__ void scatter_step(const cpx (&x)[16], int p, cpx* __restrict__ sl) {
    constexpr int P = L / 16, NB = 16 / R;
#pragma unroll
    for (int q = 0; q < NB; q++) {
        int j = p + P * q;
        int k = j & (NS - 1);
        int base = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; r++) sl[pad_idx(base + r * NS)] = x[q + NB * r];
    }
}
template <int L>
__device__ __forceinline__ void gather_step(cpx (&x)[16], int p, const cpx* __restrict__ sl) {
    constexpr int P = L / 16;
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = sl[pad_idx(p + P * i)];
}

template <int LOG2L, int T, bool GENERIC>
__global__ void __launch_bounds__(T * PassShape<LOG2L>::P, (T * PassShape<LOG2L>::P >= 512) ? 1 : 2)
fft_pass_kernel(const PassParams a) {
    using SH = PassShape<LOG2L>;
    constexpr int L = SH::L, PPT = SH::PPT, P = SH::P;
    constexpr int LS = line_stride(L, T);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* sm = reinterpret_cast<cpx*>(smem_raw);

    const int tid = threadIdx.x;
    const long long ntiles = (a.nlines + T - 1) / T;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- thread -> (line, p) for the load side
        int ell, p;
        if (a.in_mode == MODE_COL) { ell = tid % T; p = tid / T; } else { p = tid % P; ell = tid / P; }
        long long line = tile * T + ell;
        bool valid = line < a.nlines;
        long long q = 0, ii = 0;
        if (valid) { q = line / a.inner; ii = line - q * a.inner; }

        cpx x[16];
        // ---- load
        {
            const long long loc0 = ii * a.in_is, base = q * a.in_qs;
#pragma unroll
            for (int i = 0; i < PPT; i++) {
                long long loc = loc0 + (long long)(p + P * i) * a.in_es;
                cpx v = make_double2(0.0, 0.0);
                if constexpr (GENERIC) {
                    bool ok = valid && !((a.ld_flags & LD_PAD) && loc >= a.n_valid_in);
                    if (ok) {
                        long long src = loc;
                        if ((a.ld_flags & LD_REVERSE) && loc != 0) src = a.n_valid_in - loc;
                        if (a.ld_flags & LD_REAL) v.x = __ldg(reinterpret_cast<const double*>(a.in) + base + src);
                        else v = __ldg(reinterpret_cast<const cpx*>(a.in) + base + src);
                        if (a.ld_flags & LD_MULAUX) v = cmul(v, __ldg(a.aux_in + loc));
                    }
                } else {
                    if (valid) v = __ldg(reinterpret_cast<const cpx*>(a.in) + base + loc);
                }
                if (a.ld_flags & LD_SWAP) v = cswap(v);
                x[i] = v;
            }
        }

        // ---- transform
        if constexpr (LOG2L <= 4) {
            dft<L, 1>(x);
        } else {
            cpx* sl = sm + ell * LS;
            butterfly_step<L, 16, 1>(x, p, a.wl);
            scatter_step<L, 16, 1>(x, p, sl);
            if constexpr (SH::NSTEP == 2) {
                __syncthreads();
                if (a.out_mode != a.in_mode) {
                    if (a.out_mode == MODE_COL) { ell = tid % T; p = tid / T; } else { p = tid % P; ell = tid / P; }
                    sl = sm + ell * LS;
                }
                gather_step<L>(x, p, sl);
                butterfly_step<L, SH::LASTR, 16>(x, p, a.wl);
            } else {
                __syncthreads();
                gather_step<L>(x, p, sl);
                __syncthreads();
                butterfly_step<L, 16, 16>(x, p, a.wl);
                scatter_step<L, 16, 16>(x, p, sl);
                __syncthreads();
                if (a.out_mode != a.in_mode) {
                    if (a.out_mode == MODE_COL) { ell = tid % T; p = tid / T; } else { p = tid % P; ell = tid / P; }
                    sl = sm + ell * LS;
                }
                gather_step<L>(x, p, sl);
                butterfly_step<L, SH::LASTR, 256>(x, p, a.wl);
            }
            if (a.out_mode != a.in_mode) {
                line = tile * T + ell;
                valid = line < a.nlines;
                q = 0; ii = 0;
                if (valid) { q = line / a.inner; ii = line - q * a.inner; }
            }
        }

        // ---- fused output twiddle  w_M^(mult * k),  k = p + P*i
        if (a.st_flags & ST_TWIDDLE) {
            unsigned long long mask = (1ULL << a.tw_log2m) - 1ULL;
            unsigned long long mult = (unsigned long long)(a.tw_sel ? q : ii);
            cpx t0 = tw_lookup(a, (mult * (unsigned long long)p) & mask);
            if constexpr (PPT == 1) {
                x[0] = cmul(x[0], t0);
            } else {
                cpx s1 = tw_lookup(a, (mult * (unsigned long long)P) & mask);
                cpx t[PPT];
                t[0] = t0; t[1] = cmul(t0, s1);
                if constexpr (PPT >= 4) {
                    cpx s2 = csqr(s1);
                    t[2] = cmul(t[0], s2); t[3] = cmul(t[1], s2);
                    if constexpr (PPT >= 8) {
                        cpx s4 = csqr(s2);
#pragma unroll
                        for (int i = 0; i < 4; i++) t[4 + i] = cmul(t[i], s4);
                        if constexpr (PPT >= 16) {
                            cpx s8 = csqr(s4);
#pragma unroll
                            for (int i = 0; i < 8; i++) t[8 + i] = cmul(t[i], s8);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < PPT; i++) x[i] = cmul(x[i], t[i]);
            }
        }

        // ---- store
        if (valid) {
            const long long loc0 = ii * a.out_is, base = q * a.out_qs;
#pragma unroll
            for (int i = 0; i < PPT; i++) {
                long long loc = loc0 + (long long)(p + P * i) * a.out_es;
                cpx v = x[i];
                if (a.st_flags & ST_SWAP) v = cswap(v);
                if (a.st_flags & ST_SCALE) { v.x *= a.scale; v.y *= a.scale; }
                if constexpr (GENERIC) {
                    if ((a.st_flags & ST_TRUNC) && loc >= a.n_valid_out) continue;
                    if (a.st_flags & ST_MULAUX) v = cmul(v, __ldg(a.aux_out + loc));
                    if (a.st_flags & ST_DIV) { v.x /= a.div; v.y /= a.div; }
                }
                reinterpret_cast<cpx*>(a.out)[base + loc] = v;
            }
        }
        if constexpr (LOG2L > 4) __syncthreads();   // shared memory is reused by the next tile
    }
}

}  // namespace gd

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
//
// The lean variant (GENERIC = false) software-pipelines tiles: as soon as the last
// shared-memory gather of tile t is done, the inputs of tile t+grid are fetched with
// cp.async (LDGSTS, L2-only) into the now idle exchange buffer, so their latency is
// hidden behind the last butterfly step, the fused twiddle and the stores of tile t.
#pragma once
#include "fft_core.cuh"

namespace gd {

enum : int { MODE_ROW = 0, MODE_COL = 1 };
enum : int {
    LD_REAL = 1,     // input is float64, imag = 0                     (dsputils.go:25-31 ToComplex)
    LD_PAD = 2,      // elements with local offset >= n_valid_in read as 0 (dsputils.go:49-58 ZeroPad)
    LD_MULAUX = 4,   // multiply by aux_in[local offset]                (bluestein.go:74-76)
    LD_CONJ = 8,     // conjugate after the other load ops (inverse = conj . forward . conj)
    LD_REVERSE = 16, // read x[(n_valid_in - n) mod n_valid_in] instead of x[n]  (fft.go:39-43 IFFT index reversal)
};
enum : int {
    ST_TWIDDLE = 1,  // multiply output k of the line by w_M^(mult*k)   (four-step twiddle)
    ST_MULAUX = 2,   // multiply by aux_out[local offset]               (fft.go:63-66, bluestein.go:89-91)
    ST_SCALE = 4,    // multiply by scale (exact 1/N for power-of-two N; fft.go:47-50)
    ST_DIV = 8,      // divide by div (non power-of-two N keeps the reference's true division)
    ST_TRUNC = 16,   // skip outputs with local offset >= n_valid_out   (bluestein.go:93)
    ST_CONJ = 32,    // conjugate before the other store ops
};

struct PassParams {
    const void* in;
    void* out;
    long long nlines;          // number of lines in this launch
    long long inner;           // line id = q*inner + i
    long long in_qs, in_is;    // element strides of the line base: base = q*qs + i*is
    long long out_qs, out_is;
    int in_es, out_es;         // element stride inside a line (offsets inside a line fit 31 bits)
    int in_mode, out_mode;     // MODE_ROW: lanes run along the line; MODE_COL: lanes run across adjacent lines
    // Tile-major four-step intermediate: element (k, line i) of a transform lives at
    // (i/T)*T*tiled_len + k*T + i%T, so a pass-1 tile (T adjacent columns) is one contiguous block.
    // out_tiled: this pass writes it (lines = columns i, element k).  in_tiled: this pass reads it
    // (lines = rows k, element i); both need MODE_COL. tiled_len = number of rows (N1).
    int out_tiled, in_tiled, tiled_len;
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

__device__ __forceinline__ cpx cconj_if(cpx v, unsigned mask) {   // mask = 0 or 0x80000000: one LOP3
    return make_double2(v.x, __hiloint2double(__double2hiint(v.y) ^ (int)mask, __double2loint(v.y)));
}

// 16-byte async copy global -> shared, L2-only; bytes = 0 zero-fills the slot
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

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

// The same step with its twiddle bases loaded separately (so the loads can be issued ahead of a
// burst of prefetch copies instead of queueing behind them in the load/store unit).
template <int L, int R, int NS>
__device__ __forceinline__ void load_step_twiddles(cpx (&w)[16 / R], int p, const cpx* __restrict__ wl) {
    constexpr int P = L / 16, NB = 16 / R;
#pragma unroll
    for (int q = 0; q < NB; q++) {
        int k = (p + P * q) & (NS - 1);
        w[q] = __ldg(wl + k * (L / (NS * R)));
    }
}
template <int L, int R, int NS>
__device__ __forceinline__ void butterfly_step_w(cpx (&x)[16], const cpx (&w)[16 / R]) {
    constexpr int NB = 16 / R;
#pragma unroll
    for (int q = 0; q < NB; q++) {
        twiddle_powers<R, NB>(&x[q], w[q]);
        dft<R, NB>(&x[q]);
    }
}

// scatter the step's outputs into the line's shared-memory region (Stockham auto-sort).
// pad_idx(base + r*NS) = pad_idx(base) + r*(NS + NS/16) for NS >= 16, and = pad_idx(base) + r
// for NS == 1 with R == 16: one base index per butterfly, compile-time offsets per output.
template <int L, int R, int NS>
__device__ __forceinline__ void scatter_step(const cpx (&x)[16], int p, cpx* __restrict__ sl) {
    constexpr int P = L / 16, NB = 16 / R;
#pragma unroll
    for (int q = 0; q < NB; q++) {
        int j = p + P * q;
        int k = j & (NS - 1);
        int base = (j - k) * R + k;
        if constexpr (NS >= 16 || (NS == 1 && R == 16)) {
            constexpr int RS = NS >= 16 ? NS + (NS >> 4) : 1;
            cpx* b = sl + pad_idx(base);
#pragma unroll
            for (int r = 0; r < R; r++) b[r * RS] = x[q + NB * r];
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) sl[pad_idx(base + r * NS)] = x[q + NB * r];
        }
    }
}
template <int L>
__device__ __forceinline__ void gather_step(cpx (&x)[16], int p, const cpx* __restrict__ sl) {
    constexpr int P = L / 16;
    if constexpr (P >= 16) {
        const cpx* b = sl + pad_idx(p);
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = b[i * (P + (P >> 4))];
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = sl[pad_idx(p + P * i)];
    }
}

// One whole line transform on data already in registers: thread p of the line holds x[n = p + P i] on entry and
// X[k = p + P i] on return (the row-mode layout of fft_pass_kernel on both sides). Every thread of the CTA must call it
// (CTA-wide barriers); the line's exchange region sl is free again on return.
template <int LOG2L>
__device__ __forceinline__ void line_transform(cpx (&x)[16], int p, cpx* __restrict__ sl, const cpx* __restrict__ wl) {
    using SH = PassShape<LOG2L>;
    constexpr int L = SH::L;
    if constexpr (LOG2L <= 4) {
        dft<L, 1>(x);
    } else {
        butterfly_step<L, 16, 1>(x, p, wl);
        scatter_step<L, 16, 1>(x, p, sl);
        __syncthreads();
        gather_step<L>(x, p, sl);
        if constexpr (SH::NSTEP >= 3) {
            __syncthreads();
            butterfly_step<L, 16, 16>(x, p, wl);
            scatter_step<L, 16, 16>(x, p, sl);
            __syncthreads();
            gather_step<L>(x, p, sl);
        }
        if constexpr (SH::NSTEP == 4) {                   // L = 8192 (bluestein_small.cuh only): 16 x 16 x 16 x 2
            __syncthreads();
            butterfly_step<L, 16, 256>(x, p, wl);
            scatter_step<L, 16, 256>(x, p, sl);
            __syncthreads();
            gather_step<L>(x, p, sl);
        }
        __syncthreads();
        butterfly_step<L, SH::LASTR, (SH::NSTEP == 2 ? 16 : SH::NSTEP == 3 ? 256 : 4096)>(x, p, wl);
    }
}

// per-tile addressing of one thread
struct LineRef {
    long long q, ii;
    bool valid;
};
__device__ __forceinline__ LineRef line_ref(const PassParams& a, long long line) {
    LineRef r;
    r.valid = line < a.nlines;
    r.q = 0; r.ii = 0;
    if (r.valid) {
        if (a.inner == 1) r.q = line;
        else { r.q = line / a.inner; r.ii = line - r.q * a.inner; }
    }
    return r;
}

template <int LOG2L, int T, bool GENERIC>
__global__ void __launch_bounds__(T * PassShape<LOG2L>::P, (T * PassShape<LOG2L>::P >= 512) ? 1 : 2)
fft_pass_kernel(const PassParams a) {
    using SH = PassShape<LOG2L>;
    constexpr int L = SH::L, PPT = SH::PPT, P = SH::P, NT = T * P;
    constexpr int LS = line_stride(L, T);
    constexpr bool PREFETCH = !GENERIC && LOG2L >= 8;     // T*L == 16*NT and the exchange buffer can hold a tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* sm = reinterpret_cast<cpx*>(smem_raw);

    const int tid = threadIdx.x;
    const long long ntiles = (a.nlines + T - 1) / T;
    const unsigned ld_conj = (a.ld_flags & LD_CONJ) ? 0x80000000u : 0u;
    // thread -> (line, p) on the load side / store side
    int ell_in, p_in, ell_out, p_out;
    if (a.in_mode == MODE_COL) { ell_in = tid % T; p_in = tid / T; } else { p_in = tid % P; ell_in = tid / P; }
    if (a.out_mode == MODE_COL) { ell_out = tid % T; p_out = tid / T; } else { p_out = tid % P; ell_out = tid / P; }
    int in_off0 = p_in * a.in_es, in_step = P * a.in_es;
    int out_off0 = p_out * a.out_es, out_step = P * a.out_es;
    if (a.in_tiled) { in_off0 = (p_in / T) * (T * a.tiled_len) + (p_in % T); in_step = P * a.tiled_len; }
    if (a.out_tiled) { out_off0 = p_out * T; out_step = P * T; }
    // line-base offset inside a transform: plain i*is, or the tile-major forms
    auto in_line = [&](long long ii) -> long long { return a.in_tiled ? ii * T : ii * a.in_is; };
    auto out_line = [&](long long ii) -> long long {
        return a.out_tiled ? (ii / T) * ((long long)T * a.tiled_len) + (ii % T) : ii * a.out_is;
    };

    auto prefetch = [&](long long tile) {
        LineRef lr = line_ref(a, tile * T + ell_in);
        const cpx* src = reinterpret_cast<const cpx*>(a.in) + lr.q * a.in_qs + in_line(lr.ii) + in_off0;
        const int bytes = lr.valid ? 16 : 0;
        if (!lr.valid) src = reinterpret_cast<const cpx*>(a.in);
#pragma unroll
        for (int i = 0; i < 16; i++) cp_async16(sm + i * NT + tid, lr.valid ? src + i * in_step : src, bytes);
        cp_async_commit();
    };

    long long tile = blockIdx.x;
    if constexpr (PREFETCH) { if (tile < ntiles) prefetch(tile); }

    for (; tile < ntiles; tile += gridDim.x) {
        cpx x[16];
        // ---- load
        if constexpr (PREFETCH) {
            cp_async_wait_all();
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = cconj_if(sm[i * NT + tid], ld_conj);
        } else {
            LineRef lr = line_ref(a, tile * T + ell_in);
            const long long base = lr.q * a.in_qs, loc0 = in_line(lr.ii) + in_off0;
#pragma unroll
            for (int i = 0; i < PPT; i++) {
                cpx v = make_double2(0.0, 0.0);
                if constexpr (GENERIC) {
                    const long long loc = loc0 + (long long)i * in_step;
                    bool ok = lr.valid && !((a.ld_flags & LD_PAD) && loc >= a.n_valid_in);
                    if (ok) {
                        long long src = loc;
                        if ((a.ld_flags & LD_REVERSE) && loc != 0) src = a.n_valid_in - loc;
                        if (a.ld_flags & LD_REAL) v.x = __ldg(reinterpret_cast<const double*>(a.in) + base + src);
                        else v = __ldg(reinterpret_cast<const cpx*>(a.in) + base + src);
                        if (a.ld_flags & LD_MULAUX) v = cmul(v, __ldg(a.aux_in + loc));
                    }
                } else {
                    if (lr.valid) v = __ldg(reinterpret_cast<const cpx*>(a.in) + base + loc0 + i * in_step);
                }
                x[i] = cconj_if(v, ld_conj);
            }
        }

        // ---- transform
        if constexpr (LOG2L <= 4) {
            dft<L, 1>(x);
        } else {
            cpx* sl = sm + ell_in * LS;
            butterfly_step<L, 16, 1>(x, p_in, a.wl);
            if constexpr (PREFETCH) __syncthreads();      // every thread has taken its prefetched inputs
            scatter_step<L, 16, 1>(x, p_in, sl);
            __syncthreads();
            if constexpr (SH::NSTEP == 2) {
                gather_step<L>(x, p_out, sm + ell_out * LS);
            } else {
                gather_step<L>(x, p_in, sl);
                __syncthreads();
                butterfly_step<L, 16, 16>(x, p_in, a.wl);
                scatter_step<L, 16, 16>(x, p_in, sl);
                __syncthreads();
                gather_step<L>(x, p_out, sm + ell_out * LS);
            }
            __syncthreads();                              // exchange buffer is free again
            if constexpr (PREFETCH) { if (tile + gridDim.x < ntiles) prefetch(tile + gridDim.x); }
            butterfly_step<L, SH::LASTR, (SH::NSTEP == 2 ? 16 : 256)>(x, p_out, a.wl);
        }

        LineRef lo = line_ref(a, tile * T + ell_out);

        // ---- fused output twiddle  w_M^(mult * k),  k = p + P*i
        if (a.st_flags & ST_TWIDDLE) {
            unsigned long long mask = (1ULL << a.tw_log2m) - 1ULL;
            unsigned long long mult = (unsigned long long)(a.tw_sel ? lo.q : lo.ii);
            cpx t0 = tw_lookup(a, (mult * (unsigned long long)p_out) & mask);
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
        if (lo.valid) {
            const long long base = lo.q * a.out_qs, loc0 = out_line(lo.ii) + out_off0;
            double sx = 1.0, sy = 1.0;
            if (a.st_flags & ST_SCALE) { sx = a.scale; sy = a.scale; }
            if (a.st_flags & ST_CONJ) sy = -sy;
            const bool scaled = a.st_flags & (ST_SCALE | ST_CONJ);
            cpx* dst = reinterpret_cast<cpx*>(a.out) + base + loc0;
            if constexpr (GENERIC) {
#pragma unroll
                for (int i = 0; i < PPT; i++) {
                    cpx v = x[i];
                    if (scaled) { v.x *= sx; v.y *= sy; }
                    const long long loc = loc0 + (long long)i * out_step;
                    if ((a.st_flags & ST_TRUNC) && loc >= a.n_valid_out) continue;
                    if (a.st_flags & ST_MULAUX) v = cmul(v, __ldg(a.aux_out + loc));
                    if (a.st_flags & ST_DIV) { v.x /= a.div; v.y /= a.div; }
                    dst[i * out_step] = v;
                }
            } else if (scaled) {                  // uniform branch: no per-element selects
#pragma unroll
                for (int i = 0; i < PPT; i++) dst[i * out_step] = make_double2(x[i].x * sx, x[i].y * sy);
            } else {
#pragma unroll
                for (int i = 0; i < PPT; i++) dst[i * out_step] = x[i];
            }
        }
        if constexpr (LOG2L > 4 && !PREFETCH) __syncthreads();   // shared memory is reused by the next tile
    }
}

}  // namespace gd

// fft_w32.cuh -- length-1024 lines as 32 x 32: 32 points per thread, two radix-32 steps, ONE
// shared-memory exchange per line (the 16-point kernel of fft_pass.cuh needs two exchanges plus a
// cp.async staging copy, and is bound by the shared-memory / L1 data path: 96 B per point and pass
// against 32 B here). Inputs are loaded straight into registers; latency is hidden by co-resident
// CTAs instead of a staging buffer. Used for both passes of the 2^20-point four-step
// (replaces the 20 radix-2 sweeps of fft/radix2.go:131-151) and for any lean pass over 1024-point lines.
#pragma once
#include "fft_pass.cuh"

namespace gd {

// multiply by exp(-2*pi*i*M/32), M a compile-time constant
template <int M>
__device__ __forceinline__ cpx mul_w32(cpx a) {
    constexpr int m = ((M % 32) + 32) % 32;
    if constexpr ((m & 1) == 0) return mul_w16<m / 2>(a);
    else {
        // cos/sin(pi*k/16), k odd
        constexpr double C[4] = { 0.98078528040323044913, 0.83146961230254523708, 0.55557023301960222474, 0.19509032201612826785 };
        // angle index a = m (odd, 1..31); fold into the first octant pair by symmetry
        constexpr int q = m / 8;                 // which 45-degree sector pair (0..3)
        constexpr int r = m % 8;                 // 1,3,5,7
        // exp(-i*pi*m/16) = exp(-i*pi*q/2) * exp(-i*pi*r/16)
        constexpr double c = (r == 1) ? C[0] : (r == 3) ? C[1] : (r == 5) ? C[2] : C[3];
        constexpr double s = (r == 1) ? C[3] : (r == 3) ? C[2] : (r == 5) ? C[1] : C[0];
        cpx t = cmul(a, make_double2(c, -s));
        if constexpr (q == 0) return t;
        else if constexpr (q == 1) return mul_mi(t);
        else if constexpr (q == 2) return make_double2(-t.x, -t.y);
        else return make_double2(-t.y, t.x);
    }
}

// forward 32-point DFT in registers, natural order in and out: v[k] = sum_n v[n] w32^(n k)
__device__ __forceinline__ void dft32(cpx (&v)[32]) {
    dft32_fma(v);                                // generated FMA-form codelet (tools/gen_codelets.py): 376 FP64 instructions
}

// v[r] *= t0 * s^r, r = 0..31: four interleaved product chains (depth 8 instead of 31)
__device__ __forceinline__ void mul_geometric32(cpx (&v)[32], cpx t0, cpx s) {
    const cpx s2 = csqr(s), s4 = csqr(s2);
    cpx t[4];
    t[0] = t0; t[1] = cmul(t0, s); t[2] = cmul(t0, s2); t[3] = cmul(t[1], s2);
#pragma unroll
    for (int b = 0; b < 8; b++) {
#pragma unroll
        for (int a = 0; a < 4; a++) {
            v[4 * b + a] = cmul(v[4 * b + a], t[a]);
            if (b < 7) t[a] = cmul(t[a], s4);
        }
    }
}
// v[r] *= w^r, r = 0..31 (v[0] untouched)
__device__ __forceinline__ void mul_powers32(cpx (&v)[32], cpx w) {
    const cpx w2 = csqr(w), w4 = csqr(w2);
    cpx t[4];
    t[0] = w4; t[1] = w; t[2] = w2; t[3] = cmul(w, w2);
#pragma unroll
    for (int a = 1; a < 4; a++) v[a] = cmul(v[a], t[a]);
#pragma unroll
    for (int b = 1; b < 8; b++) {
#pragma unroll
        for (int a = 0; a < 4; a++) {
            if (a > 0) t[a] = cmul(t[a], w4);
            v[4 * b + a] = cmul(v[4 * b + a], t[a]);
        }
        if (b < 7) t[0] = cmul(t[0], w4);
    }
}

constexpr int W32_L = 1024;
// per-line shared-memory stride (16-byte units): element e of a line lives at e + e/32; T adjacent lines land in
// distinct 16-byte bank groups when a quarter warp spans T lines x 8/T neighbouring p
__host__ __device__ constexpr int w32_line_stride(int T) { return 1056 + (T >= 8 ? 1 : (T == 4 ? 2 : 4)); }

// The exchange between the two radix-32 steps: thread p holds y[32 p + r] and needs y[p + 32 r].
// (p_in, sl_in) is the thread's place on the load side, (p_out, sl_out) on the store side (they differ when
// a pass reads rows and writes columns). Barrier A: every thread has taken its staged inputs / the previous
// tile's gathers are done; B: the scatter is visible; C (STAGED only): the buffer is free for the next prefetch.
template <bool STAGED>
__device__ __forceinline__ void w32_exchange(cpx (&x)[32], int p_in, cpx* __restrict__ sl_in, int p_out,
                                             const cpx* __restrict__ sl_out) {
    __syncthreads();                             // A
    {
        cpx* b = sl_in + 33 * p_in;
#pragma unroll
        for (int r = 0; r < 32; r++) b[r] = x[r];
    }
    __syncthreads();                             // B
    {
        const cpx* b = sl_out + p_out;
#pragma unroll
        for (int r = 0; r < 32; r++) x[r] = b[33 * r];   // y[p + 32 r]
    }
    if constexpr (STAGED) __syncthreads();       // C
}

// lean pass over 1024-point lines, T lines per CTA (32 T threads); the PassParams subset of the lean
// fft_pass_kernel: modes, tiled intermediate, LD_CONJ, ST_TWIDDLE, ST_SCALE, ST_CONJ.
// STAGED: the next tile's inputs are fetched with cp.async (L2-only) into the exchange buffer as soon as the
// gathers of this tile are done, each thread into its own slots, so their latency hides behind the second
// radix-32 step, the fused twiddle and the stores. !STAGED: plain loads into registers.
template <int T, int MINB, bool STAGED>
__global__ void __launch_bounds__(32 * T, MINB) fft_pass32_kernel(const PassParams a) {
    constexpr int P = 32, NT = T * P;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* sm = reinterpret_cast<cpx*>(smem_raw);
    const int tid = threadIdx.x;
    const long long ntiles = (a.nlines + T - 1) / T;
    const unsigned ld_conj = (a.ld_flags & LD_CONJ) ? 0x80000000u : 0u;
    int ell_in, p_in, ell_out, p_out;
    if (a.in_mode == MODE_COL) { ell_in = tid % T; p_in = tid / T; } else { p_in = tid % P; ell_in = tid / P; }
    if (a.out_mode == MODE_COL) { ell_out = tid % T; p_out = tid / T; } else { p_out = tid % P; ell_out = tid / P; }
    int in_off0 = p_in * a.in_es, in_step = P * a.in_es;
    int out_off0 = p_out * a.out_es, out_step = P * a.out_es;
    if (a.in_tiled) { in_off0 = (p_in / T) * (T * a.tiled_len) + (p_in % T); in_step = P * a.tiled_len; }
    if (a.out_tiled) { out_off0 = p_out * T; out_step = P * T; }
    auto in_line = [&](long long ii) -> long long { return a.in_tiled ? ii * T : ii * a.in_is; };
    auto out_line = [&](long long ii) -> long long {
        return a.out_tiled ? (ii / T) * ((long long)T * a.tiled_len) + (ii % T) : ii * a.out_is;
    };
    cpx* sl_in = sm + ell_in * w32_line_stride(T);
    const cpx* sl_out = sm + ell_out * w32_line_stride(T);

    auto prefetch = [&](long long tile) {
        LineRef lr = line_ref(a, tile * T + ell_in);
        const cpx* src = reinterpret_cast<const cpx*>(a.in) + lr.q * a.in_qs + in_line(lr.ii) + in_off0;
        const int bytes = lr.valid ? 16 : 0;
        if (!lr.valid) src = reinterpret_cast<const cpx*>(a.in);
#pragma unroll
        for (int i = 0; i < 32; i++) cp_async16(sm + i * NT + tid, lr.valid ? src + (long long)i * in_step : src, bytes);
        cp_async_commit();
    };

    long long tile = blockIdx.x;
    if constexpr (STAGED) { if (tile < ntiles) prefetch(tile); }
    for (; tile < ntiles; tile += gridDim.x) {
        cpx x[32];
        if constexpr (STAGED) {
            cp_async_wait_all();
#pragma unroll
            for (int i = 0; i < 32; i++) x[i] = cconj_if(sm[i * NT + tid], ld_conj);
        } else {
            LineRef lr = line_ref(a, tile * T + ell_in);
            const cpx* src = reinterpret_cast<const cpx*>(a.in) + lr.q * a.in_qs + in_line(lr.ii) + in_off0;
            if (lr.valid) {
#pragma unroll
                for (int i = 0; i < 32; i++) x[i] = cconj_if(__ldcg(src + (long long)i * in_step), ld_conj);
            } else {
#pragma unroll
                for (int i = 0; i < 32; i++) x[i] = make_double2(0.0, 0.0);
            }
        }
        dft32(x);                                // y[32 p + r]
        w32_exchange<STAGED>(x, p_in, sl_in, p_out, sl_out);
        const cpx w = __ldg(a.wl + p_out);       // exp(-2 pi i p / 1024)
        if constexpr (STAGED) { if (tile + gridDim.x < ntiles) prefetch(tile + gridDim.x); }
        mul_powers32(x, w);
        dft32(x);                                // X[p + 32 r]

        LineRef lo = line_ref(a, tile * T + ell_out);
        if (a.st_flags & ST_TWIDDLE) {
            const unsigned long long mask = (1ULL << a.tw_log2m) - 1ULL;
            const unsigned long long mult = (unsigned long long)(a.tw_sel ? lo.q : lo.ii);
            const cpx t0 = tw_lookup(a, (mult * (unsigned long long)p_out) & mask);
            const cpx s1 = tw_lookup(a, (mult * (unsigned long long)P) & mask);
            mul_geometric32(x, t0, s1);
        }
        if (lo.valid) {
            cpx* dst = reinterpret_cast<cpx*>(a.out) + lo.q * a.out_qs + out_line(lo.ii) + out_off0;
            if (a.st_flags & (ST_SCALE | ST_CONJ)) {
                double sx = 1.0, sy = 1.0;
                if (a.st_flags & ST_SCALE) { sx = a.scale; sy = a.scale; }
                if (a.st_flags & ST_CONJ) sy = -sy;
#pragma unroll
                for (int i = 0; i < 32; i++) dst[(long long)i * out_step] = make_double2(x[i].x * sx, x[i].y * sy);
            } else {
#pragma unroll
                for (int i = 0; i < 32; i++) dst[(long long)i * out_step] = x[i];
            }
        }
    }
}

}  // namespace gd

// bluestein_small.cuh -- a whole Bluestein transform (fft/bluestein.go:67-96) of a line in ONE kernel, for padded
// lengths la <= 8192 (n <= 4096): chirp multiply and zero-pad on load, FFT_la, product with the cached FFT(b),
// inverse FFT_la as conj . FFT . conj with 1/la, chirp multiply, truncation to n -- the padded sequence never leaves
// the SM. Memory traffic is the n inputs and n outputs of each line (32 B per point) instead of two la-point
// transforms through memory; the arithmetic, operation by operation, is that of the two fft_pass_kernel<GENERIC>
// launches it replaces (LD_PAD | LD_MULAUX [| LD_REVERSE] -> ST_MULAUX, then LD_CONJ -> ST_CONJ | ST_SCALE | ST_MULAUX |
// ST_TRUNC [| ST_DIV]), so both paths give the same bits.
#pragma once
#include "fft_pass.cuh"

namespace gd {

struct BluesteinSmallParams {
    const void* in;              // batch lines of n complex (or real) values, in_dist elements apart
    cpx* out;                    // batch lines of n complex values, out_dist apart
    long long in_dist, out_dist, n, batch;
    const cpx* chirp;            // conj chirp, n entries        (bluestein.go:26-45)
    const cpx* bhat;             // FFT_la(b), la entries        (bluestein.go:78-87)
    const cpx* wl;               // exp(-2 pi i e / la), e < la
    double scale, div;           // 1 / la (exact); n (inverse only: the reference divides, fft.go:47-50)
    int real_in, inverse;        // inverse: the reference transforms the index-reversed input (fft.go:39-43)
};

template <int LOG2L, int T>
__global__ void __launch_bounds__(T * PassShape<LOG2L>::P, (T * PassShape<LOG2L>::P >= 512) ? 1 : 2)
bluestein_small_kernel(const BluesteinSmallParams a) {
    using SH = PassShape<LOG2L>;
    constexpr int L = SH::L, PPT = SH::PPT, P = SH::P;
    constexpr int LS = line_stride(L, T);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* sm = reinterpret_cast<cpx*>(smem_raw);
    const int tid = threadIdx.x, p = tid % P, ell = tid / P;
    cpx* sl = sm + ell * LS;
    const long long ntiles = (a.batch + T - 1) / T;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long line = tile * T + ell;
        const bool valid = line < a.batch;
        cpx x[16];
#pragma unroll
        for (int i = 0; i < PPT; i++) {                      // a[k] = x[k] * conj(chirp)[k], zero beyond n
            const long long loc = p + P * i;
            cpx v = make_double2(0.0, 0.0);
            if (valid && loc < a.n) {
                const long long src = (a.inverse && loc != 0) ? a.n - loc : loc;
                if (a.real_in) v.x = __ldg(reinterpret_cast<const double*>(a.in) + line * a.in_dist + src);
                else v = __ldg(reinterpret_cast<const cpx*>(a.in) + line * a.in_dist + src);
                v = cmul(v, __ldg(a.chirp + loc));
            }
            x[i] = v;
        }
        line_transform<LOG2L>(x, p, sl, a.wl);
#pragma unroll
        for (int i = 0; i < PPT; i++) {                      // A = FFT(a) * FFT(b); conjugate for the inverse transform
            const cpx v = cmul(x[i], __ldg(a.bhat + p + P * i));
            x[i] = make_double2(v.x, -v.y);
        }
        line_transform<LOG2L>(x, p, sl, a.wl);
        if (valid) {
            cpx* dst = a.out + line * a.out_dist;
#pragma unroll
            for (int i = 0; i < PPT; i++) {                  // r[k] / la, * conj(chirp)[k], k < n
                const long long loc = p + P * i;
                if (loc >= a.n) continue;
                cpx v = make_double2(x[i].x * a.scale, x[i].y * -a.scale);
                v = cmul(v, __ldg(a.chirp + loc));
                if (a.inverse) { v.x /= a.div; v.y /= a.div; }
                dst[loc] = v;
            }
        }
    }
}

}  // namespace gd

// fft_core.cuh -- register-level complex128 FFT building blocks for sm_100a.
//
// Replaces the butterfly loop of the reference (fft/radix2.go:104-121, one radix-2
// stage per full memory sweep, bit-reversed input from radix2.go:157-168) with
// auto-sorting Stockham steps of radix 16/8/4/2 held in registers: 16 points per
// thread, one shared-memory exchange between steps, no bit-reversal pass at all.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gd {

typedef double2 cpx;

__device__ __forceinline__ cpx cmul(cpx a, cpx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cpx csqr(cpx a) {
    return make_double2((a.x - a.y) * (a.x + a.y), 2.0 * a.x * a.y);
}
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cpx cswap(cpx a) { return make_double2(a.y, a.x); }
// multiply by -i  (forward-transform quarter turn)
__device__ __forceinline__ cpx mul_mi(cpx a) { return make_double2(a.y, -a.x); }

// multiply by exp(-2*pi*i*M/16), M a compile-time constant
template <int M>
__device__ __forceinline__ cpx mul_w16(cpx a) {
    constexpr double C1 = 0.92387953251128675613;   // cos(pi/8)
    constexpr double S1 = 0.38268343236508977173;   // sin(pi/8)
    constexpr double H = 0.70710678118654752440;    // sqrt(2)/2
    constexpr int m = ((M % 16) + 16) % 16;
    if constexpr (m == 0) return a;
    else if constexpr (m == 4) return mul_mi(a);
    else if constexpr (m == 8) return make_double2(-a.x, -a.y);
    else if constexpr (m == 12) return make_double2(-a.y, a.x);
    else if constexpr (m == 2) return make_double2(H * (a.x + a.y), H * (a.y - a.x));
    else if constexpr (m == 6) return make_double2(H * (a.y - a.x), -H * (a.x + a.y));
    else if constexpr (m == 10) return make_double2(-H * (a.x + a.y), H * (a.x - a.y));
    else if constexpr (m == 14) return make_double2(H * (a.x - a.y), H * (a.x + a.y));
    else if constexpr (m == 1) return cmul(a, make_double2(C1, -S1));
    else if constexpr (m == 3) return cmul(a, make_double2(S1, -C1));
    else if constexpr (m == 5) return cmul(a, make_double2(-S1, -C1));
    else if constexpr (m == 7) return cmul(a, make_double2(-C1, -S1));
    else if constexpr (m == 9) return cmul(a, make_double2(-C1, S1));
    else if constexpr (m == 11) return cmul(a, make_double2(-S1, C1));
    else if constexpr (m == 13) return cmul(a, make_double2(S1, C1));
    else return cmul(a, make_double2(C1, S1));       // m == 15
}

// In-register forward DFTs, natural-order in, natural-order out, elements at v[0], v[S], v[2S], ...
template <int S>
__device__ __forceinline__ void dft2(cpx* v) {
    cpx a = v[0], b = v[S];
    v[0] = cadd(a, b); v[S] = csub(a, b);
}
template <int S>
__device__ __forceinline__ void dft4(cpx* v) {
    cpx a0 = cadd(v[0], v[2 * S]), a1 = csub(v[0], v[2 * S]);
    cpx a2 = cadd(v[S], v[3 * S]), a3 = mul_mi(csub(v[S], v[3 * S]));
    v[0] = cadd(a0, a2); v[2 * S] = csub(a0, a2);
    v[S] = cadd(a1, a3); v[3 * S] = csub(a1, a3);
}
template <int S>
__device__ __forceinline__ void dft8(cpx* v) {
    // 2 x 4: n = 2*n1 + n2 ; k = k1 + 4*k2
    cpx e[4] = { v[0], v[2 * S], v[4 * S], v[6 * S] };
    cpx o[4] = { v[S], v[3 * S], v[5 * S], v[7 * S] };
    dft4<1>(e); dft4<1>(o);
    o[1] = mul_w16<2>(o[1]); o[2] = mul_w16<4>(o[2]); o[3] = mul_w16<6>(o[3]);
#pragma unroll
    for (int k = 0; k < 4; k++) { v[k * S] = cadd(e[k], o[k]); v[(k + 4) * S] = csub(e[k], o[k]); }
}
template <int S>
__device__ __forceinline__ void dft16(cpx* v) {
    // 4 x 4: n = 4*n1 + n2 ; k = k1 + 4*k2
    cpx c[4][4];
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) {
        c[n2][0] = v[n2 * S]; c[n2][1] = v[(4 + n2) * S]; c[n2][2] = v[(8 + n2) * S]; c[n2][3] = v[(12 + n2) * S];
        dft4<1>(c[n2]);                       // c[n2][k1]
    }
    c[1][1] = mul_w16<1>(c[1][1]); c[1][2] = mul_w16<2>(c[1][2]); c[1][3] = mul_w16<3>(c[1][3]);
    c[2][1] = mul_w16<2>(c[2][1]); c[2][2] = mul_w16<4>(c[2][2]); c[2][3] = mul_w16<6>(c[2][3]);
    c[3][1] = mul_w16<3>(c[3][1]); c[3][2] = mul_w16<6>(c[3][2]); c[3][3] = mul_w16<9>(c[3][3]);
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) {
        cpx r[4] = { c[0][k1], c[1][k1], c[2][k1], c[3][k1] };
        dft4<1>(r);                           // r[k2]
        v[k1 * S] = r[0]; v[(k1 + 4) * S] = r[1]; v[(k1 + 8) * S] = r[2]; v[(k1 + 12) * S] = r[3];
    }
}
}  // namespace gd
#include "fft_codelets.cuh"      // generated FMA-form codelets (tools/gen_codelets.py): dft8_fma 52, dft16_fma 144, dft32_fma 376 instructions
namespace gd {
template <int R, int S>
__device__ __forceinline__ void dft(cpx* v) {
    if constexpr (R == 2) dft2<S>(v);
    else if constexpr (R == 4) dft4<S>(v);
    else if constexpr (R == 8) dft8_fma<S>(v);       // 52 FP64 instructions (hand-written dft8 above: 64)
    else if constexpr (R == 16) dft16_fma<S>(v);     // 144 (dft16 above: 160)
}

// v[r*S] *= w^r for r = 1..R-1, powers built by squaring/products from w (depth <= 4 multiplies).
template <int R, int S>
__device__ __forceinline__ void twiddle_powers(cpx* v, cpx w) {
    v[S] = cmul(v[S], w);
    if constexpr (R >= 4) {
        cpx w2 = csqr(w), w3 = cmul(w, w2);
        v[2 * S] = cmul(v[2 * S], w2); v[3 * S] = cmul(v[3 * S], w3);
        if constexpr (R >= 8) {
            cpx w4 = csqr(w2);
            v[4 * S] = cmul(v[4 * S], w4);
            v[5 * S] = cmul(v[5 * S], cmul(w, w4));
            v[6 * S] = cmul(v[6 * S], cmul(w2, w4));
            cpx w7 = cmul(w3, w4);
            v[7 * S] = cmul(v[7 * S], w7);
            if constexpr (R >= 16) {
                cpx w8 = csqr(w4);
                v[8 * S] = cmul(v[8 * S], w8);
                v[9 * S] = cmul(v[9 * S], cmul(w, w8));
                v[10 * S] = cmul(v[10 * S], cmul(w2, w8));
                v[11 * S] = cmul(v[11 * S], cmul(w3, w8));
                cpx w12 = cmul(w4, w8);
                v[12 * S] = cmul(v[12 * S], w12);
                v[13 * S] = cmul(v[13 * S], cmul(w, w12));
                v[14 * S] = cmul(v[14 * S], cmul(w2, w12));
                v[15 * S] = cmul(v[15 * S], cmul(w7, w8));
            }
        }
    }
}

// Shared-memory index padding: one extra 16-byte slot every 16 elements keeps the
// radix-16 scatter (stride 16) and the stride-1 gather both conflict-free.
__host__ __device__ __forceinline__ constexpr int pad_idx(int i) { return i + (i >> 4); }

// Per-line shared-memory stride (in cpx units): padded length, adjusted so that T
// adjacent lines land in distinct 16-byte bank groups when lanes run across lines.
__host__ __device__ constexpr int line_stride(int L, int T) {
    int lp = L + (L >> 4);
    int want = T >= 8 ? 1 : (T == 4 ? 2 : (T == 2 ? 4 : 0));   // lp + adj == want (mod 8)
    int adj = ((want - (lp % 8)) % 8 + 8) % 8;
    return lp + adj;
}

__host__ __device__ constexpr int ilog2c(int v) { return v <= 1 ? 0 : 1 + ilog2c(v >> 1); }

// SplitMix64 counter generator of the synthetic benchmark inputs.
__host__ __device__ __forceinline__ double splitmix_unit(uint64_t seed, uint64_t i) {
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

}  // namespace gd

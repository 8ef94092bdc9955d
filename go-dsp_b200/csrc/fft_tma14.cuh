// fft_tma14.cuh -- 2^14-point complex128 transforms (the lines of fft.FFT2 on a 16384 x 16384 matrix, config C3, and any
// batch of 2^14-point transforms) as ONE persistent, TMA-fed launch per axis: both passes of the N = 128 x 128 four-step,
// intermediate resident in L2. Same machinery as fft_tma.cuh (in-order tile queue, loader / storer / watcher lanes,
// D = 2 phases of pass-1 run-ahead over S = 3 scratch slots of 16 MiB); what differs is the tile and the arithmetic.
// Replaces, per axis, the two launches of the pass kernel whose inter-pass array went through HBM
// (fft/fft.go:138-151 is two sweeps of 14 radix-2 stages each in the reference).
//
// A tile is 32 adjacent lines of 128 points: 128 rows x 32 complex (512 contiguous bytes), landing in halves of 64 rows.
//   ROWS (transform t contiguous, x[t][128 n1 + n2]):   lines = 32 adjacent n2 (pass 1) / 32 adjacent k1 (pass 2)
//       P1: rows n1 of x[t]          -> Int[t][n2][k1] (32 rows of 2 KiB, one bulk copy each)
//       P2: rows n2 of Int[t][.][k1] -> X[t][k1 + 128 k2] (tile store)
//   COLS (transform = column t of a row-major R x C matrix, R = 2^14): lines = 32 adjacent columns t in both passes
//       P1: rows n1 of M[128 n1 + n2][t]   -> Int[tb][n2][k1][32 t] (64 KiB contiguous)
//       P2: rows n2 of Int[tb][.][k1][.]   -> M'[k1 + 128 k2][t]
// A phase (the unit of the dependency counters) is 256 tiles = 2^20 points = one 16 MiB slot: 64 transforms (ROWS) or
// 64 columns (COLS).
// PROF: cycle counters of the consumer groups (tools/exp_fft2_axes.py --prof only).
// Consumer group = 4 warps: lane = line, warp j = residue of the point index mod 4. 128 = 32 x 4: a radix-32 step on
// points j + 4 i, the twiddle w_128^(j k) (warp-uniform, read from the kernel parameters: no product chains), one
// shared-memory exchange, eight radix-4 butterflies.
#pragma once
#include "fft_tma.cuh"

namespace gd {

// The same kernel serves N = 2^16 = 256 x 256 (LEN = 256): a tile is then 16 adjacent lines of 256 points (256 rows x 256
// bytes), lane = (line, low bit of the residue mod 8), 256 = 32 x 8: radix-32 step, twiddle w_256^(j k) by a product chain
// (j differs inside a warp), one exchange, four radix-8 butterflies. Everything else (phases of 256 tiles = 2^20 points,
// halves of 32 KiB, slots of 16 MiB) is unchanged.
template <int LEN>
struct T14Shape {
    static_assert(LEN == 128 || LEN == 256, "sub-line length");
    static constexpr int LINES = 4096 / LEN;                          // lines per tile: 32 / 16
    static constexpr int NJ = LEN / 32;                               // residues of the point index handled by different threads: 4 / 8
    static constexpr int KB = 32 / NJ;                                // outputs k = KB j' + k_lo + 32 m per thread: 8 / 4
    static constexpr int TPT = LEN / LINES;                           // ROWS: tiles per transform: 4 / 16
    static constexpr int TPP = 256 / TPT;                             // ROWS: transforms per phase: 64 / 16
    static constexpr int TBP = 256 / LEN;                             // COLS: blocks of LINES columns per phase: 2 / 1
    static constexpr int CPP = TBP * LINES;                           // COLS: columns per phase: 64 / 16
    static constexpr int ROWPITCH = LEN + 1;                          // ROWS pass-1 staging: LINES rows of LEN + 1 elements (skew: conflict-free)
    static constexpr int LOG2N = LEN == 128 ? 14 : 16;
};
constexpr int T14_HALF_BYTES = 32768;                                 // half a tile: LEN / 2 rows of LINES elements
constexpr int T14_WBYTES = 32 * 129 * 16;                             // 66048 >= 65536 (LEN = 256: 16 * 257 * 16 = 65792)
constexpr int T14_WELEMS = T14_WBYTES / 16;
constexpr int T14_SMEM = TMA_NSLOT * T14_HALF_BYTES + 2 * T14_WBYTES + 1024;     // 231424
enum : int { T14_ROWS = 0, T14_COLS = 1 };

struct Tma14Params {
    int batch;                   // phases (groups of 256 tiles) in this launch
    int delay, nslots;
    cpx* scratch;                // nslots slots of 2^20 elements
    int* done1;
    int* done2;
    int* queue;
    const cpx* tw_lo;            // w_N^e = hi[e >> 12] * lo[e & 4095], N = LEN^2
    const cpx* tw_hi;
    double scale;                // inverse: 1/N folded into the four-step twiddle
    cpx w128[3][32];             // LEN = 128: w_128^(j k), j = 1..3, k < 32
    const cpx* wl;               // LEN = 256: exp(-2 pi i p / 256), p < 256
    cpx* out;                    // opt bit 0 only: pass-2 output from registers (512-byte rows: one warp store each)
    int opt;                     // experiments: bit 0 = pass-2 output straight from registers (no staging, no TMA tile store)
    long long out_dist;          //   ROWS: transform t at out + t * out_dist;  COLS: row pitch of the matrix (columns)
    long long* prof;             // measurement: [gridDim.x][16] cycle counters of the consumer groups (null in the product)
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3, const void* src) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];\n"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(src)) : "memory");
}

// tile c of phase (type, grp), half h -> tensor coordinates (in doubles along dim 0)
template <int LEN, int MODE>
__device__ __forceinline__ void t14_coords(int type, int grp, int c, int h, int S, int& c0, int& c1, int& c2, int& c3, bool in) {
    using SH = T14Shape<LEN>;
    if constexpr (MODE == T14_ROWS) {
        const int tl = c / SH::TPT, q = c % SH::TPT;      // transform within the group, block of LINES lines
        c0 = 2 * SH::LINES * q; c1 = (LEN / 2) * h; c3 = 0;
        c2 = (type == 1 && in) ? (grp % S) * SH::TPP + tl : grp * SH::TPP + tl;
    } else {
        const int tbl = c / LEN, r = c % LEN;             // t-block within the group, n2 (pass 1) or k1 (pass 2)
        if (type == 1 && in) { c0 = 0; c1 = r; c2 = (LEN / 2) * h; c3 = (grp % S) * SH::TBP + tbl; }
        else { c0 = 2 * SH::LINES * (grp * SH::TBP + tbl); c1 = r; c2 = (LEN / 2) * h; c3 = 0; }
    }
}

template <int LEN, int MODE, bool INV, bool PROF>
__global__ void __launch_bounds__(TMA_THREADS, 1)
fft_tma14_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_int,
                 const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ Tma14Params a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    cpx* land = reinterpret_cast<cpx*>(smem_raw);
    cpx* work = reinterpret_cast<cpx*>(smem_raw + TMA_NSLOT * T14_HALF_BYTES);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + TMA_NSLOT * T14_HALF_BYTES + 2 * T14_WBYTES);
    unsigned long long* full_h = bars;                     // [3 slots][2 groups]
    unsigned long long* freed_h = bars + 6;                // [3]
    unsigned long long* rd = bars + 9;                     // [2]
    unsigned long long* staged = bars + 11;                // [2]
    unsigned long long* drained = bars + 13;               // [2 groups][2 halves of the work buffer]
    volatile int* log = reinterpret_cast<volatile int*>(bars + 22);        // [32]
    volatile int* log_count = reinterpret_cast<volatile int*>(bars + 38);
    volatile int* ready_sh = reinterpret_cast<volatile int*>(bars + 39);
    using SH = T14Shape<LEN>;
    constexpr int TPT = 256;                               // tiles per phase
    constexpr int HALF_ELEMS = T14_HALF_BYTES / 16;        // 2048
    constexpr int T14_LINES = SH::LINES, T14_LEN = LEN, T14_ROWPITCH = SH::ROWPITCH, NJ = SH::NJ, KB = SH::KB;

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < 6; i++) mbar_init(full_h + i, 1);
        for (int i = 0; i < 3; i++) mbar_init(freed_h + i, TMA_GROUP);
        for (int i = 0; i < 2; i++) { mbar_init(rd + i, TMA_GROUP); mbar_init(staged + i, TMA_GROUP); }
        for (int i = 0; i < 4; i++) mbar_init(drained + i, 1);
        *log_count = 0;
        *ready_sh = 0;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int B = a.batch, D = a.delay, S = a.nslots;
    const int nitems = 2 * B * TPT;

    if (warp >= 2 * TMA_GROUP / 32) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (tid == 2 * TMA_GROUP) {
            // ------------------------------------------------------------ loader (see fft_tma.cuh)
            int tokens = 0, ready_tf = -1;
            long long hidx = 0;
            int cur = atomicAdd(a.queue, 2);
            for (int it = 0; tokens < 2; it++) {
                if (hidx >= TMA_NSLOT) mbar_wait(freed_h + (int)(hidx % TMA_NSLOT), (unsigned)(((hidx - TMA_NSLOT) / TMA_NSLOT) & 1));
                const int item = tokens ? nitems : cur + (it & 1);
                const bool token = item >= nitems;
                TmaItem w;
                w.type = 0; w.tf = 0; w.c = 0;
                if (!token) {
                    w = tma_decode(item, B, D);
                    if (w.type == 1 && w.tf > ready_tf) {
                        while (ld_volatile_shared(ready_sh) <= w.tf) __nanosleep(20);
                        __threadfence_block();
                        asm volatile("fence.proxy.async;\n" ::: "memory");
                        ready_tf = ld_volatile_shared(ready_sh) - 1;
                    }
                }
                log[it & 31] = token ? -1 : item;
                __threadfence_block();
                *log_count = it + 1;
                if (token) tokens++;
#pragma unroll
                for (int h = 0; h < 2; h++, hidx++) {
                    const int s = (int)(hidx % TMA_NSLOT);
                    if (token && h == 1) continue;
                    if (h == 1 && hidx >= TMA_NSLOT) mbar_wait(freed_h + s, (unsigned)(((hidx - TMA_NSLOT) / TMA_NSLOT) & 1));
                    unsigned long long* fb = full_h + 2 * s + (it & 1);
                    if (token) { mbar_arrive(fb); continue; }
                    mbar_expect_tx(fb, T14_HALF_BYTES);
                    int c0, c1, c2, c3;
                    t14_coords<LEN, MODE>(w.type, w.tf, w.c, h, S, c0, c1, c2, c3, true);
                    tma_load_4d(land + (size_t)s * HALF_ELEMS, w.type == 0 ? &tm_x : &tm_int, c0, c1, c2, c3, fb);
                }
                if (!tokens && (it & 1)) cur = atomicAdd(a.queue, 2);
            }
        } else if (tid == 2 * TMA_GROUP + 96) {
            // ------------------------------------------------------------ watcher
            for (int tf = 0; tf < B; tf++) {
                while (ld_relaxed_gpu(a.done1 + tf) < TPT) __nanosleep(64);
                asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
                asm volatile("fence.proxy.async;\n" ::: "memory");
                *ready_sh = tf + 1;
            }
        } else if (tid == 2 * TMA_GROUP + 32 || tid == 2 * TMA_GROUP + 64) {
            // ------------------------------------------------------------ storer of consumer group g
            const int g = tid == 2 * TMA_GROUP + 32 ? 0 : 1;
            const unsigned long long pol_last = policy_evict_last();
            unsigned ns = 0;
            int free_tf = S - 1;
            long long c_staged = 0, c_slot = 0, c_read = 0, c_pub = 0;
            for (int it = g;; it += 2) {
                while (ld_volatile_shared(log_count) <= it) __nanosleep(64);
                __threadfence_block();
                const int item = log[it & 31];
                if (item < 0) break;
                const TmaItem pi = tma_decode(item, B, D);
                if (LEN == 128 && pi.type == 1 && (a.opt & 1)) continue;  // experiment: pass-2 tiles stored by the consumers themselves (nothing staged)
                long long t0 = 0;
                if (PROF) t0 = clock64();
                mbar_wait(staged + g, ns & 1);
                ns++;
                if (PROF) { const long long t1 = clock64(); c_staged += t1 - t0; t0 = t1; }
                const cpx* srcb = work + (size_t)g * T14_WELEMS;
                if (pi.type == 1) {
                    int c0, c1, c2, c3;
                    t14_coords<LEN, MODE>(1, pi.tf, pi.c, 0, S, c0, c1, c2, c3, false);
                    tma_store_4d(&tm_out, c0, c1, c2, c3, srcb);
                    tma_commit();
                    t14_coords<LEN, MODE>(1, pi.tf, pi.c, 1, S, c0, c1, c2, c3, false);
                    tma_store_4d(&tm_out, c0, c1, c2, c3, srcb + HALF_ELEMS);
                    tma_commit();
                } else {
                    // the slot is free once pass 2 of the phase that used it is complete (polling earlier, while the tile is
                    // still being computed, was slower: cols 2.14 -> 2.48 ms)
                    if (pi.tf > free_tf) {
                        while (ld_relaxed_gpu(a.done2 + (pi.tf - S)) < TPT) __nanosleep(32);
                        free_tf = pi.tf;
                        if (PROF) { const long long t1 = clock64(); c_slot += t1 - t0; t0 = t1; }
                    }
                    cpx* slot = a.scratch + (size_t)(pi.tf % S) * ((size_t)1 << 20);
                    if constexpr (MODE == T14_ROWS) {
                        // Int[t][n2 = 32 q + ell][k1]: 32 rows of 2 KiB, contiguous in memory, 129-element pitch in shared memory
                        cpx* dst = slot + (size_t)(pi.c / SH::TPT) * (T14_LEN * T14_LEN) + (size_t)(pi.c % SH::TPT) * T14_LINES * T14_LEN;
#pragma unroll 1
                        for (int l = 0; l < T14_LINES / 2; l++) bulk_store_1d_hint(dst + l * T14_LEN, srcb + l * T14_ROWPITCH, T14_LEN * 16, pol_last);
                        tma_commit();
#pragma unroll 1
                        for (int l = T14_LINES / 2; l < T14_LINES; l++) bulk_store_1d_hint(dst + l * T14_LEN, srcb + l * T14_ROWPITCH, T14_LEN * 16, pol_last);
                        tma_commit();
                    } else {
                        // Int[tb][n2][k1][32 t]: the tile is 64 KiB contiguous
                        cpx* dst = slot + (size_t)pi.c * (T14_LEN * T14_LINES);
                        bulk_store_1d_hint(dst, srcb, T14_HALF_BYTES, pol_last);
                        tma_commit();
                        bulk_store_1d_hint(dst + HALF_ELEMS, srcb + HALF_ELEMS, T14_HALF_BYTES, pol_last);
                        tma_commit();
                    }
                }
                tma_wait_read1();
                mbar_arrive(drained + 2 * g);
                tma_wait_read0();
                mbar_arrive(drained + 2 * g + 1);
                if (PROF) { const long long t1 = clock64(); c_read += t1 - t0; t0 = t1; }
                if (pi.type == 0) {
                    tma_wait_all0();
                    asm volatile("fence.proxy.async.global;\n" ::: "memory");
                    red_release_gpu(a.done1 + pi.tf, 1);
                    if (PROF) c_pub += clock64() - t0;
                }
            }
            tma_wait_all0();
            if (PROF) { long long* q = a.prof + (size_t)blockIdx.x * TMA_PROF_SLOTS + 16 + 4 * g; q[0] = c_staged; q[1] = c_slot; q[2] = c_read; q[3] = c_pub; }
        }
        return;
    }

    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    const int g = warp >> 2;
    // LEN = 128: point residue mod 4 = warp (uniform), line = lane; LEN = 256: residue mod 8 = 2 warp + (lane >> 4), line = lane & 15
    const int j = LEN == 128 ? (warp & 3) : 2 * (warp & 3) + ((tid >> 4) & 1), ell = tid & (T14_LINES - 1);
    cpx wj = make_double2(1.0, 0.0);                        // LEN = 256: w_256^j, the base of the twiddle chain
    if constexpr (LEN == 256) wj = __ldg(a.wl + j);
    cpx* wbuf = work + (size_t)g * T14_WELEMS;
    unsigned nrd = 0, nst = 0;
    bool prev_staged = false;                               // the previous tile left staged output in the work buffer (pass 1 only)
    unsigned fph = 0;
    long long c_full0_p1 = 0, c_full0_p2 = 0, c_full1 = 0, c_drain = 0, c_rd = 0, c_bar = 0, n_tiles = 0, t_start = 0, t0 = 0;
    if (PROF) t_start = clock64();
    for (int it = g;; it += 2) {
        const long long h0 = 2LL * it;
        const int s0 = (int)(h0 % TMA_NSLOT), s1 = (int)((h0 + 1) % TMA_NSLOT);
        if (PROF) t0 = clock64();
        mbar_wait(full_h + 2 * s0 + g, (fph >> s0) & 1);
        fph ^= 1u << s0;
        const int item = log[it & 31];
        if (item < 0) break;
        const TmaItem wi = tma_decode(item, B, D);
        if (PROF) { const long long dt = clock64() - t0; if (wi.type == 0) c_full0_p1 += dt; else c_full0_p2 += dt; n_tiles++; }
        const unsigned ld_conj = (INV && wi.type == 0) ? 0x80000000u : 0u;
        cpx x[32];
        {   // points j + 4 i of line ell: rows j + 4 i of the tile, i < 16 in the first half
            const cpx* s = land + (size_t)s0 * HALF_ELEMS + j * T14_LINES + ell;
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = INV ? cconj_if(s[i * NJ * T14_LINES], ld_conj) : s[i * NJ * T14_LINES];
        }
        mbar_arrive(freed_h + s0);
        if (PROF) t0 = clock64();
        mbar_wait(full_h + 2 * s1 + g, (fph >> s1) & 1);
        fph ^= 1u << s1;
        if (PROF) c_full1 += clock64() - t0;
        if (wi.type == 1 && tid == g * TMA_GROUP) red_relaxed_gpu(a.done2 + wi.tf, 1);
        {
            const cpx* s = land + (size_t)s1 * HALF_ELEMS + j * T14_LINES + ell;
#pragma unroll
            for (int i = 0; i < 16; i++) x[16 + i] = INV ? cconj_if(s[i * NJ * T14_LINES], ld_conj) : s[i * NJ * T14_LINES];
        }
        mbar_arrive(freed_h + s1);
        dft32(x);                                           // Y_j[k] = sum_i x[j + NJ i] w_32^(i k)
        if constexpr (LEN == 128) {
            if (j != 0) {                                   // warp-uniform: w_128^(j k) from the parameter bank
#pragma unroll
                for (int k = 1; k < 32; k++) x[k] = cmul(x[k], a.w128[j - 1][k]);
            }
        } else {
            asm volatile("" : "+d"(wj.x), "+d"(wj.y));
            mul_powers32(x, wj);                            // w_256^(j k): j differs between the halves of a warp
        }
        if (PROF) t0 = clock64();
        if (prev_staged) mbar_wait(drained + 2 * g, (nst - 1) & 1);
        if (PROF) c_drain += clock64() - t0;
        // exchange: Y_j[k] -> row NJ k + j of the work buffer, column ell; thread (ell, j') then takes rows 32 j' .. 32 j' + 31,
        // i.e. k = KB j' + k_lo, all NJ residues
        {
            cpx* s = wbuf + j * T14_LINES + ell;
#pragma unroll
            for (int k = 0; k < 16; k++) s[k * NJ * T14_LINES] = x[k];                       // rows < LEN / 2
            if (PROF) t0 = clock64();
            if (prev_staged) mbar_wait(drained + 2 * g + 1, (nst - 1) & 1);
            if (PROF) c_drain += clock64() - t0;
#pragma unroll
            for (int k = 16; k < 32; k++) s[k * NJ * T14_LINES] = x[k];
        }
        // four-step twiddle bases of this line (pass 1): w^(n2 * KB j'), w^(n2), w^(32 n2)
        cpx tb0, tb1, tb32;
        if (wi.type == 0) {
            constexpr unsigned NMASK = (unsigned)(LEN * LEN - 1);
            const unsigned n2 = MODE == T14_ROWS ? (unsigned)((wi.c % SH::TPT) * T14_LINES + ell) : (unsigned)(wi.c % LEN);
            const unsigned e0 = (n2 * (unsigned)KB * (unsigned)j) & NMASK, e1 = n2 & NMASK, e32 = (n2 * 32u) & NMASK;
            tb0 = cmul(__ldg(a.tw_hi + (e0 >> 12)), __ldg(a.tw_lo + (e0 & 4095u)));
            tb1 = cmul(__ldg(a.tw_hi + (e1 >> 12)), __ldg(a.tw_lo + (e1 & 4095u)));
            tb32 = cmul(__ldg(a.tw_hi + (e32 >> 12)), __ldg(a.tw_lo + (e32 & 4095u)));
            if (INV) tb0 = make_double2(tb0.x * a.scale, tb0.y * a.scale);
        }
        if (PROF) t0 = clock64();
        group_bar(1 + g);
        if (PROF) c_bar += clock64() - t0;
        {
            const cpx* s = wbuf + (32 * j) * T14_LINES + ell;
#pragma unroll
            for (int r = 0; r < 32; r++) x[r] = s[r * T14_LINES];                            // x[NJ k_lo + jj] = Y_jj[KB j + k_lo]
        }
        mbar_arrive(rd + g);
#pragma unroll
        for (int kl = 0; kl < KB; kl++) {                                                    // x[NJ k_lo + m] = X[KB j + k_lo + 32 m]
            if constexpr (NJ == 4) dft4<1>(&x[4 * kl]);
            else dft8_fma<1>(&x[8 * kl]);
        }
        if (PROF) t0 = clock64();
        mbar_wait(rd + g, nrd & 1);                         // every gather of this tile is done: the buffer may be overwritten
        nrd++;
        if (PROF) c_rd += clock64() - t0;
        if (wi.type == 0) {
            // x[NJ k_lo + m] *= w^(n2 (KB j + k_lo + 32 m)) = tb0 * tb1^k_lo * tb32^m
            cpx c[KB];
            c[0] = tb0;
#pragma unroll
            for (int kl = 1; kl < KB; kl++) c[kl] = cmul(c[kl - 1], tb1);
#pragma unroll
            for (int m = 0; m < NJ; m++) {
#pragma unroll
                for (int kl = 0; kl < KB; kl++) {
                    x[NJ * kl + m] = cmul(x[NJ * kl + m], c[kl]);
                    if (m < NJ - 1) c[kl] = cmul(c[kl], tb32);
                }
            }
            if constexpr (MODE == T14_ROWS) {
                cpx* s = wbuf + ell * T14_ROWPITCH + KB * j; // Int[n2 = line][k1 = KB j + k_lo + 32 m]
#pragma unroll
                for (int kl = 0; kl < KB; kl++)
#pragma unroll
                    for (int m = 0; m < NJ; m++) s[kl + 32 * m] = x[NJ * kl + m];
            } else {
                cpx* s = wbuf + (KB * j) * T14_LINES + ell;  // Int[k1][LINES t]: row k1, column = line
#pragma unroll
                for (int kl = 0; kl < KB; kl++)
#pragma unroll
                    for (int m = 0; m < NJ; m++) s[(kl + 32 * m) * T14_LINES] = x[NJ * kl + m];
            }
            fence_proxy_async();
            mbar_arrive(staged + g);
            nst++;
            prev_staged = true;
        } else if (LEN != 128 || !(a.opt & 1)) {
            cpx* s = wbuf + (KB * j) * T14_LINES + ell;      // X[k2 = KB j + k_lo + 32 m]: row k2 of the tile, column = line
#pragma unroll
            for (int kl = 0; kl < KB; kl++)
#pragma unroll
                for (int m = 0; m < NJ; m++) {
                    const cpx v = x[NJ * kl + m];
                    s[(kl + 32 * m) * T14_LINES] = INV ? make_double2(v.x, -v.y) : v;
                }
            fence_proxy_async();
            mbar_arrive(staged + g);
            nst++;
            prev_staged = true;
        } else {
            // experiment (tma_opt=16): X[k2 = 8 j + k_lo + 32 m] of line ell, the 32 lanes of a warp write one 512-byte row per
            // store. Measured equal to the staged path within run-to-run noise (profiles/r2_exp_fft2_axes_prof.jsonl): the
            // wait for the work buffer comes from the pass-1 bulk stores, not from the pass-2 tile stores.
            cpx* dst;
            long long pitch;
            if constexpr (MODE == T14_ROWS) {
                dst = a.out + (long long)(wi.tf * 64 + (wi.c >> 2)) * a.out_dist + (wi.c & 3) * 32 + ell;
                pitch = T14_LEN;
            } else {
                dst = a.out + (long long)(wi.c & 127) * a.out_dist + (long long)(wi.tf * 2 + (wi.c >> 7)) * 32 + ell;
                pitch = (long long)T14_LEN * a.out_dist;
            }
            dst += (long long)(8 * j) * pitch;
#pragma unroll
            for (int kl = 0; kl < 8; kl++)
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const cpx v = x[4 * kl + m];
                    __stcs(reinterpret_cast<double2*>(dst + (long long)(kl + 32 * m) * pitch), INV ? make_double2(v.x, -v.y) : v);
                }
            prev_staged = false;
        }
    }
    if (PROF && (tid & (TMA_GROUP - 1)) == 0) {
        long long* q = a.prof + (size_t)blockIdx.x * TMA_PROF_SLOTS + 8 * g;
        q[0] = c_full0_p1; q[1] = c_full0_p2; q[2] = c_full1; q[3] = c_drain; q[4] = c_rd; q[5] = c_bar; q[6] = clock64() - t_start; q[7] = n_tiles;
    }
}

}  // namespace gd

// fft_tma14.cuh -- complex128 transforms of N = LA x LB points, 2^13 <= N <= 2^19 (LA, LB in {64, 128, 256, 512}, LA >= LB; 2^19 = 1024 x 512, rows only), as
// ONE persistent, TMA-fed launch: both passes of the four-step, intermediate resident in L2. N = 2^14 = 128 x 128 are the
// lines of fft.FFT2 on a 16384 x 16384 matrix (config C3); N = 2^16 = 256 x 256 the local lines of the sharded 2^32-point
// transform (config C5). Same machinery as fft_tma.cuh (in-order tile queue, loader / storer / watcher lanes, D = 2 phases
// of pass-1 run-ahead over S = 3 scratch slots of 16 MiB); what differs is the tile and the arithmetic.
// Replaces, per axis, the two launches of the pass kernel whose inter-pass array went through HBM or an L2-sized chunk
// (fft/fft.go:138-151 and fft/radix2.go:131-151 are log2 N radix-2 sweeps in the reference).
//
// x[n = LB n1 + n2] -> pass 1: lines over n1 (length LA), twiddle w_N^(n2 k1) -> Int -> pass 2: lines over n2 (length LB)
// -> X[k1 + LA k2]. A tile is 4096 points: LINES = 4096 / LEN adjacent lines of LEN points, stored as LEN rows of LINES
// complex (LINES * 16 contiguous bytes), landing in halves of LEN / 2 rows (32 KiB). A phase (the unit of the dependency
// counters) is 256 tiles = 2^20 points = one 16 MiB slot.
//   ROWS (transform t contiguous):   lines = LINES_A adjacent n2 (pass 1) / LINES_B adjacent k1 (pass 2)
//       P1: rows n1 of x[t]          -> Int[t][n2][k1] (LINES_A rows of LA elements, one bulk copy each)
//       P2: rows n2 of Int[t][.][k1] -> X[t][k1 + LA k2] (tile store)
//   COLS (transform = column t of a row-major N x C matrix; LB <= 256): a block tb is LINES_A adjacent columns
//       P1: rows n1 of M[LB n1 + n2][tb]       -> Int[tb][n2][k1][LINES_A t] (64 KiB contiguous)
//       P2: rows n2 of Int[tb][.][k1 ..][.]    -> M'[k1 + LA k2][tb]; a tile takes LINES_B / LINES_A adjacent k1
// Consumer group = 4 warps, 32 points per thread; LEN = 32 NJ: thread (line, j) holds the points j + NJ i of its line,
// lane = line (+ LINES x low bits of j when LINES < 32): a radix-32 step, the twiddle w_LEN^(j k), one shared-memory
// exchange, 32 / NJ butterflies of radix NJ. For LEN = 128 j is warp-uniform and the twiddles come from the kernel
// parameters; otherwise they are a product chain from w_LEN^j.
// PROF: cycle counters of the consumer groups (tools/exp_fft2_axes.py --prof only).
// TW2 (COLS only): the lines are the length-N1 columns of ONE transform of M = N1 x N2 points laid out as [N1][N2]; output k1 of
// column n2 leaves multiplied by w_M^(n2 k1), so the outer four-step's twiddle sweep disappears (fft_pow2_huge in engine.cu).
// The three bases per thread and tile come from sincospi of exactly reduced exponents (no table above 2^24 entries).
// TW2 = 2: the same, and the stores go to the ranks of a sharded transform (T14Peers below).
// TW2 = 3 (ROWS): output k of every transform leaves multiplied by aux[k] (Bluestein's product with the cached FFT(b),
// fft/fft.go:63-66: one sweep less than a separate product kernel; aux is L2 resident).
#pragma once
#include <type_traits>
#include "fft_tma.cuh"

namespace gd {

template <int LEN>
struct T14Len {
    static_assert(LEN == 64 || LEN == 128 || LEN == 256 || LEN == 512 || LEN == 1024, "sub-line length");
    static constexpr int LINES = 4096 / LEN;                          // lines per tile: 64 / 32 / 16 / 8 / 4
    static constexpr int NJ = LEN / 32;                               // residues of the point index held by different threads: 2 / 4 / 8 / 16 / 32
    static constexpr int KB = 32 / NJ;                                // outputs k = KB j' + k_lo + 32 m per thread
    static constexpr int LOG2 = LEN == 64 ? 6 : LEN == 128 ? 7 : LEN == 256 ? 8 : LEN == 512 ? 9 : 10;
};
template <int LA, int LB>
struct T14Shape {
    static_assert(LA >= LB, "LA >= LB");
    static constexpr int N = LA * LB;
    static constexpr int LOG2N = T14Len<LA>::LOG2 + T14Len<LB>::LOG2;
    static constexpr int LINES_A = T14Len<LA>::LINES, LINES_B = T14Len<LB>::LINES;
    static constexpr int RA = LINES_B / LINES_A;                      // COLS: adjacent k1 per pass-2 tile (= LA / LB)
    static constexpr int TPT = N / 4096;                              // ROWS: tiles per transform
    static constexpr int UNIT = (1 << 20) / N;                        // transforms (ROWS) / columns (COLS) per phase
    static constexpr int TBP = LB <= 256 ? 256 / LB : 1;              // COLS: blocks of LINES_A columns per phase (LB <= 256 only)
    static constexpr int ROWPITCH = LA + 1;                           // ROWS pass-1 staging: LINES_A rows of LA + 1 elements (skew: conflict-free)
};
constexpr int T14_HALF_BYTES = 32768;                                 // half a tile: LEN / 2 rows of LINES elements
constexpr int T14_WBYTES = 32 * 129 * 16;                             // 66048 >= LINES_A * (LA + 1) * 16 for LA >= 128
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
    const cpx* tw_lo;            // w_N^e = hi[e >> 12] * lo[e & 4095]
    const cpx* tw_hi;
    double scale;                // inverse: 1/N folded into the four-step twiddle
    cpx w128[3][32];             // LEN = 128: w_128^(j k), j = 1..3, k < 32
    const cpx* wla;              // exp(-2 pi i p / LA), p < LA
    const cpx* wlb;              // exp(-2 pi i p / LB)
    long long* prof;             // measurement: [gridDim.x][TMA_PROF_SLOTS] cycle counters (null in the product)
    int tw2_log2m;               // TW2 (COLS): the stores of pass 2 carry an outer four-step twiddle w_M^(column * k), M = 2^tw2_log2m,
    long long tw2_col0;          //   column = tw2_col0 + the column's index in this launch, k = the output index in the line
    int bshift;                  // COLS: a launch over several matrices (dimension 3 of the maps) has 2^bshift phases per matrix; 31 = one matrix
    int npeer;                   // TW2 == 2: number of ranks G the rows of the output are spread over (T14Peers)
    int prot;                    //   this rank: the boxes of a half tile go out in rotated order, so the ranks do not all store to the same peer at once
    const cpx* aux;              // TW2 == 3 (ROWS): N factors, output k of every transform is multiplied by aux[k]
    int seg;                     // ROWS, > 0: a transform's LA rows of LB points are `seg` segments of LA / seg rows (dimension 2 of the input map)
};
// TW2 == 2 (COLS): the output rows k = k1 + LA k2 of the slab belong to rank k / (N / G): instead of one local output map the stores of
// pass 2 go through one map per rank, straight into that rank's receive buffer over NVLink (a half tile = max(1, G / 2) boxes of
// LB / G rows of k2). The sharded four-step's exchange is then the store phase of its first line pass.
struct T14Peers { CUtensorMap m[8]; };

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3, const void* src) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];\n"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(src)) : "memory");
}

// tile c of phase (type, grp), half h -> tensor coordinates (in doubles along dim 0); in: the load side (pass 2 reads Int)
template <int LA, int LB, int MODE>
__device__ __forceinline__ void t14_coords(int type, int grp, int c, int h, int S, int bshift, int& c0, int& c1, int& c2, int& c3, bool in) {
    using SH = T14Shape<LA, LB>;
    if constexpr (MODE == T14_ROWS) {
        const int tl = c / SH::TPT, q = c % SH::TPT;      // transform within the group, block of LINES lines
        c3 = 0;
        c2 = (type == 1 && in) ? (grp % S) * SH::UNIT + tl : grp * SH::UNIT + tl;
        if (type == 0) { c0 = 2 * SH::LINES_A * q; c1 = (LA / 2) * h; }
        else { c0 = 2 * SH::LINES_B * q; c1 = (LB / 2) * h; }
    } else {
        const int tbl = c / LB, r = c % LB;               // column block within the group; n2 (pass 1) or k1 group (pass 2)
        const int mat = grp >> bshift, gl = grp & (int)((1u << bshift) - 1u);      // matrix of the launch, phase within the matrix
        if (type == 0) { c0 = 2 * (gl * SH::UNIT + tbl * SH::LINES_A); c1 = r; c2 = (LA / 2) * h; c3 = mat; }
        else if (in) { c0 = 0; c1 = r * SH::RA; c2 = (LB / 2) * h; c3 = (grp % S) * SH::TBP + tbl; }
        else { c0 = 2 * (gl * SH::UNIT + tbl * SH::LINES_A); c1 = r * SH::RA; c2 = (LB / 2) * h; c3 = mat; }
    }
}

template <int LA, int LB, int MODE, bool INV, bool PROF, int TW2 = 0>
__global__ void __launch_bounds__(TMA_THREADS, 1)
fft_tma14_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_int,
                 const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ Tma14Params a,
                 const __grid_constant__ std::conditional_t<TW2 == 2, T14Peers, int> peers) {
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
    using SH = T14Shape<LA, LB>;
    constexpr int TPT = 256;                               // tiles per phase
    constexpr int HALF_ELEMS = T14_HALF_BYTES / 16;        // 2048

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
                    t14_coords<LA, LB, MODE>(w.type, w.tf, w.c, h, S, a.bshift, c0, c1, c2, c3, true);
                    if (MODE == T14_ROWS && a.seg > 0 && w.type == 0) { c3 = c2; c1 = 0; c2 = (a.seg / 2) * h; }   // rows (segment, row in segment)
                    if constexpr (LA == 1024) {
                        // a box has at most 256 rows: the 512 rows of a pass-1 half arrive as two copies on the same barrier
                        if (w.type == 0) {
                            tma_load_4d(land + (size_t)s * HALF_ELEMS, &tm_x, c0, c1, c2, c3, fb);
                            tma_load_4d(land + (size_t)s * HALF_ELEMS + HALF_ELEMS / 2, &tm_x, c0, c1 + 256, c2, c3, fb);
                            continue;
                        }
                    }
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
                long long t0 = 0;
                if (PROF) t0 = clock64();
                mbar_wait(staged + g, ns & 1);
                ns++;
                if (PROF) { const long long t1 = clock64(); c_staged += t1 - t0; t0 = t1; }
                const cpx* srcb = work + (size_t)g * T14_WELEMS;
                if (pi.type == 1) {
                    int c0, c1, c2, c3;
                    if constexpr (TW2 == 2) {
                        // rows k2 of the tile go to rank k2 / (LB / G), local row k2 % (LB / G): boxes of min(LB / 2, LB / G) rows
                        const int G = a.npeer, per = LB / G, rows = per < LB / 2 ? per : LB / 2, nbox = (LB / 2) / rows;
#pragma unroll 1
                        for (int hf = 0; hf < 2; hf++) {
                            t14_coords<LA, LB, MODE>(1, pi.tf, pi.c, hf, S, a.bshift, c0, c1, c2, c3, false);
#pragma unroll 1
                            for (int jq = 0; jq < nbox; jq++) {
                                const int jb = (jq + a.prot) % nbox;
                                const int k2 = c2 + jb * rows;
                                tma_store_4d(&peers.m[k2 / per], c0, c1, k2 % per, c3, srcb + (size_t)hf * HALF_ELEMS + (size_t)jb * rows * (4096 / LB));
                            }
                            tma_commit();
                        }
                    } else {
                        t14_coords<LA, LB, MODE>(1, pi.tf, pi.c, 0, S, a.bshift, c0, c1, c2, c3, false);
                        tma_store_4d(&tm_out, c0, c1, c2, c3, srcb);
                        tma_commit();
                        t14_coords<LA, LB, MODE>(1, pi.tf, pi.c, 1, S, a.bshift, c0, c1, c2, c3, false);
                        tma_store_4d(&tm_out, c0, c1, c2, c3, srcb + HALF_ELEMS);
                        tma_commit();
                    }
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
                        // Int[t][n2 = LINES_A q + ell][k1]: LINES_A rows of LA elements, contiguous in memory, pitch LA + 1 in shared memory
                        cpx* dst = slot + (size_t)(pi.c / SH::TPT) * SH::N + (size_t)(pi.c % SH::TPT) * (SH::LINES_A * LA);
#pragma unroll 1
                        for (int l = 0; l < SH::LINES_A / 2; l++) bulk_store_1d_hint(dst + l * LA, srcb + l * SH::ROWPITCH, LA * 16, pol_last);
                        tma_commit();
#pragma unroll 1
                        for (int l = SH::LINES_A / 2; l < SH::LINES_A; l++) bulk_store_1d_hint(dst + l * LA, srcb + l * SH::ROWPITCH, LA * 16, pol_last);
                        tma_commit();
                    } else {
                        // Int[tb][n2][k1][LINES_A t]: the tile is 64 KiB contiguous
                        cpx* dst = slot + (size_t)pi.c * 4096;
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
    const int tig = tid & (TMA_GROUP - 1);
    cpx* wbuf = work + (size_t)g * T14_WELEMS;
    unsigned nrd = 0, nst = 0;
    bool prev_staged = false;                               // the previous tile left staged output in the work buffer
    unsigned fph = 0;
    long long c_full0_p1 = 0, c_full0_p2 = 0, c_full1 = 0, c_drain = 0, c_rd = 0, c_bar = 0, n_tiles = 0, t_start = 0, t0 = 0;
    if (PROF) t_start = clock64();

    // one tile of sub-line length LEN; wi.type is a compile-time constant after inlining when LA != LB
    auto tile = [&](auto len_c, const TmaItem wi, const int s0, const int s1) {
        constexpr int LEN = decltype(len_c)::value;
        using SL = T14Len<LEN>;
        constexpr int LINES = SL::LINES, NJ = SL::NJ, KB = SL::KB;
        // line = low bits of the thread index, j = the rest: LEN = 128: j = warp (uniform); 256: 2 warp + (lane >> 4); ...
        const int ell = tig & (LINES - 1), j = tig / LINES;
        const unsigned ld_conj = (INV && wi.type == 0) ? 0x80000000u : 0u;
        cpx x[32];
        {   // points j + NJ i of line ell: rows j + NJ i of the tile, i < 16 in the first half
            const cpx* s = land + (size_t)s0 * HALF_ELEMS + j * LINES + ell;
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = INV ? cconj_if(s[i * NJ * LINES], ld_conj) : s[i * NJ * LINES];
        }
        mbar_arrive(freed_h + s0);
        if (PROF) t0 = clock64();
        mbar_wait(full_h + 2 * s1 + g, (fph >> s1) & 1);
        fph ^= 1u << s1;
        if (PROF) c_full1 += clock64() - t0;
        if (wi.type == 1 && tig == 0) red_relaxed_gpu(a.done2 + wi.tf, 1);
        {
            const cpx* s = land + (size_t)s1 * HALF_ELEMS + j * LINES + ell;
#pragma unroll
            for (int i = 0; i < 16; i++) x[16 + i] = INV ? cconj_if(s[i * NJ * LINES], ld_conj) : s[i * NJ * LINES];
        }
        mbar_arrive(freed_h + s1);
        dft32(x);                                           // Y_j[k] = sum_i x[j + NJ i] w_32^(i k)
        if constexpr (LEN == 128) {
            if (j != 0) {                                   // warp-uniform: w_128^(j k) from the parameter bank
#pragma unroll
                for (int k = 1; k < 32; k++) x[k] = cmul(x[k], a.w128[j - 1][k]);
            }
        } else {
            const cpx wj = __ldg((wi.type == 0 ? a.wla : a.wlb) + j);
            mul_powers32(x, wj);                            // w_LEN^(j k): j differs inside a warp
        }
        if (PROF) t0 = clock64();
        if (prev_staged) mbar_wait(drained + 2 * g, (nst - 1) & 1);
        if (PROF) c_drain += clock64() - t0;
        // exchange: Y_j[k] -> row NJ k + j of the work buffer, column ell; thread (ell, j') then takes rows 32 j' .. 32 j' + 31,
        // i.e. k = KB j' + k_lo, all NJ residues
        {
            cpx* s = wbuf + j * LINES + ell;
#pragma unroll
            for (int k = 0; k < 16; k++) s[k * NJ * LINES] = x[k];                           // rows < LEN / 2
            if (PROF) t0 = clock64();
            if (prev_staged) mbar_wait(drained + 2 * g + 1, (nst - 1) & 1);
            if (PROF) c_drain += clock64() - t0;
#pragma unroll
            for (int k = 16; k < 32; k++) s[k * NJ * LINES] = x[k];
        }
        // four-step twiddle bases of this line (pass 1): w^(n2 * KB j'), w^(n2), w^(32 n2)
        cpx tb0, tb1, tb32;
        if (wi.type == 0) {
            constexpr unsigned NMASK = (unsigned)(SH::N - 1);
            const unsigned n2 = MODE == T14_ROWS ? (unsigned)((wi.c % SH::TPT) * LINES + ell) : (unsigned)(wi.c % LB);
            const unsigned e0 = (n2 * (unsigned)KB * (unsigned)j) & NMASK, e1 = n2 & NMASK, e32 = (n2 * 32u) & NMASK;
            tb0 = cmul(__ldg(a.tw_hi + (e0 >> 12)), __ldg(a.tw_lo + (e0 & 4095u)));
            tb1 = cmul(__ldg(a.tw_hi + (e1 >> 12)), __ldg(a.tw_lo + (e1 & 4095u)));
            tb32 = cmul(__ldg(a.tw_hi + (e32 >> 12)), __ldg(a.tw_lo + (e32 & 4095u)));
            if (INV) tb0 = make_double2(tb0.x * a.scale, tb0.y * a.scale);
        }
        if constexpr (TW2 != 0 && MODE == T14_COLS) {
            if (wi.type == 1 && a.tw2_log2m > 0) {
                // outer twiddle of this line's outputs k = k1 + LA k2, k2 = KB j' + k_lo + 32 m: w^(col (k1 + LA KB j')), w^(col LA), w^(32 col LA)
                const unsigned long long mask = (1ULL << a.tw2_log2m) - 1ULL;
                const unsigned long long col = (unsigned long long)a.tw2_col0 +
                                               (unsigned long long)((wi.tf & (int)((1u << a.bshift) - 1u)) * SH::UNIT + (wi.c / LB) * SH::LINES_A + (ell % SH::LINES_A));
                const unsigned long long k1 = (unsigned long long)((wi.c % LB) * SH::RA + ell / SH::LINES_A);
                const double sc2 = 2.0 / (double)(1ULL << a.tw2_log2m);
                double sn, cs;
                sincospi((double)((col * (k1 + (unsigned long long)(LA * KB * j))) & mask) * sc2, &sn, &cs);
                tb0 = make_double2(cs, -sn);
                sincospi((double)((col * (unsigned long long)LA) & mask) * sc2, &sn, &cs);
                tb1 = make_double2(cs, -sn);
                sincospi((double)((col * (unsigned long long)(32 * LA)) & mask) * sc2, &sn, &cs);
                tb32 = make_double2(cs, -sn);
            }
        }
        if (PROF) t0 = clock64();
        group_bar(1 + g);
        if (PROF) c_bar += clock64() - t0;
        {
            const cpx* s = wbuf + (32 * j) * LINES + ell;
#pragma unroll
            for (int r = 0; r < 32; r++) x[r] = s[r * LINES];                                // x[NJ k_lo + jj] = Y_jj[KB j + k_lo]
        }
        mbar_arrive(rd + g);
        if constexpr (NJ == 32) dft32(x);
        else {
#pragma unroll
            for (int kl = 0; kl < KB; kl++) dft<NJ, 1>(&x[NJ * kl]);                         // x[NJ k_lo + m] = X[KB j + k_lo + 32 m]
        }
        if (PROF) t0 = clock64();
        mbar_wait(rd + g, nrd & 1);                         // every gather of this tile is done: the buffer may be overwritten
        nrd++;
        if (PROF) c_rd += clock64() - t0;
        if (wi.type == 0) {
            // x[NJ k_lo + m] *= w^(n2 (KB j + k_lo + 32 m)) = tb0 * tb1^k_lo * tb32^m
            if constexpr (NJ == 32) mul_geometric32(x, tb0, tb32);
            else {
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
            }
            if constexpr (MODE == T14_ROWS) {
                cpx* s = wbuf + ell * SH::ROWPITCH + KB * j; // Int[n2 = line][k1 = KB j + k_lo + 32 m]
#pragma unroll
                for (int kl = 0; kl < KB; kl++)
#pragma unroll
                    for (int m = 0; m < NJ; m++) s[kl + 32 * m] = x[NJ * kl + m];
            } else {
                cpx* s = wbuf + (KB * j) * LINES + ell;      // Int[k1][LINES t]: row k1, column = line
#pragma unroll
                for (int kl = 0; kl < KB; kl++)
#pragma unroll
                    for (int m = 0; m < NJ; m++) s[(kl + 32 * m) * LINES] = x[NJ * kl + m];
            }
        } else {
            if (TW2 != 0 && MODE == T14_COLS && a.tw2_log2m > 0) {       // tw2_log2m = 0 with TW2 = 2: peer stores without a twiddle (FFT2)
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
            }
            if constexpr (TW2 == 3 && MODE == T14_ROWS) {
                // k = k1 + LA k2, k1 = LINES (tile of the transform) + line, k2 = KB j' + k_lo + 32 m
                const cpx* ap = a.aux + (size_t)((wi.c % SH::TPT) * LINES + ell) + (size_t)LA * (KB * j);
#pragma unroll
                for (int kl = 0; kl < KB; kl++)
#pragma unroll
                    for (int m = 0; m < NJ; m++) x[NJ * kl + m] = cmul(x[NJ * kl + m], __ldg(ap + (size_t)LA * (kl + 32 * m)));
            }
            cpx* s = wbuf + (KB * j) * LINES + ell;          // X[k2 = KB j + k_lo + 32 m]: row k2 of the tile, column = line
#pragma unroll
            for (int kl = 0; kl < KB; kl++)
#pragma unroll
                for (int m = 0; m < NJ; m++) {
                    const cpx v = x[NJ * kl + m];
                    s[(kl + 32 * m) * LINES] = INV ? make_double2(v.x, -v.y) : v;
                }
        }
        fence_proxy_async();
        mbar_arrive(staged + g);
        nst++;
        prev_staged = true;
    };

    for (int it = g;; it += 2) {
        const long long h0 = 2LL * it;
        const int s0 = (int)(h0 % TMA_NSLOT), s1 = (int)((h0 + 1) % TMA_NSLOT);
        if (PROF) t0 = clock64();
        mbar_wait(full_h + 2 * s0 + g, (fph >> s0) & 1);
        fph ^= 1u << s0;
        const int item = log[it & 31];
        if (item < 0) break;
        TmaItem wi = tma_decode(item, B, D);
        if (PROF) { const long long dt = clock64() - t0; if (wi.type == 0) c_full0_p1 += dt; else c_full0_p2 += dt; n_tiles++; }
        if constexpr (LA == LB) tile(std::integral_constant<int, LA>{}, wi, s0, s1);
        else if (wi.type == 0) { wi.type = 0; tile(std::integral_constant<int, LA>{}, wi, s0, s1); }
        else { wi.type = 1; tile(std::integral_constant<int, LB>{}, wi, s0, s1); }
    }
    if (PROF && tig == 0) {
        long long* q = a.prof + (size_t)blockIdx.x * TMA_PROF_SLOTS + 8 * g;
        q[0] = c_full0_p1; q[1] = c_full0_p2; q[2] = c_full1; q[3] = c_drain; q[4] = c_rd; q[5] = c_bar; q[6] = clock64() - t_start; q[7] = n_tiles;
    }
}

}  // namespace gd

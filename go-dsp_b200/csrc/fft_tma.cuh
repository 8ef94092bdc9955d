// fft_tma.cuh -- batched 2^20-point complex128 transforms (BASELINE config C3): both passes of the N = 1024 x 1024
// four-step in ONE persistent, TMA-fed, warp-specialised kernel; the transposed intermediate never leaves L2.
// Replaces the 20 radix-2 sweeps of fft/radix2.go:131-151 (and its bit-reversal pass, radix2.go:157-168) by
// one HBM read and one HBM write per point.
//
// A tile is four adjacent 1024-point lines in "column" form -- 1024 rows of 64 contiguous bytes, row pitch 16 KiB --
// which is what both passes read once the intermediate is kept transposed:
//     pass 1 (P1): rows n1, columns n2 of x[n1][n2]     -> lines over n1, twiddle w_N^(n2 k1), Int[n2][k1]
//     pass 2 (P2): rows n2, columns k1 of Int[n2][k1]   -> lines over n2,                      X[k1 + 1024 k2]
//
// Work is a fixed global sequence of phases over the transforms of the launch:
//     P1(0) .. P1(D)   P2(0) P1(D+1)   P2(1) P1(D+2) ...   P2(B-1),            each phase = 256 tiles,
// claimed from a device-wide in-order queue. P1(g) writes Int into scratch slot g mod S; a P2(g) tile may be loaded
// once done1[g] = 256 (every P1(g) tile published), and the rows of slot g mod S may be overwritten once
// done2[g - S] = 256 (every P2(g - S) tile has landed in shared memory). Both conditions point at earlier items of
// the sequence, every CTA is resident (one per SM) and works through what it claimed in claim order, so the
// globally earliest unfinished item can always run: no deadlock.
// Default D = 2, S = 3. About 550 tiles are in flight at any time (148 CTAs x (2 in the consumers + 1.5 landing) and
// the publication latency), more than two phases, so the two dependencies are given very different distances:
//   * P1(g) -> P2(g): 2 D = 4 phases (1024 tiles). With D = 1 (2 phases) every loader reached the first P2(g) tile
//     before the last P1(g) tile was published and the consumers waited a quarter of the time.
//   * P2(g) -> P1(g + S): 2 (S - D) - 2 = 0 phases. That is enough because the two ends are asymmetric: a P2 tile has
//     landed ~0.75 tile times after its claim, while a P1 tile needs its slot only when its OUTPUT is stored, ~1.75
//     tile times after its claim. The test is therefore made by the storer just before it overwrites the rows, not
//     by the loader, and S stays 3: the intermediate (48 MiB) keeps fitting the L2 set-aside.
// The S scratch slots (16 MiB each) sit under a persisting L2 access-policy window set by the host.
//
// Roles (384 threads): warps 0-3 / 4-7 = two consumer groups, 32 points per thread, two radix-32 steps
// (generated FMA-form codelet, fft_codelets.cuh) with ONE shared-memory exchange per line; warps 8-11 = producer
// warpgroup trimmed to 40 registers (setmaxnreg), four working lanes: the loader (claims pairs of adjacent tiles,
// issues cp.async.bulk.tensor loads), one storer per consumer group (bulk stores of the staged tiles, device-scope
// publication of finished P1 tiles once cp.async.bulk.wait_group says the writes are complete) and the watcher
// (polls done1 transform by transform and does the acquire fences, so that the loader -- whose own loads in flight a
// device-scope fence would have to wait for, about 7000 cycles per transform -- only reads a shared-memory flag).
//   landing   3 x 32 KiB   a tile lands in halves (512 rows = two 8-double x 256-row boxes) and is copied to
//                          registers at once, so a slot is busy only from issue to landing
//   work      2 x 64 KiB   per group: the exchange between the two radix-32 steps, then staging of the output
// Barriers: full[slot][group] (tx bytes), freed[slot] (128), rd[g] (128: gathers done), staged[g] (128),
// drained[g][half] (1: staged output read out of shared memory).
#pragma once
#include <cuda.h>
#include "fft_w32.cuh"

namespace gd {

constexpr int TMA_T = 4;                          // lines per tile
constexpr int TMA_L = 1024;
constexpr int TMA_BOX_ROWS = 256;                 // TMA box: 8 doubles x 256 rows
constexpr int TMA_GROUP = 128;                    // consumer threads per group
constexpr int TMA_THREADS = 3 * TMA_GROUP;        // + a producer warpgroup: the register file is per scheduler, so a 9th warp
                                                  // would cap everyone at 168 registers; a whole warpgroup can hand its
                                                  // registers to the consumers with setmaxnreg
constexpr int TMA_TILE_BYTES = TMA_T * TMA_L * 16;            // 65536
constexpr int TMA_HALF_BYTES = TMA_TILE_BYTES / 2;            // 32768
constexpr int TMA_NSLOT = 3;
// a work buffer also stages pass-1 output as 4 rows of 1024 + 2 elements (the 32-byte skew makes the 4-lines x
// 2-residues store pattern conflict-free): 65664 bytes, a multiple of 128
constexpr int TMA_ROWLINE = TMA_L + 2;
constexpr int TMA_WBYTES = TMA_T * TMA_ROWLINE * 16;
constexpr int TMA_WELEMS = TMA_WBYTES / 16;
constexpr int TMA_SMEM = TMA_NSLOT * TMA_HALF_BYTES + 2 * TMA_WBYTES + 1024;      // 230656
constexpr int TMA_PROF_SLOTS = 32;                // long long counters per CTA (profiling instantiation only)

struct TmaFusedParams {
    int batch;                   // transforms in this launch (<= 128: one tensor map covers the batch)
    int delay;                   // D
    int nslots;                  // S
    cpx* scratch;                // S slots of 2^20 elements
    int* done1;                  // [batch] zeroed by the host
    int* done2;                  // [batch]
    int* queue;                  // next item of the sequence (zeroed by the host)
    const cpx* wl;               // exp(-2 pi i p / 1024)
    const cpx* tw_lo;            // four-step twiddle tables: w_N^e = hi[e >> 12] * lo[e & 4095]
    const cpx* tw_hi;
    int tw_log2m;
    double scale;                // inverse: 1/N, folded into the four-step twiddle (exact: a power of two)
    int opt;                     // measurement switch: bit 1 = claim single tiles instead of pairs
    long long* prof;             // PROF instantiation: [gridDim.x][TMA_PROF_SLOTS] cycle counters
};

// ---- mbarrier / TMA primitives (PTX ISA 8.x, sm_90+) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, int c0, int c1, int c2, const void* src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];\n"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void bulk_store_1d_hint(void* gdst, const void* ssrc, unsigned bytes, unsigned long long pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\n"
                 ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void group_bar(int id) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "n"(TMA_GROUP) : "memory"); }
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_gpu(int* p, int v) {
    asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// the intermediate is written once and read once shortly after: evict-last keeps it under the persisting window
// (without the hint, bulk stores into the window crawl: 73 instead of 106 GS/s)
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ int ld_volatile_shared(const volatile int* p) { return *p; }

struct TmaItem { int type, tf, c; };

// item gi of the global sequence -> (pass, transform, tile)
__device__ __forceinline__ TmaItem tma_decode(long long gi, int B, int D) {
    constexpr int TPT = TMA_L / TMA_T;
    const int f = (int)(gi / TPT);
    TmaItem it;
    it.c = (int)(gi % TPT);
    if (B <= D + 1) {
        if (f < B) { it.type = 0; it.tf = f; } else { it.type = 1; it.tf = f - B; }
    } else if (f <= D) { it.type = 0; it.tf = f; }
    else {
        const int m = f - D - 1, npairs = B - D - 1;
        if (m < 2 * npairs) {
            if (m & 1) { it.type = 0; it.tf = D + 1 + (m >> 1); } else { it.type = 1; it.tf = m >> 1; }
        } else { it.type = 1; it.tf = npairs + (m - 2 * npairs); }
    }
    return it;
}

// INV: the inverse transform as conj . forward . conj with 1/N folded into the four-step twiddle (fft/fft.go:35-52; exact
// for N = 2^20, where x/N and x * 2^-20 are the same bits). PROF: cycle counters, tools/exp_tma_prof.py only.
template <bool INV, bool PROF>
__global__ void __launch_bounds__(TMA_THREADS, 1)
fft_tma_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_int,
                     const __grid_constant__ CUtensorMap tm_out, const TmaFusedParams a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    cpx* land = reinterpret_cast<cpx*>(smem_raw);
    cpx* work = reinterpret_cast<cpx*>(smem_raw + TMA_NSLOT * TMA_HALF_BYTES);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + TMA_NSLOT * TMA_HALF_BYTES + 2 * TMA_WBYTES);
    // full is per (slot, consumer group): a parity wait is only safe for a waiter that observes EVERY phase of its
    // barrier in order. With one barrier per slot, group B could reach its wait for phase k+1 while phase k -- a half
    // of group A's tile, issued earlier but landing later (HBM vs L2) -- was still open; the parity test then
    // succeeds at once and B reads a slot that has not landed.
    unsigned long long* full_h = bars;                     // [3 slots][2 groups]
    unsigned long long* freed_h = bars + 6;                // [3]
    unsigned long long* rd = bars + 9;                     // [2]
    unsigned long long* staged = bars + 11;                // [2]
    unsigned long long* drained = bars + 13;               // [2 groups][2 halves of the work buffer]
    volatile int* log = reinterpret_cast<volatile int*>(bars + 22);        // [32] item id by local step (-1 = no more work)
    volatile int* log_count = reinterpret_cast<volatile int*>(bars + 38);  // local steps published by the loader
    volatile int* ready_sh = reinterpret_cast<volatile int*>(bars + 39);   // transforms < *ready_sh are published and acquired (watcher)
    constexpr int TPT = TMA_L / TMA_T;
    constexpr int HALF_ELEMS = TMA_HALF_BYTES / 16;        // 2048
    constexpr int TILE_ELEMS = TMA_WELEMS;                 // work buffer pitch (4104 elements)

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
    long long* prof = PROF ? a.prof + (size_t)blockIdx.x * TMA_PROF_SLOTS : nullptr;

    if (warp >= 2 * TMA_GROUP / 32) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (tid == 2 * TMA_GROUP) {
            // ------------------------------------------------------------ loader
            int tokens = 0, ready_tf = -1;                  // transforms known to be published
            long long hidx = 0;                             // halves issued so far
            long long c_claim = 0, c_d1 = 0, c_d2 = 0, c_freed = 0, n_d1wait = 0, t_start = 0;
            if (PROF) t_start = clock64();
            // Tiles are claimed in sequence order, two adjacent tiles per atomic (one for each consumer group: the two
            // 64-byte halves of every 128-byte line are then fetched by one SM at about the same time). The next pair is
            // claimed right after the last load of this one has been issued and is first looked at after the wait for a
            // free landing slot, so the atomic's round trip (~1100 cycles) is off the path without claiming a whole pair
            // early: every tile claimed but not yet published delays the first P2 tile of its transform.
            const bool pairs = !(a.opt & 2);
            int cur = atomicAdd(a.queue, pairs ? 2 : 1);
            for (int it = 0; tokens < 2; it++) {
                long long t0 = 0;
                if (PROF) t0 = clock64();
                // the landing slot of this tile's first half (whatever tile it will be)
                if (hidx >= TMA_NSLOT) mbar_wait(freed_h + (int)(hidx % TMA_NSLOT), (unsigned)(((hidx - TMA_NSLOT) / TMA_NSLOT) & 1));
                if (PROF) { const long long t1 = clock64(); c_freed += t1 - t0; t0 = t1; }
                int item;
                if (tokens) item = nitems;
                else if (!pairs) item = cur;
                else item = cur + (it & 1);
                const bool token = item >= nitems;
                if (PROF) { const long long t1 = clock64(); c_claim += t1 - t0; t0 = t1; }
                TmaItem w;
                w.type = 0; w.tf = 0; w.c = 0;
                const CUtensorMap* tm = &tm_x;
                int tfc = 0;
                if (!token) {
                    w = tma_decode(item, B, D);
                    if (w.type == 0) {
                        if (PROF) { const long long t1 = clock64(); c_d2 += t1 - t0; t0 = t1; }
                        tfc = w.tf;
                    } else {
                        if (w.tf > ready_tf) {
                            if (PROF && ld_volatile_shared(ready_sh) <= w.tf) n_d1wait++;
                            while (ld_volatile_shared(ready_sh) <= w.tf) __nanosleep(20);
                            __threadfence_block();
                            // the watcher has acquired at device scope; this lane still has to order its own async-proxy
                            // loads after its generic read of the flag (without this fence: one stale tile in 1100 stress runs)
                            asm volatile("fence.proxy.async;\n" ::: "memory");
                            ready_tf = ld_volatile_shared(ready_sh) - 1;
                        }
                        if (PROF) { const long long t1 = clock64(); c_d1 += t1 - t0; t0 = t1; }
                        tm = &tm_int; tfc = w.tf % S;
                    }
                }
                log[it & 31] = token ? -1 : item;
                __threadfence_block();
                *log_count = it + 1;
                if (token) tokens++;
#pragma unroll
                for (int h = 0; h < 2; h++, hidx++) {
                    const int s = (int)(hidx % TMA_NSLOT);
                    if (token && h == 1) continue;          // only the first half of a token is ever looked at (and never freed)
                    if (h == 1) {
                        if (PROF) t0 = clock64();
                        if (hidx >= TMA_NSLOT) mbar_wait(freed_h + s, (unsigned)(((hidx - TMA_NSLOT) / TMA_NSLOT) & 1));
                        if (PROF) c_freed += clock64() - t0;
                    }
                    unsigned long long* fb = full_h + 2 * s + (it & 1);
                    if (token) { mbar_arrive(fb); continue; }
                    mbar_expect_tx(fb, TMA_HALF_BYTES);
#pragma unroll
                    for (int j = 0; j < 2; j++)
                        tma_load_3d(land + (size_t)s * HALF_ELEMS + j * TMA_BOX_ROWS * TMA_T, tm, w.c * 2 * TMA_T,
                                    (2 * h + j) * TMA_BOX_ROWS, tfc, fb);
                }
                if (!tokens && (!pairs || (it & 1))) cur = atomicAdd(a.queue, pairs ? 2 : 1);
            }
            if (PROF) {
                prof[16] = c_claim; prof[17] = c_d1; prof[18] = c_d2; prof[19] = c_freed; prof[20] = clock64() - t_start; prof[21] = n_d1wait;
            }
        } else if (tid == 2 * TMA_GROUP + 96) {
            // ------------------------------------------------------------ watcher: P1 phases complete in transform order
            for (int tf = 0; tf < B; tf++) {
                while (ld_relaxed_gpu(a.done1 + tf) < TPT) __nanosleep(64);
                // Acquire with full fences: the P1 rows were written through the async proxy by other CTAs and are read
                // through the async proxy by this CTA's loader. (An acquire load instead of the device-scope fence showed a
                // stale tile about once in 1500 runs of 256 transforms.) This lane has nothing in flight, so the fence is cheap.
                asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
                asm volatile("fence.proxy.async;\n" ::: "memory");
                *ready_sh = tf + 1;
            }
        } else if (tid == 2 * TMA_GROUP + 32 || tid == 2 * TMA_GROUP + 64) {
            // ------------------------------------------------------------ storer of consumer group g
            const int g = tid == 2 * TMA_GROUP + 32 ? 0 : 1;
            const unsigned long long pol_last = policy_evict_last();
            unsigned ns = 0;
            int free_tf = S - 1;                            // transforms whose slot is known to be free
            long long c_staged = 0, c_read = 0, c_all = 0, c_slot = 0;
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
                const cpx* srcb = work + (size_t)g * TILE_ELEMS;
                if (pi.type == 1) {
                    // two bulk groups, rows 0..511 and 512..1023: the group's next exchange may refill the first half of
                    // the work buffer while the second is still being read out
#pragma unroll
                    for (int j = 0; j < 2; j++)
                        tma_store_3d(&tm_out, pi.c * 2 * TMA_T, j * TMA_BOX_ROWS, pi.tf, srcb + j * TMA_BOX_ROWS * TMA_T);
                    tma_commit();
#pragma unroll
                    for (int j = 2; j < 4; j++)
                        tma_store_3d(&tm_out, pi.c * 2 * TMA_T, j * TMA_BOX_ROWS, pi.tf, srcb + j * TMA_BOX_ROWS * TMA_T);
                    tma_commit();
                } else {
                    if (pi.tf > free_tf) {                  // one poll per transform, not per tile
                        // the rows of slot tf mod S may be overwritten once every P2(tf - S) tile has landed in shared memory
                        while (ld_relaxed_gpu(a.done2 + (pi.tf - S)) < TPT) __nanosleep(32);
                        free_tf = pi.tf;
                        if (PROF) { const long long t1 = clock64(); c_slot += t1 - t0; t0 = t1; }
                    }
                    cpx* dst = a.scratch + (size_t)(pi.tf % S) * ((size_t)TMA_L * TMA_L) + (size_t)(pi.c * TMA_T) * TMA_L;
#pragma unroll
                    for (int l = 0; l < 2; l++) bulk_store_1d_hint(dst + (size_t)l * TMA_L, srcb + l * TMA_ROWLINE, TMA_L * 16, pol_last);
                    tma_commit();
#pragma unroll
                    for (int l = 2; l < 4; l++) bulk_store_1d_hint(dst + (size_t)l * TMA_L, srcb + l * TMA_ROWLINE, TMA_L * 16, pol_last);
                    tma_commit();
                }
                tma_wait_read1();
                mbar_arrive(drained + 2 * g);
                tma_wait_read0();
                mbar_arrive(drained + 2 * g + 1);
                if (PROF) { const long long t1 = clock64(); c_read += t1 - t0; t0 = t1; }
                if (pi.type == 0) {
                    // Publication: this lane issued the bulk stores, waits until the writes are complete, and releases
                    // itself -- one thread, no hand-off. (A release by another lane after an mbarrier hand-off, and
                    // register stores fenced by one thread per group, both let P2 tiles read stale rows in the stress test.)
                    tma_wait_all0();
                    asm volatile("fence.proxy.async.global;\n" ::: "memory");
                    red_release_gpu(a.done1 + pi.tf, 1);
                    if (PROF) { const long long t1 = clock64(); c_all += t1 - t0; t0 = t1; }
                }
            }
            tma_wait_all0();
            if (PROF) { prof[24 + 4 * g] = c_staged; prof[25 + 4 * g] = c_read; prof[26 + 4 * g] = c_all; prof[27 + 4 * g] = c_slot; }
        }
        return;
    }

    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    const int g = warp >> 2;
    const int tig = tid & (TMA_GROUP - 1);
    const int ell = tig & (TMA_T - 1), p = tig >> 2;
    const int odd = p & 1;
    cpx* wbuf = work + (size_t)g * TILE_ELEMS;
    cpx w = __ldg(a.wl + p);
    // The exchange: thread p scatters its first-step outputs y_p[r] to row p + 32 r, thread q gathers rows 32 q + j.
    // Lanes pair up as (4 lines) x (2 values of q): rows 32 q + j and 32 (q + 1) + j are 2 KiB apart, a 2-way bank
    // conflict. Rows are therefore placed at row ^ ((row >> 5) & 1): every other block of 32 rows has its row pairs
    // swapped, which moves the second lane group by 64 bytes; scatter and gather stay dense and conflict-free.
    cpx* sc_even = wbuf + p * TMA_T + ell;                   // rows p + 32 r, r even
    cpx* sc_odd = wbuf + (p ^ 1) * TMA_T + ell;              // rows (p ^ 1) + 32 r, r odd
    const cpx* ga_even = wbuf + (32 * p + odd) * TMA_T + ell;  // rows 32 p + (j ^ odd), j even
    const cpx* ga_odd = wbuf + (32 * p - odd) * TMA_T + ell;   //                         j odd
    unsigned nrd = 0, nst = 0;                              // phases of rd[g] waited so far; staged tiles of this group so far
    bool prev_staged = false, first = true;
    unsigned fph = 0;                                       // phase bit of full_h[slot][g], one per slot
    long long c_full0_p1 = 0, c_full0_p2 = 0, c_full1 = 0, c_wbuf = 0, c_rd = 0, n_p1 = 0, n_p2 = 0, t_start = 0;
    if (PROF) t_start = clock64();
    for (int it = g;; it += 2) {
        const long long h0 = 2LL * it;
        const int s0 = (int)(h0 % TMA_NSLOT), s1 = (int)((h0 + 1) % TMA_NSLOT);
        long long t0 = 0;
        if (PROF) t0 = clock64();
        mbar_wait(full_h + 2 * s0 + g, (fph >> s0) & 1);
        fph ^= 1u << s0;
        const int item = log[it & 31];
        if (item < 0) break;
        const TmaItem wi = tma_decode(item, B, D);
        if (PROF) { const long long t1 = clock64(); if (wi.type == 0) { c_full0_p1 += t1 - t0; n_p1++; } else { c_full0_p2 += t1 - t0; n_p2++; } }
        const unsigned ld_conj = (INV && wi.type == 0) ? 0x80000000u : 0u;
        cpx x[32];
        {
            const cpx* s = land + (size_t)s0 * HALF_ELEMS + p * TMA_T + ell;
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = INV ? cconj_if(s[i * 32 * TMA_T], ld_conj) : s[i * 32 * TMA_T];
        }
        mbar_arrive(freed_h + s0);
        if (PROF) t0 = clock64();
        mbar_wait(full_h + 2 * s1 + g, (fph >> s1) & 1);
        fph ^= 1u << s1;
        if (PROF) c_full1 += clock64() - t0;
        if (wi.type == 1 && tig == 0) red_relaxed_gpu(a.done2 + wi.tf, 1);      // this tile of Int has been read
        {
            const cpx* s = land + (size_t)s1 * HALF_ELEMS + p * TMA_T + ell;
#pragma unroll
            for (int i = 0; i < 16; i++) x[16 + i] = INV ? cconj_if(s[i * 32 * TMA_T], ld_conj) : s[i * 32 * TMA_T];
        }
        mbar_arrive(freed_h + s1);
        dft32(x);
        // the work buffer is free once the previous tile's gathers are done or its staged outputs have drained
        if (PROF) t0 = clock64();
        if (!first) {
            if (prev_staged) mbar_wait(drained + 2 * g, (nst - 1) & 1);
            else { mbar_wait(rd + g, nrd & 1); nrd++; }
        }
        if (PROF) c_wbuf += clock64() - t0;
#pragma unroll
        for (int r = 0; r < 16; r++) ((r & 1) ? sc_odd : sc_even)[r * 32 * TMA_T] = x[r];        // rows < 512
        if (PROF) t0 = clock64();
        if (!first && prev_staged) mbar_wait(drained + 2 * g + 1, (nst - 1) & 1);
        if (PROF) c_wbuf += clock64() - t0;
#pragma unroll
        for (int r = 16; r < 32; r++) ((r & 1) ? sc_odd : sc_even)[r * 32 * TMA_T] = x[r];
        first = false;
        cpx t_lo0, t_hi0, t_lo1, t_hi1;
        if (wi.type == 0) {
            const unsigned long long mask = (1ULL << a.tw_log2m) - 1ULL;
            const unsigned long long n2 = (unsigned long long)(wi.c * TMA_T + ell);
            const unsigned long long e0 = (n2 * (unsigned long long)p) & mask, e1 = (n2 * 32ULL) & mask;
            t_lo0 = __ldg(a.tw_lo + (e0 & 4095ULL)); t_hi0 = __ldg(a.tw_hi + (e0 >> 12));
            t_lo1 = __ldg(a.tw_lo + (e1 & 4095ULL)); t_hi1 = __ldg(a.tw_hi + (e1 >> 12));
        }
        group_bar(1 + g);
#pragma unroll
        for (int j = 0; j < 32; j++) x[j] = ((j & 1) ? ga_odd : ga_even)[j * TMA_T];
        mbar_arrive(rd + g);
        // (a 16 KiB table of these powers read through L1 instead of the product chains: 81 instead of 126 GS/s)
        asm volatile("" : "+d"(w.x), "+d"(w.y));
        mul_powers32(x, w);
        dft32(x);
        if (wi.type == 0) {
            cpx t0c = cmul(t_hi0, t_lo0);
            if (INV) t0c = make_double2(t0c.x * a.scale, t0c.y * a.scale);
            mul_geometric32(x, t0c, cmul(t_hi1, t_lo1));
            if (PROF) t0 = clock64();
            mbar_wait(rd + g, nrd & 1);                      // every gather of this tile is done: the buffer may be overwritten
            nrd++;
            if (PROF) c_rd += clock64() - t0;
            cpx* s = wbuf + ell * TMA_ROWLINE + p;           // Int[n2 = 4c + ell][k1 = p + 32 r], rows skewed by 2 elements
#pragma unroll
            for (int r = 0; r < 32; r++) s[r * 32] = x[r];
            fence_proxy_async();
            mbar_arrive(staged + g);
            nst++;
            prev_staged = true;
        } else {
            if (PROF) t0 = clock64();
            mbar_wait(rd + g, nrd & 1);
            nrd++;
            if (PROF) c_rd += clock64() - t0;
            cpx* s = wbuf + p * TMA_T + ell;                 // X[k1 = 4c + ell + 1024 (p + 32 r)]: row p + 32 r of the tile
#pragma unroll
            for (int r = 0; r < 32; r++) s[r * 32 * TMA_T] = INV ? make_double2(x[r].x, -x[r].y) : x[r];
            fence_proxy_async();
            mbar_arrive(staged + g);
            nst++;
            prev_staged = true;
        }
    }
    if (PROF && tig == 0) {
        long long* q = prof + 8 * g;
        q[0] = c_full0_p1; q[1] = c_full0_p2; q[2] = c_full1; q[3] = c_wbuf; q[4] = c_rd; q[5] = clock64() - t_start; q[6] = n_p1; q[7] = n_p2;
    }
}

}  // namespace gd

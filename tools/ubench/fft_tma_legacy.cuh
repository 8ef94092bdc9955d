// fft_tma.cuh -- both passes of the N = 1024 x 1024 four-step as TMA-fed, warp-specialised kernels.
//
// One persistent CTA per SM: a producer warp moves tiles with the tensor-memory accelerator
// (cp.async.bulk.tensor + mbarrier), two consumer groups of four warps transform them. A tile is four
// adjacent 1024-point lines in "column" form -- 1024 rows of 64 contiguous bytes, row pitch 16 KiB -- which is
// what both passes read once the intermediate is kept transposed:
//     pass 1: rows n1, columns n2 of x[n1][n2]      -> lines over n1, twiddle w_N^(n2 k1), Int[n2][k1]
//     pass 2: rows n2, columns k1 of Int[n2][k1]    -> lines over n2,                     X[k1 + 1024 k2]
// (index maps of the four-step in engine.cu; replaces the 20 radix-2 sweeps of fft/radix2.go:131-151).
// The tile lands densely in one 64 KiB shared-memory buffer of a ring of three; each thread holds 32 points
// (two radix-32 steps, fft_w32.cuh) and the exchange between the steps happens IN PLACE in that buffer:
// thread p reads rows p + 32 i, writes its results back to the same slots, and thread q then gathers rows
// 32 q + p'. Lanes pair up as (4 lines) x (2 values of q), so the gather would be a 2-way bank conflict; odd q
// read p' ^ 1 instead (rows one apart = 64 bytes apart) and swap register pairs afterwards.
// Nothing on the load side touches the LSU global path or registers, so the loads of the next tiles are in
// flight for a whole tile time; the store side is either plain 128-byte row stores (pass 1: Int rows are
// contiguous in k1) or a TMA store of the tile staged in place (pass 2).
#pragma once
#include <cuda.h>
#include "fft_w32.cuh"

namespace gd {

constexpr int TMA_T = 4;                          // lines per tile
constexpr int TMA_L = 1024;
constexpr int TMA_NBUF = 3;
constexpr int TMA_TILE_BYTES = TMA_T * TMA_L * 16;   // 65536
constexpr int TMA_BOX_ROWS = 256;                 // TMA box: 8 doubles x 256 rows
constexpr int TMA_GROUP = 128;                    // consumer threads per group
constexpr int TMA_THREADS = 3 * TMA_GROUP;        // + a producer warpgroup (one working lane): the register file is per
                                                  // scheduler, so a 9th warp would cap everyone at 168 registers; a whole
                                                  // warpgroup can hand its registers to the consumers with setmaxnreg
constexpr int TMA_SMEM = TMA_NBUF * TMA_TILE_BYTES + 1024;
// fused kernel: a buffer also stages pass-1 output as 4 lines of 1024 + 2 elements (the 32-byte skew makes the
// 4-lines x 2-residues store pattern conflict-free), 128-byte aligned
constexpr int TMA_ROWLINE = TMA_L + 2;
constexpr int TMA_FBUF_BYTES = ((TMA_T * TMA_ROWLINE * 16 + 1023) / 1024) * 1024;   // 66560
constexpr int TMA_FBUF_ELEMS = TMA_FBUF_BYTES / 16;
constexpr int TMA_FUSED_SMEM = TMA_NBUF * TMA_FBUF_BYTES + 1024;

enum { TMA_OUT_ROWS = 0, TMA_OUT_TILE = 1 };

struct TmaPassParams {
    long long ntiles;            // transforms * 256
    cpx* out;                    // TMA_OUT_ROWS: Int base; line (tf, c*4 + ell) is out + tf*out_dist + (c*4+ell)*1024
    long long out_dist;
    const cpx* wl;               // exp(-2 pi i p / 1024), p < 1024 (32 used)
    const cpx* tw_lo;            // four-step twiddle tables (w_N^e = hi[e >> 12] * lo[e & 4095]); null = no twiddle
    const cpx* tw_hi;
    int tw_log2m;
    int ld_conj;                 // conjugate on load (inverse transforms)
    int st_conj;                 // conjugate on store
    double scale;                // multiply on store (1.0 = none)
    int tf_mask;                 // timing experiments only: transform index & tf_mask (-1 = off) keeps the working set in L2
    int dbg_noload, dbg_nostore; // timing experiments only: skip the TMA loads / every store (results are garbage)
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
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {      // non-blocking
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, int c0, int c1, int c2, const void* src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];\n"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* ssrc, unsigned bytes) {   // contiguous shared -> global
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_1d_hint(void* gdst, const void* ssrc, unsigned bytes, unsigned long long pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;\n"
                 ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void group_bar(int id) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "n"(TMA_GROUP) : "memory"); }

// NOTE: fft_tma_kernel (one pass per launch) and fft_tma_fused_kernel (first fused version) are measurement kernels for
// tools/ubench/tma_pass.cu and are not reachable from the C ABI. The product is fft_tma_fused2_kernel below; in
// particular the first fused version hands P1 tiles to its publisher without a device-scope fence by the storing
// warps and lets helper lanes follow mbarrier phases they do not own, both of which the stress test showed to be unsafe.
template <int OUTMODE>
__global__ void __launch_bounds__(TMA_THREADS, 1)
fft_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, const TmaPassParams a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // buffers first (TMA destinations: 128-byte aligned), barriers after them
    cpx* bufs = reinterpret_cast<cpx*>(smem_raw);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + TMA_NBUF * TMA_TILE_BYTES);
    // full is per (buffer, consumer group): a parity wait is only safe when the waiter observes every phase of its barrier
    unsigned long long* full = bars;              // [NBUF][2] tile landed (tx bytes)
    unsigned long long* freed = bars + 2 * TMA_NBUF;  // [NBUF] OUT_ROWS: inputs consumed; OUT_TILE: outputs staged   (128 arrivals)
    unsigned long long* rd = bars + 3 * TMA_NBUF; // [2] per group: every gather of the tile is done             (128 arrivals)

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < TMA_NBUF; i++) { mbar_init(full + 2 * i, 1); mbar_init(full + 2 * i + 1, 1); mbar_init(freed + i, TMA_GROUP); }
        mbar_init(rd + 0, TMA_GROUP); mbar_init(rd + 1, TMA_GROUP);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const long long my_tiles = (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;   // this CTA's share (>= 0)
    constexpr int TPT = TMA_L / TMA_T;           // tiles per transform: 256

    if (warp >= 2 * TMA_GROUP / 32) {
        // ------------------------------------------------ producer: one elected lane
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (tid == 2 * TMA_GROUP) {
            for (long long it = 0; it < my_tiles; it++) {
                const int b = (int)(it % TMA_NBUF);
                if (it >= TMA_NBUF) {
                    const long long prev = it - TMA_NBUF;
                    mbar_wait(freed + b, (unsigned)((prev / TMA_NBUF) & 1));
                    if (OUTMODE == TMA_OUT_TILE && !a.dbg_nostore) {
                        const long long tile = blockIdx.x + prev * gridDim.x;
                        const int tf = (int)(tile / TPT) & a.tf_mask, c = (int)(tile % TPT);
#pragma unroll
                        for (int j = 0; j < TMA_L / TMA_BOX_ROWS; j++)
                            tma_store_3d(&tm_out, c * 2 * TMA_T, j * TMA_BOX_ROWS, tf, bufs + (size_t)b * (TMA_TILE_BYTES / 16) + j * TMA_BOX_ROWS * TMA_T);
                        tma_commit();
                        tma_wait_read0();
                    }
                }
                const long long tile = blockIdx.x + it * gridDim.x;
                const int tf = (int)(tile / TPT) & a.tf_mask, c = (int)(tile % TPT);
                unsigned long long* fb = full + 2 * b + (int)(it & 1);
                if (a.dbg_noload) { mbar_arrive(fb); continue; }
                mbar_expect_tx(fb, TMA_TILE_BYTES);
#pragma unroll
                for (int j = 0; j < TMA_L / TMA_BOX_ROWS; j++)
                    tma_load_3d(bufs + (size_t)b * (TMA_TILE_BYTES / 16) + j * TMA_BOX_ROWS * TMA_T, &tm_in, c * 2 * TMA_T, j * TMA_BOX_ROWS, tf, fb);
            }
            if (OUTMODE == TMA_OUT_TILE && !a.dbg_nostore) {
                for (long long it = my_tiles > TMA_NBUF ? my_tiles - TMA_NBUF : 0; it < my_tiles; it++) {
                    const int b = (int)(it % TMA_NBUF);
                    mbar_wait(freed + b, (unsigned)((it / TMA_NBUF) & 1));
                    const long long tile = blockIdx.x + it * gridDim.x;
                    const int tf = (int)(tile / TPT) & a.tf_mask, c = (int)(tile % TPT);
#pragma unroll
                    for (int j = 0; j < TMA_L / TMA_BOX_ROWS; j++)
                        tma_store_3d(&tm_out, c * 2 * TMA_T, j * TMA_BOX_ROWS, tf, bufs + (size_t)b * (TMA_TILE_BYTES / 16) + j * TMA_BOX_ROWS * TMA_T);
                    tma_commit();
                }
                tma_wait_all0();
            }
        }
        return;
    }

    // ---------------------------------------------------- consumers: group g takes this CTA's tiles g, g+2, ...
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    const int g = warp >> 2;
    const int tig = tid & (TMA_GROUP - 1);
    const int ell = tig & (TMA_T - 1), p = tig >> 2;             // line in the tile; residue / output-residue index
    const unsigned ld_conj = a.ld_conj ? 0x80000000u : 0u;
    const int odd = p & 1;
    cpx w = __ldg(a.wl + p);                                     // exp(-2 pi i p / 1024)
    unsigned use = 0;                                            // uses of rd[g] so far
    for (long long it = g; it < my_tiles; it += 2, use++) {
        const int b = (int)(it % TMA_NBUF);
        cpx* base = bufs + (size_t)b * (TMA_TILE_BYTES / 16);
        const long long tile = blockIdx.x + it * gridDim.x;
        const int tf = (int)(tile / TPT) & a.tf_mask, c = (int)(tile % TPT);
        mbar_wait(full + 2 * b + g, (unsigned)((it / (2 * TMA_NBUF)) & 1));   // (buffer, group) recurs every 6 steps
        cpx x[32];
        {
            const cpx* s = base + p * TMA_T + ell;
#pragma unroll
            for (int i = 0; i < 32; i++) x[i] = cconj_if(s[i * 32 * TMA_T], ld_conj);       // x[p + 32 i]
        }
        dft32(x);                                                                           // y[p][r]
        {
            cpx* s = base + p * TMA_T + ell;
#pragma unroll
            for (int r = 0; r < 32; r++) s[r * 32 * TMA_T] = x[r];                          // slot p + 32 r (own slots)
        }
        // four-step twiddle operands: fetched here, where no transform registers are live
        cpx t_lo0, t_hi0, t_lo1, t_hi1;
        const bool twiddle = a.tw_lo != nullptr;
        if (twiddle) {
            const unsigned long long mask = (1ULL << a.tw_log2m) - 1ULL;
            const unsigned long long n2 = (unsigned long long)(c * TMA_T + ell);
            const unsigned long long e0 = (n2 * (unsigned long long)p) & mask, e1 = (n2 * 32ULL) & mask;
            t_lo0 = __ldg(a.tw_lo + (e0 & 4095ULL)); t_hi0 = __ldg(a.tw_hi + (e0 >> 12));
            t_lo1 = __ldg(a.tw_lo + (e1 & 4095ULL)); t_hi1 = __ldg(a.tw_hi + (e1 >> 12));
        }
        group_bar(1 + g);
        {
            const cpx* s = base + (32 * p) * TMA_T + ell;
#pragma unroll
            for (int j = 0; j < 32; j++) x[j] = s[(j ^ odd) * TMA_T];                       // y[j ^ odd][q = p]
        }
        if constexpr (OUTMODE == TMA_OUT_TILE) mbar_arrive(rd + g); else mbar_arrive(freed + b);
        if (odd) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) { cpx t = x[j]; x[j] = x[j + 1]; x[j + 1] = t; }
        }
        // keep the compiler from hoisting the 31 powers of w out of the tile loop (124 registers -> spills)
        asm volatile("" : "+d"(w.x), "+d"(w.y));
        mul_powers32(x, w);
        dft32(x);                                                                           // X[p + 32 s]
        if (twiddle) mul_geometric32(x, cmul(t_hi0, t_lo0), cmul(t_hi1, t_lo1));
        if (a.st_conj || a.scale != 1.0) {                     // uniform; the forward transform stores as is
            const double sx = a.scale, sy = a.st_conj ? -a.scale : a.scale;
#pragma unroll
            for (int s = 0; s < 32; s++) x[s] = make_double2(x[s].x * sx, x[s].y * sy);
        }
        if constexpr (OUTMODE == TMA_OUT_ROWS) {
            cpx* dst = a.out + (long long)tf * a.out_dist + (long long)(c * TMA_T + ell) * TMA_L + p;
            if (a.dbg_nostore) { if (x[0].x == 1.2345e-300) dst[0] = x[31]; }
            else {
#pragma unroll
                for (int s = 0; s < 32; s++) dst[32 * s] = x[s];
            }
        } else {
            mbar_wait(rd + g, use & 1);                        // every thread of the group has gathered: slots may be overwritten
            cpx* s = base + p * TMA_T + ell;
#pragma unroll
            for (int r = 0; r < 32; r++) s[r * 32 * TMA_T] = x[r];                          // row p + 32 r
            fence_proxy_async();
            mbar_arrive(freed + b);
        }
    }
}


// =====================================================================================================
// Both passes in ONE persistent launch, with the transposed intermediate kept in L2.
//
// Work is a fixed global sequence of phases over the transforms of the batch (a "group" is one transform,
// 16 MiB of intermediate):  P1(0) .. P1(D)  P2(0) P1(D+1)  P2(1) P1(D+2) ...  P2(B-1),  each phase = 256
// tiles; item i of the sequence belongs to CTA i mod gridDim.x, and every CTA walks its items in order.
// P1(g) writes Int into scratch slot g mod S (S = D + 2); P2(g) may load a tile once done1[g] says all 256
// P1(g) tiles are published, and P1(g) may overwrite a slot once done2[g - S] says all 256 P2(g - S) tiles
// have been read. Both conditions point at earlier items of the sequence and all CTAs are co-resident (one
// per SM), so the globally earliest unfinished item can always run: no deadlock. The producer lane does the
// polling before it issues a tile's TMA load, so consumers never spin on global memory.
// The scratch slots (S x 16 MiB) sit under a persisting L2 access-policy window set by the host; HBM then
// sees the input once and the output once: 32 B per point, the algorithmic minimum.
struct TmaFusedParams {
    int batch;                   // transforms in this launch (< 256: one tensor map covers the batch)
    int delay;                   // D
    int nslots;                  // S = D + 2
    cpx* scratch;                // S slots of 2^20 elements
    int* done1;                  // [batch] zeroed by the host
    int* done2;                  // [batch]
    int* queue;                  // next item of the sequence; two_queues: queue[0] = next P1 tile, queue[1] = next P2 tile (zeroed by the host)
    int two_queues;
    int dbg_acqload;             // measurement only: acquire with a load instead of fence.acq_rel.gpu
    int dbg_nosplit;             // measurement only: free the whole work buffer at once after a P2 store
    int dbg_wproxy;              // measurement only: writer-side generic->async proxy fence after the P1 stores
    int dbg_nopubfence;          // measurement only: drop the storing warps' device-scope fence (then results are occasionally wrong)
    int dbg_out_alias, dbg_in_alias;   // measurement only: mask of transform-index bits cleared for the output / input (keeps them in L2)
    int dbg_nodeps;              // timing experiments only: ignore the global dependencies (results are garbage)
    int dbg_nop1st;              // timing experiments only: skip the pass-1 stores
    int dbg_nop2st, dbg_noload;  // timing experiments only: skip the pass-2 tile stores / the tile loads
    const cpx* wl;
    const cpx* tw_lo;
    const cpx* tw_hi;
    int tw_log2m;
    int ld_conj, st_conj;
    double scale;
    long long* stats;            // optional [gridDim.x][8] cycle counters (tools/ubench/tma_pass.cu); null in the product
    int hints;                   // L2 eviction-priority hints: 1 = Int bulk stores evict-last, 2 = x loads evict-first,
                                 // 4 = Int loads evict-last, 8 = output stores evict-first
    int p2_stg;                  // pass 2 stores straight from registers (64-byte chunks) instead of staging a TMA store:
                                 // the tile buffer is free again right after the gather
    cpx* out;                    // p2_stg: output base, transform tf at out + tf * out_dist
    long long out_dist;
};

struct TmaItem { int type, tf, c; };

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
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
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

// L2 eviction-priority hints (createpolicy): the input is read once and the output written once (evict first),
// so they do not displace the intermediate, which is written once and read once shortly after (evict last)
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_3d_hint(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, unsigned long long* bar,
                                                 unsigned long long pol) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;\n"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* tm, int c0, int c1, int c2, const void* src, unsigned long long pol) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;\n"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src)), "l"(pol) : "memory");
}

// Roles: warps 0-3 / 4-7 = consumer groups; warp 8 = loader (claims items from the device-wide queue, waits for
// their global dependencies, issues the TMA loads), warp 9 = storer (TMA stores of staged P2 tiles), warp 10 =
// publisher (device-scope release of finished P1 tiles); one working lane each, all off the consumers' path.
// Items are claimed in sequence order with one atomicAdd each, so a CTA that falls behind simply takes fewer
// tiles: with a static round-robin every phase ran at the pace of the slowest of the 148 CTAs. A CTA processes
// the items it claimed in claim order, which keeps the no-deadlock argument above intact.
//   full[b]   (1 + tx)  tile of local step `it` (buffer it % 3) landed; log[it & 31] holds its item id (-1 = no more work)
//   freed[b]  (128)     P1: inputs gathered, buffer reusable; P2: outputs staged in the buffer     consumers -> loader/storer/publisher
//   empty[b]  (1)       the staged P2 tile has been read out of shared memory                      storer -> loader
//   rd[g]     (128)     every thread of group g has gathered (slots may be overwritten)            group-internal
//   pd[g][j]  (128)     the group's stores of a P1 tile have been issued (j alternates)            consumers -> publisher
__global__ void __launch_bounds__(TMA_THREADS, 1)
fft_tma_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_int,
                     const __grid_constant__ CUtensorMap tm_out, const TmaFusedParams a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    cpx* bufs = reinterpret_cast<cpx*>(smem_raw);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + TMA_NBUF * TMA_TILE_BYTES);
    unsigned long long* full = bars;                       // [NBUF][2 groups]: see fft_tma_fused2_kernel for why per group
    unsigned long long* freed = bars + 2 * TMA_NBUF;
    unsigned long long* empty = bars + 3 * TMA_NBUF;
    unsigned long long* rd = bars + 4 * TMA_NBUF;
    unsigned long long* pd = bars + 4 * TMA_NBUF + 2;      // [2 groups][2]
    volatile int* log = reinterpret_cast<volatile int*>(bars + 20);   // [32] item ids by local step
    volatile long long* tissue = reinterpret_cast<volatile long long*>(bars + 36);   // [32] load issue times (stats only)
    constexpr int TPT = TMA_L / TMA_T;

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < TMA_NBUF; i++) { mbar_init(full + 2 * i, 1); mbar_init(full + 2 * i + 1, 1); mbar_init(freed + i, TMA_GROUP); mbar_init(empty + i, 1); }
        mbar_init(rd + 0, TMA_GROUP); mbar_init(rd + 1, TMA_GROUP);
        for (int i = 0; i < 4; i++) mbar_init(pd + i, TMA_GROUP);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int nitems = 2 * a.batch * TPT;                  // host keeps this below 2^31
    const int B = a.batch, D = a.delay, S = a.nslots;

    if (warp >= 2 * TMA_GROUP / 32) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (tid == 2 * TMA_GROUP) {
            // ------------------------------------------------------------ loader
            const unsigned long long pol_first = policy_evict_first(), pol_last = policy_evict_last();
            long long c_buf = 0, c_dep1 = 0, c_dep2 = 0, ntaken = 0;
            const long long t_begin = clock64();
            unsigned p2uses = 0;                            // 2 bits per buffer would do; kept as 3 x 8-bit counters
            unsigned prev_p2 = 0;                           // bit b: the tile last put in buffer b was a P2 tile
            int tokens = 0, ready_tf = -1;
            for (long long it = 0; tokens < 2; it++) {
                const int b = (int)(it % TMA_NBUF);
                // claim the next item and settle its dependencies first, the buffer wait comes last: the queue
                // round trip, the polling and the fences then overlap the time the ring is full anyway
                const long long t0 = clock64();
                const int item = tokens ? nitems : atomicAdd(a.queue, 1);
                const bool token = item >= nitems;
                TmaItem w;
                w.type = 0; w.tf = 0; w.c = 0;
                const CUtensorMap* tm = &tm_x;
                int tfc = 0;
                if (!token) {
                    ntaken++;
                    w = tma_decode(item, B, D);
                    if (w.type == 0) {
                        if (w.tf >= S && !a.dbg_nodeps) { while (ld_relaxed_gpu(a.done2 + (w.tf - S)) < TPT) __nanosleep(32); }
                        tm = &tm_x; tfc = w.tf;
                        c_dep2 += clock64() - t0;
                    } else {
                        if (w.tf != ready_tf) {             // one poll + fence pair per transform, not per tile
                            if (!a.dbg_nodeps) { while (ld_relaxed_gpu(a.done1 + w.tf) < TPT) __nanosleep(32); }
                            asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
                            asm volatile("fence.proxy.async;\n" ::: "memory");  // other CTAs' generic-proxy stores -> this async-proxy read
                            ready_tf = w.tf;
                        }
                        tm = &tm_int; tfc = w.tf % S;
                        c_dep1 += clock64() - t0;
                    }
                }
                const long long t2 = clock64();
                if (it >= TMA_NBUF) {
                    if ((prev_p2 >> b) & 1) {
                        mbar_wait(empty + b, (p2uses >> (8 * b)) & 1);
                        p2uses += 1u << (8 * b);
                        p2uses &= ~(0xFEu << (8 * b));      // keep one parity bit per buffer
                    } else mbar_wait(freed + b, (unsigned)(((it - TMA_NBUF) / TMA_NBUF) & 1));
                }
                c_buf += clock64() - t2;
                if (token) {                                // no more work: one token per consumer group
                    log[it & 31] = -1;
                    prev_p2 &= ~(1u << b);
                    mbar_arrive(full + 2 * b + (int)(it & 1));
                    tokens++;
                    continue;
                }
                if (w.type == 0) prev_p2 &= ~(1u << b); else prev_p2 |= 1u << b;
                log[it & 31] = item;
                if (a.stats) tissue[it & 31] = clock64();
                unsigned long long* fb = full + 2 * b + (int)(it & 1);
                mbar_expect_tx(fb, TMA_TILE_BYTES);         // release: the log entry is visible to whoever sees the phase
                cpx* dstb = bufs + (size_t)b * (TMA_TILE_BYTES / 16);
                if (a.hints) {
                    const unsigned long long pol = w.type == 0 ? pol_first : pol_last;
#pragma unroll
                    for (int j = 0; j < TMA_L / TMA_BOX_ROWS; j++)
                        tma_load_3d_hint(dstb + j * TMA_BOX_ROWS * TMA_T, tm, w.c * 2 * TMA_T, j * TMA_BOX_ROWS, tfc, fb, pol);
                } else {
#pragma unroll
                    for (int j = 0; j < TMA_L / TMA_BOX_ROWS; j++)
                        tma_load_3d(dstb + j * TMA_BOX_ROWS * TMA_T, tm, w.c * 2 * TMA_T, j * TMA_BOX_ROWS, tfc, fb);
                }
            }
            if (a.stats) {
                long long* st = a.stats + blockIdx.x * 8;
                st[0] = clock64() - t_begin; st[1] = c_buf; st[3] = c_dep1; st[4] = c_dep2; st[7] = ntaken;
            }
        } else if (tid == 2 * TMA_GROUP + 32) {
            // ------------------------------------------------------------ storer of P2 tiles
            const unsigned long long pol_first = policy_evict_first();
            long long c_drain = 0;
            for (long long it = 0;; it++) {
                const int b = (int)(it % TMA_NBUF);
                mbar_wait(freed + b, (unsigned)((it / TMA_NBUF) & 1));      // every phase is observed, in order
                const int item = log[it & 31];
                if (item < 0) break;
                const TmaItem pi = tma_decode(item, B, D);
                if (pi.type != 1) continue;
                const long long t0 = clock64();
                const cpx* srcb = bufs + (size_t)b * (TMA_TILE_BYTES / 16);
                if (a.hints) {
#pragma unroll
                    for (int j = 0; j < TMA_L / TMA_BOX_ROWS; j++)
                        tma_store_3d_hint(&tm_out, pi.c * 2 * TMA_T, j * TMA_BOX_ROWS, pi.tf, srcb + j * TMA_BOX_ROWS * TMA_T, pol_first);
                } else {
#pragma unroll
                    for (int j = 0; j < TMA_L / TMA_BOX_ROWS; j++)
                        tma_store_3d(&tm_out, pi.c * 2 * TMA_T, j * TMA_BOX_ROWS, pi.tf, srcb + j * TMA_BOX_ROWS * TMA_T);
                }
                tma_commit();
                tma_wait_read0();
                mbar_arrive(empty + b);
                c_drain += clock64() - t0;
            }
            tma_wait_all0();
            if (a.stats) a.stats[blockIdx.x * 8 + 2] = c_drain;
        } else if (tid == 2 * TMA_GROUP + 64) {
            // ------------------------------------------------------------ publisher of P1 tiles
            unsigned np1[2] = {0, 0};
            for (long long it = 0;; it++) {
                const int b = (int)(it % TMA_NBUF);
                mbar_wait(freed + b, (unsigned)((it / TMA_NBUF) & 1));
                const int item = log[it & 31];
                if (item < 0) break;
                const TmaItem pi = tma_decode(item, B, D);
                if (pi.type != 0) continue;
                const int g = (int)(it & 1);
                mbar_wait(pd + 2 * g + (np1[g] & 1), (np1[g] >> 1) & 1);
                np1[g]++;
                red_release_gpu(a.done1 + pi.tf, 1);       // cumulative: publishes the group's stores (ordered by the mbarrier)
            }
        } else if (tid == 2 * TMA_GROUP + 96 && a.stats) {
            // ------------------------------------------------------------ stats only: issue -> landed latency of the tile loads
            long long lat1 = 0, lat2 = 0, n1 = 0, n2 = 0;
            for (long long it = 0;; it++) {
                const int b = (int)(it % TMA_NBUF);
                mbar_wait(full + 2 * b + (int)(it & 1), (unsigned)((it / (2 * TMA_NBUF)) & 1));
                const long long t = clock64();
                const int item = log[it & 31];
                if (item < 0) break;
                const TmaItem pi = tma_decode(item, B, D);
                if (pi.type == 0) { lat1 += t - tissue[it & 31]; n1++; } else { lat2 += t - tissue[it & 31]; n2++; }
            }
            long long* st = a.stats + (size_t)gridDim.x * 8 + blockIdx.x * 4;
            st[0] = lat1; st[1] = n1; st[2] = lat2; st[3] = n2;
        }
        return;
    }

    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    const int g = warp >> 2;
    const int tig = tid & (TMA_GROUP - 1);
    const int ell = tig & (TMA_T - 1), p = tig >> 2;
    const int odd = p & 1;
    cpx w = __ldg(a.wl + p);
    unsigned use = 0, np1 = 0;
    long long c_full = 0;
    for (long long it = g;; it += 2) {
        const int b = (int)(it % TMA_NBUF);
        cpx* base = bufs + (size_t)b * (TMA_TILE_BYTES / 16);
        if (a.stats) {
            const long long t0 = clock64();
            mbar_wait(full + 2 * b + g, (unsigned)((it / (2 * TMA_NBUF)) & 1));
            c_full += clock64() - t0;
        } else mbar_wait(full + 2 * b + g, (unsigned)((it / (2 * TMA_NBUF)) & 1));
        const int item = log[it & 31];
        if (item < 0) { mbar_arrive(freed + b); break; }   // lets the storer and the publisher see the token too
        const TmaItem wi = tma_decode(item, B, D);
        const unsigned ld_conj = (a.ld_conj && wi.type == 0) ? 0x80000000u : 0u;
        if (wi.type == 1 && tig == 0) red_relaxed_gpu(a.done2 + wi.tf, 1);      // this tile of Int has been read
        cpx x[32];
        {
            const cpx* s = base + p * TMA_T + ell;
#pragma unroll
            for (int i = 0; i < 32; i++) x[i] = cconj_if(s[i * 32 * TMA_T], ld_conj);
        }
        dft32(x);
        {
            cpx* s = base + p * TMA_T + ell;
#pragma unroll
            for (int r = 0; r < 32; r++) s[r * 32 * TMA_T] = x[r];
        }
        cpx t_lo0, t_hi0, t_lo1, t_hi1;
        if (wi.type == 0) {
            const unsigned long long mask = (1ULL << a.tw_log2m) - 1ULL;
            const unsigned long long n2 = (unsigned long long)(wi.c * TMA_T + ell);
            const unsigned long long e0 = (n2 * (unsigned long long)p) & mask, e1 = (n2 * 32ULL) & mask;
            t_lo0 = __ldg(a.tw_lo + (e0 & 4095ULL)); t_hi0 = __ldg(a.tw_hi + (e0 >> 12));
            t_lo1 = __ldg(a.tw_lo + (e1 & 4095ULL)); t_hi1 = __ldg(a.tw_hi + (e1 >> 12));
        }
        group_bar(1 + g);
        {
            const cpx* s = base + (32 * p) * TMA_T + ell;
#pragma unroll
            for (int j = 0; j < 32; j++) x[j] = s[(j ^ odd) * TMA_T];
        }
        if (wi.type == 1) mbar_arrive(rd + g); else mbar_arrive(freed + b);
        if (odd) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) { cpx t = x[j]; x[j] = x[j + 1]; x[j + 1] = t; }
        }
        asm volatile("" : "+d"(w.x), "+d"(w.y));
        mul_powers32(x, w);
        dft32(x);
        if (wi.type == 0) {
            // x[r] *= t0 * s1^r (four interleaved chains), each value stored as soon as it is ready
            cpx* dst = a.scratch + (size_t)(wi.tf % S) * ((size_t)TMA_L * TMA_L) + (size_t)(wi.c * TMA_T + ell) * TMA_L + p;
            const cpx t0 = cmul(t_hi0, t_lo0), s1 = cmul(t_hi1, t_lo1);
            const cpx s2 = csqr(s1), s4 = csqr(s2);
            cpx t[4];
            t[0] = t0; t[1] = cmul(t0, s1); t[2] = cmul(t0, s2); t[3] = cmul(t[1], s2);
            if (a.dbg_nop1st) {
                cpx acc = make_double2(0.0, 0.0);
#pragma unroll
                for (int bb = 0; bb < 8; bb++) {
#pragma unroll
                    for (int aa = 0; aa < 4; aa++) {
                        acc = cadd(acc, cmul(x[4 * bb + aa], t[aa]));
                        if (bb < 7) t[aa] = cmul(t[aa], s4);
                    }
                }
                if (acc.x == 1.2345e-300) dst[0] = acc;
            } else {
#pragma unroll
            for (int bb = 0; bb < 8; bb++) {
#pragma unroll
                for (int aa = 0; aa < 4; aa++) {
                    dst[32 * (4 * bb + aa)] = cmul(x[4 * bb + aa], t[aa]);
                    if (bb < 7) t[aa] = cmul(t[aa], s4);
                }
            }
            }
            mbar_arrive(pd + 2 * g + (np1 & 1));               // release.cta: the publisher makes it device-visible
            np1++;
        } else {
            if (a.st_conj || a.scale != 1.0) {
                const double sx = a.scale, sy = a.st_conj ? -a.scale : a.scale;
#pragma unroll
                for (int s = 0; s < 32; s++) x[s] = make_double2(x[s].x * sx, x[s].y * sy);
            }
            mbar_wait(rd + g, use & 1);
            use++;
            cpx* s = base + p * TMA_T + ell;                   // X[k1 = 4c + ell + 1024 (p + 32 r)]: row p + 32 r of the tile
#pragma unroll
            for (int r = 0; r < 32; r++) s[r * 32 * TMA_T] = x[r];
            fence_proxy_async();
            mbar_arrive(freed + b);
        }
    }
    if (a.stats && tig == 0) a.stats[blockIdx.x * 8 + 5 + g] = c_full;
}


// =====================================================================================================
// Second fused kernel: landing slots decoupled from the work buffers.
//
// In fft_tma_fused_kernel a tile occupies one of three 64 KiB buffers from the moment its load is issued until
// its outputs have drained (P2), so at most one tile is ever in flight per CTA and the consumers wait for data
// about a quarter of the time (ncu: the `full` wait is the top stall). Here a tile lands in HALVES (512 rows =
// two TMA boxes) in a ring of three 32 KiB slots and is copied to registers at once -- a slot is busy only
// from issue to landing -- while the exchange and the P2 output staging use a 64 KiB work buffer owned by the
// consumer group. The loader can therefore run a whole tile time ahead of each group.
//   landing   3 x 32 KiB      full_h[s] (1 + tx), freed_h[s] (128)
//   work      2 x 64 KiB      rd[g] (128): gathers done;  staged[g] (128): P2 outputs staged;  drained[g] (1): read out
//   log[32] item ids by local step, log_count (monotonic, written by the loader) for the helper warps
constexpr int TMA2_HALF_BYTES = TMA_TILE_BYTES / 2;          // 32768
constexpr int TMA2_NSLOT = 3;
// a work buffer also stages pass-1 output as 4 rows of 1024 + 2 elements (P1BULK; the 32-byte skew makes the 4-lines x
// 2-residues store pattern conflict-free): 65664 bytes, a multiple of 128
constexpr int TMA2_ROWLINE = TMA_L + 2;
constexpr int TMA2_WBYTES = TMA_T * TMA2_ROWLINE * 16;
constexpr int TMA2_WELEMS = TMA2_WBYTES / 16;
constexpr int TMA2_SMEM = TMA2_NSLOT * TMA2_HALF_BYTES + 2 * TMA2_WBYTES + 1024;      // 230656

__device__ __forceinline__ int ld_volatile_shared(const volatile int* p) { return *p; }

// Two-queue scheduling (a.two_queues): pass-1 tiles and pass-2 tiles are claimed from separate in-order queues.
// A loader takes a P2 tile when the head transform of the P2 queue is fully published, otherwise a P1 tile when
// its scratch slot is free, so the P1 stream runs as far ahead as the S slots allow (about two transforms with
// S = 3) instead of the fixed one-phase distance of the single sequence -- the P1 -> P2 dependency then has slack
// to spare without a fourth 16 MiB slot in L2. Item code: bit 30 = pass, low bits = tile index in that pass.
// No deadlock: a P1 tile waits only for P2 tiles of an older transform whose P1 tiles are all claimed already,
// and a P2 tile waits only for claimed P1 tiles; claimed tiles always finish.
__device__ __forceinline__ TmaItem tma_decode2(int code) {
    constexpr int TPT = TMA_L / TMA_T;
    TmaItem it;
    const int idx = code & 0x3FFFFFFF;
    it.type = (code >> 30) & 1;
    it.tf = idx / TPT;
    it.c = idx % TPT;
    return it;
}

// P1BULK: pass-1 output is staged in the work buffer too and leaves through the async proxy (four 16 KiB bulk copies, one per
// row of Int); the storer lane of the group publishes the tile after cp.async.bulk.wait_group 0, i.e. when the writes are
// complete -- no device-scope fence by the four storing warps, no publisher lane.
template <bool P1BULK>
__global__ void __launch_bounds__(TMA_THREADS, 1)
fft_tma_fused2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_int,
                      const __grid_constant__ CUtensorMap tm_out, const TmaFusedParams a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    cpx* land = reinterpret_cast<cpx*>(smem_raw);
    cpx* work = reinterpret_cast<cpx*>(smem_raw + TMA2_NSLOT * TMA2_HALF_BYTES);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + TMA2_NSLOT * TMA2_HALF_BYTES + 2 * TMA2_WBYTES);
    // full_h is per (slot, consumer group): a parity wait is only safe for a waiter that observes EVERY phase of its
    // barrier in order. With one barrier per slot, group B could reach its wait for phase k+1 while phase k -- a half
    // of group A's tile, issued earlier but landing later (HBM vs L2) -- was still open; the parity test then
    // succeeds at once and B reads a slot that has not landed.
    unsigned long long* full_h = bars;                     // [3 slots][2 groups]
    unsigned long long* freed_h = bars + 6;                // [3]
    unsigned long long* rd = bars + 9;                     // [2]
    unsigned long long* staged = bars + 11;                // [2]
    unsigned long long* drained = bars + 13;               // [2 groups][2 halves of the work buffer]
    unsigned long long* pd = bars + 17;                    // [2][2]
    volatile int* log = reinterpret_cast<volatile int*>(bars + 22);        // [32]
    volatile int* log_count = reinterpret_cast<volatile int*>(bars + 38);  // local steps published by the loader
    constexpr int TPT = TMA_L / TMA_T;
    constexpr int HALF_ELEMS = TMA2_HALF_BYTES / 16;       // 2048
    constexpr int TILE_ELEMS = TMA2_WELEMS;                // work buffer pitch (4104 elements)

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < 6; i++) mbar_init(full_h + i, 1);
        for (int i = 0; i < 3; i++) mbar_init(freed_h + i, TMA_GROUP);
        for (int i = 0; i < 2; i++) { mbar_init(rd + i, TMA_GROUP); mbar_init(staged + i, TMA_GROUP); }
        for (int i = 0; i < 4; i++) mbar_init(drained + i, 1);
        for (int i = 0; i < 4; i++) mbar_init(pd + i, TMA_GROUP);
        *log_count = 0;
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int nitems = 2 * a.batch * TPT;
    const int B = a.batch, D = a.delay, S = a.nslots;

    if (warp >= 2 * TMA_GROUP / 32) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
        if (tid == 2 * TMA_GROUP) {
            // ------------------------------------------------------------ loader
            int tokens = 0, ready_tf = -1;
            const unsigned long long ld_pol_first = policy_evict_first(), ld_pol_last = policy_evict_last();
            long long hidx = 0;                             // halves issued so far
            int ready_tf2 = -1, free_tf1 = S - 1;       // transforms known to be published / whose slot is known to be free
            const int ntiles = B * TPT;
            for (int it = 0; tokens < 2; it++) {
                int item = nitems;
                bool token = tokens != 0;
                TmaItem w;
                w.type = 0; w.tf = 0; w.c = 0;
                const CUtensorMap* tm = &tm_x;
                int tfc = 0;
                if (a.two_queues) {
                    // ---- two in-order queues; prefer the pass the previous step did not take
                    int code = -1;
                    bool fence_needed = false;
                    while (!token && code < 0) {
                        const int q2 = ld_relaxed_gpu(a.queue + 1), q1 = ld_relaxed_gpu(a.queue);
                        if (q1 >= ntiles && q2 >= ntiles) { token = true; break; }
                        bool ok2 = false, ok1 = false;
                        if (q2 < ntiles) {
                            const int gq = q2 / TPT;
                            if (gq <= ready_tf2) ok2 = true;
                            else if (ld_relaxed_gpu(a.done1 + gq) >= TPT) { ready_tf2 = gq; fence_needed = true; ok2 = true; }
                        }
                        if (q1 < ntiles) {
                            const int hq = q1 / TPT;
                            if (hq <= free_tf1) ok1 = true;
                            else if (ld_relaxed_gpu(a.done2 + (hq - S)) >= TPT) { free_tf1 = hq; ok1 = true; }
                        }
                        const bool prefer2 = a.two_queues == 1 ? (it & 1) == 0 : a.two_queues == 2 ? true : a.two_queues == 3 ? false : ((it >> 1) & 1) == 0;
                        int pick = -1;
                        if (ok2 && (prefer2 || !ok1)) pick = 1; else if (ok1) pick = 0;
                        if (pick < 0) { __nanosleep(64); continue; }
                        const int idx = atomicAdd(a.queue + pick, 1);
                        if (idx >= ntiles) continue;            // lost the race for the last tile of that pass
                        code = (pick << 30) | idx;
                        const int tfq = idx / TPT;
                        if (pick == 1 && tfq > ready_tf2) {    // the head moved on to the next transform meanwhile: wait for it
                            while (ld_relaxed_gpu(a.done1 + tfq) < TPT) __nanosleep(32);
                            ready_tf2 = tfq; fence_needed = true;
                        }
                        if (pick == 0 && tfq > free_tf1) {
                            while (ld_relaxed_gpu(a.done2 + (tfq - S)) < TPT) __nanosleep(32);
                            free_tf1 = tfq;
                        }
                    }
                    if (!token) {
                        w = tma_decode2(code);
                        item = code;
                        if (w.type == 1) {
                            if (fence_needed) {
                                asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
                                asm volatile("fence.proxy.async;\n" ::: "memory");
                            }
                            tm = &tm_int; tfc = w.tf % S;
                        } else tfc = w.tf;
                    }
                } else {
                    item = tokens ? nitems : atomicAdd(a.queue, 1);
                    token = item >= nitems;
                    if (!token) {
                        w = tma_decode(item, B, D);
                        if (w.type == 0) {
                            if (w.tf >= S && !(a.dbg_nodeps & 2)) { while (ld_relaxed_gpu(a.done2 + (w.tf - S)) < TPT) __nanosleep(32); }
                            tfc = w.tf;
                        } else {
                            if (w.tf != ready_tf) {             // one poll + fence pair per transform, not per tile
                                if (!(a.dbg_nodeps & 1)) { while (ld_relaxed_gpu(a.done1 + w.tf) < TPT) __nanosleep(32); }
                                // Acquire with a full fence. An acquire load + fence.proxy.async.global is about 2.5 % faster (the
                                // MEMBAR also waits for this lane's own tile loads in flight, ~2000 cycles on the phase boundary), but
                                // the stress test still showed a stale tile about once per 1500 runs of 256 transforms at delay 2 with
                                // it (0 of 2000 with the fences), so the fences stay; `dbg_acqload` keeps the variant measurable.
                                if (a.dbg_acqload) {
                                    (void)ld_acquire_gpu(a.done1 + w.tf);
                                    asm volatile("fence.proxy.async.global;\n" ::: "memory");
                                } else {
                                    asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
                                    asm volatile("fence.proxy.async;\n" ::: "memory");   // other CTAs' generic-proxy stores -> this async-proxy read
                                }
                                ready_tf = w.tf;
                            }
                            tm = &tm_int; tfc = w.tf % S;
                        }
                    }
                }
                log[it & 31] = token ? -1 : item;
                __threadfence_block();
                *log_count = it + 1;
                if (token) tokens++;
#pragma unroll
                for (int h = 0; h < 2; h++, hidx++) {
                    const int s = (int)(hidx % TMA2_NSLOT);
                    if (token) {                            // only the first half of a token is ever looked at (and never freed)
                        if (h == 0) {
                            if (hidx >= TMA2_NSLOT) mbar_wait(freed_h + s, (unsigned)(((hidx - TMA2_NSLOT) / TMA2_NSLOT) & 1));
                            mbar_arrive(full_h + 2 * s + (it & 1));
                        }
                        continue;
                    }
                    if (hidx >= TMA2_NSLOT) mbar_wait(freed_h + s, (unsigned)(((hidx - TMA2_NSLOT) / TMA2_NSLOT) & 1));
                    unsigned long long* fb = full_h + 2 * s + (it & 1);
                    if (a.dbg_noload) { mbar_arrive(fb); continue; }
                    mbar_expect_tx(fb, TMA2_HALF_BYTES);
                    const int hint_bit = w.type == 0 ? 2 : 4;
                    if (a.hints & hint_bit) {
                        const unsigned long long pol = w.type == 0 ? ld_pol_first : ld_pol_last;
#pragma unroll
                        for (int j = 0; j < 2; j++)
                            tma_load_3d_hint(land + (size_t)s * HALF_ELEMS + j * TMA_BOX_ROWS * TMA_T, tm, w.c * 2 * TMA_T,
                                             (2 * h + j) * TMA_BOX_ROWS, tfc, fb, pol);
                    } else {
#pragma unroll
                    for (int j = 0; j < 2; j++)
                        tma_load_3d(land + (size_t)s * HALF_ELEMS + j * TMA_BOX_ROWS * TMA_T, tm, w.c * 2 * TMA_T,
                                    (2 * h + j) * TMA_BOX_ROWS, (w.type == 0 ? tfc & ~a.dbg_in_alias : tfc), fb);
                    }
                }
            }
        } else if (P1BULK && (tid == 2 * TMA_GROUP + 32 || tid == 2 * TMA_GROUP + 64)) {
            // ------------------------------------------------------------ P1BULK: one storer lane per consumer group, every tile
            const int g = tid == 2 * TMA_GROUP + 32 ? 0 : 1;
            const unsigned long long pol_last = policy_evict_last(), pol_first = policy_evict_first();
            unsigned ns = 0;
            for (int it = g;; it += 2) {
                while (ld_volatile_shared(log_count) <= it) __nanosleep(64);
                __threadfence_block();
                const int item = log[it & 31];
                if (item < 0) break;
                const TmaItem pi = a.two_queues ? tma_decode2(item) : tma_decode(item, B, D);
                mbar_wait(staged + g, ns & 1);
                ns++;
                const cpx* srcb = work + (size_t)g * TILE_ELEMS;
                if (pi.type == 1) {
                    if (a.hints & 8) {
#pragma unroll
                        for (int j = 0; j < 2; j++)
                            tma_store_3d_hint(&tm_out, pi.c * 2 * TMA_T, j * TMA_BOX_ROWS, pi.tf, srcb + j * TMA_BOX_ROWS * TMA_T, pol_first);
                        tma_commit();
#pragma unroll
                        for (int j = 2; j < 4; j++)
                            tma_store_3d_hint(&tm_out, pi.c * 2 * TMA_T, j * TMA_BOX_ROWS, pi.tf, srcb + j * TMA_BOX_ROWS * TMA_T, pol_first);
                        tma_commit();
                    } else {
#pragma unroll
                    for (int j = 0; j < 2; j++)
                        tma_store_3d(&tm_out, pi.c * 2 * TMA_T, j * TMA_BOX_ROWS, pi.tf, srcb + j * TMA_BOX_ROWS * TMA_T);
                    tma_commit();
#pragma unroll
                    for (int j = 2; j < 4; j++)
                        tma_store_3d(&tm_out, pi.c * 2 * TMA_T, j * TMA_BOX_ROWS, pi.tf, srcb + j * TMA_BOX_ROWS * TMA_T);
                    tma_commit();
                    }
                } else {
                    cpx* dst = a.scratch + (size_t)(pi.tf % S) * ((size_t)TMA_L * TMA_L) + (size_t)(pi.c * TMA_T) * TMA_L;
#pragma unroll
                    for (int l = 0; l < 2; l++) {
                        if (a.hints & 1) bulk_store_1d_hint(dst + (size_t)l * TMA_L, srcb + l * TMA2_ROWLINE, TMA_L * 16, pol_last);
                        else bulk_store_1d(dst + (size_t)l * TMA_L, srcb + l * TMA2_ROWLINE, TMA_L * 16);
                    }
                    tma_commit();
#pragma unroll
                    for (int l = 2; l < 4; l++) {
                        if (a.hints & 1) bulk_store_1d_hint(dst + (size_t)l * TMA_L, srcb + l * TMA2_ROWLINE, TMA_L * 16, pol_last);
                        else bulk_store_1d(dst + (size_t)l * TMA_L, srcb + l * TMA2_ROWLINE, TMA_L * 16);
                    }
                    tma_commit();
                }
                asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
                mbar_arrive(drained + 2 * g);
                tma_wait_read0();
                mbar_arrive(drained + 2 * g + 1);
                if (pi.type == 0) {
                    tma_wait_all0();                                   // the four rows of Int are written
                    asm volatile("fence.proxy.async.global;\n" ::: "memory");
                    red_release_gpu(a.done1 + pi.tf, 1);
                }
            }
            tma_wait_all0();
        } else if (!P1BULK && tid == 2 * TMA_GROUP + 32) {
            // ------------------------------------------------------------ storer of P2 tiles
            unsigned np2[2] = {0, 0};
            for (int it = 0;; it++) {
                while (ld_volatile_shared(log_count) <= it) __nanosleep(64);
                __threadfence_block();
                const int item = log[it & 31];
                if (item < 0) break;
                const TmaItem pi = a.two_queues ? tma_decode2(item) : tma_decode(item, B, D);
                if (pi.type != 1 || a.p2_stg) continue;
                const int g = it & 1;
                mbar_wait(staged + g, np2[g] & 1);
                np2[g]++;
                const cpx* srcb = work + (size_t)g * TILE_ELEMS;
                // two bulk groups, rows 0..511 and 512..1023: the group's next exchange may refill the first half of the
                // work buffer while the second is still being read out
                if (!a.dbg_nop2st) {
#pragma unroll
                    for (int j = 0; j < 2; j++)
                        tma_store_3d(&tm_out, pi.c * 2 * TMA_T, j * TMA_BOX_ROWS, pi.tf & ~a.dbg_out_alias, srcb + j * TMA_BOX_ROWS * TMA_T);
                    tma_commit();
#pragma unroll
                    for (int j = 2; j < 4; j++)
                        tma_store_3d(&tm_out, pi.c * 2 * TMA_T, j * TMA_BOX_ROWS, pi.tf & ~a.dbg_out_alias, srcb + j * TMA_BOX_ROWS * TMA_T);
                    tma_commit();
                    if (a.dbg_nosplit) tma_wait_read0();
                    else asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
                }
                mbar_arrive(drained + 2 * g);
                if (!a.dbg_nop2st) tma_wait_read0();
                mbar_arrive(drained + 2 * g + 1);
            }
            tma_wait_all0();
        } else if (!P1BULK && tid == 2 * TMA_GROUP + 64) {
            // ------------------------------------------------------------ publisher of P1 tiles
            unsigned np1[2] = {0, 0};
            for (int it = 0;; it++) {
                while (ld_volatile_shared(log_count) <= it) __nanosleep(64);
                __threadfence_block();
                const int item = log[it & 31];
                if (item < 0) break;
                const TmaItem pi = a.two_queues ? tma_decode2(item) : tma_decode(item, B, D);
                if (pi.type != 0) continue;
                const int g = it & 1;
                mbar_wait(pd + 2 * g + (np1[g] & 1), (np1[g] >> 1) & 1);
                np1[g]++;
                red_release_gpu(a.done1 + pi.tf, 1);       // cumulative: publishes the group's stores (ordered by the mbarrier)
            }
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
    unsigned nrd = 0, np1 = 0, np2 = 0;                     // phases of rd[g] waited so far; P1 / P2 tiles of this group so far
    bool prev_p2 = false, first = true;
    unsigned fph = 0;                                       // phase bit of full_h[slot][g], one per slot
    for (int it = g;; it += 2) {
        const long long h0 = 2LL * it;
        const int s0 = (int)(h0 % TMA2_NSLOT), s1 = (int)((h0 + 1) % TMA2_NSLOT);
        mbar_wait(full_h + 2 * s0 + g, (fph >> s0) & 1);
        fph ^= 1u << s0;
        const int item = log[it & 31];
        if (item < 0) break;
        const TmaItem wi = a.two_queues ? tma_decode2(item) : tma_decode(item, B, D);
        const unsigned ld_conj = (a.ld_conj && wi.type == 0) ? 0x80000000u : 0u;
        cpx x[32];
        {
            const cpx* s = land + (size_t)s0 * HALF_ELEMS + p * TMA_T + ell;
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = cconj_if(s[i * 32 * TMA_T], ld_conj);
        }
        mbar_arrive(freed_h + s0);
        mbar_wait(full_h + 2 * s1 + g, (fph >> s1) & 1);
        fph ^= 1u << s1;
        if (wi.type == 1 && tig == 0) red_relaxed_gpu(a.done2 + wi.tf, 1);      // this tile of Int has been read
        {
            const cpx* s = land + (size_t)s1 * HALF_ELEMS + p * TMA_T + ell;
#pragma unroll
            for (int i = 0; i < 16; i++) x[16 + i] = cconj_if(s[i * 32 * TMA_T], ld_conj);
        }
        mbar_arrive(freed_h + s1);
        dft32(x);
        // the work buffer is free once the previous tile's gathers are done (P1) or its staged outputs have drained (P2)
        if (!first) {
            if (prev_p2) mbar_wait(drained + 2 * g, (np2 - 1) & 1);
            else { mbar_wait(rd + g, nrd & 1); nrd++; }
        }
        {
            cpx* s = wbuf + p * TMA_T + ell;
#pragma unroll
            for (int r = 0; r < 16; r++) s[r * 32 * TMA_T] = x[r];                          // rows p + 32 r < 512
            if (!first && prev_p2) mbar_wait(drained + 2 * g + 1, (np2 - 1) & 1);
#pragma unroll
            for (int r = 16; r < 32; r++) s[r * 32 * TMA_T] = x[r];
        }
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
        {
            const cpx* s = wbuf + (32 * p) * TMA_T + ell;
#pragma unroll
            for (int j = 0; j < 32; j++) x[j] = s[(j ^ odd) * TMA_T];
        }
        mbar_arrive(rd + g);
        if (odd) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) { cpx t = x[j]; x[j] = x[j + 1]; x[j + 1] = t; }
        }
        asm volatile("" : "+d"(w.x), "+d"(w.y));
        mul_powers32(x, w);
        dft32(x);
        if (wi.type == 0 && P1BULK) {
            mul_geometric32(x, cmul(t_hi0, t_lo0), cmul(t_hi1, t_lo1));
            mbar_wait(rd + g, nrd & 1);                      // every gather of this tile is done: the buffer may be overwritten
            nrd++;
            cpx* s = wbuf + ell * TMA2_ROWLINE + p;          // Int[n2 = 4c + ell][k1 = p + 32 r], rows skewed by 2 elements
#pragma unroll
            for (int r = 0; r < 32; r++) s[r * 32] = x[r];
            fence_proxy_async();
            mbar_arrive(staged + g);
            np2++;
            prev_p2 = true;
        } else if (wi.type == 0) {
            cpx* dst = a.scratch + (size_t)(wi.tf % S) * ((size_t)TMA_L * TMA_L) + (size_t)(wi.c * TMA_T + ell) * TMA_L + p;
            const cpx t0 = cmul(t_hi0, t_lo0), s1c = cmul(t_hi1, t_lo1);
            const cpx s2 = csqr(s1c), s4 = csqr(s2);
            cpx t[4];
            t[0] = t0; t[1] = cmul(t0, s1c); t[2] = cmul(t0, s2); t[3] = cmul(t[1], s2);
            if (a.dbg_nop1st) {
                cpx acc = make_double2(0.0, 0.0);
#pragma unroll
                for (int bb = 0; bb < 8; bb++) {
#pragma unroll
                    for (int aa = 0; aa < 4; aa++) {
                        acc = cadd(acc, cmul(x[4 * bb + aa], t[aa]));
                        if (bb < 7) t[aa] = cmul(t[aa], s4);
                    }
                }
                if (acc.x == 1.2345e-300) dst[0] = acc;
            } else {
#pragma unroll
                for (int bb = 0; bb < 8; bb++) {
#pragma unroll
                    for (int aa = 0; aa < 4; aa++) {
                        dst[32 * (4 * bb + aa)] = cmul(x[4 * bb + aa], t[aa]);
                        if (bb < 7) t[aa] = cmul(t[aa], s4);
                    }
                }
            }
            // device-scope fence by every storing warp before the hand-off. Measured with tools/stress_fft.py (400 x 256
            // transforms each): a release issued only by the publisher lane (another warp, mbarrier hand-off) gave wrong
            // rows in 15 of 400 runs at delay 2 and 1 of 150 at delay 1 -- P2 tiles loaded rows of Int that had not reached
            // L2 -- and so did "bar.sync, then one thread fences and releases" (17 of 400); with this fence 0 of 1100.
            // It costs about 6 % (the warp waits for its stores to be acknowledged before it can start the next tile).
            if (!a.dbg_nopubfence) asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
            if (a.dbg_wproxy) asm volatile("fence.proxy.async.global;\n" ::: "memory");
            mbar_arrive(pd + 2 * g + (np1 & 1));
            np1++;
            prev_p2 = false;
        } else {
            if (a.st_conj || a.scale != 1.0) {
                const double sx = a.scale, sy = a.st_conj ? -a.scale : a.scale;
#pragma unroll
                for (int s = 0; s < 32; s++) x[s] = make_double2(x[s].x * sx, x[s].y * sy);
            }
            if (a.p2_stg) {
                cpx* dst = a.out + (long long)wi.tf * a.out_dist + (long long)p * TMA_L + wi.c * TMA_T + ell;   // X[k1 + 1024 k2], k2 = p + 32 r
#pragma unroll
                for (int r = 0; r < 32; r++) __stcs(reinterpret_cast<double2*>(dst + (long long)r * 32 * TMA_L), x[r]);
                prev_p2 = false;                             // nothing staged: the work buffer is free once the gathers are done
            } else {
                mbar_wait(rd + g, nrd & 1);                  // every gather of this tile is done: the slots may be overwritten
                nrd++;
                cpx* s = wbuf + p * TMA_T + ell;
#pragma unroll
                for (int r = 0; r < 32; r++) s[r * 32 * TMA_T] = x[r];
                fence_proxy_async();
                mbar_arrive(staged + g);
                np2++;
                prev_p2 = true;
            }
        }
    }
}

}  // namespace gd

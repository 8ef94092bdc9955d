// fft_fused.cuh -- both four-step passes of a batch of N = L*L point transforms in ONE
// persistent kernel, with the inter-pass array kept in the 126 MB L2 instead of HBM.
//
// The batch is cut into groups of J transforms. Work is a fixed sequence of phases
//     P1(0) .. P1(D) P2(0) P1(D+1) P2(1) ... P1(G-1) P2(G-D-1) ... P2(G-1)
// (P1 = column transforms + fused twiddle, P2 = row transforms + transposed store), each
// phase = J*L/T tiles, handed out in order from a device-wide queue to the resident CTAs. P1(g) writes the
// group's intermediate into scratch slot g mod (D+2) (tile-major layout, so both its stores and
// P2's loads are fully coalesced); P2(g) starts when a device-scope counter says every
// P1(g) tile has been published, and P1(g) reuses a slot only after P2(g-D-2) has drained
// it. Every dependency points at an earlier phase, a CTA never blocks while it holds an
// unfinished tile, and all CTAs are co-resident, so the schedule cannot deadlock.
// The scratch slots are covered by a persisting L2 access-policy window (set by the host) and the
// output stores are streaming (st.global.cs), so HBM sees ~32 B/point (the algorithmic minimum), not 64.
#pragma once
#include <stdio.h>
#include "fft_pass.cuh"

namespace gd {

struct FusedParams {
    const cpx* in;
    cpx* out;
    cpx* scratch;              // NSLOT slots of group_tf * N elements
    long long in_dist, out_dist;
    int batch;                 // transforms
    int group_tf;              // J: transforms per group
    int ngroups;
    int* done1;                // [ngroups] tiles of P1(g) published
    int* done2;                // [ngroups] tiles of P2(g) finished
    int* next_item;            // work queue head (zeroed by the host): item = gridDim.x + atomicAdd(next_item, 1)
    int log2n;
    int ld_conj, st_flags;     // LD_CONJ on the first pass; ST_CONJ | ST_SCALE on the last
    double scale;
    const cpx* tw_lo;
    const cpx* tw_hi;
    const cpx* wl;
    int debug;
    int delay;                 // D: P2(g) is issued right after P1(g + D)
    int nslots;                // D + 2 scratch slots
};


// NOTE: no per-instruction L2 cache hints here. With `.L2::cache_hint` operands ptxas 12.9 placed the
// policy descriptor of one LDGSTS group in an odd uniform-register pair (desc[UR1]), which traps with
// "illegal instruction" on sm_100a. L2 residency of the scratch slots is requested from the host instead
// (stream access-policy window, engine.cu).
// Counter reads are relaxed (strong, device scope): every load that consumes the published data is an
// L2-only cp.async issued after the read returned, so no L1 invalidation (ld.acquire's CCTL.IVALL) is needed.
__device__ __forceinline__ int ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// release-increment: MEMBAR + REDG, without the L1 invalidation a full fence would add
__device__ __forceinline__ void red_release_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

struct FusedItem {
    int type;      // 0 = P1, 1 = P2
    int group;
    int tf;        // transform index in the batch
    int tile;      // tile inside the transform
    bool exists;
};

template <int LOG2L, int T>
__global__ void __launch_bounds__(T * PassShape<LOG2L>::P, (T * PassShape<LOG2L>::P >= 512) ? 1 : 2)
fft_fused_kernel(const FusedParams a) {
    using SH = PassShape<LOG2L>;
    constexpr int L = SH::L, P = SH::P, NT = T * P;
    constexpr int LS = line_stride(L, T);
    constexpr int TPT = L / T;                    // tiles per transform per pass
    static_assert(LOG2L >= 8, "fused kernel needs T*L == 16*NT");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cpx* sm = reinterpret_cast<cpx*>(smem_raw);
    __shared__ int s_ready, s_next;

    const int tid = threadIdx.x;
    const int ell = tid % T, p = tid / T;         // lanes run across the T adjacent lines in both passes
    cpx* sl = sm + ell * LS;
    const unsigned ld_conj = a.ld_conj ? 0x80000000u : 0u;
    const long long N = (long long)L * L;
    const int phase_items = a.group_tf * TPT;       // host guarantees 2 * ngroups * phase_items < 2^31
    const int total = 2 * a.ngroups * phase_items;
    const long long slot_elems = (long long)a.group_tf * N;

    auto decode = [&](int g) {
        FusedItem it;
        it.exists = false;
        it.type = 0; it.group = 0; it.tf = 0; it.tile = 0;
        if (g >= total) return it;
        int f = g / phase_items;
        int w = g - f * phase_items;
        // phase order: P1(0..D), then pairs P2(i), P1(D+1+i), then the remaining P2s
        const int D = a.delay, G = a.ngroups;
        if (G <= D + 1) {
            if (f < G) { it.type = 0; it.group = f; } else { it.type = 1; it.group = f - G; }
        } else if (f <= D) { it.type = 0; it.group = f; }
        else {
            const int m = f - D - 1, npairs = G - D - 1;
            if (m < 2 * npairs) {
                if (m & 1) { it.type = 0; it.group = D + 1 + (m >> 1); } else { it.type = 1; it.group = m >> 1; }
            } else { it.type = 1; it.group = npairs + (m - 2 * npairs); }
        }
        int j = w / TPT;
        it.tile = w - j * TPT;
        it.tf = it.group * a.group_tf + j;
        it.exists = it.tf < a.batch;              // the last group may be partial
        return it;
    };
    // number of tiles a finished phase of group g publishes
    auto phase_count = [&](int g) {
        int ntf = a.batch - g * a.group_tf;
        if (ntf > a.group_tf) ntf = a.group_tf;
        return ntf * TPT;
    };
    auto dep_ready = [&](const FusedItem& it) -> bool {   // one thread
        if (it.type == 1) return ld_relaxed(a.done1 + it.group) >= phase_count(it.group);
        if (it.group >= a.nslots) return ld_relaxed(a.done2 + it.group - a.nslots) >= phase_count(it.group - a.nslots);
        return true;
    };
    auto prefetch = [&](const FusedItem& it) {
        const int j = it.tf - it.group * a.group_tf;
        if (it.type == 0) {
            // column n2 = tile*T + ell of transform tf: element n1 = p + P*i at n1*L + n2
            const cpx* src = a.in + (long long)it.tf * a.in_dist + (long long)p * L + it.tile * T + ell;
#pragma unroll
            for (int i = 0; i < 16; i++) cp_async16(sm + i * NT + tid, src + (long long)i * (P * L), 16);
        } else {
            // row k1 = tile*T + ell: element n2 = p + P*i lives at (n2/T)*(T*L) + k1*T + n2%T (tile-major)
            const cpx* src = a.scratch + (long long)(it.group % a.nslots) * slot_elems + (long long)j * N +
                             (long long)(p / T) * (T * L) + (it.tile * T + ell) * T + (p % T);
#pragma unroll
            for (int i = 0; i < 16; i++) cp_async16(sm + i * NT + tid, src + (long long)i * (P * L), 16);
        }
        cp_async_commit();
    };
    auto wait_dep = [&](const FusedItem& it) {    // whole CTA; blocks until the item's inputs exist
        if (tid == 0 && !(a.debug & 4)) {
            while (!dep_ready(it)) __nanosleep(200);
        }
        __syncthreads();
    };

    // address + target of the counter an item depends on (nullptr: no dependency)
    auto dep_counter = [&](const FusedItem& it, int* need) -> const int* {
        if (it.type == 1) { *need = phase_count(it.group); return a.done1 + it.group; }
        if (it.group >= a.nslots) { *need = phase_count(it.group - a.nslots); return a.done2 + it.group - a.nslots; }
        *need = 0;
        return nullptr;
    };
    auto publish = [&](const FusedItem& it) {     // thread 0, after a CTA barrier that follows the tile's stores
        red_release_add((it.type == 0 ? a.done1 : a.done2) + it.group, 1);
    };

    // the first item of every CTA is its block index; later ones come from the queue
    if ((int)blockIdx.x >= total) return;
    FusedItem cur = decode(blockIdx.x);
    if (!cur.exists) return;                      // cannot happen: grid <= items of the first (full) group
    wait_dep(cur);
    prefetch(cur);
    FusedItem pend = cur;
    bool pend_valid = false;

    while (true) {
        // claim the next item of this CTA from the device-wide queue (dynamic: a slow CTA takes fewer tiles
        // instead of holding every phase back); the atomic's round trip overlaps the first butterfly step
        int claimed = 0;
        if (tid == 0) claimed = (int)gridDim.x + atomicAdd(a.next_item, 1);
#ifdef GD_FUSED_TRACE
        long long ck[8];
        const bool tr = (a.debug & 16) && tid == 64 && (blockIdx.x == 0 || blockIdx.x == 200);
        if (tr) ck[0] = clock64();
#define GD_CK(i) if (tr) ck[i] = clock64()
#else
#define GD_CK(i)
#endif

        cpx x[16];
        cp_async_wait_all();
        GD_CK(1);
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = sm[i * NT + tid];
        if (cur.type == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = cconj_if(x[i], ld_conj);
        }
        butterfly_step<L, 16, 1>(x, p, a.wl);
        int probe = 0, need = 0;
        if (tid == 0) {
            FusedItem n = decode(claimed);
            while (claimed < total && !n.exists) { claimed = (int)gridDim.x + atomicAdd(a.next_item, 1); n = decode(claimed); }
            s_next = claimed;
            if (claimed < total) {                // probe its dependency now; the answer is consumed after the last gather
                const int* c = dep_counter(n, &need);
                if (c) probe = ld_relaxed(c);
            }
        }
        GD_CK(2);
        __syncthreads();                          // every thread has taken its prefetched inputs; s_next visible
        GD_CK(3);
        const int gn = s_next;
        const bool have_next = gn < total;
        const FusedItem nxt = decode(gn);
        scatter_step<L, 16, 1>(x, p, sl);
        __syncthreads();
        // publish the previous tile from the LAST warp (the first one carries the queue / probe duties), here
        // where a long compute section follows: the release fence's round trip hides behind step 2
        if (pend_valid && tid == NT - 32) publish(pend);
        gather_step<L>(x, p, sl);
        if constexpr (SH::NSTEP == 3) {
            __syncthreads();
            butterfly_step<L, 16, 16>(x, p, a.wl);
            scatter_step<L, 16, 16>(x, p, sl);
            __syncthreads();
            gather_step<L>(x, p, sl);
        }
        if (tid == 0) {
            if (have_next && probe < need) {          // not ready a tile ago: look again (costs one L2 round trip)
                const int* c = dep_counter(nxt, &need);
                probe = ld_relaxed(c);
            }
            s_ready = (have_next && probe >= need && !(a.debug & 1)) ? 1 : 0;
        }
        GD_CK(4);
        __syncthreads();                          // exchange buffer is free again; s_ready visible
        GD_CK(5);
        const bool early = s_ready != 0;
        // twiddle loads of the tail go out BEFORE the prefetch burst (they would queue behind it otherwise)
        constexpr int LNS = SH::NSTEP == 2 ? 16 : 256;
        cpx wlast[16 / SH::LASTR];
        load_step_twiddles<L, SH::LASTR, LNS>(wlast, p, a.wl);
        cpx tw0 = make_double2(1.0, 0.0), tws = make_double2(1.0, 0.0);
        if (cur.type == 0) {
            const unsigned long long mask = (1ULL << a.log2n) - 1ULL;
            const unsigned long long n2 = (unsigned long long)(cur.tile * T + ell);
            PassParams tp;                         // only the twiddle fields are used by tw_lookup
            tp.tw_log2m = a.log2n; tp.tw_lo = a.tw_lo; tp.tw_hi = a.tw_hi;
            tw0 = tw_lookup(tp, (n2 * (unsigned long long)p) & mask);
            tws = tw_lookup(tp, (n2 * (unsigned long long)P) & mask);
        }
        if (early) prefetch(nxt);
        butterfly_step_w<L, SH::LASTR, LNS>(x, wlast);

        const int j = cur.tf - cur.group * a.group_tf;
        if (cur.type == 0) {
            // fused twiddle w_N^(n2*k1), k1 = p + P*i  (base tw0 and step tws were looked up above)
            cpx t[16];
            t[0] = tw0;
            const cpx s1 = tws;
            t[1] = cmul(t[0], s1);
            cpx s2 = csqr(s1);
            t[2] = cmul(t[0], s2); t[3] = cmul(t[1], s2);
            cpx s4 = csqr(s2);
#pragma unroll
            for (int i = 0; i < 4; i++) t[4 + i] = cmul(t[i], s4);
            cpx s8 = csqr(s4);
#pragma unroll
            for (int i = 0; i < 8; i++) t[8 + i] = cmul(t[i], s8);
            // tile-major scratch: element (k1, lane) of this tile at tile*(T*L) + k1*T + ell
            cpx* dst = a.scratch + (long long)(cur.group % a.nslots) * slot_elems + (long long)j * N +
                       (long long)cur.tile * (T * L) + p * T + ell;
#pragma unroll
            for (int i = 0; i < 16; i++) dst[i * (P * T)] = cmul(x[i], t[i]);
        } else {
            // X[k1 + L*k2], k1 = tile*T + ell, k2 = p + P*i
            cpx* dst = a.out + (long long)cur.tf * a.out_dist + (long long)p * L + cur.tile * T + ell;
            double sx = 1.0, sy = 1.0;
            if (a.st_flags & ST_SCALE) { sx = a.scale; sy = a.scale; }
            if (a.st_flags & ST_CONJ) sy = -sy;
            if (a.st_flags & (ST_SCALE | ST_CONJ)) {
#pragma unroll
                for (int i = 0; i < 16; i++) __stcs(dst + (long long)i * (P * L), make_double2(x[i].x * sx, x[i].y * sy));
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) __stcs(dst + (long long)i * (P * L), x[i]);
            }
        }
#ifdef GD_FUSED_TRACE
        if (tr) {
            ck[6] = clock64();
            printf("cta %d type %d grp %d tile %d early %d | wait %lld ld+bf1 %lld bar1 %lld mid %lld bar5 %lld tail %lld total %lld\n", blockIdx.x, cur.type,
                   cur.group, cur.tile, (int)early, ck[1] - ck[0], ck[2] - ck[1], ck[3] - ck[2], ck[4] - ck[3], ck[5] - ck[4], ck[6] - ck[5], ck[6] - ck[0]);
        }
#endif
        pend = cur;
        pend_valid = true;
        if (!have_next) break;
        if (!early) {
            // never block while holding an unpublished tile
            __syncthreads();
            if (tid == NT - 32) publish(pend);
            pend_valid = false;
            wait_dep(nxt);
            prefetch(nxt);
        }
        cur = nxt;
    }
    __syncthreads();
    if (tid == NT - 32) publish(pend);
}

}  // namespace gd

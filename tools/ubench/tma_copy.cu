// TMA tile-copy throughput per SM as a function of the box row width: one thread per CTA streams tiles
// global -> shared (cp.async.bulk.tensor) -> global through a ring of 3 x 64 KiB buffers, no compute.
// Box = ROWB bytes x (65536 / ROWB / nbox) rows; the tensor is [rows][16 KiB pitch] like the FFT tiles.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../go-dsp_b200/csrc tma_copy.cu -o tma_copy
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "fft_tma.cuh"
using namespace gd;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// mode: 1 = load only, 2 = store only, 3 = load + store
__global__ void __launch_bounds__(32, 1) copy_kernel(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout,
                                                     int ntiles, int rowd /*doubles per box row*/, int boxrows, int nbox, int mode, int tf_mask) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + 3 * 65536);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 3; i++) mbar_init(full + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int tiles_per_tf = 2048 / rowd;          // column tiles across a 16 KiB row
    const int my = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    for (int it = 0; it < my + 2; it++) {
        // stage A: load tile it into buffer it % 3 (its previous store has been waited for below)
        if (it < my) {
            const int b = it % 3;
            const int tile = blockIdx.x + it * gridDim.x;
            const int tf = (tile / tiles_per_tf) & tf_mask, c = tile % tiles_per_tf;
            if (mode & 1) {
                mbar_expect_tx(full + b, 65536);
                for (int j = 0; j < nbox; j++)
                    tma_load_3d(smem_raw + b * 65536 + j * (65536 / nbox), &tin, c * rowd, j * boxrows, tf, full + b);
            }
        }
        // stage B: store tile it - 2
        const int st = it - 2;
        if (st >= 0 && st < my) {
            const int b = st % 3;
            const int tile = blockIdx.x + st * gridDim.x;
            const int tf = (tile / tiles_per_tf) & tf_mask, c = tile % tiles_per_tf;
            if (mode & 1) mbar_wait(full + b, (st / 3) & 1);
            if (mode & 2) {
                for (int j = 0; j < nbox; j++)
                    tma_store_3d(&tout, c * rowd, j * boxrows, tf, smem_raw + b * 65536 + j * (65536 / nbox));
                tma_commit();
                tma_wait_read0();
            }
        }
    }
    tma_wait_all0();
}

int main(int argc, char** argv) {
    int batch = argc > 1 ? atoi(argv[1]) : 64;
    int mask = argc > 2 ? atoi(argv[2]) : -1;
    const long long N = 1 << 20;
    int nbuf = mask >= 0 ? mask + 1 : batch;
    double *x, *y;
    CK(cudaMalloc(&x, (size_t)nbuf * N * 16)); CK(cudaMalloc(&y, (size_t)nbuf * N * 16));
    CK(cudaMemset(x, 1, (size_t)nbuf * N * 16)); CK(cudaMemset(y, 0, (size_t)nbuf * N * 16));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fn;
    CK(cudaFuncSetAttribute(copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 65536 + 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rowb : {32, 64, 128, 256, 1024}) {
        const int rowd = rowb / 8, rows = 65536 / rowb;           // rows per tile
        int boxrows = rows > 256 ? 256 : rows, nbox = rows / boxrows;
        if (rows > 1024) continue;                                // tile taller than the 1024-row transform
        auto mk = [&](void* base) {
            CUtensorMap m;
            cuuint64_t dims[3] = {2048, 1024, (cuuint64_t)nbuf};
            cuuint64_t strides[2] = {16384, (cuuint64_t)N * 16};
            cuuint32_t box[3] = {(cuuint32_t)rowd, (cuuint32_t)boxrows, 1};
            cuuint32_t es[3] = {1, 1, 1};
            CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
            return m;
        };
        CUtensorMap mi = mk(x), mo = mk(y);
        // a tile = `rows` rows of rowb bytes; tiles per transform = (16384 / rowb) column tiles x (1024 / rows) row bands;
        // to keep it simple only the first band of each column tile is used when rows < 1024 (same bytes per tile)
        const int ntiles = batch * (2048 / rowd);
        for (int mode : {1, 2, 3}) {
            float best = 1e9;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                copy_kernel<<<148, 32, 3 * 65536 + 1024>>>(mi, mo, ntiles, rowd, boxrows, nbox, mode, mask);
                cudaEventRecord(e1);
                CK(cudaDeviceSynchronize());
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            const double bytes = (double)ntiles * 65536;
            const double cyc_per_tile = best * 1e-3 * 1.965e9 / ((double)ntiles / 148);
            printf("row %4d B  box %3d rows x %d  mode %s  %.3f ms  %.0f GB/s per direction  %.0f cycles per 64 KiB tile per SM  (%.2f cycles/row)\n", rowb, boxrows, nbox,
                   mode == 1 ? "load " : mode == 2 ? "store" : "both ", best, bytes / best / 1e6, cyc_per_tile, cyc_per_tile / rows / (mode == 3 ? 2 : 1));
        }
    }
    printf("last error %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// Raw PCIe copy bandwidth of the box: H2D alone, D2H alone, both directions at once (pinned memory),
// for several chunk sizes. Sets the ceiling for the end-to-end (host-buffer) entry points.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 pcie_bw.cu -o pcie_bw
#include <cstdio>
#include <cuda_runtime.h>
#include <chrono>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    const size_t total = 2ull << 30;
    char *h_in, *h_out, *d_in, *d_out;
    cudaHostAlloc(&h_in, total, cudaHostAllocPortable);
    cudaHostAlloc(&h_out, total, cudaHostAllocPortable);
    cudaMalloc(&d_in, total); cudaMalloc(&d_out, total);
    memset(h_in, 1, total); memset(h_out, 2, total);
    cudaStream_t s0, s1;
    cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
    for (size_t chunk : {8ull << 20, 32ull << 20, 128ull << 20, 512ull << 20, 2048ull << 20}) {
        for (int mode = 0; mode < 3; mode++) {
            double best = 1e9;
            for (int rep = 0; rep < 3; rep++) {
                cudaDeviceSynchronize();
                double t0 = now();
                for (size_t o = 0; o < total; o += chunk) {
                    if (mode != 1) cudaMemcpyAsync(d_in + o, h_in + o, chunk, cudaMemcpyHostToDevice, s0);
                    if (mode != 0) cudaMemcpyAsync(h_out + o, d_out + o, chunk, cudaMemcpyDeviceToHost, s1);
                }
                cudaStreamSynchronize(s0); cudaStreamSynchronize(s1);
                double t = now() - t0;
                if (t < best) best = t;
            }
            const char* nm[3] = {"H2D only", "D2H only", "both    "};
            printf("chunk %5zu MiB  %s  %.2f ms  %.1f GB/s per direction\n", chunk >> 20, nm[mode], best * 1e3, total / best / 1e9);
        }
    }
    int v = 0; cudaDeviceGetAttribute(&v, cudaDevAttrAsyncEngineCount, 0); printf("async engines %d\n", v);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

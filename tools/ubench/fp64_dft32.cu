// Microbenchmark: how much of the FP64 pipe can register-resident radix-32 butterflies use with 8 / 12 warps per SM?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../go-dsp_b200/csrc fp64_dft32.cu -o fp64_dft32
#include <cstdio>
#include "fft_w32.cuh"
using namespace gd;

template <int MINB, int MODE>
__global__ void __launch_bounds__(128, MINB) k(double* out, int iters, cpx w0) {
    cpx x[32];
#pragma unroll
    for (int i = 0; i < 32; i++) x[i] = make_double2(threadIdx.x * 0.001 + i, blockIdx.x * 0.002 - i);
    extern __shared__ __align__(16) unsigned char smraw[];
    cpx* sm = reinterpret_cast<cpx*>(smraw);
    const int tid = threadIdx.x, ell = tid % 4, p = tid / 4;
    cpx* sl = sm + ell * w32_line_stride(4);
    cpx w = w0;
    for (int it = 0; it < iters; it++) {
        dft32(x);
        if (MODE >= 1) w32_exchange<false>(x, p, sl, p, sl);
        mul_powers32(x, w);
        dft32(x);
        if (MODE >= 2) mul_geometric32(x, w, w0);
        w.x += 1e-9;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <int MINB, int MODE>
void run(const char* name, double ops_per_iter) {
    int iters = 200;
    int smem = 4 * w32_line_stride(4) * 16;
    cudaFuncSetAttribute(k<MINB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<MINB, MODE>, 128, smem);
    int grid = 148 * nb;
    double* out;
    cudaMalloc(&out, grid * 128 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MINB, MODE><<<grid, 128, smem>>>(out, 10, make_double2(0.999, -0.01));
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MINB, MODE><<<grid, 128, smem>>>(out, iters, make_double2(0.999, -0.01));
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double pts = (double)grid * 128 * 32 * iters;     // point-passes
    printf("%-28s blocks/SM %d  %.3f ms  %.1f Gpt-pass/s (= %.1f GS/s for 2 passes)  err=%s\n", name, nb, ms, pts / ms / 1e6, pts / ms / 2e6,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    run<2, 0>("regs244 compute only", 0);
    run<3, 0>("regs168 compute only", 0);
    run<2, 1>("regs244 + exchange", 0);
    run<3, 1>("regs168 + exchange", 0);
    run<2, 2>("regs244 + exch + 4step tw", 0);
    run<3, 2>("regs168 + exch + 4step tw", 0);
    return 0;
}

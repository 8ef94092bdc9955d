// Microbenchmark: FP64 pipe peak on B200 (DFMA / DADD / DMUL warp-instruction issue rate) as a function of resident
// warps per SM and independent chains per thread; the denominator of `roofline.fp64_frac` in bench.py.
// Also: cost of LDS.128 when the lanes of a warp read 8 distinct 16-byte words (4-way broadcast) vs 32 distinct.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp64_peak.cu -o fp64_peak && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void __launch_bounds__(1024) fp64_kernel(double* out, int iters, double a, double b) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (OP == 0) v[i] = fma(v[i], a, b);
                else if (OP == 1) v[i] = v[i] + a;
                else v[i] = v[i] * a;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP, int OP>
double run_fp64(int warps_per_sm, int sms, double* out) {
    const int iters = 2000, threads = warps_per_sm * 32;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fp64_kernel<ILP, OP><<<sms, threads>>>(out, 10, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    fp64_kernel<ILP, OP><<<sms, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return (double)sms * threads * iters * 16.0 * ILP / (ms * 1e-3);      // thread-instructions per second
}

// mode 0: lanes read 32 distinct 16-byte words (512 B); mode 1: 8 distinct words, 4 lanes each (128 B contiguous);
// mode 2: all lanes the same word
template <int MODE>
__global__ void __launch_bounds__(256) lds_kernel(double* out, int iters) {
    __shared__ double2 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_double2(i, -i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int idx = MODE == 0 ? lane : (MODE == 1 ? (lane >> 2) : 0);
    double2 acc = make_double2(0, 0);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 32; r++) {
            double2 v = sm[(idx + 32 * r) & 2047];
            acc.x += v.x; acc.y += v.y;
        }
        idx = (idx + (int)acc.x) & 2047 & ~31 | (idx & 31);      // keep the compiler from hoisting; pattern unchanged
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}
template <int MODE>
double run_lds(int sms, double* out) {
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    lds_kernel<MODE><<<sms, 256>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    lds_kernel<MODE><<<sms, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return (double)sms * 8 * iters * 32.0 / (ms * 1e-3);                 // warp-level LDS.128 per second
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double* out;
    cudaMalloc(&out, (size_t)sms * 1024 * 8);
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_max_mhz\": %d,\n", p.name, sms, khz / 1000);
    const char* ops[3] = {"dfma", "dadd", "dmul"};
    double best = 0;
    for (int op = 0; op < 3; op++) {
        for (int w : {4, 8, 16, 32}) {
            double r4 = op == 0 ? run_fp64<4, 0>(w, sms, out) : op == 1 ? run_fp64<4, 1>(w, sms, out) : run_fp64<4, 2>(w, sms, out);
            double r8 = op == 0 ? run_fp64<8, 0>(w, sms, out) : op == 1 ? run_fp64<8, 1>(w, sms, out) : run_fp64<8, 2>(w, sms, out);
            printf(" \"%s_w%d\": {\"ilp4_tinst_per_s\": %.4g, \"ilp8_tinst_per_s\": %.4g, \"ilp8_per_clk_per_sm_at_max_clock\": %.2f},\n", ops[op], w, r4,
                   r8, r8 / sms / (khz * 1e3));
            if (op == 0 && r8 > best) best = r8;
            if (op == 0 && r4 > best) best = r4;
        }
    }
    printf(" \"dfma_peak_tinst_per_s\": %.5g, \"dfma_peak_tflops\": %.3f,\n", best, 2 * best / 1e12);
    double l0 = run_lds<0>(sms, out), l1 = run_lds<1>(sms, out), l2 = run_lds<2>(sms, out);
    printf(" \"lds128_warp_inst_per_s\": {\"distinct32\": %.4g, \"distinct8_bcast4\": %.4g, \"same\": %.4g},\n", l0, l1, l2);
    printf(" \"lds128_clk_per_inst_per_sm_at_max_clock\": {\"distinct32\": %.2f, \"distinct8_bcast4\": %.2f, \"same\": %.2f}}\n",
           sms * khz * 1e3 / l0, sms * khz * 1e3 / l1, sms * khz * 1e3 / l2);
    return 0;
}

#include "fft_w32.cuh"
#include "pass_launch.cuh"
namespace gd {

template <int T, int MINB, bool STAGED>
static cudaError_t launch_pass32_impl(const PassParams& a, int num_sms, cudaStream_t st) {
    static KernelInfoPerDevice per_dev;
    KernelInfo& info = per_dev.current();
    auto kern = fft_pass32_kernel<T, MINB, STAGED>;
    if (!info.ready) {
        info.threads = 32 * T;
        info.smem = T * w32_line_stride(T) * (int)sizeof(cpx);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, info.smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&info.blocks_per_sm, kern, info.threads, info.smem);
        if (e != cudaSuccess) return e;
        if (info.blocks_per_sm < 1) return cudaErrorLaunchOutOfResources;
        info.ready = true;
    }
    long long ntiles = (a.nlines + T - 1) / T;
    if (ntiles <= 0) return cudaSuccess;
    long long cap = (long long)num_sms * info.blocks_per_sm;
    int grid = (int)(ntiles < cap ? ntiles : cap);
    kern<<<grid, info.threads, info.smem, st>>>(a);
    return cudaGetLastError();
}

// variant: 1 = 4 lines x 3 CTAs/SM (168 registers), staged; 2 = 4 lines x 2 CTAs/SM, staged; 3 = 8 lines x 1 CTA/SM, staged;
// 4..6 = the same shapes with plain loads into registers (no cp.async staging)
cudaError_t launch_pass32(int variant, const PassParams& a, int num_sms, cudaStream_t st) {
    switch (variant) {
        case 1: return launch_pass32_impl<4, 3, true>(a, num_sms, st);
        case 2: return launch_pass32_impl<4, 2, true>(a, num_sms, st);
        case 3: return launch_pass32_impl<8, 1, true>(a, num_sms, st);
        case 4: return launch_pass32_impl<4, 3, false>(a, num_sms, st);
        case 5: return launch_pass32_impl<4, 2, false>(a, num_sms, st);
        case 6: return launch_pass32_impl<8, 1, false>(a, num_sms, st);
    }
    return cudaErrorInvalidValue;
}
int pass32_tile_lines(int variant) { return (variant == 3 || variant == 6) ? 8 : 4; }
}  // namespace gd

#include "fft_fused.cuh"
#include "pass_launch.cuh"
namespace gd {

template <int LOG2L, int T>
static cudaError_t launch_fused_impl(const FusedParams& a, long long total_items, int num_sms, cudaStream_t st) {
    using SH = PassShape<LOG2L>;
    static KernelInfoPerDevice per_dev;
    KernelInfo& info = per_dev.current();
    auto kern = fft_fused_kernel<LOG2L, T>;
    if (!info.ready) {
        info.threads = T * SH::P;
        info.smem = T * line_stride(SH::L, T) * (int)sizeof(cpx);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, info.smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&info.blocks_per_sm, kern, info.threads, info.smem);
        if (e != cudaSuccess) return e;
        if (info.blocks_per_sm < 1) return cudaErrorLaunchOutOfResources;
        info.ready = true;
    }
    // every CTA must be resident at once (CTAs wait on each other's tiles): grid <= SMs * occupancy
    long long cap = (long long)num_sms * info.blocks_per_sm;
    int grid = (int)(total_items < cap ? total_items : cap);
    kern<<<grid, info.threads, info.smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_fused(int log2l, bool wide, const FusedParams& a, long long total_items, int num_sms, cudaStream_t st) {
    if (wide && log2l == 10) return launch_fused_impl<10, 8>(a, total_items, num_sms, st);
    switch (log2l) {
        case 8: return launch_fused_impl<8, 16>(a, total_items, num_sms, st);
        case 9: return launch_fused_impl<9, 8>(a, total_items, num_sms, st);
        case 10: return launch_fused_impl<10, 4>(a, total_items, num_sms, st);
        case 11: return launch_fused_impl<11, 2>(a, total_items, num_sms, st);
        case 12: return launch_fused_impl<12, 1>(a, total_items, num_sms, st);
    }
    return cudaErrorInvalidValue;
}
int fused_tile_lines(int log2l, bool wide) { return pass_tile_lines(log2l, wide && log2l == 10); }
}  // namespace gd

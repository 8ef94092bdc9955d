#include "bluestein_small.cuh"
#include "pass_launch.cuh"
namespace gd {

template <int LOG2L, int T>
static cudaError_t launch_bs(const BluesteinSmallParams& a, int num_sms, cudaStream_t st) {
    using SH = PassShape<LOG2L>;
    static KernelInfoPerDevice per_dev;
    KernelInfo& info = per_dev.current();
    auto kern = bluestein_small_kernel<LOG2L, T>;
    if (!info.ready) {
        info.threads = T * SH::P;
        info.smem = LOG2L > 4 ? T * line_stride(SH::L, T) * (int)sizeof(cpx) : 0;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, info.smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&info.blocks_per_sm, kern, info.threads, info.smem);
        if (e != cudaSuccess) return e;
        if (info.blocks_per_sm < 1) return cudaErrorLaunchOutOfResources;
        info.ready = true;
    }
    const long long ntiles = (a.batch + T - 1) / T;
    if (ntiles <= 0) return cudaSuccess;
    const long long cap = (long long)num_sms * info.blocks_per_sm;
    kern<<<(int)(ntiles < cap ? ntiles : cap), info.threads, info.smem, st>>>(a);
    return cudaGetLastError();
}

// lines per CTA as in pass_tile_lines()
cudaError_t launch_bluestein_small(int log2la, const BluesteinSmallParams& a, int num_sms, cudaStream_t st) {
    switch (log2la) {
        case 1: return launch_bs<1, 128>(a, num_sms, st);
        case 2: return launch_bs<2, 128>(a, num_sms, st);
        case 3: return launch_bs<3, 128>(a, num_sms, st);
        case 4: return launch_bs<4, 128>(a, num_sms, st);
        case 5: return launch_bs<5, 64>(a, num_sms, st);
        case 6: return launch_bs<6, 32>(a, num_sms, st);
        case 7: return launch_bs<7, 16>(a, num_sms, st);
        case 8: return launch_bs<8, 16>(a, num_sms, st);
        case 9: return launch_bs<9, 8>(a, num_sms, st);
        case 10: return launch_bs<10, 4>(a, num_sms, st);
        case 11: return launch_bs<11, 2>(a, num_sms, st);
        case 12: return launch_bs<12, 1>(a, num_sms, st);
        case 13: return launch_bs<13, 1>(a, num_sms, st);      // 512 threads, the 8192-point sequence (139 KiB) in shared memory
    }
    return cudaErrorInvalidValue;
}

}  // namespace gd

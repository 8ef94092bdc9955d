// pass_launch.cuh -- instantiation + launch glue for fft_pass_kernel<LOG2L, T, GENERIC>.
#pragma once
#include "fft_pass.cuh"
#include <cuda_runtime.h>
#include <stdlib.h>

namespace gd {

struct KernelInfo {
    bool ready = false;
    int blocks_per_sm = 0;
    int smem = 0;
    int threads = 0;
};
// The dynamic shared-memory opt-in (cudaFuncSetAttribute) and the occupancy are per DEVICE, and one process may drive
// several GPUs (gd_init(ndev), gd_use_device): one slot per device, filled on first use on that device. Callers hold
// their device's mutex, so a slot is only ever touched by one thread at a time.
constexpr int GD_MAX_DEVICES = 16;
struct KernelInfoPerDevice {
    KernelInfo slot[GD_MAX_DEVICES];
    KernelInfo& current() {
        int dev = 0;
        cudaGetDevice(&dev);
        return slot[dev >= 0 && dev < GD_MAX_DEVICES ? dev : 0];
    }
};

// one launcher per (LOG2L, T); `wide` picks the 512-thread variant where one exists.
typedef cudaError_t (*PassLauncher)(const PassParams&, bool generic, int num_sms, cudaStream_t);

template <int LOG2L, int T, bool GENERIC>
cudaError_t launch_pass_impl(const PassParams& a, int num_sms, cudaStream_t st) {
    using SH = PassShape<LOG2L>;
    static KernelInfoPerDevice per_dev;      // one per instantiation
    KernelInfo& info = per_dev.current();
    auto kern = fft_pass_kernel<LOG2L, T, GENERIC>;
    if (!info.ready) {
        info.threads = T * SH::P;
        info.smem = LOG2L > 4 ? T * line_stride(SH::L, T) * (int)sizeof(cpx) : 0;
        if (const char* ex = getenv("GD_EXTRA_SMEM_KB")) info.smem += atoi(ex) * 1024;   // occupancy experiments
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

template <int LOG2L, int T>
cudaError_t launch_pass_t(const PassParams& a, bool generic, int num_sms, cudaStream_t st) {
    return generic ? launch_pass_impl<LOG2L, T, true>(a, num_sms, st)
                   : launch_pass_impl<LOG2L, T, false>(a, num_sms, st);
}

// defined in pass_inst_*.cu
cudaError_t launch_pass_small(int log2l, const PassParams& a, bool generic, int num_sms, cudaStream_t st);   // 1..8
cudaError_t launch_pass_mid(int log2l, bool wide, const PassParams& a, bool generic, int num_sms, cudaStream_t st);   // 9,10
cudaError_t launch_pass_big(int log2l, bool wide, const PassParams& a, bool generic, int num_sms, cudaStream_t st);   // 11,12

// lines per CTA of the variant the launchers above pick
inline int pass_tile_lines(int log2l, bool wide) {
    switch (log2l) {
        case 1: case 2: case 3: case 4: return 128;
        case 5: return 64; case 6: return 32; case 7: return 16; case 8: return 16; case 9: return 8;
        case 10: return wide ? 8 : 4; case 11: return wide ? 4 : 2; case 12: return wide ? 2 : 1;
    }
    return 0;
}

}  // namespace gd

#include "pass_launch.cuh"
namespace gd {
cudaError_t launch_pass_small(int log2l, const PassParams& a, bool generic, int num_sms, cudaStream_t st) {
    switch (log2l) {
        case 1: return launch_pass_t<1, 128>(a, generic, num_sms, st);
        case 2: return launch_pass_t<2, 128>(a, generic, num_sms, st);
        case 3: return launch_pass_t<3, 128>(a, generic, num_sms, st);
        case 4: return launch_pass_t<4, 128>(a, generic, num_sms, st);
        case 5: return launch_pass_t<5, 64>(a, generic, num_sms, st);
        case 6: return launch_pass_t<6, 32>(a, generic, num_sms, st);
        case 7: return launch_pass_t<7, 16>(a, generic, num_sms, st);
        case 8: return launch_pass_t<8, 16>(a, generic, num_sms, st);
    }
    return cudaErrorInvalidValue;
}
}  // namespace gd

#include "pass_launch.cuh"
namespace gd {
cudaError_t launch_pass_mid(int log2l, bool wide, const PassParams& a, bool generic, int num_sms, cudaStream_t st) {
    switch (log2l) {
        case 9: return launch_pass_t<9, 8>(a, generic, num_sms, st);
        case 10: return wide ? launch_pass_t<10, 8>(a, generic, num_sms, st) : launch_pass_t<10, 4>(a, generic, num_sms, st);
    }
    return cudaErrorInvalidValue;
}
}  // namespace gd

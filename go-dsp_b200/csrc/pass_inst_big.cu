#include "pass_launch.cuh"
namespace gd {
cudaError_t launch_pass_big(int log2l, bool wide, const PassParams& a, bool generic, int num_sms, cudaStream_t st) {
    switch (log2l) {
        case 11: return wide ? launch_pass_t<11, 4>(a, generic, num_sms, st) : launch_pass_t<11, 2>(a, generic, num_sms, st);
        case 12: return wide ? launch_pass_t<12, 2>(a, generic, num_sms, st) : launch_pass_t<12, 1>(a, generic, num_sms, st);
    }
    return cudaErrorInvalidValue;
}
}  // namespace gd

// 2^14-point transforms (N = 128 x 128) through the TMA-fed fused four-step (fft_tma14.cuh): instantiations and entry points
#include "tma14_host.cuh"

namespace gd {
GD_TMA2D_ENTRY(14, 128, 128)
}  // namespace gd

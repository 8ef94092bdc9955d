// 2^15-point transforms (N = 256 x 128) through the TMA-fed fused four-step (fft_tma14.cuh): instantiations and entry points
#include "tma14_host.cuh"

namespace gd {
GD_TMA2D_ENTRY(15, 256, 128)
}  // namespace gd

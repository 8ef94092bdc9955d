// 2^18-point transforms (N = 512 x 512) through the TMA-fed fused four-step (fft_tma14.cuh): instantiations and entry points
#include "tma14_host.cuh"

namespace gd {
GD_TMA2D_ENTRY(18, 512, 512)
}  // namespace gd

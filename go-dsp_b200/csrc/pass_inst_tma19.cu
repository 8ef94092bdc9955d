// 2^19-point transforms (N = 1024 x 512, rows of a batch only) through the TMA-fed fused four-step (fft_tma14.cuh): instantiations and entry points
#include "tma14_host.cuh"

namespace gd {
GD_TMA2D_ENTRY(19, 1024, 512)
}  // namespace gd

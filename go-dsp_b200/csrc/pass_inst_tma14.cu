// 2^14-point transforms (N = 128 x 128) through the TMA-fed fused four-step: instantiations and entry points (tma14_host.cuh)
#include "tma14_host.cuh"

namespace gd {

bool tma14_rows_applicable(const void* in, long long in_dist, const cpx* out, long long out_dist, long long batch, int ld_conj, int st_conj,
                           double scale) {
    return tma2d_rows_applicable<128>(in, in_dist, out, out_dist, batch, ld_conj, st_conj, scale);
}
bool tma14_cols_applicable(const cpx* src, const cpx* dst, long long len, long long s) { return tma2d_cols_applicable<128>(src, dst, len, s); }
Status fft_tma_2p14(Device& d, int mode, const cpx* in, long long in_dist, cpx* out, long long out_dist, long long count, bool inv, double scale,
                    cudaStream_t st) {
    return fft_tma_2d<128>(d, mode, in, in_dist, out, out_dist, count, inv, scale, st);
}

}  // namespace gd

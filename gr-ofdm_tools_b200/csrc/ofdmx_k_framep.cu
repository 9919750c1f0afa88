// ofdmx_k_framep.cu -- the rx_framep_kernel instantiations (fft_len 2048, a pair of warps per frame).
#include "ofdmx_launch.h"
#include "ofdmx_frame2048p.cuh"

template <int B>
static cudaError_t fp_conf(size_t smem)
{
    cudaError_t e = ofdmx_raise_smem_limit(rx_framep_kernel<B, false>, smem);
    cudaError_t e2 = ofdmx_raise_smem_limit(rx_framep_kernel<B, true>, smem);
    return e != cudaSuccess ? e : e2;
}

cudaError_t ofdmx_fp_configure(int bps, size_t smem)
{
    switch (bps) {
    case 1: return fp_conf<1>(smem);
    case 2: return fp_conf<2>(smem);
    case 3: return fp_conf<3>(smem);
    case 4: return fp_conf<4>(smem);
    case 6: return fp_conf<6>(smem);
    default: return cudaErrorInvalidValue;
    }
}

template <int B>
static void fp_go(unsigned grid, size_t smem, cudaStream_t st, const FwArgs &a, const uint16_t *pair_tab, int hsz)
{
    if (a.z_out)
        rx_framep_kernel<B, true><<<grid, FP_THREADS, smem, st>>>(a.kp, a.samples, a.n, a.stride, a.trig, a.trig_stream, a.cfo,
            a.stream_start, a.n_trig, a.spec, a.bytes_out, a.byte_stride, a.z_out, a.z_stride, a.x_2048, pair_tab, hsz);
    else
        rx_framep_kernel<B, false><<<grid, FP_THREADS, smem, st>>>(a.kp, a.samples, a.n, a.stride, a.trig, a.trig_stream, a.cfo,
            a.stream_start, a.n_trig, a.spec, a.bytes_out, a.byte_stride, a.z_out, a.z_stride, a.x_2048, pair_tab, hsz);
}

bool ofdmx_fp_launch(int bps, unsigned grid, size_t smem, cudaStream_t st, const FwArgs &a, const uint16_t *pair_tab, int hsz)
{
    switch (bps) {
    case 1: fp_go<1>(grid, smem, st, a, pair_tab, hsz); return true;
    case 2: fp_go<2>(grid, smem, st, a, pair_tab, hsz); return true;
    case 3: fp_go<3>(grid, smem, st, a, pair_tab, hsz); return true;
    case 4: fp_go<4>(grid, smem, st, a, pair_tab, hsz); return true;
    case 6: fp_go<6>(grid, smem, st, a, pair_tab, hsz); return true;
    default: return false;
    }
}

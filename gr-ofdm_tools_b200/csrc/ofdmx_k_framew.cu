// ofdmx_k_framew.cu -- the rx_framew_kernel instantiations of ONE fft_len (compiled once per fft_len with
// -DOFDMX_FW_N=<fft_len>; see ofdm_tools/build.py).
#include "ofdmx_launch.h"
#include "ofdmx_frame1024w.cuh"

#ifndef OFDMX_FW_N
#error "compile with -DOFDMX_FW_N=<fft_len>"
#endif
#define FW_CAT2(a, b) a##b
#define FW_CAT(a, b) FW_CAT2(a, b)

template <int B>
static cudaError_t fw_conf(size_t smem, int threads, int *occ)
{
    cudaError_t e = ofdmx_raise_smem_limit(rx_framew_kernel<OFDMX_FW_N, B, false>, smem);
    cudaError_t e2 = ofdmx_raise_smem_limit(rx_framew_kernel<OFDMX_FW_N, B, true>, smem);
    if (e == cudaSuccess) e = e2;
    if (e == cudaSuccess && occ) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, rx_framew_kernel<OFDMX_FW_N, B, false>, threads, smem);
    return e;
}

cudaError_t FW_CAT(ofdmx_fw_configure_, OFDMX_FW_N)(int bps, size_t smem, int threads, int *occ)
{
    switch (bps) {
    case 1: return fw_conf<1>(smem, threads, occ);
    case 2: return fw_conf<2>(smem, threads, occ);
    case 3: return fw_conf<3>(smem, threads, occ);
    case 4: return fw_conf<4>(smem, threads, occ);
    case 6: return fw_conf<6>(smem, threads, occ);
    default: return cudaErrorInvalidValue;
    }
}

template <int B>
static void fw_go(unsigned grid, unsigned threads, size_t smem, cudaStream_t st, const FwArgs &a)
{
    if (a.z_out)
        rx_framew_kernel<OFDMX_FW_N, B, true><<<grid, threads, smem, st>>>(a.kp, a.samples, a.n, a.stride, a.trig, a.trig_stream, a.cfo,
            a.stream_start, a.n_trig, a.spec, a.bytes_out, a.byte_stride, a.z_out, a.z_stride, a.x_2048, a.dec_off, a.dec_all);
    else
        rx_framew_kernel<OFDMX_FW_N, B, false><<<grid, threads, smem, st>>>(a.kp, a.samples, a.n, a.stride, a.trig, a.trig_stream, a.cfo,
            a.stream_start, a.n_trig, a.spec, a.bytes_out, a.byte_stride, a.z_out, a.z_stride, a.x_2048, a.dec_off, a.dec_all);
}

bool FW_CAT(ofdmx_fw_launch_, OFDMX_FW_N)(int bps, unsigned grid, unsigned threads, size_t smem, cudaStream_t st, const FwArgs &a)
{
    switch (bps) {
    case 1: fw_go<1>(grid, threads, smem, st, a); return true;
    case 2: fw_go<2>(grid, threads, smem, st, a); return true;
    case 3: fw_go<3>(grid, threads, smem, st, a); return true;
    case 4: fw_go<4>(grid, threads, smem, st, a); return true;
    case 6: fw_go<6>(grid, threads, smem, st, a); return true;
    default: return false;
    }
}

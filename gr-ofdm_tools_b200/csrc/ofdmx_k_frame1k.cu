// ofdmx_k_frame1k.cu -- the rx_frame1024_kernel instantiations (CTA per frame, fft_len 1024: general carrier plans).
#include "ofdmx_launch.h"
#include "ofdmx_frame1024.cuh"

template <int B>
static cudaError_t f1k_conf(size_t smem)
{
    cudaError_t e = cudaSuccess, e2;
#define F1K_ATTR(S, Z) e2 = ofdmx_raise_smem_limit(rx_frame1024_kernel<B, S, Z>, smem); if (e2 != cudaSuccess) e = e2;
    F1K_ATTR(true, false) F1K_ATTR(true, true) F1K_ATTR(false, false) F1K_ATTR(false, true)
#undef F1K_ATTR
    return e;
}

cudaError_t ofdmx_f1k_configure(int bps, size_t smem)
{
    switch (bps) {
    case 1: return f1k_conf<1>(smem);
    case 2: return f1k_conf<2>(smem);
    case 3: return f1k_conf<3>(smem);
    case 4: return f1k_conf<4>(smem);
    case 6: return f1k_conf<6>(smem);
    default: return cudaErrorInvalidValue;
    }
}

template <int B, bool S, bool Z>
static void f1k_go3(unsigned grid, size_t smem, cudaStream_t st, const F1kArgs &a)
{
    rx_frame1024_kernel<B, S, Z><<<grid, F1K_THREADS, smem, st>>>(a.kp, a.warps, a.samples, a.n, a.stride, a.trig, a.trig_stream, a.cfo,
        a.stream_start, a.n_trig, a.spec, a.bytes_out, a.byte_stride, a.z_out, a.z_stride);
}
template <int B>
static void f1k_go(bool simple, unsigned grid, size_t smem, cudaStream_t st, const F1kArgs &a)
{
    if (simple) { if (a.z_out) f1k_go3<B, true, true>(grid, smem, st, a); else f1k_go3<B, true, false>(grid, smem, st, a); }
    else { if (a.z_out) f1k_go3<B, false, true>(grid, smem, st, a); else f1k_go3<B, false, false>(grid, smem, st, a); }
}

void ofdmx_f1k_launch(int bps, bool simple, unsigned grid, size_t smem, cudaStream_t st, const F1kArgs &a)
{
    switch (bps) {
    case 1: f1k_go<1>(simple, grid, smem, st, a); break;
    case 2: f1k_go<2>(simple, grid, smem, st, a); break;
    case 3: f1k_go<3>(simple, grid, smem, st, a); break;
    case 4: f1k_go<4>(simple, grid, smem, st, a); break;
    default: f1k_go<6>(simple, grid, smem, st, a); break;
    }
}

// ofdmx_launch.h -- internal interface between ofdmx_api.cu (host logic + the small kernels) and the
// translation units that hold the big template kernels.  One object per (kernel, fft_len): they build in
// parallel and a change to one kernel recompiles one object.  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ofdmx.h"
#include "ofdmx_dev.cuh"

#pragma GCC visibility push(hidden)     // internal: not exported from libofdmx.so

#include <map>
#include <mutex>
// cudaFuncAttributeMaxDynamicSharedMemorySize is per function and process-wide: several contexts (or two plans of
// one context) with different needs share it, so it only ever grows -- a launch may use less than the limit.
template <typename F>
static inline cudaError_t ofdmx_raise_smem_limit(F *func, size_t bytes)
{
    static std::mutex mu;
    static std::map<const void *, size_t> high;
    std::lock_guard<std::mutex> lk(mu);
    size_t &h = high[(const void *)func];
    if (bytes <= h) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) h = bytes;
    return e;
}

// ---- rx_framew_kernel<NFFT, BPS, WANT_Z> (ofdmx_frame1024w.cuh): one warp per frame
struct FwArgs {
    KP kp;
    const float2 *samples;
    long long n, stride;
    const long long *trig;
    const int *trig_stream;
    const float *cfo;
    const int *stream_start;
    const int *n_trig;
    ofdmx_frame *spec;
    uint8_t *bytes_out;
    long long byte_stride;
    float2 *z_out;
    long long z_stride;
    uint32_t x_2048;
    int dec_off, dec_all;
};
// sets the dynamic shared-memory limit of the <NFFT, bps, *> instantiations; *occ = resident CTAs per SM
cudaError_t ofdmx_fw_configure(int nfft, int bps, size_t smem, int threads, int *occ);
// false: no instantiation for (nfft, bps)
bool ofdmx_fw_launch(int nfft, int bps, unsigned grid, unsigned threads, size_t smem, cudaStream_t st, const FwArgs &a);
#define OFDMX_FW_DECL(N)                                                                   \
    cudaError_t ofdmx_fw_configure_##N(int bps, size_t smem, int threads, int *occ);       \
    bool ofdmx_fw_launch_##N(int bps, unsigned grid, unsigned threads, size_t smem, cudaStream_t st, const FwArgs &a);
OFDMX_FW_DECL(64) OFDMX_FW_DECL(128) OFDMX_FW_DECL(256) OFDMX_FW_DECL(512) OFDMX_FW_DECL(1024) OFDMX_FW_DECL(2048)
#undef OFDMX_FW_DECL

// ---- rx_framep_kernel<BPS, WANT_Z> (ofdmx_frame2048p.cuh): fft_len 2048, a pair of warps per frame
cudaError_t ofdmx_fp_configure(int bps, size_t smem);
bool ofdmx_fp_launch(int bps, unsigned grid, size_t smem, cudaStream_t st, const FwArgs &a, const uint16_t *pair_tab, int hsz);

// ---- rx_frame1024_kernel<BPS, SIMPLE, WANT_Z> (ofdmx_frame1024.cuh): one CTA per frame, fft_len 1024
struct F1kArgs {
    KP kp;
    int warps;
    const float2 *samples;
    long long n, stride;
    const long long *trig;
    const int *trig_stream;
    const float *cfo;
    const int *stream_start;
    const int *n_trig;
    ofdmx_frame *spec;
    uint8_t *bytes_out;
    long long byte_stride;
    float2 *z_out;
    long long z_stride;
};
cudaError_t ofdmx_f1k_configure(int bps, size_t smem);
void ofdmx_f1k_launch(int bps, bool simple, unsigned grid, size_t smem, cudaStream_t st, const F1kArgs &a);

// ---- tx_framew_kernel<NFFT, BPS> (ofdmx_tx1024w.cuh): one warp per packet
struct TxwArgs {
    KP kp;
    const uint8_t *payload;
    const long long *pkt_off;
    long long n_pkts;
    int first_num;
    float2 *out;
    long long cap;
    const long long *sample_off;
    const uint16_t *tx_map;
    const float2 *sync_td;
    uint32_t x_2048;
    int pb_bytes;
};
cudaError_t ofdmx_txw_configure(int nfft, int bps, size_t smem);
bool ofdmx_txw_launch(int nfft, int bps, unsigned grid, unsigned threads, size_t smem, cudaStream_t st, const TxwArgs &a);
#define OFDMX_TXW_DECL(N)                                                                  \
    cudaError_t ofdmx_txw_configure_##N(int bps, size_t smem);                             \
    bool ofdmx_txw_launch_##N(int bps, unsigned grid, unsigned threads, size_t smem, cudaStream_t st, const TxwArgs &a);
OFDMX_TXW_DECL(64) OFDMX_TXW_DECL(128) OFDMX_TXW_DECL(256) OFDMX_TXW_DECL(512) OFDMX_TXW_DECL(1024)
#undef OFDMX_TXW_DECL

#pragma GCC visibility pop

// ofdmx_k_txw.cu -- the tx_framew_kernel instantiations of ONE fft_len (compiled once per fft_len with
// -DOFDMX_TXW_N=<fft_len>; see ofdm_tools/build.py).
#include "ofdmx_launch.h"
#include "ofdmx_tx1024w.cuh"

#ifndef OFDMX_TXW_N
#error "compile with -DOFDMX_TXW_N=<fft_len>"
#endif
#define TXW_CAT2(a, b) a##b
#define TXW_CAT(a, b) TXW_CAT2(a, b)

cudaError_t TXW_CAT(ofdmx_txw_configure_, OFDMX_TXW_N)(int bps, size_t smem)
{
#define TXW_ATTR(B) case B: { cudaError_t e = ofdmx_raise_smem_limit(tx_framew_kernel<OFDMX_TXW_N, B, false>, smem); \
                               cudaError_t e2 = ofdmx_raise_smem_limit(tx_framew_kernel<OFDMX_TXW_N, B, true>, smem); return e != cudaSuccess ? e : e2; }
    switch (bps) {
    TXW_ATTR(1) TXW_ATTR(2) TXW_ATTR(3) TXW_ATTR(4) TXW_ATTR(6)
    default: return cudaErrorInvalidValue;
    }
#undef TXW_ATTR
}

bool TXW_CAT(ofdmx_txw_launch_, OFDMX_TXW_N)(int bps, unsigned grid, unsigned threads, size_t smem, cudaStream_t st, const TxwArgs &a)
{
#define TXW_GO(B) case B: if (a.kp.roll) tx_framew_kernel<OFDMX_TXW_N, B, true><<<grid, threads, smem, st>>>(a.kp, a.payload, a.pkt_off, a.n_pkts, a.first_num, a.out, \
        a.cap, a.sample_off, a.tx_map, a.sync_td, a.x_2048, a.pb_bytes); \
        else tx_framew_kernel<OFDMX_TXW_N, B, false><<<grid, threads, smem, st>>>(a.kp, a.payload, a.pkt_off, a.n_pkts, a.first_num, a.out, \
        a.cap, a.sample_off, a.tx_map, a.sync_td, a.x_2048, a.pb_bytes); return true;
    switch (bps) {
    TXW_GO(1) TXW_GO(2) TXW_GO(3) TXW_GO(4) TXW_GO(6)
    default: return false;
    }
#undef TXW_GO
}

// ofdmx_cond.cuh -- signal conditioning around the OFDM chain.
//
// analog.agc2_cc in front of ofdm_rx (python/ofdm_tx_rx_hier.py:75-76, python/ofdm_radio_hier.py:180-181) is a
// per-sample NON-LINEAR recurrence on the loop gain:
//     out = in * g;  e = |out| - reference;  g -= e * (e > g ? attack : decay);  clamp g
// so a stream cannot be split without changing its results.  The parallel axis is the stream: one lane per
// stream, 32 streams per warp.  Streams are rows of the sample matrix, so a warp moves tiles of 32 streams x
// 32 samples through shared memory: coalesced 256-byte row segments on the global side, a conflict-free
// transposed walk (row stride 33 float2) on the recurrence side.  Arithmetic: float32 with one rounding per
// operation in GNU Radio's order (no FMA contraction), IEEE square root -- bit-exact against the oracle.
#pragma once
#include "ofdmx_dev.cuh"

#define AGC_WARPS 4

__global__ void __launch_bounds__(AGC_WARPS * 32)
agc2_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, long long n, long long stride, int n_streams,
            float attack, float decay, float reference, float max_gain, float *__restrict__ gain_io)
{
    __shared__ float2 tile[AGC_WARPS][32][33];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s0 = (blockIdx.x * AGC_WARPS + w) * 32;
    if (s0 >= n_streams) return;
    const int ns = min(32, n_streams - s0);
    float g = (lane < ns) ? gain_io[s0 + lane] : 1.0f;
    const float2 *src = in + (long long)s0 * stride + lane;
    float2 *dst = out + (long long)s0 * stride + lane;
    // the next tile travels in registers while the recurrence walks the current one (the walk is a chain of
    // dependent float operations ~180 cycles per sample long: nothing else hides the global latency)
    float2 nxt[32];
#pragma unroll
    for (int r = 0; r < 32; r++)
        nxt[r] = (r < ns && lane < n) ? __ldcs(src + (long long)r * stride) : make_float2(0.f, 0.f);
    for (long long i0 = 0; i0 < n; i0 += 32) {
        const int nc = (int)min((long long)32, n - i0);
#pragma unroll
        for (int r = 0; r < 32; r++) tile[w][r][lane] = nxt[r];
        __syncwarp();
        const long long i1 = i0 + 32;
#pragma unroll
        for (int r = 0; r < 32; r++)
            if (r < ns && i1 + lane < n) nxt[r] = __ldcs(src + (long long)r * stride + i1);
        if (lane < ns) {
#pragma unroll 4
            for (int k = 0; k < nc; k++) {
                const float2 x = tile[w][lane][k];
                const float re = __fmul_rn(x.x, g), im = __fmul_rn(x.y, g);
                tile[w][lane][k] = make_float2(re, im);
                const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
                const float tmp = __fadd_rn(-reference, mag);
                const float rate = (tmp > g) ? attack : decay;
                g = __fsub_rn(g, __fmul_rn(tmp, rate));
                if (g < 0.0f) g = 10e-5f;
                if (max_gain > 0.0f && g > max_gain) g = max_gain;
            }
        }
        __syncwarp();
        if (lane < nc) {
#pragma unroll 8
            for (int r = 0; r < ns; r++) __stcs(dst + (long long)r * stride + i0, tile[w][r][lane]);
        }
        __syncwarp();
    }
    if (lane < ns) gain_io[s0 + lane] = g;
}

// ofdmx_sync_warpn.cuh -- K2, fft_len 32 .. 512: Schmidl & Cox metric with WARP-AUTONOMOUS streaming.
//
// The short-window counterpart of ofdmx_sync_warp.cuh.  On frame-dense streams with a 32-sample window the
// chunk-level rejection of the other kernels never fires, so every sample pays the sliding pass and the kernel is
// bound by instruction issue (ncu, configs[1] on the TMA ring kernel: 74 thread-instructions per sample at 70 %
// issue utilisation, a third of them in the block-wide float64 scan, its barriers and the ring bookkeeping).
// For fft_len <= 512 the whole window (fft_len samples = at most 32 chunks of 16) lies inside the current and
// the previous tile of a warp, which removes the scan altogether:
//   * a warp walks its own span of one stream in tiles of 512 samples = 32 lanes x 16 samples (same ring of 4 KB
//     slots filled by cp.async as the fft_len 1024 kernel; one warm-up tile per span);
//   * the window sums at the start of a lane's chunk are sums of the HC = fft_len/32 (products) and
//     NC = fft_len/16 (energies) chunk totals in front of it: log2 doubling steps over the lanes
//     (W2(l) = T(l) + T(l-1), W4(l) = W2(l) + W2(l-2), ...), lanes near the start of the tile taking the
//     previous tile's partial windows from registers carried over.  All float32, and every partial sum covers
//     terms of the window itself only, so the local error bound of ofdmx_sync.cuh holds with room to spare
//     (16-term chunk totals + <= 5 doubling adds + 16 sliding updates against a budget of 100 ulp);
//   * products and energies of the chunk stay in registers between the totals and the sliding pass.
// Same filtered predicate and the same exact float64 re-evaluation as the other sync kernels: the detect bits do
// not depend on the summation order.
//
// Preconditions (host): fft_len a power of two in 32 .. 512, sample pointer 16-byte aligned, even stream stride.
#pragma once
#include "ofdmx_sync_warp.cuh"

template <int LOG>   // sum of the `1 << LOG` values ending at this lane (virtual index: previous tile below lane 0)
struct SwnWin {
    // lv[j]: windows of 2^j chunks ending at each lane of the current tile; pv[j]: the same of the previous tile
    __device__ __forceinline__ static void build(float t, float (&lv)[6], const float (&pv)[6], int lane)
    {
#pragma unroll
        for (int j = 0; j < 6; j++) lv[j] = 0.f;
        lv[0] = t;
#pragma unroll
        for (int j = 0; j < LOG; j++) {
            const int o = 1 << j;
            const float up = __shfl_up_sync(0xffffffffu, lv[j], o);
            const float pr = __shfl_sync(0xffffffffu, pv[j], (lane - o) & 31);
            lv[j + 1] = lv[j] + (lane >= o ? up : pr);
        }
    }
};

template <int N>
__global__ void __launch_bounds__(SW_WARPS * 32, 1)
sync_metric_warpn_kernel(const float2 *__restrict__ samples, long long n, long long stride, float thr_f, double thr_d,
                         uint32_t *__restrict__ detmask, uint32_t *__restrict__ trigmask, long long wps,
                         int tiles_per_stream, int span, int spans_per_stream, int total_spans)
{
    static_assert(N >= 32 && N <= 512 && (N & (N - 1)) == 0, "fft_len 32 .. 512, power of two");
    constexpr int HC = N / 32, NC = N / 16;                // chunks per half window / per window
    constexpr int LX = (HC == 1) ? 0 : (HC == 2) ? 1 : (HC == 4) ? 2 : (HC == 8) ? 3 : 4;
    constexpr int LE = LX + 1;
    extern __shared__ __align__(128) unsigned char sw_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char *ring = sw_smem + (size_t)wid * SW_RING_BYTES;
    const float thr4 = 0.25f * thr_f;
    const int gw = blockIdx.x * (blockDim.x >> 5) + wid, nw = gridDim.x * (blockDim.x >> 5);
    const int full_tiles = (int)(n / SW_TILE);
    const int off_e = ((lane >> 3) << 7) + (((lane & 7) ^ (lane >> 3)) << 4);
    const int off_o = (((lane >> 3) + 4) << 7) + (((lane & 7) ^ ((lane >> 3) + 4)) << 4);
    // rows of the chunks fft_len/2 and fft_len behind this lane's: same tile, or the previous one for the first lanes
    const bool dprev = lane < HC, nprev = lane < NC;
    const int rowd = (lane - HC) & 31, rown = (lane - NC) & 31;
    const int sown = (lane & 7) << 4, sd = (rowd & 7) << 4, sn = (rown & 7) << 4;

    for (int sp = gw; sp < total_spans; sp += nw) {
        const int s = sp / spans_per_stream;
        const int k0 = (sp - s * spans_per_stream) * span;
        const int k1 = min(k0 + span, tiles_per_stream);
        const float2 *r = samples + (long long)s * stride;
        float pxr[6], pxi[6], pe[6];                       // previous tile's windows of 1, 2, 4, .. chunks per lane
#pragma unroll
        for (int j = 0; j < 6; j++) { pxr[j] = 0.f; pxi[j] = 0.f; pe[j] = 0.f; }
        __syncwarp();
        // slot 3 stands for "the tile before the warm-up tile": its samples only feed sums that are never used,
        // but keep the arithmetic on defined data
#pragma unroll
        for (int q = 0; q < 8; q++) *reinterpret_cast<float4 *>(ring + (3 << 12) + (lane << 7) + (q << 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
        sw_fill(ring, 0, r, n, k0 - 1, full_tiles, lane, off_e, off_o);
        int it = 0;
        for (int k = k0 - 1; k < k1; k++, it++) {
            const int sc = it & 3, s1 = (it + 3) & 3;      // current tile, previous tile
            if (k + 1 < k1) {
                sw_fill(ring, (it + 1) & 3, r, n, k + 1, full_tiles, lane, off_e, off_o);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncwarp();
            const unsigned char *po = ring + (sc << 12) + (lane << 7);
            const unsigned char *pd = ring + ((dprev ? s1 : sc) << 12) + (rowd << 7);
            // ---- products and energies of this lane's 16 samples, and their totals
            float2 x[SV_C];
            float e[SV_C];
            float sxr = 0.f, sxi = 0.f, se = 0.f;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float4 a = *reinterpret_cast<const float4 *>(po + ((q << 4) ^ sown));
                const float4 b = *reinterpret_cast<const float4 *>(pd + ((q << 4) ^ sd));
                x[2 * q] = make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -(a.x * b.y)));
                x[2 * q + 1] = make_float2(fmaf(a.z, b.z, a.w * b.w), fmaf(a.w, b.z, -(a.z * b.w)));
                e[2 * q] = fmaf(a.x, a.x, a.y * a.y);
                e[2 * q + 1] = fmaf(a.z, a.z, a.w * a.w);
                sxr += x[2 * q].x; sxi += x[2 * q].y; se += e[2 * q];
                sxr += x[2 * q + 1].x; sxi += x[2 * q + 1].y; se += e[2 * q + 1];
            }
            // ---- windows of 2, 4, .. chunks ending at each lane
            float cxr[6], cxi[6], ce[6];
            SwnWin<LX>::build(sxr, cxr, pxr, lane);
            SwnWin<LX>::build(sxi, cxi, pxi, lane);
            SwnWin<LE>::build(se, ce, pe, lane);

            if (k >= k0) {
                // window sums at the start of this chunk = the windows ending at the chunk in front of it
                float Pr, Pi, E;
                {
                    const float u0 = __shfl_up_sync(0xffffffffu, cxr[LX], 1), v0 = __shfl_sync(0xffffffffu, pxr[LX], 31);
                    const float u1 = __shfl_up_sync(0xffffffffu, cxi[LX], 1), v1 = __shfl_sync(0xffffffffu, pxi[LX], 31);
                    const float u2 = __shfl_up_sync(0xffffffffu, ce[LE], 1), v2 = __shfl_sync(0xffffffffu, pe[LE], 31);
                    Pr = lane ? u0 : v0; Pi = lane ? u1 : v1; E = lane ? u2 : v2;
                }
                float cD, cN;                                   // energies of the chunks fft_len/2 and fft_len behind
                {
                    const float u0 = __shfl_sync(0xffffffffu, se, rowd), v0 = __shfl_sync(0xffffffffu, pe[0], rowd);
                    const float u1 = __shfl_sync(0xffffffffu, se, rown), v1 = __shfl_sync(0xffffffffu, pe[0], rown);
                    cD = dprev ? v0 : u0; cN = nprev ? v1 : u1;
                }
                const float cJ = se;
                const float A = E + cJ + cD + cN;                // local bound (see ofdmx_sync.cuh)
                const float eps = 6.0e-6f * A;
                const float e3 = 3.5f * eps, e33 = 3.0f * eps * eps;
                unsigned det = 0, unc = 0;
                {
                    const unsigned char *pn = ring + ((nprev ? s1 : sc) << 12) + (rown << 7);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const float4 b = *reinterpret_cast<const float4 *>(pd + ((q << 4) ^ sd));   // r[n - N/2]
                        const float4 c = *reinterpret_cast<const float4 *>(pn + ((q << 4) ^ sn));   // r[n - N]
#pragma unroll
                        for (int t2 = 0; t2 < 2; t2++) {
                            const float br = t2 ? b.z : b.x, bi = t2 ? b.w : b.y, cr = t2 ? c.z : c.x, ci = t2 ? c.w : c.y;
                            const float xdr = fmaf(br, cr, bi * ci), xdi = fmaf(bi, cr, -(br * ci));
                            const float ed = fmaf(cr, cr, ci * ci);
                            const int kk = 2 * q + t2;
                            Pr += x[kk].x - xdr;
                            Pi += x[kk].y - xdi;
                            E += e[kk] - ed;
                            const float d = fmaf(Pr, Pr, Pi * Pi) - thr4 * E * E;
                            const float aE = fabsf(E);
                            const float err = fmaf(aE, fmaf(5.0e-7f, aE, e3), e33);
                            if (d > err) det |= 1u << kk;
                            if (fabsf(d) <= err) unc |= 1u << kk;
                        }
                    }
                }
                if (A == 0.0f) { det = 0; unc = 0; }
                if (k >= full_tiles) {
                    const long long firsts = ((long long)k * 32 + lane) * SV_C;
                    if (firsts + SV_C > n) {
                        const int valid = (n > firsts) ? (int)(n - firsts) : 0;
                        const unsigned m = (valid >= 16) ? 0xffffu : ((1u << valid) - 1u);
                        det &= m; unc &= m;
                    }
                }
                // ---- exact re-evaluation (float64, whole warp per sample) of the samples inside the band
                unsigned pending = __ballot_sync(0xffffffffu, unc != 0);
                while (pending) {
                    const int src = __ffs(pending) - 1;
                    pending &= pending - 1;
                    unsigned m = __shfl_sync(0xffffffffu, unc, src);
                    while (m) {
                        const int kk = __ffs(m) - 1;
                        m &= m - 1;
                        const int i = (src << 4) + kk;                 // sample position relative to the tile start
                        double sr = 0.0, si = 0.0, sen = 0.0;
                        for (int t2 = lane; t2 < N; t2 += 32) {
                            const int sa = i - t2;                     // -(N-1) .. 511
                            const int wa = sa & 511;
                            const float2 a = *reinterpret_cast<const float2 *>(
                                ring + sw_unit((it + 4 + (sa >> 9)) & 3, wa >> 4, (wa >> 1) & 7) + ((wa & 1) << 3));
                            sen += (double)a.x * a.x + (double)a.y * a.y;
                            if (t2 < N / 2) {
                                const int sb = sa - N / 2;
                                const int wb = sb & 511;
                                const float2 b = *reinterpret_cast<const float2 *>(
                                    ring + sw_unit((it + 4 + (sb >> 9)) & 3, wb >> 4, (wb >> 1) & 7) + ((wb & 1) << 3));
                                sr += (double)a.x * b.x + (double)a.y * b.y;
                                si += (double)a.y * b.x - (double)a.x * b.y;
                            }
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            sr += __shfl_xor_sync(0xffffffffu, sr, o);
                            si += __shfl_xor_sync(0xffffffffu, si, o);
                            sen += __shfl_xor_sync(0xffffffffu, sen, o);
                        }
                        const double R = 0.5 * sen, R2 = R * R, pm2 = sr * sr + si * si;
                        const bool dd = (R2 > 0.0) && (pm2 >= thr_d * R2);
                        if (lane == src) det = dd ? (det | (1u << kk)) : (det & ~(1u << kk));
                    }
                }
                // ---- 16 bits per lane -> 32-bit words
                const unsigned hi = __shfl_down_sync(0xffffffffu, det, 1);
                if (!(lane & 1)) {
                    const long long w = (long long)k * 16 + (lane >> 1);
                    if (w < wps) {
                        detmask[(long long)s * wps + w] = (det & 0xffffu) | (hi << 16);
                        trigmask[(long long)s * wps + w] = 0u;         // cleared here: saves a memset pass
                    }
                }
            }
            // ---- carry: this tile becomes the previous one
#pragma unroll
            for (int j = 0; j < 6; j++) { pxr[j] = cxr[j]; pxi[j] = cxi[j]; pe[j] = ce[j]; }
            __syncwarp();      // every lane is done with the slot the next iteration's cp.async overwrites
        }
    }
}

// ofdmx_sync_warp.cuh -- K2, fft_len 1024 and 2048: Schmidl & Cox metric with WARP-AUTONOMOUS streaming.
//
// The TMA ring kernel (ofdmx_sync_tma.cuh) spends a fifth of its warp time at block barriers: eight warps
// share a 4096-sample tile, exchange chunk totals through shared memory and scan them as a block.  For
// fft_len 1024 the structure of the problem removes every cross-warp dependency:
//   * a warp walks its own span of the stream in tiles of 512 samples = fft_len/2 = 32 lanes x 16 samples;
//   * the sample fft_len/2 behind a lane's chunk is the SAME lane's chunk of the previous tile, the sample
//     fft_len behind it the same lane's chunk two tiles back;
//   * the window sums at the start of a chunk are "suffix of the previous tile(s) from this lane on + prefix of
//     this tile up to this lane": one warp-shuffle scan per tile, the suffixes are carried in registers.
// So there is no block barrier, no shared scan array and no halo exchange: each warp owns a private ring of four
// 4 KB tile slots (current, fft_len/2 back, fft_len back, one being filled by cp.async) and only ever executes
// __syncwarp.  Same arithmetic and the same filtered predicate as ofdmx_sync.cuh: float32 window sums with an
// error bound, chunk-level rejection, exact float64 re-evaluation of the samples inside the uncertainty band --
// the detect bits are exact whatever the summation order.
//
// The kernel is a template of the chunk size C (samples per lane and tile): C = 16 is fft_len 1024 (tiles of 512
// samples, 4 KB slots, 14 warps per SM), C = 32 is fft_len 2048 (tiles of 1024 samples, 8 KB slots, 7 warps per SM --
// the same 56 KB of samples in flight per SM); fft_len is always 64 C, so the tile is half a window in both.
//
// Preconditions (host): fft_len == 64 C, sample pointer 16-byte aligned, even stream stride.
#pragma once
#include "ofdmx_sync.cuh"

#define SW_WARPS 14
#define SW_TILE 512                    // samples per warp tile (C = 16)
#define SW_SLOT_BYTES 4096
#define SW_RING_BYTES (4 * SW_SLOT_BYTES)
#define SW_SMEM_BYTES (SW_WARPS * SW_RING_BYTES)   // per CTA, whatever C: warps per CTA = SW_WARPS * 16 / C

// byte offset, inside a warp's ring, of 16-byte unit q of row `row` of slot `slot` (rows are 128 bytes = 16
// samples; units XOR-swizzled by the row so that 16-byte reads of one row per lane are conflict free)
__device__ __forceinline__ int sw_unit(int slot, int row, int q) { return (slot << 12) + (row << 7) + ((q ^ (row & 7)) << 4); }

__device__ __forceinline__ void sw_cp_async16(uint32_t dst, const void *src, int src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

// fill slot `slot` with tile k of the stream (samples [512 k, 512 k + 512)); out-of-range samples become zeros.
// Unit u = 32 i + lane of the tile (16 bytes = samples 2u, 2u+1) lands in row 4 i + lane/8, unit lane%8: the
// swizzled offset depends on i only through its parity, so the two per-lane offsets are computed once (off_e for
// even i, off_o for odd i) and an interior tile costs one address add per copy.
__device__ __forceinline__ void sw_fill(unsigned char *ring, int slot, const float2 *__restrict__ r, long long n, int k,
                                        int full_tiles, int lane, int off_e, int off_o)
{
    const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(ring) + (slot << 12);
    if (k >= 0 && k < full_tiles) {
        const float2 *src = r + (long long)k * SW_TILE + 2 * lane;
#pragma unroll
        for (int i = 0; i < 8; i++) sw_cp_async16(d0 + (i >> 1) * 1024 + ((i & 1) ? off_o : off_e), src + 64 * i, 16);
    } else {
        const long long base = (long long)k * SW_TILE;
#pragma unroll 1
        for (int i = 0; i < 8; i++) {
            const long long idx = base + 64 * i + 2 * lane;
            int bytes = 0;
            if (idx >= 0 && idx < n) bytes = (idx + 1 < n) ? 16 : 8;
            sw_cp_async16(d0 + (i >> 1) * 1024 + ((i & 1) ? off_o : off_e), r + (bytes ? idx : 0), bytes);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// ---- chunk size C: rows of 8 C bytes, slots of 256 C bytes, tiles of 32 C samples
template <int C>
__device__ __forceinline__ int swc_unit(int slot, int row, int q) { return slot * (256 * C) + row * (8 * C) + ((q ^ (row & 7)) << 4); }

// fill slot `slot` with tile k of the stream (samples [32 C k, 32 C (k + 1))); out-of-range samples become zeros.
// Unit u = 32 i + lane of the tile (i < C/2) lands in row u / (C/2), unit u % (C/2), swizzled by the row.  For
// C = 16 this is sw_fill; for C = 32 the row is 2 i + lane/16, so the swizzled offset depends on i through i % 4:
// four per-lane offsets (offs[]) are computed once and an interior tile costs one address add per copy.
template <int C>
__device__ __forceinline__ void swc_fill(unsigned char *ring, int slot, const float2 *__restrict__ r, long long n, int k,
                                         int full_tiles, int lane, const int (&offs)[4])
{
    constexpr int TILE = 32 * C, NCP = C / 2;              // copies of 16 bytes per lane
    const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(ring) + slot * (256 * C);
    // i -> destination: C = 16: offs[i & 1] + (i >> 1) * 1024 (8 rows of 128 bytes per pair of copies);
    //                   C = 32: offs[i & 3] + (i >> 2) * 2048 (8 rows of 256 bytes per four copies)
    if (k >= 0 && k < full_tiles) {
        const float2 *src = r + (long long)k * TILE + 2 * lane;
#pragma unroll
        for (int i = 0; i < NCP; i++)
            sw_cp_async16(d0 + (C == 16 ? (i >> 1) * 1024 + offs[i & 1] : (i >> 2) * 2048 + offs[i & 3]), src + 64 * i, 16);
    } else {
        const long long base = (long long)k * TILE;
#pragma unroll 1
        for (int i = 0; i < NCP; i++) {
            const long long idx = base + 64 * i + 2 * lane;
            int bytes = 0;
            if (idx >= 0 && idx < n) bytes = (idx + 1 < n) ? 16 : 8;
            sw_cp_async16(d0 + (C == 16 ? (i >> 1) * 1024 + offs[i & 1] : (i >> 2) * 2048 + offs[i & 3]), r + (bytes ? idx : 0), bytes);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int C>
__global__ void __launch_bounds__(SW_WARPS * 16 / C * 32, 1)
sync_metric_warp_kernel(const float2 *__restrict__ samples, long long n, long long stride, float thr_f, double thr_d,
                        uint32_t *__restrict__ detmask, uint32_t *__restrict__ trigmask, long long wps,
                        int tiles_per_stream, int span, int spans_per_stream, int total_spans)
{
    static_assert(C == 16 || C == 32, "chunk of 16 (fft_len 1024) or 32 (fft_len 2048) samples");
    constexpr int TILE = 32 * C, LT = (C == 16) ? 9 : 10, NU = C / 2, ROWB = 8 * C, SLOTB = 256 * C, N = 64 * C;
    constexpr unsigned FULL = (C == 32) ? 0xffffffffu : 0xffffu;
    extern __shared__ __align__(128) unsigned char sw_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char *ring = sw_smem + (size_t)wid * (4 * SLOTB);
    const float thr4 = 0.25f * thr_f;
    const int gw = blockIdx.x * (blockDim.x >> 5) + wid, nw = gridDim.x * (blockDim.x >> 5);
    const int full_tiles = (int)(n / TILE);                // tiles that lie completely inside the stream
    const int sown = (lane & 7) << 4;                      // swizzle of this lane's row
    int offs[4];
    if (C == 16) {
        offs[0] = ((lane >> 3) << 7) + (((lane & 7) ^ (lane >> 3)) << 4);              // rows 8j + lane/8
        offs[1] = (((lane >> 3) + 4) << 7) + (((lane & 7) ^ ((lane >> 3) + 4)) << 4);  // rows 8j + 4 + lane/8
        offs[2] = offs[0]; offs[3] = offs[1];
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int row = 2 * j + (lane >> 4);                                       // rows 8m + 2j + lane/16
            offs[j] = row * ROWB + (((lane & 15) ^ (row & 7)) << 4);
        }
    }

    for (int sp = gw; sp < total_spans; sp += nw) {
        const int s = sp / spans_per_stream;
        const int k0 = (sp - s * spans_per_stream) * span;
        const int k1 = min(k0 + span, tiles_per_stream);
        const float2 *r = samples + (long long)s * stride;
        // carried per-lane state: suffix sums (this lane's chunk .. lane 31) of the previous tiles, in float64
        double sufPr = 0.0, sufPi = 0.0;                   // products, tile t-1
        double sufE1 = 0.0, sufE2 = 0.0, totE1 = 0.0;      // energies: suffix of t-1, suffix of t-2, total of t-1
        float ce1 = 0.f, ce2 = 0.f;                        // this lane's chunk energy 1 and 2 tiles back
        __syncwarp();
        // slot 3 stands for "the tile before the first one" in the first iteration: its products are never used,
        // but keep the arithmetic on defined data
#pragma unroll
        for (int q = 0; q < NU; q++) *reinterpret_cast<float4 *>(ring + 3 * SLOTB + lane * ROWB + (q << 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
        swc_fill<C>(ring, 0, r, n, k0 - 2, full_tiles, lane, offs);
        int it = 0;
        for (int k = k0 - 2; k < k1; k++, it++) {
            const int sc = it & 3, s1 = (it + 3) & 3, s2 = (it + 2) & 3;     // current, fft_len/2 back, fft_len back
            if (k + 1 < k1) {
                swc_fill<C>(ring, (it + 1) & 3, r, n, k + 1, full_tiles, lane, offs);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncwarp();
            // ---- chunk totals of this lane's C samples (the products themselves are not kept: most tiles are
            //      rejected as a whole below, the others recompute them in the sliding pass)
            const unsigned char *po = ring + sc * SLOTB + lane * ROWB, *pd = ring + s1 * SLOTB + lane * ROWB;
            float sxr = 0.f, sxi = 0.f, se = 0.f;
#pragma unroll
            for (int q = 0; q < NU; q++) {
                const float4 a = *reinterpret_cast<const float4 *>(po + ((q << 4) ^ sown));
                const float4 b = *reinterpret_cast<const float4 *>(pd + ((q << 4) ^ sown));
                // same operations, same order as the sliding pass (x[0], x[1], ... summed left to right)
                sxr += fmaf(a.x, b.x, a.y * b.y); sxi += fmaf(a.y, b.x, -(a.x * b.y)); se += fmaf(a.x, a.x, a.y * a.y);
                sxr += fmaf(a.z, b.z, a.w * b.w); sxi += fmaf(a.w, b.z, -(a.z * b.w)); se += fmaf(a.z, a.z, a.w * a.w);
            }
            // ---- inclusive scan of the chunk totals over the lanes (float64)
            double ia = sxr, ib = sxi, ic = se;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double pa = __shfl_up_sync(0xffffffffu, ia, o);
                const double pb = __shfl_up_sync(0xffffffffu, ib, o);
                const double pc = __shfl_up_sync(0xffffffffu, ic, o);
                if (lane >= o) { ia += pa; ib += pb; ic += pc; }
            }
            const double ta = __shfl_sync(0xffffffffu, ia, 31), tb = __shfl_sync(0xffffffffu, ib, 31),
                         tc = __shfl_sync(0xffffffffu, ic, 31);

            if (k >= k0) {
                // ---- window sums at the start of this chunk: 32 chunks of products, 64 chunks of energy
                float Pr = (float)(sufPr + (ia - (double)sxr));
                float Pi = (float)(sufPi + (ib - (double)sxi));
                float E = (float)(sufE2 + totE1 + (ic - (double)se));
                const float cJ = se, cD = ce1, cN = ce2;
                const float A = E + cJ + cD + cN;                // local bound (see ofdmx_sync.cuh)
                const float eps = 6.0e-6f * A;
                const float e3 = 3.5f * eps, e33 = 3.0f * eps * eps;
                unsigned det = 0, unc = 0;
                bool skip;
                {   // chunk-level rejection (see ofdmx_sync.cuh)
                    const float pmax = fabsf(Pr) + fabsf(Pi) + 0.5f * (cJ + 2.0f * cD + cN) + 4.0f * eps;
                    const float emin = E - cN - 4.0f * eps;
                    skip = (emin > 0.0f) && (pmax * pmax < 0.999f * thr4 * emin * emin);
                }
                if (!__all_sync(0xffffffffu, skip)) {
                    const unsigned char *pn = ring + s2 * SLOTB + lane * ROWB;
#pragma unroll 2
                    for (int q = 0; q < NU; q++) {
                        const float4 a = *reinterpret_cast<const float4 *>(po + ((q << 4) ^ sown));   // r[n]
                        const float4 b = *reinterpret_cast<const float4 *>(pd + ((q << 4) ^ sown));   // r[n - N/2]
                        const float4 c = *reinterpret_cast<const float4 *>(pn + ((q << 4) ^ sown));   // r[n - N]
#pragma unroll
                        for (int t2 = 0; t2 < 2; t2++) {
                            const float ar = t2 ? a.z : a.x, ai = t2 ? a.w : a.y;
                            const float br = t2 ? b.z : b.x, bi = t2 ? b.w : b.y, cr = t2 ? c.z : c.x, ci = t2 ? c.w : c.y;
                            const float xr = fmaf(ar, br, ai * bi), xi = fmaf(ai, br, -(ar * bi));
                            const float en = fmaf(ar, ar, ai * ai);
                            const float xdr = fmaf(br, cr, bi * ci), xdi = fmaf(bi, cr, -(br * ci));
                            const float ed = fmaf(cr, cr, ci * ci);
                            const int kk = 2 * q + t2;
                            Pr += xr - xdr;
                            Pi += xi - xdi;
                            E += en - ed;
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
                    const long long firsts = ((long long)k * 32 + lane) * C;
                    if (firsts + C > n) {
                        const int valid = (n > firsts) ? (int)(n - firsts) : 0;
                        const unsigned m = (valid >= C) ? FULL : ((1u << valid) - 1u);
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
                        const int i = src * C + kk;                    // sample position relative to the tile start
                        double sr = 0.0, si = 0.0, sen = 0.0;
                        for (int t2 = lane; t2 < N; t2 += 32) {
                            const int sa = i - t2;                     // -(N-1) .. TILE-1
                            const int wa = sa & (TILE - 1);
                            const float2 a = *reinterpret_cast<const float2 *>(
                                ring + swc_unit<C>((it + 4 + (sa >> LT)) & 3, wa / C, (wa % C) >> 1) + ((wa & 1) << 3));
                            sen += (double)a.x * a.x + (double)a.y * a.y;
                            if (t2 < N / 2) {
                                const int sb = sa - N / 2;             // -N .. -1
                                const int wb = sb & (TILE - 1);
                                const float2 b = *reinterpret_cast<const float2 *>(
                                    ring + swc_unit<C>((it + 4 + (sb >> LT)) & 3, wb / C, (wb % C) >> 1) + ((wb & 1) << 3));
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
                // ---- C bits per lane -> 32-bit words
                if (C == 16) {
                    const unsigned hi = __shfl_down_sync(0xffffffffu, det, 1);
                    if (!(lane & 1)) {
                        const long long w = (long long)k * 16 + (lane >> 1);
                        if (w < wps) {
                            detmask[(long long)s * wps + w] = (det & 0xffffu) | (hi << 16);
                            trigmask[(long long)s * wps + w] = 0u;         // cleared here: saves a memset pass
                        }
                    }
                } else {
                    const long long w = (long long)k * 32 + lane;
                    if (w < wps) {
                        detmask[(long long)s * wps + w] = det;
                        trigmask[(long long)s * wps + w] = 0u;
                    }
                }
            }
            // ---- carry: this tile becomes "one back"
            sufE2 = sufE1;
            sufE1 = tc - ic + (double)se;
            totE1 = tc;
            sufPr = ta - ia + (double)sxr;
            sufPi = tb - ib + (double)sxi;
            ce2 = ce1;
            ce1 = se;
            __syncwarp();      // every lane is done with the slot the next iteration's cp.async overwrites
        }
    }
}

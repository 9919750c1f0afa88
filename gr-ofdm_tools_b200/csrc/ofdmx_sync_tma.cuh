// ofdmx_sync_tma.cuh -- K2 streaming version: Schmidl & Cox metric with TMA-staged sample tiles.
//
// Same arithmetic and the same filtered predicate as ofdmx_sync.cuh (float32 window sums with an error
// bound, exact float64 re-evaluation inside the band), restructured around the memory system:
//   * persistent CTAs walk spans of 16 consecutive tiles (4096 samples each) of one stream;
//   * samples arrive by TMA (cp.async.bulk.tensor, 3-D map {32 floats, rows of 16 samples, streams},
//     64-row boxes, SWIZZLE_128B) into a shared-memory ring and complete on an mbarrier -- no thread
//     issues per-sample load/store instructions, out-of-range rows are zero-filled by the TMA unit;
//   * the ring keeps the last fft_len samples (and their chunk totals) of the previous tile, so the
//     window history is never re-read from HBM or recomputed inside a span;
//   * the 128-byte swizzle makes every thread's 16-byte reads of its own 128-byte chunk, and of the
//     chunks fft_len/2 and fft_len behind it, bank-conflict free.
#pragma once
#include <cuda.h>
#include "ofdmx_sync.cuh"

#define ST_BOX_ROWS 64                  // rows (of 16 samples) per TMA box: 8 KB
#define ST_BOX_BYTES (ST_BOX_ROWS * 128)
#define ST_TILE_BOXES 4                 // 4096 samples per tile
#define ST_SPAN_TILES 32

__device__ __forceinline__ uint32_t st_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void st_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = st_smem(bar);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void st_tma_box(const CUtensorMap *tmap, uint32_t dst, uint64_t *bar, int row, int stream)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(0), "r"(row), "r"(stream), "r"(st_smem(bar))
                 : "memory");
}

// byte offset (inside the ring) of 16-byte unit q of ring row rho (SWIZZLE_128B)
__device__ __forceinline__ int st_unit(int rho, int q) { return (rho << 7) + ((q ^ (rho & 7)) << 4); }

__device__ __forceinline__ void st_products(const unsigned char *ring, int rown, int rdel, bool has_del, float2 *x, float *e)
{
    const int sown = (rown & 7) << 4, sdel = (rdel & 7) << 4;
    const unsigned char *po = ring + (rown << 7), *pd = ring + (rdel << 7);
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const float4 a = *reinterpret_cast<const float4 *>(po + ((q << 4) ^ sown));
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_del) b = *reinterpret_cast<const float4 *>(pd + ((q << 4) ^ sdel));
        x[2 * q] = make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -(a.x * b.y)));
        x[2 * q + 1] = make_float2(fmaf(a.z, b.z, a.w * b.w), fmaf(a.w, b.z, -(a.z * b.w)));
        e[2 * q] = fmaf(a.x, a.x, a.y * a.y);
        e[2 * q + 1] = fmaf(a.z, a.z, a.w * a.w);
    }
}

__global__ void __launch_bounds__(SV_THREADS, 2)
sync_metric_tma_kernel(const __grid_constant__ CUtensorMap tmap, const float2 *__restrict__ samples, long long n,
                       long long stride, int N, float thr_f, double thr_d, uint32_t *__restrict__ detmask,
                       uint32_t *__restrict__ trigmask, long long wps, long long tiles_per_stream, long long spans_per_stream, long long total_spans)
{
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    // SWIZZLE_128B needs the ring 1024-byte aligned in the shared address space
    unsigned char *smem_raw = smem_dyn + ((1024u - (st_smem(smem_dyn) & 1023u)) & 1023u);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int h = N >> 1;
    const int nhc = N >> 4;                               // halo rows (chunks)
    const int hc = h >> 4;
    const int nch = nhc + SV_THREADS;
    const int hb = (nhc + ST_BOX_ROWS - 1) / ST_BOX_ROWS;  // halo boxes
    const int RB = 2 * ST_TILE_BOXES + hb;                 // ring boxes: current tile + prefetched tile + halo
    const int RR = RB * ST_BOX_ROWS;                       // ring rows
    unsigned char *ring = smem_raw;                        // RB * 8 KB, 1024-byte aligned
    float *CXr = reinterpret_cast<float *>(ring + RB * ST_BOX_BYTES);   // chunk totals, indexed by ring row
    float *CXi = CXr + RR;
    float *CE = CXi + RR;
    double *TXr = reinterpret_cast<double *>(CE + RR);     // float64 prefix sums in logical chunk order, nch + 2 each
    double *TXi = TXr + nch + 2;
    double *TE = TXi + nch + 2;
    uint64_t *bar = reinterpret_cast<uint64_t *>(TE + nch + 2);   // two barriers
    __shared__ double wsum[3][SV_THREADS / 32 + 1];

    if (tid == 0) { st_mbar_init(bar, 1); st_mbar_init(bar + 1, 1); }
    __syncthreads();
    uint32_t cseq = 0;                                     // tiles processed by this CTA: barrier cseq&1, parity (cseq>>1)&1
    const float thr4 = 0.25f * thr_f;
    const long long tail_row = (n & 15) ? (n >> 4) : -1;   // partially filled last row (not covered by the map)

    for (long long sp = blockIdx.x; sp < total_spans; sp += gridDim.x) {
        const int s = (int)(sp / spans_per_stream);
        const long long k0 = (sp - (long long)s * spans_per_stream) * ST_SPAN_TILES;
        const long long k1 = min(k0 + ST_SPAN_TILES, tiles_per_stream);
        const float2 *r = samples + (long long)s * stride;
        int slot0 = (int)((k0 * ST_TILE_BOXES + 1024LL * RB) % RB);   // ring slot of the tile's first box
        if (tid == 0) {
            // start of a span: halo boxes + first tile on this tile's barrier
            uint64_t *b0 = bar + (cseq & 1);
            const int nb = ST_TILE_BOXES + hb;
            st_mbar_expect_tx(b0, (uint32_t)nb * ST_BOX_BYTES);
            for (int i = 0; i < nb; i++) {
                int slot = slot0 - hb + i;
                if (slot < 0) slot += RB;
                if (slot >= RB) slot -= RB;
                st_tma_box(&tmap, st_smem(ring + slot * ST_BOX_BYTES), b0, (int)((k0 * ST_TILE_BOXES - hb + i) * ST_BOX_ROWS), s);
            }
        }
        for (long long k = k0; k < k1; k++, cseq++) {
            const bool first = (k == k0);
            const long long g0 = k * SV_THREADS;            // first row of the tile
            // ---- TMA prefetch of the next tile of the span into the boxes the previous tile left
            if (tid == 0 && k + 1 < k1) {
                uint64_t *bn = bar + ((cseq + 1) & 1);
                st_mbar_expect_tx(bn, (uint32_t)ST_TILE_BOXES * ST_BOX_BYTES);
                for (int i = 0; i < ST_TILE_BOXES; i++) {
                    int slot = slot0 + ST_TILE_BOXES + i;
                    if (slot >= RB) slot -= RB;
                    st_tma_box(&tmap, st_smem(ring + slot * ST_BOX_BYTES), bn, (int)(((k + 1) * ST_TILE_BOXES + i) * ST_BOX_ROWS), s);
                }
            }
            st_mbar_wait(bar + (cseq & 1), (cseq >> 1) & 1);
            const int rho0 = slot0 * ST_BOX_ROWS;           // ring row of the tile's first row
            slot0 += ST_TILE_BOXES;
            if (slot0 >= RB) slot0 -= RB;
            if (tail_row >= g0 - nhc && tail_row < g0 + SV_THREADS) {
                // the map covers whole rows only: bring in the last n%16 samples by hand
                if (first || tail_row >= g0) {
                    if (tid < (int)(n & 15)) {
                        int rho = rho0 + (int)(tail_row - g0);
                        if (rho < 0) rho += RR;
                        if (rho >= RR) rho -= RR;
                        const float2 v = __ldg(&r[tail_row * 16 + tid]);
                        *reinterpret_cast<float2 *>(ring + st_unit(rho, tid >> 1) + ((tid & 1) << 3)) = v;
                    }
                }
                __syncthreads();
            }
            // ---- phase 1: products and chunk totals of this thread's row
            int rown = rho0 + tid;
            if (rown >= RR) rown -= RR;
            int rdel = rown - hc;
            if (rdel < 0) rdel += RR;
            int rdn = rown - nhc;
            if (rdn < 0) rdn += RR;
            float2 x[SV_C];
            float e[SV_C];
            st_products(ring, rown, rdel, true, x, e);
            {
                float sxr = 0.f, sxi = 0.f, se = 0.f;
#pragma unroll
                for (int q = 0; q < SV_C; q++) { sxr += x[q].x; sxi += x[q].y; se += e[q]; }
                CXr[rown] = sxr; CXi[rown] = sxi; CE[rown] = se;
            }
            if (first) {
                // start of a span: totals of the halo rows (afterwards they are carried in the ring)
                for (int jh = tid; jh < nhc; jh += SV_THREADS) {
                    int rh = rho0 - nhc + jh;
                    if (rh < 0) rh += RR;
                    int rhd = rh - hc;
                    if (rhd < 0) rhd += RR;
                    float2 hx[SV_C];
                    float he[SV_C];
                    st_products(ring, rh, rhd, jh >= hc, hx, he);
                    float sxr = 0.f, sxi = 0.f, se = 0.f;
#pragma unroll
                    for (int q = 0; q < SV_C; q++) { sxr += hx[q].x; sxi += hx[q].y; se += he[q]; }
                    CXr[rh] = sxr; CXi[rh] = sxi; CE[rh] = se;
                }
            }
            __syncthreads();
            // ---- phase 2: exclusive scan of the chunk totals in logical order (halo rows first)
            {
                const int j0 = 2 * tid, j1 = 2 * tid + 1;
                int q0 = rho0 - nhc + j0, q1 = rho0 - nhc + j1;
                if (q0 < 0) q0 += RR;
                if (q0 >= RR) q0 -= RR;
                if (q1 < 0) q1 += RR;
                if (q1 >= RR) q1 -= RR;
                const double a0 = (j0 < nch) ? CXr[q0] : 0.0, a1 = (j1 < nch) ? CXr[q1] : 0.0;
                const double b0 = (j0 < nch) ? CXi[q0] : 0.0, b1 = (j1 < nch) ? CXi[q1] : 0.0;
                const double c0 = (j0 < nch) ? CE[q0] : 0.0, c1 = (j1 < nch) ? CE[q1] : 0.0;
                double ia = a0 + a1, ib = b0 + b1, ic = c0 + c1;
                const double ta = ia, tb = ib, tc = ic;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double pa = __shfl_up_sync(0xffffffffu, ia, o);
                    const double pb = __shfl_up_sync(0xffffffffu, ib, o);
                    const double pc = __shfl_up_sync(0xffffffffu, ic, o);
                    if (lane >= o) { ia += pa; ib += pb; ic += pc; }
                }
                if (lane == 31) { wsum[0][wid] = ia; wsum[1][wid] = ib; wsum[2][wid] = ic; }
                __syncthreads();
                double oa = 0.0, ob = 0.0, oc = 0.0;
                for (int w = 0; w < wid; w++) { oa += wsum[0][w]; ob += wsum[1][w]; oc += wsum[2][w]; }
                const double ea = oa + ia - ta, eb = ob + ib - tb, ec = oc + ic - tc;
                if (j0 < nch) { TXr[j0] = ea; TXi[j0] = eb; TE[j0] = ec; }
                if (j1 < nch) { TXr[j1] = ea + a0; TXi[j1] = eb + b0; TE[j1] = ec + c0; }
                if (tid == ((nch - 1) >> 1)) { TXr[nch] = ea + a0 + a1; TXi[nch] = eb + b0 + b1; TE[nch] = ec + c0 + c1; }
            }
            __syncthreads();
            // ---- phase 3: sliding window inside the chunk, filtered comparison
            const int J = nhc + tid;
            float Pr = (float)(TXr[J] - TXr[J - hc]);
            float Pi = (float)(TXi[J] - TXi[J - hc]);
            float E = (float)(TE[J] - TE[J - nhc]);
            const float cJ = CE[rown], cD = CE[rdel], cN = CE[rdn];
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
                const int sd = (rdel & 7) << 4, sn = (rdn & 7) << 4;
                const unsigned char *pd = ring + (rdel << 7), *pn = ring + (rdn << 7);
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
                        // |Pr| + |Pi| <= 0.71 E (Cauchy-Schwarz), pm2 + rhs <= 0.5 E^2
                        const float aE = fabsf(E);
                        const float err = fmaf(aE, fmaf(5.0e-7f, aE, e3), e33);
                        if (d > err) det |= 1u << kk;
                        if (fabsf(d) <= err) unc |= 1u << kk;
                    }
                }
            }
            if (A == 0.0f) { det = 0; unc = 0; }
            {
                const long long firsts = (g0 + tid) * SV_C;
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
                const int jsrc = (wid << 5) + src;                 // row within the tile
                while (m) {
                    const int kk = __ffs(m) - 1;
                    m &= m - 1;
                    const int i = (jsrc << 4) + kk;                // sample position relative to the tile start
                    double sr = 0.0, si = 0.0, se = 0.0;
                    for (int t2 = lane; t2 < N; t2 += 32) {
                        const int sa = i - t2;                     // may be negative: halo rows
                        int ra = rho0 + (sa >> 4);                 // arithmetic shift: floor
                        if (ra < 0) ra += RR;
                        if (ra >= RR) ra -= RR;
                        const float2 a = *reinterpret_cast<const float2 *>(ring + st_unit(ra, (sa >> 1) & 7) + ((sa & 1) << 3));
                        se += (double)a.x * a.x + (double)a.y * a.y;
                        if (t2 < h) {
                            const int sb = sa - h;
                            int rb = rho0 + (sb >> 4);
                            if (rb < 0) rb += RR;
                            if (rb >= RR) rb -= RR;
                            const float2 b = *reinterpret_cast<const float2 *>(ring + st_unit(rb, (sb >> 1) & 7) + ((sb & 1) << 3));
                            sr += (double)a.x * b.x + (double)a.y * b.y;
                            si += (double)a.y * b.x - (double)a.x * b.y;
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        sr += __shfl_xor_sync(0xffffffffu, sr, o);
                        si += __shfl_xor_sync(0xffffffffu, si, o);
                        se += __shfl_xor_sync(0xffffffffu, se, o);
                    }
                    const double R = 0.5 * se, R2 = R * R, pm2 = sr * sr + si * si;
                    const bool dd = (R2 > 0.0) && (pm2 >= thr_d * R2);
                    if (lane == src) det = dd ? (det | (1u << kk)) : (det & ~(1u << kk));
                }
            }
            // ---- 16 bits per thread -> 32-bit words
            const unsigned hi = __shfl_down_sync(0xffffffffu, det, 1);
            if (!(tid & 1)) {
                const long long w = (g0 >> 1) + (tid >> 1);
                if (w < wps) {
                    detmask[(long long)s * wps + w] = (det & 0xffffu) | (hi << 16);
                    trigmask[(long long)s * wps + w] = 0u;         // cleared here: saves a memset pass
                }
            }
            // all generic-proxy reads of the ring are done before the next TMA overwrites dead boxes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
        }
    }
}

static inline size_t sync_tma_smem_bytes(int N)
{
    const int nhc = N >> 4, nch = nhc + SV_THREADS;
    const int hb = (nhc + ST_BOX_ROWS - 1) / ST_BOX_ROWS, RB = 2 * ST_TILE_BOXES + hb, RR = RB * ST_BOX_ROWS;
    return (size_t)RB * ST_BOX_BYTES + 3 * (size_t)RR * 4 + 3 * (size_t)(nch + 2) * 8 + 16 + 16 + 1024;
}

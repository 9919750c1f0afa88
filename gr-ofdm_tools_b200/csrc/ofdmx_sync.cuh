// ofdmx_sync.cuh -- K2: Schmidl & Cox timing metric, fast kernel (fft_len >= 32).
//
// Replaces the per-sample chain of ofdm_sync_sc_cfb (delay, conjugate, multiply, two moving-average
// FIRs, squares, divide; python/ofdm_txrx_modules.py:324) up to the plateau detector input, producing
// one detect bit per sample:   detect[n] = R[n]^2 > 0  &&  |P[n]|^2 >= thr * R[n]^2.
//
// The decision is defined in exact arithmetic (the oracle evaluates it in float64).  This kernel gets
// the same bits at float32 speed with a filtered predicate:
//   * fast path: float32 window sums from a two-level scheme -- per-thread chunk totals (16
//     consecutive samples), a block scan of the chunk totals gives each chunk's initial window sums,
//     then a 16-step sliding recurrence inside the chunk;
//   * every comparison carries a bound `err` on what float32 rounding can have done to it (derived
//     from the tile's total energy); |lhs - rhs| > err decides immediately;
//   * the rare samples inside the band (a few per plateau edge) are re-evaluated by the whole warp in
//     float64 directly from the staged samples -- exactly the oracle's sum.
//
// Tile = 4096 samples + fft_len halo per CTA, 256 threads, one 16-sample chunk per thread.  Samples
// are staged once in shared memory in a 144-byte-per-chunk layout, so that every thread reads its
// chunk (and the chunks fft_len/2 and fft_len behind it) with conflict-free 16-byte loads.
#pragma once
#include "ofdmx_dev.cuh"

#define SV_C 16              // samples per chunk (per thread)
#define SV_T 4096            // samples per tile
#define SV_THREADS 256       // = SV_T / SV_C

// float4 index of 16-byte unit q (0..7) of chunk j
__device__ __forceinline__ int sv_off(int j, int q) { return 9 * j + q; }
// float2 (sample) index of sample s (relative to the start of the staged region)
__device__ __forceinline__ int sv_sample(int s) { return (9 * (s >> 4) + ((s >> 1) & 7)) * 2 + (s & 1); }

__device__ __forceinline__ void sv_products(const float4 *__restrict__ r4, int jown, int jdel, float2 *x, float *e)
{
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const float4 a = r4[sv_off(jown, q)];
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (jdel >= 0) b = r4[sv_off(jdel, q)];
        x[2 * q] = make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -(a.x * b.y)));
        x[2 * q + 1] = make_float2(fmaf(a.z, b.z, a.w * b.w), fmaf(a.w, b.z, -(a.z * b.w)));
        e[2 * q] = fmaf(a.x, a.x, a.y * a.y);
        e[2 * q + 1] = fmaf(a.z, a.z, a.w * a.w);
    }
}

template <int NT>   // NT = fft_len at compile time, or 0 to use the runtime value
__global__ void __launch_bounds__(SV_THREADS, 4)
sync_metric_fast_kernel(const float2 *__restrict__ samples, long long n, long long stride, int Nrt, float thr_f,
                        double thr_d, uint32_t *__restrict__ detmask, uint32_t *__restrict__ trigmask, long long wps)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int N = NT ? NT : Nrt;
    const int h = N >> 1;
    const int nhc = N >> 4;                 // halo chunks
    const int hc = h >> 4;                  // chunks per half window
    const int nch = nhc + SV_THREADS;       // chunks staged
    float4 *r4 = reinterpret_cast<float4 *>(smem_raw);            // 9 * nch float4
    // chunk totals and their exclusive prefix live in float64: the scan covers the whole tile, and a
    // float32 prefix would make every window inherit the rounding error of the strongest burst in the tile
    double *TXr = reinterpret_cast<double *>(r4 + 9 * nch);       // nch + 2 each
    double *TXi = TXr + nch + 2;
    double *TE = TXi + nch + 2;
    float *CE = reinterpret_cast<float *>(TE + nch + 2);          // chunk energies, nch
    __shared__ double wsum[3][SV_THREADS / 32 + 1];

    const long long ts = (long long)blockIdx.x * SV_T;
    const float2 *r = samples + (long long)blockIdx.y * stride;
    const bool aligned = ((reinterpret_cast<unsigned long long>(r) & 15ull) == 0);

    // ---- phase 0: stage [ts - N, ts + SV_T) (zeros outside the stream).  Loads are issued in batches
    // of 8 per thread before any store, so one HBM latency covers the whole batch.
    const int n_units = nch * 8;
    const bool interior = aligned && (ts - N >= 0) && (ts + SV_T <= n);
    for (int u0 = 0; u0 < n_units; u0 += 8 * SV_THREADS) {
        float4 v[8];
        if (interior) {
            const float4 *src = reinterpret_cast<const float4 *>(r + (ts - N)) + u0 + tid;
#pragma unroll
            for (int i = 0; i < 8; i++)
                v[i] = (u0 + i * SV_THREADS + tid < n_units) ? __ldg(src + i * SV_THREADS) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int u = u0 + i * SV_THREADS + tid;
                const long long m = ts - N + 2LL * u;
                float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (u < n_units && m >= 0 && m + 1 < n) {
                    if (aligned) t4 = __ldg(reinterpret_cast<const float4 *>(r + m));
                    else { const float2 a = __ldg(&r[m]), b = __ldg(&r[m + 1]); t4 = make_float4(a.x, a.y, b.x, b.y); }
                } else if (u < n_units && m >= 0 && m < n) {
                    const float2 a = __ldg(&r[m]);
                    t4.x = a.x; t4.y = a.y;
                }
                v[i] = t4;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int u = u0 + i * SV_THREADS + tid;
            if (u < n_units) r4[u + (u >> 3)] = v[i];
        }
    }
    __syncthreads();

    // ---- phase 1: products and chunk totals
    const int J = nhc + tid;                // this thread's tile chunk
    {
        // only the totals are kept: most warps never need the per-sample products again (chunk-level
        // rejection below), the others recompute them from shared memory
        float2 x[SV_C];
        float e[SV_C];
        sv_products(r4, J, J - hc, x, e);
        float sxr = 0.f, sxi = 0.f, se = 0.f;
#pragma unroll
        for (int q = 0; q < SV_C; q++) { sxr += x[q].x; sxi += x[q].y; se += e[q]; }
        TXr[J] = (double)sxr; TXi[J] = (double)sxi; TE[J] = (double)se; CE[J] = se;
    }
    for (int j = tid; j < nhc; j += SV_THREADS) {      // halo chunks: totals only
        float2 hx[SV_C];
        float he[SV_C];
        sv_products(r4, j, j - hc, hx, he);
        float sxr = 0.f, sxi = 0.f, se = 0.f;
#pragma unroll
        for (int q = 0; q < SV_C; q++) { sxr += hx[q].x; sxi += hx[q].y; se += he[q]; }
        TXr[j] = (double)sxr; TXi[j] = (double)sxi; TE[j] = (double)se; CE[j] = se;
    }
    __syncthreads();

    // ---- phase 2: exclusive scan of the chunk totals (<= 512 entries, 2 per thread)
    {
        const int j0 = 2 * tid, j1 = 2 * tid + 1;
        const double a0 = (j0 < nch) ? TXr[j0] : 0.0, a1 = (j1 < nch) ? TXr[j1] : 0.0;
        const double b0 = (j0 < nch) ? TXi[j0] : 0.0, b1 = (j1 < nch) ? TXi[j1] : 0.0;
        const double c0 = (j0 < nch) ? TE[j0] : 0.0, c1 = (j1 < nch) ? TE[j1] : 0.0;
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
        // exclusive prefix over the 8 warp totals: lanes 0..7 hold them, 3 shuffle steps, pick lane wid
        double oa = (lane < SV_THREADS / 32) ? wsum[0][lane] : 0.0;
        double ob = (lane < SV_THREADS / 32) ? wsum[1][lane] : 0.0;
        double oc = (lane < SV_THREADS / 32) ? wsum[2][lane] : 0.0;
        {
            const double sa = oa, sb = ob, sc = oc;
#pragma unroll
            for (int o = 1; o < SV_THREADS / 32; o <<= 1) {
                const double pa = __shfl_up_sync(0xffffffffu, oa, o);
                const double pb = __shfl_up_sync(0xffffffffu, ob, o);
                const double pc = __shfl_up_sync(0xffffffffu, oc, o);
                if (lane >= o) { oa += pa; ob += pb; oc += pc; }
            }
            oa = __shfl_sync(0xffffffffu, oa - sa, wid);
            ob = __shfl_sync(0xffffffffu, ob - sb, wid);
            oc = __shfl_sync(0xffffffffu, oc - sc, wid);
        }
        const double ea = oa + ia - ta, eb = ob + ib - tb, ec = oc + ic - tc;   // exclusive prefix at j0
        if (j0 < nch) { TXr[j0] = ea; TXi[j0] = eb; TE[j0] = ec; }
        if (j1 < nch) { TXr[j1] = ea + a0; TXi[j1] = eb + b0; TE[j1] = ec + c0; }
        if (tid == ((nch - 1) >> 1)) {               // element nch = grand total (a1/b1/c1 are 0 past the end)
            TXr[nch] = ea + a0 + a1; TXi[nch] = eb + b0 + b1; TE[nch] = ec + c0 + c1;
        }
    }
    __syncthreads();

    // ---- phase 3: sliding window inside the chunk, filtered comparison
    float Pr = (float)(TXr[J] - TXr[J - hc]);
    float Pi = (float)(TXi[J] - TXi[J - hc]);
    float E = (float)(TE[J] - TE[J - nhc]);
    const float cJ = CE[J], cD = CE[J - hc], cN = CE[J - nhc];
    // every partial sum formed for this chunk is bounded by the window energy plus the three chunk
    // energies that enter or leave: float32 chunk totals (16 terms), <= 32 sliding updates
    const float A = E + cJ + cD + cN;
    const float eps = 6.0e-6f * A;                   // bound on the float32 error of Pr, Pi, E
    const float thr4 = 0.25f * thr_f;
    const float e3 = 3.5f * eps, e33 = 3.0f * eps * eps;
    unsigned det = 0, unc = 0;
    const int jd = J - hc, jn = J - nhc;
    // Chunk-level rejection: inside the chunk |P| can grow by at most the energy entering and leaving
    // the two half windows, and R can shrink by at most the energy leaving.  If even those extremes stay
    // below the threshold, no sample of the chunk detects (or is uncertain) and the sliding pass is
    // skipped -- decided per warp, so noise and payload regions cost no phase-3 work.
    bool skip;
    {
        const float pmax = fabsf(Pr) + fabsf(Pi) + 0.5f * (cJ + 2.0f * cD + cN) + 4.0f * eps;
        const float emin = E - cN - 4.0f * eps;
        skip = (emin > 0.0f) && (pmax * pmax < 0.999f * thr4 * emin * emin);
    }
    if (!__all_sync(0xffffffffu, skip)) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const float4 a = r4[sv_off(J, q)];           // r[n]
        const float4 b = r4[sv_off(jd, q)];          // r[n - N/2]
        const float4 c = r4[sv_off(jn, q)];          // r[n - N]
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const float ar = s ? a.z : a.x, ai = s ? a.w : a.y;
            const float br = s ? b.z : b.x, bi = s ? b.w : b.y, cr = s ? c.z : c.x, ci = s ? c.w : c.y;
            const float xnr = fmaf(ar, br, ai * bi), xni = fmaf(ai, br, -(ar * bi));   // x[n]
            const float en = fmaf(ar, ar, ai * ai);                                       // e[n]
            const float xdr = fmaf(br, cr, bi * ci), xdi = fmaf(bi, cr, -(br * ci));   // x[n - N/2]
            const float ed = fmaf(cr, cr, ci * ci);                                       // e[n - N]
            const int k = 2 * q + s;
            Pr += xnr - xdr;
            Pi += xni - xdi;
            E += en - ed;
            const float d = fmaf(Pr, Pr, Pi * Pi) - thr4 * E * E;
            // |Pr| + |Pi| <= 0.71 E (Cauchy-Schwarz on the two half windows), pm2 + rhs <= 0.5 E^2
            const float aE = fabsf(E);
            const float err = fmaf(aE, fmaf(5.0e-7f, aE, e3), e33);
            if (d > err) det |= 1u << k;
            if (fabsf(d) <= err) unc |= 1u << k;
        }
    }
    }
    if (A == 0.0f) { det = 0; unc = 0; }             // all-zero windows: R^2 > 0 is false everywhere
    {   // samples beyond the end of the stream never detect
        const long long first = ts + (long long)tid * SV_C;
        if (first + SV_C > n) {
            const int valid = (n > first) ? (int)(n - first) : 0;
            const unsigned m = (valid >= 16) ? 0xffffu : ((1u << valid) - 1u);
            det &= m; unc &= m;
        }
    }

    // ---- exact re-evaluation (float64, whole warp per sample) of the samples inside the band
    unsigned pending = __ballot_sync(0xffffffffu, unc != 0);
    const float2 *r2 = reinterpret_cast<const float2 *>(r4);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        unsigned m = __shfl_sync(0xffffffffu, unc, src);
        const int jsrc = nhc + (wid << 5) + src;
        while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            const int i = (jsrc << 4) + k;           // sample position in the staged region
            double sr = 0.0, si = 0.0, se = 0.0;
            for (int t = lane; t < h; t += 32) {
                const float2 a = r2[sv_sample(i - t)], b = r2[sv_sample(i - t - h)];
                sr += (double)a.x * b.x + (double)a.y * b.y;
                si += (double)a.y * b.x - (double)a.x * b.y;
            }
            for (int t = lane; t < N; t += 32) {
                const float2 a = r2[sv_sample(i - t)];
                se += (double)a.x * a.x + (double)a.y * a.y;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sr += __shfl_xor_sync(0xffffffffu, sr, o);
                si += __shfl_xor_sync(0xffffffffu, si, o);
                se += __shfl_xor_sync(0xffffffffu, se, o);
            }
            const double R = 0.5 * se, R2 = R * R, pm2 = sr * sr + si * si;
            const bool dd = (R2 > 0.0) && (pm2 >= thr_d * R2);
            if (lane == src) det = dd ? (det | (1u << k)) : (det & ~(1u << k));
        }
    }

    // ---- 16 bits per thread -> 32-bit words
    const unsigned hi = __shfl_down_sync(0xffffffffu, det, 1);
    if (!(tid & 1)) {
        const long long w = (ts >> 5) + (tid >> 1);
        if (w < wps) {
            detmask[(long long)blockIdx.y * wps + w] = (det & 0xffffu) | (hi << 16);
            trigmask[(long long)blockIdx.y * wps + w] = 0u;      // cleared here: saves a memset pass
        }
    }
}

static inline size_t sync_fast_smem_bytes(int N)
{
    const int nch = (N >> 4) + SV_THREADS;
    return (size_t)nch * 9 * 16 + 3 * (size_t)(nch + 2) * 8 + (size_t)nch * 4 + 16;
}

// ofdmx_kernels.cuh -- the OFDM PHY kernels (sm_100a).  Included once by ofdmx_api.cu.
//
//   K2  sync_metric_kernel   Schmidl & Cox timing metric -> 1 detect bit per sample
//       plateau_kernel       plateau_detector_fb, cluster-parallel -> 1 trigger bit per sample
//       trig_* kernels       ordered compaction of trigger bits -> (stream, index) list
//       cfo_kernel           arg(P) at each trigger (sample-and-hold input)
//   K1+K3+K4 rx_frame_kernel one CTA per trigger: NCO derotation + CP strip + FFT, chanest,
//                            DFE equaliser + demap + bit pack + descramble + CRC-32
//       chain_* kernels      header_payload_demux acceptance chain (pointer jumping) + emit
//   K5+K1 tx_frame_kernel    CRC append, header, scramble, map, allocate, IFFT, cyclic prefix
#pragma once
#include "ofdmx_dev.cuh"
#include "ofdmx.h"

#define TWO_PI_D 6.283185307179586476925286766559
#define TRIG_WPT 4   // trigger-mask words per thread in the compaction kernels

// K2 (the Schmidl & Cox metric kernels) lives in ofdmx_sync*.cuh; the kernels below turn its detect bits into
// triggers.
#ifdef OFDMX_GENERIC_KERNELS   // the non-template kernels are compiled by ofdmx_api.cu only

// ---------------------------------------------------------------------------------------------
// plateau_detector_fb(max_len = cp_len, threshold) evaluated over the whole stream
// (blocks/plateau_detector_fb_impl.cc semantics), parallel over independent
// clusters: a run start preceded by >= cp+1 clear bits is reached by the sequential scan in its
// fresh state, so each such start can be walked by its own thread.
// ---------------------------------------------------------------------------------------------
// Index type I of the walks: int when the stream length fits (the cluster walks are instruction bound and 64-bit
// index arithmetic is two instructions per operation), long long otherwise.
template <typename I>
__device__ __forceinline__ uint32_t mword(const uint32_t *m, I w, I wps)
{
    return (w < 0 || w >= wps) ? 0u : m[w];
}
template <typename I>
__device__ __forceinline__ int mbit(const uint32_t *m, I i, I n, I wps)
{
    if (i < 0 || i >= n) return 0;
    return (mword<I>(m, i >> 5, wps) >> (i & 31)) & 1u;
}
// are all bits in [lo, hi) clear?  (indices < 0 count as clear)
template <typename I>
__device__ bool bits_clear(const uint32_t *m, I lo, I hi, I wps)
{
    if (lo < 0) lo = 0;
    if (hi <= lo) return true;
    I w0 = lo >> 5, w1 = (hi - 1) >> 5;
    for (I w = w0; w <= w1; w++) {
        uint32_t x = mword<I>(m, w, wps);
        if (w == w0) x &= 0xffffffffu << (lo & 31);
        if (w == w1 && ((hi & 31) != 0)) x &= 0xffffffffu >> (32 - (hi & 31));
        if (x) return false;
    }
    return true;
}
// first index >= i whose bit equals `val`, searching no further than limit (exclusive); returns limit if none
template <typename I>
__device__ I next_bit(const uint32_t *m, I i, I limit, int val, I wps)
{
    while (i < limit) {
        uint32_t x = mword<I>(m, i >> 5, wps);
        if (!val) x = ~x;
        x &= 0xffffffffu << (i & 31);
        if (x) {
            I j = ((i >> 5) << 5) + (__ffs(x) - 1);
            return j < limit ? j : limit;
        }
        i = ((i >> 5) + 1) << 5;
    }
    return limit;
}

// Measured and not kept (round 2, headline workload, 70 us): staging the CTA's words in shared memory (101 us), a
// persistent grid with the next trip prefetched (83 us), a per-thread 16-word window around each flank (136 us) --
// the kernel is bound by the instructions of the divergent cluster walks (ncu: 70 % issue utilisation, 1-2 lanes
// active per warp), not by the latency of their loads.
#define PL_WPT 8   // detect words per thread
template <typename I>
__global__ void __launch_bounds__(OFDMX_THREADS)
plateau_kernel(const uint32_t *__restrict__ detmask, uint32_t *__restrict__ trigmask, long long n_ll,
               long long wps_ll, long long n_streams, int cp, int *__restrict__ blocksum)
{
    const long long wps = wps_ll;
    const I n = (I)n_ll, wpsI = (I)wps_ll;
    const long long total = wps * n_streams;
    const long long g0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * PL_WPT;
    if (g0 >= total) return;
    // the thread's eight words as two 16-byte loads (a warp reads 1 KB contiguous); almost all of them are zero
    uint32_t wv[PL_WPT];
    if (g0 + PL_WPT <= total && (reinterpret_cast<uintptr_t>(detmask + g0) & 15) == 0) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(detmask + g0)), b = __ldg(reinterpret_cast<const uint4 *>(detmask + g0) + 1);
        wv[0] = a.x; wv[1] = a.y; wv[2] = a.z; wv[3] = a.w; wv[4] = b.x; wv[5] = b.y; wv[6] = b.z; wv[7] = b.w;
    } else {
#pragma unroll
        for (int qq = 0; qq < PL_WPT; qq++) wv[qq] = (g0 + qq < total) ? detmask[g0 + qq] : 0u;
    }
    if (!(wv[0] | wv[1] | wv[2] | wv[3] | wv[4] | wv[5] | wv[6] | wv[7])) return;
    // stream and word-in-stream of the thread's first word: one 64-bit division per thread, the following words
    // step from it (a thread's eight words may cross into the next streams)
    long long s = g0 / wps, w = g0 - s * wps;
#pragma unroll
    for (int qq = 0; qq < PL_WPT; qq++, w++) {
        while (w >= wps) { w -= wps; s++; }
        const uint32_t word = wv[qq];
        if (!word) continue;
        const uint32_t *m = detmask + s * wps;
        uint32_t *tm = trigmask + s * wps;
        const I wI = (I)w;
        const uint32_t prev = (w > 0) ? (m[w - 1] >> 31) : 0u;
        uint32_t rising = word & ~((word << 1) | prev);
        while (rising) {
            const int b = __ffs(rising) - 1;
            rising &= rising - 1;
            I i = (wI << 5) + b;
            if (!bits_clear<I>(m, i - (cp + 1), i, wpsI)) continue;   // reached in non-fresh state: owned by an earlier cluster
            // sequential walk of this cluster
            for (;;) {
                if (n - i < 2 * (I)cp) break;                       // "come back later": never at stream end
                const I start = i;
                i = next_bit<I>(m, i, n, 0, wpsI);                     // end of run (exclusive)
                if (i - start > 1) {
                    const I tp = start + (i - start) / 2;
                    atomicOr(&tm[tp >> 5], 1u << (tp & 31));
                    // per-block trigger count for the ordered compaction (same partition as trig_scatter_kernel)
                    atomicAdd(&blocksum[(s * wps + (long long)(tp >> 5)) / (OFDMX_THREADS * TRIG_WPT)], 1);
                    i = (i + cp < n - 1) ? i + cp : n - 1;
                }
                i++;                                               // the for-loop increment
                if (i >= n) break;
                // next flank the sequential scan would see
                const I lim = (i + cp + 2 < n) ? i + cp + 2 : n;
                const I j = next_bit<I>(m, i, lim, 1, wpsI);
                if (j >= lim) break;                               // >= cp+1 clear bits follow: next start is independent
                const bool rising_j = !mbit<I>(m, j - 1, n, wpsI);
                if (rising_j && bits_clear<I>(m, j - (cp + 1), j, wpsI)) break;   // independent start: its own thread walks it
                i = j;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Ordered compaction of the trigger bitmask.
// ---------------------------------------------------------------------------------------------
// single CTA: exclusive scan of blocksum[nb] in place; writes totals
__global__ void __launch_bounds__(1024)
trig_scan_kernel(int *__restrict__ blocksum, int nb, int max_trig, ofdmx_counts *__restrict__ counts,
                 int *__restrict__ n_trig_dev, int *__restrict__ stream_start, long long n_streams)
{
    __shared__ int wt[33];
    int carry = 0;
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        int idx = b0 + threadIdx.x;
        int v = (idx < nb) ? blocksum[idx] : 0;
        int total;
        int ex = block_excl_scan(v, wt, total);
        if (idx < nb) blocksum[idx] = carry + ex;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int nt = carry < max_trig ? carry : max_trig;
        *n_trig_dev = nt;
        counts->n_triggers = carry;
        counts->overflow = (carry > max_trig) ? 1 : 0;
        counts->n_frames = 0;
        counts->reserved = 0;
        stream_start[n_streams] = nt;
    }
}

__global__ void __launch_bounds__(OFDMX_THREADS)
trig_scatter_kernel(const uint32_t *__restrict__ trigmask, long long n_words, long long wps,
                    const int *__restrict__ blocksum, int max_trig, long long *__restrict__ trig,
                    int *__restrict__ trig_stream, int *__restrict__ stream_start)
{
    __shared__ int wt[33];
    const long long base = ((long long)blockIdx.x * OFDMX_THREADS + threadIdx.x) * TRIG_WPT;
    uint32_t wv[TRIG_WPT];
    int c = 0;
    if (base + TRIG_WPT <= n_words && (reinterpret_cast<uintptr_t>(trigmask + base) & 15) == 0) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(trigmask + base));     // one 16-byte load per thread
        wv[0] = a.x; wv[1] = a.y; wv[2] = a.z; wv[3] = a.w;
    } else {
#pragma unroll
        for (int q = 0; q < TRIG_WPT; q++) wv[q] = (base + q < n_words) ? trigmask[base + q] : 0u;
    }
#pragma unroll
    for (int q = 0; q < TRIG_WPT; q++) c += __popc(wv[q]);
    int total;
    int o = blocksum[blockIdx.x] + block_excl_scan(c, wt, total);
    long long s = base / wps, w = base - s * wps;            // one division per thread; the words step from it
    for (int q = 0; q < TRIG_WPT; q++, w++) {
        const long long g = base + q;
        if (g >= n_words) break;
        while (w >= wps) { w -= wps; s++; }
        if (w == 0) stream_start[s] = o < max_trig ? o : max_trig;
        uint32_t x = wv[q];
        while (x) {
            const int b = __ffs(x) - 1;
            x &= x - 1;
            if (o < max_trig) { trig[o] = (w << 5) + b; trig_stream[o] = (int)s; }
            o++;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fine frequency estimate: complex_to_arg(P[t]) sampled at the trigger (sample_and_hold_ff).
// One warp per trigger, exact float64 evaluation of P.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OFDMX_THREADS)
cfo_kernel(const float2 *__restrict__ samples, long long n, long long stride, int N,
           const long long *__restrict__ trig, const int *__restrict__ trig_stream,
           const int *__restrict__ n_trig_dev, float *__restrict__ cfo)
{
    const int lane = threadIdx.x & 31;
    const int nt = *n_trig_dev;
    const int h = N >> 1;
    for (int j = blockIdx.x * (OFDMX_THREADS / 32) + (threadIdx.x >> 5); j < nt; j += gridDim.x * (OFDMX_THREADS / 32)) {
        const long long t = trig[j];
        const float2 *r = samples + (long long)trig_stream[j] * stride;
        double sr = 0, si = 0;
        if (t - 2 * h + 1 >= 0) {
            // interior trigger: batches of independent load pairs per lane (the rolled loop exposed one L2/HBM
            // latency per 32 products): 16 pairs when the half window has that many per lane, else 8
            const float2 *px = r + t - lane, *py = px - h;
            if (h >= 512) {
                for (int k0 = 0; k0 < h; k0 += 512) {
                    float2 x[16], y[16];
#pragma unroll
                    for (int u = 0; u < 16; u++) { x[u] = __ldg(px - (k0 + 32 * u)); y[u] = __ldg(py - (k0 + 32 * u)); }
#pragma unroll
                    for (int u = 0; u < 16; u++) {
                        sr += (double)x[u].x * y[u].x + (double)x[u].y * y[u].y;
                        si += (double)x[u].y * y[u].x - (double)x[u].x * y[u].y;
                    }
                }
            } else {
                for (int k0 = 0; k0 < h; k0 += 256) {
                    float2 x[8], y[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int k = k0 + 32 * u;
                        if (k + lane < h) { x[u] = __ldg(px - k); y[u] = __ldg(py - k); }
                        else { x[u] = make_float2(0.f, 0.f); y[u] = x[u]; }
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        sr += (double)x[u].x * y[u].x + (double)x[u].y * y[u].y;
                        si += (double)x[u].y * y[u].x - (double)x[u].x * y[u].y;
                    }
                }
            }
        } else
        for (int k = lane; k < h; k += 32) {
            long long a = t - k, b = t - k - h;
            if (b < 0) continue;
            float2 x = __ldg(&r[a]), y = __ldg(&r[b]);
            sr += (double)x.x * y.x + (double)x.y * y.y;
            si += (double)x.y * y.x - (double)x.x * y.y;
        }
        for (int o = 16; o > 0; o >>= 1) {
            sr += __shfl_xor_sync(0xffffffffu, sr, o);
            si += __shfl_xor_sync(0xffffffffu, si, o);
        }
        if (lane == 0) cfo[j] = (float)atan2(-si, -sr);
    }
    (void)n;
}

// Short windows (fft_len <= 256): a whole warp per trigger leaves most lanes with one product and pays a five-step
// float64 reduction for it, so G = fft_len/8 lanes (8 .. 32) share a trigger and a warp takes 32/G of them.
template <int G>
__global__ void __launch_bounds__(OFDMX_THREADS)
cfo_small_kernel(const float2 *__restrict__ samples, long long n, long long stride, int N,
                 const long long *__restrict__ trig, const int *__restrict__ trig_stream,
                 const int *__restrict__ n_trig_dev, float *__restrict__ cfo)
{
    constexpr int TPW = 32 / G;
    const int lane = threadIdx.x & 31, gl = lane % G, grp = lane / G;
    const int nt = *n_trig_dev;
    const int h = N >> 1;
    const int warps = gridDim.x * (OFDMX_THREADS / 32);
    for (int j0 = (blockIdx.x * (OFDMX_THREADS / 32) + (threadIdx.x >> 5)) * TPW; j0 < nt; j0 += warps * TPW) {
        const int j = j0 + grp;
        double sr = 0, si = 0;
        if (j < nt) {
            const long long t = trig[j];
            const float2 *r = samples + (long long)trig_stream[j] * stride;
#pragma unroll 4
            for (int k = gl; k < h; k += G) {
                const long long a = t - k, b = a - h;
                if (b < 0) continue;
                const float2 x = __ldg(&r[a]), y = __ldg(&r[b]);
                sr += (double)x.x * y.x + (double)x.y * y.y;
                si += (double)x.y * y.x - (double)x.x * y.y;
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            sr += __shfl_xor_sync(0xffffffffu, sr, o);
            si += __shfl_xor_sync(0xffffffffu, si, o);
        }
        if (gl == 0 && j < nt) cfo[j] = (float)atan2(-si, -sr);
    }
    (void)n;
}

#endif  // OFDMX_GENERIC_KERNELS
// =============================================================================================
// RX frame kernel: everything downstream of the trigger for one frame, in one CTA.
// =============================================================================================
struct FrameSmem {
    float2 *bufA, *bufB, *H, *zs;
    uint8_t *dec, *syms, *pk, *hb;
    uint32_t *scratch;   // 16 words
};

// NCO phase (in turns) at delayed-stream item i for the frame whose trigger ordinal is j:
// frequency_modulator_fc(-2/N) integrating the sample-and-held arg(P); the hold value changes at
// every raw trigger item, also inside a frame.  Phase reference: 0 just before this frame's trigger.
__device__ __forceinline__ double nco_turns(long long i, int j, int jend, const long long *__restrict__ trig,
                                            const float *__restrict__ cfo, int N)
{
    double acc = 0.0;
    int m = j;
    long long tm = trig[j];
    while (m + 1 < jend) {
        const long long tn = trig[m + 1];
        if (tn > i) break;
        acc += (double)__ldcg(&cfo[m]) * (double)(tn - tm);
        m++;
        tm = tn;
    }
    acc += (double)__ldcg(&cfo[m]) * (double)(i - tm + 1);   // L2 load: the frame kernel may have just written it
    return acc * (-2.0 / (double)N) * (1.0 / TWO_PI_D);
}

// blocks.delay(N+cp) -> multiply_cc(NCO) -> header_payload_demux CP strip -> (bit-reversed) smem
__device__ __forceinline__ void load_symbol(float2 *buf, const KP &p, const float2 *__restrict__ r, long long n,
                                            long long i0, int j, int jend, const long long *__restrict__ trig,
                                            const float *__restrict__ cfo)
{
    for (int m = threadIdx.x; m < p.N; m += blockDim.x) {
        const long long i = i0 + m, s = i - p.D;
        float2 v = make_float2(0.f, 0.f);
        if (s >= 0 && s < n) v = __ldg(&r[s]);
        double turns = nco_turns(i, j, jend, trig, cfo, p.N);
        turns -= rint(turns);
        float sn, cs;
        sincospif(2.0f * (float)turns, &sn, &cs);
        buf[bitrev(m, p.logN)] = cmul(v, make_float2(cs, sn));
    }
}

// shifted-order read of an FFT output held in natural order
__device__ __forceinline__ float2 ysh(const float2 *buf, int ks, int N) { return buf[ks ^ (N >> 1)]; }

// ofdm_frame_equalizer_vcvc (carrier-offset shift + phase fix) + ofdm_equalizer_simpledfe for one
// OFDM symbol.  i1 = symbol index within the equaliser's frame, plus one.
__device__ __forceinline__ void equalize_symbol(const float2 *buf, const KP &p, FrameSmem &sm, int off, int i1,
                                                int pset, int bps, const float2 *pts, const uint8_t *lut,
                                                bool want_z)
{
    const float qiw = (bps == p.bps_p && lut == p.lut_p) ? p.qiw_p : p.qiw_h;
    const float arg = (float)(-TWO_PI_D * off * p.cp / p.N * i1);
    float sn, cs;
    sincosf(arg, &sn, &cs);
    const float2 pc = make_float2(cs, sn);
    const float al = p.alpha, oma = 1.0f - p.alpha;
    for (int u = threadIdx.x; u < p.n_occ_u; u += blockDim.x) {
        const int k = p.occ_u[u];
        const int src = k + off;
        float2 y = make_float2(0.f, 0.f);
        if (src >= 0 && src < p.N) y = cmul(ysh(buf, src, p.N), pc);
        const float2 Hk = sm.H[k];
        if (p.n_pil_sets && p.pil_flag[pset * p.N + k]) {
            const float2 q = cdivf(y, p.pil_val[pset * p.N + k]);
            sm.H[k] = make_float2(al * Hk.x + oma * q.x, al * Hk.y + oma * q.y);
            const float2 pv = p.pil_val[pset * p.N + k];      // the serialiser sees the pilot value
            sm.dec[k] = (uint8_t)ofdm_decide(bps, pv.x, pv.y, lut, qiw);
            if (want_z) sm.zs[k] = make_float2(0.f, 0.f);
        } else {
            const float2 z = cdivf(y, Hk);
            const int d = ofdm_decide(bps, z.x, z.y, lut, qiw);
            const float2 q = cdivf(y, pts[d]);
            sm.H[k] = make_float2(al * Hk.x + oma * q.x, al * Hk.y + oma * q.y);
            sm.dec[k] = (uint8_t)d;
            if (want_z) sm.zs[k] = z;
        }
    }
}

#ifdef OFDMX_GENERIC_KERNELS   // the non-template kernels are compiled by ofdmx_api.cu only
__global__ void __launch_bounds__(OFDMX_THREADS)
rx_frame_kernel(const KP p, const float2 *__restrict__ samples, long long n, long long stride,
                const long long *__restrict__ trig, const int *__restrict__ trig_stream,
                const float *__restrict__ cfo, const int *__restrict__ stream_start,
                const int *__restrict__ n_trig_dev, ofdmx_frame *__restrict__ spec,
                uint8_t *__restrict__ bytes_out, long long byte_stride, float2 *__restrict__ z_out,
                long long z_stride)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameSmem sm;
    sm.bufA = reinterpret_cast<float2 *>(smem_raw);
    sm.bufB = sm.bufA + p.N;
    sm.H = sm.bufB + p.N;
    sm.zs = sm.H + p.N;
    sm.scratch = reinterpret_cast<uint32_t *>(sm.zs + p.N);
    sm.dec = reinterpret_cast<uint8_t *>(sm.scratch + 16);
    sm.hb = sm.dec + p.N;
    sm.syms = sm.hb + ((p.hl + 15) & ~15);
    sm.pk = sm.syms + ((p.max_pkt_syms + 15) & ~15);
    __shared__ float wbest[OFDMX_THREADS / 32];
    __shared__ int wbestg[OFDMX_THREADS / 32];
    __shared__ int s_off, s_ok, s_plen, s_pnum, s_psyms, s_fsyms;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nt = *n_trig_dev;
    const int N = p.N, D = p.D;
    const bool want_z = (z_out != nullptr);

    for (int j = blockIdx.x; j < nt; j += gridDim.x) {
        __syncthreads();
        const int st = trig_stream[j];
        const long long t = trig[j];
        const float2 *r = samples + (long long)st * stride;
        const int jend = stream_start[st + 1];
        ofdmx_frame rec;
        rec.trigger = t; rec.cfo = cfo[j]; rec.stream = st; rec.flags = 0; rec.pkt_len = 0; rec.pkt_num = 0;
        rec.frame_syms = 0; rec.carr_offset = 0; rec.slot = (uint32_t)j;
        const int pre = p.nsw + 1;       // OFDM symbols in front of the payload
        if (t + (long long)pre * D > n) {   // header never completes in this buffer
            if (tid == 0) spec[j] = rec;
            continue;
        }
        // ---- sync symbols -> Y1 (bufA), Y2 (bufB)
        load_symbol(sm.bufA, p, r, n, t + p.cp, j, jend, trig, cfo);
        if (p.nsw == 2) load_symbol(sm.bufB, p, r, n, t + D + p.cp, j, jend, trig, cfo);
        __syncthreads();
        fft_smem<false>(sm.bufA, N, p.logN, p.tw);
        if (p.nsw == 2) fft_smem<false>(sm.bufB, N, p.logN, p.tw);
        if (p.nsw == 1) {
            // ---- ofdm_chanest_vcvc with one sync symbol [UPSTREAM get_carr_offset, "Correlate" branch]:
            //      new_diffs[i] = |Y1[i] - Y1[i+2]|^2 (shifted order), sum_j known_diffs[j] * new_diffs[j+g], first maximum
            float *nd = reinterpret_cast<float *>(sm.bufB);
            for (int i = tid; i < N; i += blockDim.x) {
                float v = 0.f;
                if (i < N - 2) {
                    const float2 a = ysh(sm.bufA, i, N), b = ysh(sm.bufA, i + 2, N);
                    const float dr = a.x - b.x, di = a.y - b.y;
                    v = dr * dr + di * di;
                }
                nd[i] = v;
            }
            __syncthreads();
            float best = 0.f;
            int bestg = 0;
            const int ng = (p.gpos - p.gneg) / 2 + 1;
            for (int gi = wid; gi < ng; gi += OFDMX_THREADS / 32) {
                const int g = p.gneg + 2 * gi;
                float acc = 0.f;
                for (int c = lane; c < p.n_cv; c += 32) {
                    const int k = p.cv_k[c] + g;
                    if (k >= 0 && k < N) acc += p.cv_conj[c].x * nd[k];
                }
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (acc > best) { best = acc; bestg = g; }
            }
            if (lane == 0) { wbest[wid] = best; wbestg[wid] = bestg; }
            __syncthreads();
            if (tid == 0) {
                float b = 0.f;
                int g = 0;
                for (int w = 0; w < OFDMX_THREADS / 32; w++) {
                    const float v = wbest[w];
                    if (v > b || (v == b && v > 0.f && wbestg[w] < g)) { b = v; g = wbestg[w]; }
                }
                s_off = g;
            }
            __syncthreads();
        } else {
        // ---- ofdm_chanest_vcvc: integer carrier offset (Schmidl & Cox second stage)
        {
            float best = 0.f;
            int bestg = 0;
            const int ng = (p.gpos - p.gneg) / 2 + 1;
            for (int gi = wid; gi < ng; gi += OFDMX_THREADS / 32) {
                const int g = p.gneg + 2 * gi;
                float2 acc = make_float2(0.f, 0.f);
                for (int c = lane; c < p.n_cv; c += 32) {
                    const int k = p.cv_k[c] + g;
                    const float2 a = ysh(sm.bufA, k, N), b = ysh(sm.bufB, k, N);
                    // conj(Y1) * conj(cv) * Y2
                    const float2 t1 = cmul_conj(b, a);            // Y2 * conj(Y1)
                    acc = cadd(acc, cmul(t1, p.cv_conj[c]));
                }
                for (int o = 16; o > 0; o >>= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
                }
                const float v = acc.x * acc.x + acc.y * acc.y;
                if (v > best) { best = v; bestg = g; }
            }
            if (lane == 0) { wbest[wid] = best; wbestg[wid] = bestg; }
            __syncthreads();
            if (tid == 0) {
                float b = 0.f;
                int g = 0;
                for (int w = 0; w < OFDMX_THREADS / 32; w++) {
                    const float v = wbest[w];
                    if (v > b || (v == b && v > 0.f && wbestg[w] < g)) { b = v; g = wbestg[w]; }
                }
                s_off = g;
            }
            __syncthreads();
        }
        }
        const int off = s_off;
        // ---- channel taps H[k] = Y_ref[k+off] / ref[k]  (ref = sync word 2, or sync word 1 in the single-word mode)
        {
            const float2 *yref = (p.nsw == 2) ? sm.bufB : sm.bufA;
            for (int k = tid; k < N; k += blockDim.x) {
                const int src = k + off;
                float2 hv = make_float2(0.f, 0.f);
                const float2 inv = p.inv_sw2[k];
                if (src >= 0 && src < N && (inv.x != 0.f || inv.y != 0.f)) hv = cmul(ysh(yref, src, N), inv);
                sm.H[k] = hv;
            }
            __syncthreads();
            if (p.nsw == 1 && p.interp) {
                // sync word 1 occupies every second carrier: the taps in between copy their left neighbour
                // [UPSTREAM get_chan_taps, d_interpolate]
                for (int i = p.first_act + 1 + 2 * tid; i < p.last_act; i += 2 * blockDim.x) sm.H[i] = sm.H[i - 1];
                __syncthreads();
                if (tid == 0) sm.H[p.last_act] = sm.H[p.last_act - 1];
            }
        }
        __syncthreads();
        if (p.h_taps)        // debug tap: ofdm_sync_chan_taps as the header equaliser receives it
            for (int k = tid; k < N; k += blockDim.x) p.h_taps[(long long)j * p.h_stride + k] = sm.H[k];
        // ---- header symbol
        load_symbol(sm.bufA, p, r, n, t + (long long)p.nsw * D + p.cp, j, jend, trig, cfo);
        __syncthreads();
        fft_smem<false>(sm.bufA, N, p.logN, p.tw);
        equalize_symbol(sm.bufA, p, sm, off, 1, 0, p.bps_h, p.hpts, p.lut_h, want_z);
        __syncthreads();
        {
            const int b0 = p.occ_base[0];
            for (int q = tid; q < p.hl; q += blockDim.x) {
                const int bin = p.occ_bins[b0 + q];
                sm.hb[q] = sm.dec[bin] ^ p.hdr_mask[q];
                if (want_z) z_out[(long long)j * z_stride + q] = sm.zs[bin];
            }
        }
        __syncthreads();
        if (tid == 0) {
            // packet_header_default::header_parser + packet_header_ofdm::header_parser
            const int bpb = p.bps_h, msk = (1 << bpb) - 1;
            unsigned len = 0, num = 0;
            int k = 0, ok = 1;
            for (int i = 0; i < 12 && k < p.hl; i += bpb, k++) len |= ((unsigned)(sm.hb[k] & msk)) << i;
            if (k < p.hl) {
                for (int i = 0; i < 12 && k < p.hl; i += bpb, k++) num |= ((unsigned)(sm.hb[k] & msk)) << i;
                if (k < p.hl) {
                    const uint8_t crc = crc8_hdr(len, num);
                    for (int i = 0; i < 8 && k < p.hl; i += bpb, k++)
                        if ((sm.hb[k] & msk) != ((crc >> i) & msk)) ok = 0;
                }
            }
            int ps = (int)len * 8 / p.bps_p;
            if (((int)len * 8) % p.bps_p) ps++;
            int fl = 0, acc = 0, s = 0;
            while (acc < ps) { fl++; acc += p.occ_size[s]; s = (s + 1) % p.n_occ_sets; }
            s_ok = ok; s_plen = (int)len; s_pnum = (int)num; s_psyms = ps; s_fsyms = fl;
        }
        // channel state carried to the payload equaliser: H *= exp(+j 2 pi off cp / N * 1)
        {
            const float arg = (float)(TWO_PI_D * off * p.cp / N * 1);
            float sn, cs;
            sincosf(arg, &sn, &cs);
            for (int k = tid; k < N; k += blockDim.x) sm.H[k] = cmul(sm.H[k], make_float2(cs, sn));
        }
        __syncthreads();
        rec.flags = OFDMX_F_HDR_SEEN;
        rec.carr_offset = (int16_t)off;
        rec.pkt_len = (uint16_t)s_plen;
        rec.pkt_num = (uint16_t)s_pnum;
        const int fsyms = s_fsyms, psyms = s_psyms;
        rec.frame_syms = (uint16_t)fsyms;
        if (!s_ok) {
            if (tid == 0) spec[j] = rec;
            continue;
        }
        rec.flags |= OFDMX_F_HDR_OK;
        if (t + (long long)(pre + fsyms) * D > n) {
            // payload never completes in this buffer
            if (tid == 0) spec[j] = rec;
            continue;
        }
        rec.flags |= OFDMX_F_COMPLETE;
        if (s_plen > p.max_pkt_bytes) {
            // the demux consumes the declared payload and searches on behind it; the packet does not fit a slot
            rec.flags |= OFDMX_F_OVERSIZE;
            if (tid == 0) spec[j] = rec;
            continue;
        }
        // ---- payload symbols
        int cnt = 0;                      // serialised symbols so far
        int pset = p.n_pil_sets ? 1 % p.n_pil_sets : 0;
        int set = 1 % p.n_occ_sets;
        for (int i = 0; i < fsyms; i++) {
            load_symbol(sm.bufA, p, r, n, t + (long long)(pre + i) * D + p.cp, j, jend, trig, cfo);
            __syncthreads();
            fft_smem<false>(sm.bufA, N, p.logN, p.tw);
            equalize_symbol(sm.bufA, p, sm, off, i + 1, pset, p.bps_p, p.ppts, p.lut_p, want_z);
            __syncthreads();
            // ofdm_serializer_vcc(..., packet_len_key, symbols_skipped=1) + constellation_decoder_cb
            const int b0 = p.occ_base[set], sz = p.occ_size[set];
            for (int q = tid; q < sz; q += blockDim.x) {
                const int idx = cnt + q;
                if (idx < psyms) {
                    const int bin = p.occ_bins[b0 + q];
                    sm.syms[idx] = sm.dec[bin];
                    if (want_z && p.hl + idx < z_stride) z_out[(long long)j * z_stride + p.hl + idx] = sm.zs[bin];
                }
            }
            cnt = min(cnt + sz, psyms);
            set = (set + 1) % p.n_occ_sets;
            if (p.n_pil_sets) pset = (pset + 1) % p.n_pil_sets;
            __syncthreads();
        }
        // ---- repack_bits_bb(bps, 8, key, align_output=True) + additive_scrambler_bb
        const int nbytes = min(cnt * p.bps_p / 8, p.max_pkt_bytes);
        for (int mb = tid; mb < nbytes; mb += blockDim.x) {
            unsigned v = 0;
            for (int b = 0; b < 8; b++) {
                const int bi = mb * 8 + b;
                const int si = bi / p.bps_p, sb = bi - si * p.bps_p;
                v |= ((unsigned)(sm.syms[si] >> sb) & 1u) << b;
            }
            const uint8_t o = (uint8_t)v ^ p.keystream[mb];
            sm.pk[mb] = o;
            bytes_out[(long long)j * byte_stride + mb] = o;
        }
        __syncthreads();
        // ---- crc32_bb(check=True)
        bool crc_ok = true;
        if (p.crc_mode) {
            if (nbytes < 4) crc_ok = false;
            else {
                const uint32_t c = crc32_block(sm.pk, nbytes - 4, p.crc_tab, p.crc_pow, sm.scratch);
                const uint32_t got = (uint32_t)sm.pk[nbytes - 4] | ((uint32_t)sm.pk[nbytes - 3] << 8)
                                     | ((uint32_t)sm.pk[nbytes - 2] << 16) | ((uint32_t)sm.pk[nbytes - 1] << 24);
                crc_ok = (c == got);
            }
        }
        if (crc_ok) rec.flags |= OFDMX_F_CRC_OK;
        if (tid == 0) spec[j] = rec;
    }
}

// =============================================================================================
// TX
// =============================================================================================
__device__ __forceinline__ int tx_payload_ofdm_syms(const KP &p, int n_syms)
{
    int cntr = 0, acc = 0, s = 1 % p.n_occ_sets;
    while (acc < n_syms) { cntr++; acc += p.occ_size[s]; s = (s + 1) % p.n_occ_sets; }
    return cntr;
}

// sample offsets: exclusive scan of per-packet frame lengths (single CTA)
__global__ void __launch_bounds__(1024)
tx_offsets_kernel(const KP p, const long long *__restrict__ pkt_off, long long n_pkts, long long *__restrict__ sample_off)
{
    __shared__ int wt[33];
    long long carry = 0;
    for (long long b0 = 0; b0 < n_pkts; b0 += 1024) {
        const long long idx = b0 + threadIdx.x;
        int v = 0;
        if (idx < n_pkts) {
            const int lp = (int)(pkt_off[idx + 1] - pkt_off[idx]) + (p.crc_mode ? 4 : 0);
            const int ns = (lp * 8 + p.bps_p - 1) / p.bps_p;
            v = p.nsw + 1 + tx_payload_ofdm_syms(p, ns);     // OFDM symbols of this frame
        }
        int total;
        const int ex = block_excl_scan(v, wt, total);
        // with rolloff every burst carries roll-1 extra samples (the flushed down flank of its last symbol)
        if (idx < n_pkts) sample_off[idx] = (carry + ex) * (long long)p.D + idx * (long long)(p.roll ? p.roll - 1 : 0);
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) sample_off[n_pkts] = carry * (long long)p.D + n_pkts * (long long)(p.roll ? p.roll - 1 : 0);
}

__global__ void __launch_bounds__(OFDMX_THREADS)
tx_frame_kernel(const KP p, const uint8_t *__restrict__ payload, const long long *__restrict__ pkt_off,
                long long n_pkts, int first_num, float2 *__restrict__ out, long long cap,
                const long long *__restrict__ sample_off)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *buf = reinterpret_cast<float2 *>(smem_raw);
    uint32_t *scratch = reinterpret_cast<uint32_t *>(buf + p.N);
    uint8_t *pb = reinterpret_cast<uint8_t *>(scratch + 16);
    uint8_t *hdr = pb + ((p.max_pkt_bytes + 8 + 15) & ~15);
    float2 *tail = reinterpret_cast<float2 *>(hdr + ((p.hl + 15) & ~15));   // delay line of the prefixer [roll-1]
    const int tid = threadIdx.x, N = p.N;
    const int nfl = p.roll ? p.roll - 1 : 0;

    for (long long pk = blockIdx.x; pk < n_pkts; pk += gridDim.x) {
        __syncthreads();
        for (int m = tid; m < nfl; m += blockDim.x) tail[m] = make_float2(0.f, 0.f);
        const long long o0 = pkt_off[pk];
        const int len = (int)(pkt_off[pk + 1] - o0);
        const int lp = len + (p.crc_mode ? 4 : 0);
        if (lp > p.max_pkt_bytes) continue;       // rejected on the host as well
        for (int m = tid; m < len; m += blockDim.x) pb[m] = payload[o0 + m];
        __syncthreads();
        if (p.crc_mode) {   // crc32_bb(check=False): append little-endian CRC
            const uint32_t c = crc32_block(pb, len, p.crc_tab, p.crc_pow, scratch);
            if (tid < 4) pb[len + tid] = (uint8_t)(c >> (8 * tid));
        }
        if (tid == 0) {
            // packet_header_default::header_formatter + packet_header_ofdm scramble mask
            const int bpb = p.bps_h, msk = (1 << bpb) - 1;
            const unsigned plen = (unsigned)lp & 0xFFF, pnum = (unsigned)(first_num + (int)pk) & 0xFFF;
            const uint8_t crc = crc8_hdr(plen, pnum);
            for (int k = 0; k < p.hl; k++) hdr[k] = 0;
            int k = 0;
            for (int i = 0; i < 12 && k < p.hl; i += bpb, k++) hdr[k] = (uint8_t)((plen >> i) & msk);
            for (int i = 0; i < 12 && k < p.hl; i += bpb, k++) hdr[k] = (uint8_t)((pnum >> i) & msk);
            for (int i = 0; i < 8 && k < p.hl; i += bpb, k++) hdr[k] = (uint8_t)((crc >> i) & msk);
            for (int q = 0; q < p.hl; q++) hdr[q] ^= p.hdr_mask[q];
        }
        __syncthreads();
        for (int m = tid; m < lp; m += blockDim.x) pb[m] ^= p.keystream[m];   // additive_scrambler_bb
        __syncthreads();
        const int ns = (lp * 8 + p.bps_p - 1) / p.bps_p;       // repack_bits_bb(8, bps, key, False)
        const int n_ofdm = p.nsw + 1 + tx_payload_ofdm_syms(p, ns);
        const long long base = sample_off[pk];
        if (base + (long long)n_ofdm * p.D + nfl > cap) continue;
        int sym_base = 0, set = 0;
        for (int o = 0; o < n_ofdm; o++) {
            // ofdm_carrier_allocator_cvc: sync words, then data on occupied bins, pilots on top
            if (o < p.nsw) {
                const float2 *sw = (o == 0) ? p.sw1 : p.sw2;
                for (int ks = tid; ks < N; ks += blockDim.x) buf[bitrev(ks ^ (N >> 1), p.logN)] = sw[ks];
            } else {
                for (int k = tid; k < N; k += blockDim.x) buf[k] = make_float2(0.f, 0.f);
                __syncthreads();
                const int di = o - p.nsw;
                const int b0 = p.occ_base[set], sz = p.occ_size[set];
                for (int q = tid; q < sz; q += blockDim.x) {
                    const int idx = sym_base + q;
                    if (idx >= p.hl + ns) continue;
                    float2 v;
                    if (idx < p.hl) v = p.hpts[hdr[idx]];
                    else {
                        const int si = idx - p.hl;
                        unsigned c = 0;
                        for (int b = 0; b < p.bps_p; b++) {
                            const int bi = si * p.bps_p + b;
                            if (bi < lp * 8) c |= ((unsigned)(pb[bi >> 3] >> (bi & 7)) & 1u) << b;
                        }
                        v = p.ppts[c];
                    }
                    buf[bitrev(p.occ_bins[b0 + q] ^ (N >> 1), p.logN)] = v;
                }
                sym_base += sz;
                set = (set + 1) % p.n_occ_sets;
                __syncthreads();
                if (p.n_pil_sets) {
                    const int ps = di % p.n_pil_sets, pss = di % p.n_pil_sym_sets;
                    for (int q = tid; q < p.pil_size[ps]; q += blockDim.x)
                        buf[bitrev(p.pil_bins[p.pil_base[ps] + q] ^ (N >> 1), p.logN)] = p.pil_sym[p.pil_sym_base[pss] + q];
                }
            }
            __syncthreads();
            // fft_vcc(inverse, shift) + ofdm_cyclic_prefixer(rolloff 0) + multiply_const(tx_scale)
            fft_smem<true>(buf, N, p.logN, p.tw);
            float2 *dst = out + base + (long long)o * p.D;
            for (int m = tid; m < p.D; m += blockDim.x) {
                float2 v = buf[(m - p.cp + N) & (N - 1)];
                if (m < nfl) {
                    // ofdm_cyclic_prefixer with rolloff: out[i] = out[i]*up[i] + delay[i]; delay[i] = in[i]*down[i]
                    // (nfl <= cp, so thread m owns tail[m] and only reads buf)
                    const float up = p.roll_flank[m], dn = p.roll_flank[nfl + m];
                    const float2 t = tail[m], h = buf[m];
                    v = make_float2(__fadd_rn(__fmul_rn(v.x, up), t.x), __fadd_rn(__fmul_rn(v.y, up), t.y));
                    tail[m] = make_float2(__fmul_rn(h.x, dn), __fmul_rn(h.y, dn));
                }
                float2 o = make_float2(v.x * p.tx_scale, v.y * p.tx_scale);
                if (p.tx_clip > 0.f) {      // clipper: analog.rail_ff(-c, c) on re and im
                    o.x = o.x < -p.tx_clip ? -p.tx_clip : (o.x > p.tx_clip ? p.tx_clip : o.x);
                    o.y = o.y < -p.tx_clip ? -p.tx_clip : (o.y > p.tx_clip ? p.tx_clip : o.y);
                }
                dst[m] = o;
            }
            __syncthreads();
        }
        // packet mode: the delay line is flushed behind the last symbol (and cleared for the next packet)
        for (int m = tid; m < nfl; m += blockDim.x) {
            const float2 t = tail[m];
            float2 o = make_float2(t.x * p.tx_scale, t.y * p.tx_scale);
            if (p.tx_clip > 0.f) {
                o.x = o.x < -p.tx_clip ? -p.tx_clip : (o.x > p.tx_clip ? p.tx_clip : o.x);
                o.y = o.y < -p.tx_clip ? -p.tx_clip : (o.y > p.tx_clip ? p.tx_clip : o.y);
            }
            out[base + (long long)n_ofdm * p.D + m] = o;
        }
    }
}

// =============================================================================================
// single blocks
// =============================================================================================
// fft.fft_vcc(N, forward, (), shift=True): one CTA per symbol
template <bool INVERSE>
__global__ void __launch_bounds__(OFDMX_THREADS)
fft_vcc_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, long long n_syms, int N, int logN,
               const float2 *__restrict__ tw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *buf = reinterpret_cast<float2 *>(smem_raw);
    for (long long s = blockIdx.x; s < n_syms; s += gridDim.x) {
        __syncthreads();
        for (int m = threadIdx.x; m < N; m += blockDim.x) {
            // inverse: input is in shifted order -> natural index m ^ N/2
            const int nat = INVERSE ? (m ^ (N >> 1)) : m;
            buf[bitrev(nat, logN)] = in[s * N + m];
        }
        __syncthreads();
        fft_smem<INVERSE>(buf, N, logN, tw);
        for (int m = threadIdx.x; m < N; m += blockDim.x) {
            // forward: output in shifted order
            out[s * N + m] = INVERSE ? buf[m] : buf[m ^ (N >> 1)];
        }
    }
}

__global__ void __launch_bounds__(OFDMX_THREADS)
crc32_kernel(const uint8_t *__restrict__ bytes, const long long *__restrict__ pkt_off, long long n_pkts,
             uint32_t *__restrict__ crc_out, const uint32_t *__restrict__ tab, const uint32_t *__restrict__ powtab)
{
    __shared__ uint32_t scratch[16];
    for (long long pk = blockIdx.x; pk < n_pkts; pk += gridDim.x) {
        __syncthreads();
        const long long o0 = pkt_off[pk];
        const int len = (int)(pkt_off[pk + 1] - o0);
        const uint32_t c = crc32_block(bytes + o0, len, tab, powtab, scratch);
        if (threadIdx.x == 0) crc_out[pk] = c;
    }
}

#endif  // OFDMX_GENERIC_KERNELS
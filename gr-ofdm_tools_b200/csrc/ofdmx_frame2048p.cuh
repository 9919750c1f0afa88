// ofdmx_frame2048p.cuh -- K1+K3+K4 at fft_len 2048 with a PAIR OF WARPS PER FRAME: rx_framep_kernel<bps, want_z>.
//
// The one-warp-per-frame kernel needs 28 KB of shared memory per frame at fft_len 2048 (two 1024-bin half spectra and
// 1200 channel taps), which leaves 7 warps on an SM.  Here a frame belongs to two warps that split it by carrier
// parity with a decimation-in-frequency first stage:
//     a[n] = x[n] + x[n + 1024]                    ->  1024-point FFT  ->  X[2k]      (warp 0 of the pair)
//     b[n] = (x[n] - x[n + 1024]) W_2048^n         ->  1024-point FFT  ->  X[2k + 1]  (warp 1 of the pair)
// Each warp then is a fft_len-1024-sized worker: 32x32 register transform, its 600 carriers, its 600 channel taps
// (8.7 + 4.8 KB), equaliser and demapper over its own carriers -- the integer carrier offset is even, so a carrier and
// the bin it is read from always have the same parity and no spectrum data ever crosses between the two warps.  What
// does cross goes through a few bytes of pair-shared memory behind a named barrier (bar.sync id, 64):
//     the partial sums of the carrier-offset metric (once per frame), the 32 header bits (once), the decisions of a
//     symbol (one byte each, double buffered; the bytes of a symbol are packed by both warps, 3 bytes from 4 decisions
//     at 64-QAM), and the finished packet for the CRC (warp 0).
// Both warps load all 2048 samples of a symbol (the second read hits L1 / L2); the NCO costs one extra complex
// multiply per output (x[n+1024] is rotated onto x[n]'s phase by the per-frame constant exp(j 2 pi kappa 1024)), and
// the W_2048^n factor of the odd half is folded into the phasor recurrence.
//
// Preconditions (host): fft_len 2048, two sync words, one carrier set without repeated carriers, no pilot inside it,
// BPSK header of >= 32 items, bits per OFDM symbol a multiple of 8, <= 4 carrier-offset candidates.
#pragma once
#include "ofdmx_frame1024w.cuh"

// Pairs per CTA (one CTA per SM).  Seven fit the shared memory, but with more than three warps on a scheduler the
// register file caps a thread at 128 registers and this kernel then spills (19.7 ms on configs[3]); six pairs run at 168
// registers without spills (18.9 ms).
#ifndef FP_PAIRS
#define FP_PAIRS 6
#endif
#define FP_THREADS (FP_PAIRS * 64)

__device__ __forceinline__ void fp_pair_sync(int pair)
{
    asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
}

// exact NCO phasor at item i (piecewise accumulation over the raw triggers), fft_len 2048; cold path
static __device__ __noinline__ float2 f2kp_exact_phasor(long long i, int j, int jend, const long long *__restrict__ trig,
                                                        const float *__restrict__ cfo)
{
    double turns = nco_turns(i, j, jend, trig, cfo, 2048);
    turns -= rint(turns);
    float s2, c2;
    sincospif(2.0f * (float)turns, &s2, &c2);
    return make_float2(c2, s2);
}

// One half-symbol: load, derotate, fold the two halves (sum for half 0, W-weighted difference for half 1), 1024-point
// FFT.  Result: Tw[m] = X[2 m + half], natural order.  st = phasor of 32 samples of NCO advance (half 1: times
// W_2048^32), q1024 = phasor of 1024 samples.
__device__ __forceinline__ void f2kp_symbol(const KP &p, const float2 *__restrict__ r, long long n, long long i0, long long t,
                                            double kappa, float2 st, float2 q1024, bool slow, int j, int jend,
                                            const long long *__restrict__ trig, const float *__restrict__ cfo,
                                            float2 *__restrict__ Tw, const float2 *__restrict__ tws, int lane, int half)
{
    float2 v[32];
    const long long sbase = i0 - p.D + lane;          // stream index of this lane's first sample
    const float sgn = half ? -1.0f : 1.0f;
    // `slow`: another raw trigger falls inside (or before) this symbol, where the sample-and-hold value of the NCO
    // changes.  On back-to-back frames that is the NEXT FRAME's trigger arriving a few samples early (its jitter), i.e.
    // nearly every other frame's last symbol, with only the last few samples of the upper half affected: those are
    // patched after the uniform-frequency path (c = first affected sample of the upper half); anything else takes
    // the per-sample path below.
    int c = 1024;
    if (slow) {
        const long long tnx = (j + 1 < jend) ? trig[j + 1] : 0x7fffffffffffffffLL;
        c = (tnx - i0 - 1024 >= 800) ? (int)min(tnx - i0 - 1024, 1024LL) : -1;
    }
    if (c >= 0) {
        const float2 qs = make_float2(sgn * q1024.x, sgn * q1024.y);
        if (sbase - lane >= 0 && sbase - lane + 2048 <= n) {
#pragma unroll
            for (int a = 0; a < 32; a++) v[a] = half ? __ldg(&r[sbase + 32 * a]) : f1k_ld_stream(&r[sbase + 32 * a]);   // one warp of the pair allocates in L1 (the other's read of the same line merges with it), the other streams
            // (the upper half in four groups of eight: 64 more registers for it do not exist; measured alternatives --
            // cp.async of the upper half into the transpose buffer, one rolled copy of the 32-point transform -- cost
            // more in spills than they gained)
#pragma unroll
            for (int g = 0; g < 4; g++) {
                float2 hi[8];
#pragma unroll
                for (int a = 0; a < 8; a++) hi[a] = half ? __ldg(&r[sbase + 1024 + 32 * (8 * g + a)]) : f1k_ld_stream(&r[sbase + 1024 + 32 * (8 * g + a)]);
#pragma unroll
                for (int a = 0; a < 8; a++) v[8 * g + a] = cadd(v[8 * g + a], cmul(hi[a], qs));
            }
        } else {
#pragma unroll
            for (int a = 0; a < 32; a++) {
                const long long s = sbase + 32 * a, s2 = s + 1024;
                const float2 lo = (s >= 0 && s < n) ? __ldg(&r[s]) : make_float2(0.f, 0.f);
                const float2 hi = (s2 >= 0 && s2 < n) ? __ldg(&r[s2]) : make_float2(0.f, 0.f);
                v[a] = cadd(lo, cmul(hi, qs));
            }
        }
        // phase(i) = 2 pi kappa (i - t + 1) [- 2 pi m / 2048 for the odd half, m = lane + 32 a]
        double tb = kappa * (double)(i0 + lane - t + 1) - (half ? (double)lane * (1.0 / 2048.0) : 0.0);
        tb -= rint(tb);
        float sn, cs;
        sincospif(2.0f * (float)tb, &sn, &cs);
        float2 ph = make_float2(cs, sn);
#pragma unroll
        for (int a = 0; a < 32; a++) {
            v[a] = cmul(v[a], ph);
            ph = cmul(ph, st);
        }
        if (c < 1024) {
            // upper-half samples m >= c (c >= 800: rows 24 .. 31 at most): add x (P_exact - P_uniform) [W_2048^m]
            const int a0 = (c - 31) >> 5;
#pragma unroll 1
            for (int a = a0; a < 32; a++) {
                const int m = lane + 32 * a;
                float2 d = make_float2(0.f, 0.f);
                if (m >= c) {
                    const long long i = i0 + m + 1024, sidx = i - p.D;
                    const float2 x = (sidx >= 0 && sidx < n) ? __ldg(&r[sidx]) : make_float2(0.f, 0.f);
                    const float2 pe = f2kp_exact_phasor(i, j, jend, trig, cfo);
                    double tb2 = kappa * (double)(i - t + 1);
                    tb2 -= rint(tb2);
                    float s2, c2;
                    sincospif(2.0f * (float)tb2, &s2, &c2);
                    d = cmul(x, make_float2(sgn * (pe.x - c2), sgn * (pe.y - s2)));
                    if (half) {
                        sincospif(-(float)m * (1.0f / 1024.0f), &s2, &c2);     // W_2048^m
                        d = cmul(d, make_float2(c2, s2));
                    }
                }
                Tw[a * F1K_ROW + lane] = d;
            }
            __syncwarp();
#pragma unroll
            for (int a = 24; a < 32; a++)
                if (a >= a0) v[a] = cadd(v[a], Tw[a * F1K_ROW + lane]);
            __syncwarp();
        }
    } else {
        // the general case: the two halves no longer differ by a constant phasor.  Every sample gets its own phase
        // (exact piecewise accumulation behind the trigger, closed form in front of it).
        const long long tnx = (j + 1 < jend) ? trig[j + 1] : 0x7fffffffffffffffLL;
#pragma unroll 1
        for (int a = 0; a < 32; a++) {
            const int m = lane + 32 * a;
            float2 u = make_float2(0.f, 0.f);
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                const long long i = i0 + m + 1024 * hh, s = i - p.D;
                const float2 x = (s >= 0 && s < n) ? __ldg(&r[s]) : make_float2(0.f, 0.f);
                float2 pa;
                if (i >= tnx) pa = f2kp_exact_phasor(i, j, jend, trig, cfo);
                else {
                    double tb = kappa * (double)(i - t + 1);
                    tb -= rint(tb);
                    float sn, cs;
                    sincospif(2.0f * (float)tb, &sn, &cs);
                    pa = make_float2(cs, sn);
                }
                const float2 y = cmul(x, pa);
                u = hh ? make_float2(u.x + sgn * y.x, u.y + sgn * y.y) : y;
            }
            if (half) {
                float sn, cs;
                sincospif(-(float)m * (1.0f / 1024.0f), &sn, &cs);     // W_2048^m
                u = cmul(u, make_float2(cs, sn));
            }
            // v[] is a register array: park the value in the transpose buffer and fetch it statically below
            Tw[a * F1K_ROW + lane] = u;
        }
        __syncwarp();
#pragma unroll
        for (int a = 0; a < 32; a++) v[a] = Tw[a * F1K_ROW + lane];
        __syncwarp();
    }
    // 1024-point transform as 32 x 32 (the body of f1k_symbol)
    fft32_fwd(v);
#pragma unroll
    for (int q = 0; q < 32; q++) {
        const int k1 = brev5(q);
        Tw[k1 * F1K_ROW + lane] = cmul(v[q], tws[k1 * 32 + lane]);
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 16; q++) {
        const float4 t4 = *reinterpret_cast<const float4 *>(&Tw[lane * F1K_ROW + 2 * q]);
        v[2 * q] = make_float2(t4.x, t4.y);
        v[2 * q + 1] = make_float2(t4.z, t4.w);
    }
    __syncwarp();
    fft32_fwd(v);
#pragma unroll
    for (int q = 0; q < 32; q++) Tw[lane + 32 * brev5(q)] = v[q];
}

// pair_tab (host-built, uint16): [cnt0, cnt1, NH, 0, then for h = 0, 1: idx[NH] (natural bin >> 1), pos[NH] (serialiser
// position), ks[NH] (shifted bin)]: the occupied carriers of natural-bin parity h in serialiser order.
template <int BPS_P, bool WANT_Z>
__global__ void __launch_bounds__(FP_THREADS, 1)
rx_framep_kernel(const KP p, const float2 *__restrict__ samples, long long n, long long stride,
                 const long long *__restrict__ trig, const int *__restrict__ trig_stream,
                 const float *__restrict__ cfo, const int *__restrict__ stream_start,
                 const int *__restrict__ n_trig_dev, ofdmx_frame *__restrict__ spec,
                 uint8_t *__restrict__ bytes_out, long long byte_stride, float2 *__restrict__ z_out,
                 long long z_stride, uint32_t x_2048, const uint16_t *__restrict__ pair_tab, int hsz)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, NTH = blockDim.x;
    const int pair = wid >> 1, half = wid & 1;
    const int nu = p.n_occ_u;
    const int NH = pair_tab[2], NH8 = (NH + 7) & ~7;
    constexpr int N = 2048, HALF = 1024;
    // ---- CTA-shared tables
    float2 *tws = reinterpret_cast<float2 *>(smem_raw);           // [1024] W_1024^(b k1)
    float2 *ipts = tws + 1024;                                    // [64] (1 - alpha) / constellation point
    uint32_t *s_tab = reinterpret_cast<uint32_t *>(ipts + 64);    // [256]
    uint32_t *s_pow = s_tab + 256;                                // [32]
    uint8_t *lut = reinterpret_cast<uint8_t *>(s_pow + 32);       // [64]
    uint16_t *tabs = reinterpret_cast<uint16_t *>(lut + 64);      // [2][3][NH8]
    const size_t shared_bytes = 1024 * 8 + 64 * 8 + 256 * 4 + 32 * 4 + 64 + (size_t)6 * NH8 * 2;
    // ---- per-pair and per-warp buffers
    const int decN = (nu + 15) & ~15;
    const size_t pair_bytes = (size_t)2 * decN + 64 + 64;
    const size_t warp_bytes = (size_t)F1K_SLOT * 8 + (size_t)hsz * 8 + sizeof(FwState);
    unsigned char *pbase = smem_raw + ((shared_bytes + 15) & ~(size_t)15) + (size_t)pair * (pair_bytes + 2 * warp_bytes);
    uint8_t *dec2 = pbase;                                        // [2][decN] decisions of the current / previous symbol
    uint8_t *hb = dec2 + 2 * decN;                                // [64] header items
    float2 *xch = reinterpret_cast<float2 *>(hb + 64);            // [2][4] partial offset metrics of the two warps
    unsigned char *wbase = pbase + pair_bytes + (size_t)half * warp_bytes;
    float2 *Y = reinterpret_cast<float2 *>(wbase);                // this warp's half spectrum
    float2 *Hs = Y + F1K_SLOT;                                    // channel taps of this warp's carriers (parks Y1 first)
    volatile FwState *fs = reinterpret_cast<volatile FwState *>(Hs + hsz);

    for (int i = tid; i < 1024; i += NTH) {
        const int k1 = i >> 5, b = i & 31;
        float sn, cs;
        sincospif(-(float)(b * k1) * (2.0f / 1024.0f), &sn, &cs);
        tws[i] = make_float2(cs, sn);
    }
    for (int i = tid; i < 6 * NH8; i += NTH) {
        const int a = i / NH8, q = i - a * NH8;
        tabs[i] = (q < NH) ? pair_tab[4 + a * NH + q] : 0;
    }
    for (int i = tid; i < 256; i += NTH) s_tab[i] = p.crc_tab[i];
    for (int i = tid; i < 32; i += NTH) s_pow[i] = p.crc_pow64[i];
    for (int i = tid; i < 64; i += NTH) {
        lut[i] = p.lut_p[i];
        const float2 ip = (i < (1 << BPS_P)) ? p.inv_ppts[i] : make_float2(0.f, 0.f);
        ipts[i] = make_float2((1.0f - p.alpha) * ip.x, (1.0f - p.alpha) * ip.y);
    }
    __syncthreads();

    const int cnt = pair_tab[half];                               // carriers of this warp
    const uint16_t *t_idx = tabs + (size_t)(3 * half) * NH8, *t_pos = t_idx + NH8, *t_ks = t_pos + NH8;
    const int nt = *n_trig_dev;
    const int D = p.D;
    const float al = p.alpha, oma = 1.0f - p.alpha, qiw = p.qiw_p;
    const int size0 = p.occ_size[0];
    const int sym_bytes = size0 * BPS_P / 8;
    const int ng = (p.gpos - p.gneg) / 2 + 1;
    const int y1_lo = p.y1_lo;
    const unsigned hmask32 = __ballot_sync(0xffffffffu, p.hdr_mask[lane] & 1);
    const bool words_ok = ((reinterpret_cast<uintptr_t>(bytes_out) | (uintptr_t)byte_stride) & 15) == 0;

    for (int j = blockIdx.x * FP_PAIRS + pair; j < nt; j += gridDim.x * FP_PAIRS) {
        const int st = trig_stream[j];
        const long long t = trig[j];
        const float2 *r = samples + (long long)st * stride;
        const int jend = stream_start[st + 1];
        const long long rem = n - t;
        if (lane == 0) {
            const float cf = cfo[j];
            fs->rec.trigger = t; fs->rec.cfo = cf; fs->rec.stream = st; fs->rec.flags = 0; fs->rec.pkt_len = 0;
            fs->rec.pkt_num = 0; fs->rec.frame_syms = 0; fs->rec.carr_offset = 0; fs->rec.slot = (uint32_t)j;
            const double kap = (double)cf * (-2.0 / 2048.0) * (1.0 / TWO_PI_D);
            double ts = kap * 32.0 - (half ? 1.0 / 64.0 : 0.0);      // 32 samples of advance (odd half: times W_2048^32)
            ts -= rint(ts);
            float sn, cs;
            sincospif(2.0f * (float)ts, &sn, &cs);
            fs->kappa = kap; fs->stx = cs; fs->sty = sn;
            fs->tnext = (j + 1 < jend) ? trig[j + 1] : 0x7fffffffffffffffLL;
            fs->nsym = 3;
            fs->nbytes = 0;
        }
        __syncwarp();
        if (3LL * D > rem) {
            if (half == 0) fw_flush(fs, spec + j, lane);
            continue;
        }
        float2 q1024;
        {
            double tq = fs->kappa * 1024.0;
            tq -= rint(tq);
            float sn, cs;
            sincospif(2.0f * (float)tq, &sn, &cs);
            q1024 = make_float2(cs, sn);
        }
        int off = 0, psyms = 0;
        bool dead = false;
        for (int sidx = 0; sidx < fs->nsym; sidx++) {
            const long long i0 = t + (long long)sidx * D + p.cp;
            f2kp_symbol(p, r, n, i0, t, fs->kappa, make_float2(fs->stx, fs->sty), q1024, fs->tnext <= i0 + 2047, j, jend,
                        trig, cfo, Y, tws, lane, half);
            if (half == 0) {   // pull the next symbol towards L2 while this one is processed
                const long long sn = i0 + D - p.D + lane * 64;
                if (sn >= 0 && sn + 64 <= n) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r + sn));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r + sn + 16));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r + sn + 32));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r + sn + 48));
                }
            }
            __syncwarp();
            if (sidx == 0) {
                // sync word 1: park the bins of this parity that the offset search reads (shifted bins y1_lo + q)
                const int q0 = ((y1_lo ^ half) & 1);                 // first q whose bin has this warp's parity
                for (int q = q0 + 2 * lane; q < p.y1_span; q += 64) Hs[q >> 1] = Y[((y1_lo + q) ^ HALF) >> 1];
            } else if (sidx == 1) {
                // sync word 2: integer carrier offset; each warp sums the correlation terms of its parity
                float2 acc[4];
#pragma unroll
                for (int gi = 0; gi < 4; gi++) acc[gi] = make_float2(0.f, 0.f);
                for (int c = lane; c < p.n_cv; c += 32) {
                    const int kc = p.cv_k[c];
                    if (((kc ^ HALF) & 1) != half) continue;
                    const float2 cvc = p.cv_conj[c];
#pragma unroll
                    for (int gi = 0; gi < 4; gi++)
                        if (gi < ng) {
                            const int k = kc + p.gneg + 2 * gi;
                            acc[gi] = cadd(acc[gi], cmul(cmul_conj(Y[(k ^ HALF) >> 1], Hs[(k - y1_lo) >> 1]), cvc));
                        }
                }
#pragma unroll
                for (int gi = 0; gi < 4; gi++) {
                    for (int o = 16; o > 0; o >>= 1) {
                        acc[gi].x += __shfl_xor_sync(0xffffffffu, acc[gi].x, o);
                        acc[gi].y += __shfl_xor_sync(0xffffffffu, acc[gi].y, o);
                    }
                    if (lane == 0) xch[half * 4 + gi] = acc[gi];
                }
                fp_pair_sync(pair);
                float b = 0.f;
#pragma unroll
                for (int gi = 0; gi < 4; gi++) {
                    const float2 a0 = xch[gi], a1 = xch[4 + gi];
                    const float sx = a0.x + a1.x, sy = a0.y + a1.y;
                    const float v = sx * sx + sy * sy;
                    if (gi < ng && v > b) { b = v; off = p.gneg + 2 * gi; }
                }
                __syncwarp();
                // H[k] = Y2[k + off] / sw2[k] for this warp's carriers (overwrites the parked Y1 bins)
                for (int u = lane; u < cnt; u += 32) {
                    const int ks = t_ks[u], src = ks + off;
                    float2 Hk = make_float2(0.f, 0.f);
                    if (src >= 0 && src < N) Hk = cmul(Y[(src ^ HALF) >> 1], p.inv_sw2[ks]);
                    Hs[u] = Hk;
                }
                if (WANT_Z && p.h_taps) {
                    __syncwarp();
                    for (int u = lane; u < cnt; u += 32) p.h_taps[(long long)j * p.h_stride + t_ks[u]] = Hs[u];
                }
            } else if (sidx == 2) {
                // header symbol: frame equaliser (offset shift + phase fix) + simpledfe with the BPSK header
                float2 pc = make_float2(1.f, 0.f), rot = make_float2(1.f, 0.f);
                if (off != 0) {
                    float sn, cs;
                    sincospif((float)((off * p.cp) & (N - 1)) * (2.0f / N), &sn, &cs);
                    pc = make_float2(cs, -sn);
                    rot = make_float2(cs, sn);
                }
                for (int u = lane; u < cnt; u += 32) {
                    float2 y;
                    if (off == 0) y = Y[t_idx[u]];
                    else {
                        const int src = (int)t_ks[u] + off;
                        y = (src >= 0 && src < N) ? cmul(Y[(src ^ HALF) >> 1], pc) : make_float2(0.f, 0.f);
                    }
                    float2 Hk = Hs[u];
                    const float2 hn = cmul_conj(y, Hk);
                    const int d = hn.x > 0.f;
                    const float2 q = make_float2(d ? y.x : -y.x, d ? y.y : -y.y);
                    const int pos = t_pos[u];
                    if (pos < 64) hb[pos] = (uint8_t)d;
                    if (WANT_Z) {
                        const float hinv = f1k_rcp(fmaf(Hk.x, Hk.x, Hk.y * Hk.y));
                        z_out[(unsigned long long)(unsigned)j * (unsigned long long)z_stride + pos] = make_float2(hn.x * hinv, hn.y * hinv);
                    }
                    Hk = make_float2(fmaf(al, Hk.x, oma * q.x), fmaf(al, Hk.y, oma * q.y));
                    Hs[u] = (off != 0) ? cmul(Hk, rot) : Hk;
                }
                fp_pair_sync(pair);
                const unsigned bits = __ballot_sync(0xffffffffu, hb[lane] & 1) ^ hmask32;
                const int plen = (int)(bits & 0xFFFu);
                const int pnum = (int)((bits >> 12) & 0xFFFu);
                unsigned c8 = (lane < 24 && ((bits >> lane) & 1u)) ? (unsigned)p.crc8_bit[lane] : 0u;
                for (int o = 16; o > 0; o >>= 1) c8 ^= __shfl_xor_sync(0xffffffffu, c8, o);
                const bool ok = ((c8 ^ p.crc8_zero) == (bits >> 24));
                psyms = (plen * 8 + BPS_P - 1) / BPS_P;
                const int fsyms = (psyms + size0 - 1) / size0;
                const bool present = ok && (long long)(3 + fsyms) * D <= rem;
                const bool oversize = present && plen > p.max_pkt_bytes;
                const bool complete = present && !oversize;
                if (lane == 0) {
                    fs->rec.flags = OFDMX_F_HDR_SEEN | (ok ? OFDMX_F_HDR_OK : 0) | (present ? OFDMX_F_COMPLETE : 0)
                                    | (oversize ? OFDMX_F_OVERSIZE : 0);
                    fs->rec.carr_offset = (int16_t)off;
                    fs->rec.pkt_len = (uint16_t)plen;
                    fs->rec.pkt_num = (uint16_t)pnum;
                    fs->rec.frame_syms = (uint16_t)fsyms;
                    fs->nbytes = min(psyms * BPS_P / 8, p.max_pkt_bytes);
                    if (complete) fs->nsym = 3 + fsyms;
                }
                __syncwarp();
                if (!complete) { dead = true; break; }
            } else {
                // payload symbol i: this warp equalises and demaps its carriers, the pair packs the bytes
                const int i = sidx - 3;
                float2 pc = make_float2(1.f, 0.f);
                if (off != 0) {
                    float sn, cs;
                    sincospif((float)((off * p.cp * (i + 1)) & (N - 1)) * (2.0f / N), &sn, &cs);
                    pc = make_float2(cs, -sn);
                }
                const int cb = i * size0;
                uint8_t *decw = dec2 + (i & 1) * decN;
                for (int u0 = lane; u0 < cnt; u0 += 64) {
                    float2 y[2], Hk[2];
                    int pos[2];
#pragma unroll
                    for (int jj = 0; jj < 2; jj++) {
                        const int u = u0 + 32 * jj, uc = (u < cnt) ? u : lane;
                        if (off == 0) y[jj] = Y[t_idx[uc]];
                        else {
                            const int src = (int)t_ks[uc] + off;
                            y[jj] = (src >= 0 && src < N) ? cmul(Y[(src ^ HALF) >> 1], pc) : make_float2(0.f, 0.f);
                        }
                        Hk[jj] = Hs[uc];
                        pos[jj] = t_pos[uc];
                    }
                    int d[2];
                    float2 hq[2], z[2];
#pragma unroll
                    for (int jj = 0; jj < 2; jj++) {
                        const float rinv = f1k_rcp(fmaf(Hk[jj].x, Hk[jj].x, Hk[jj].y * Hk[jj].y));
                        const float2 nn = cmul_conj(y[jj], Hk[jj]);
                        if (WANT_Z) {
                            z[jj] = make_float2(nn.x * rinv, nn.y * rinv);
                            d[jj] = f1k_decide<BPS_P>(z[jj].x, z[jj].y, lut, qiw);
                        } else {
                            d[jj] = f1k_decide<BPS_P>(nn.x, nn.y, lut, rinv * qiw);
                        }
                        const float2 q = cmul(y[jj], ipts[d[jj]]);
                        hq[jj] = make_float2(fmaf(al, Hk[jj].x, q.x), fmaf(al, Hk[jj].y, q.y));
                    }
#pragma unroll
                    for (int jj = 0; jj < 2; jj++) {
                        const int u = u0 + 32 * jj;
                        if (u < cnt) {
                            Hs[u] = hq[jj];
                            decw[pos[jj]] = (uint8_t)d[jj];
                            if (WANT_Z) {
                                const int idx = cb + pos[jj];
                                if (idx < psyms && p.hl + idx < z_stride) z_out[(unsigned long long)(unsigned)j * (unsigned long long)z_stride + p.hl + idx] = z[jj];
                            }
                        }
                    }
                }
                fp_pair_sync(pair);           // all decisions of this symbol are in decw
                // repack_bits_bb(bps, 8) + additive_scrambler_bb: the two warps take alternate groups of bytes
                const int b0 = i * sym_bytes;
                const int nbytes = fs->nbytes;
                uint8_t *orow = bytes_out + (unsigned long long)(unsigned)j * (unsigned long long)byte_stride;
                if (BPS_P == 6) {
                    // 4 decisions -> 3 bytes
                    for (int g = half * 32 + lane; 3 * g < sym_bytes; g += 64) {
                        const uint32_t w4 = *reinterpret_cast<const uint32_t *>(decw + 4 * g);
                        const uint32_t v = (w4 & 0x3Fu) | ((w4 >> 2) & 0xFC0u) | ((w4 >> 4) & 0x3F000u) | ((w4 >> 6) & 0xFC0000u);
#pragma unroll
                        for (int b = 0; b < 3; b++) {
                            const int gb = b0 + 3 * g + b;
                            if (gb < nbytes && 3 * g + b < sym_bytes) orow[gb] = (uint8_t)(v >> (8 * b)) ^ __ldg(&p.keystream[gb]);
                        }
                    }
                } else {
                    for (int m = half * 32 + lane; m < sym_bytes; m += 64) {
                        const int gb = b0 + m;
                        if (gb >= nbytes) break;
                        unsigned v = 0;
                        if (BPS_P == 4) v = (unsigned)decw[2 * m] | ((unsigned)decw[2 * m + 1] << 4);
                        else if (BPS_P == 2)
                            v = (unsigned)decw[4 * m] | ((unsigned)decw[4 * m + 1] << 2) | ((unsigned)decw[4 * m + 2] << 4) | ((unsigned)decw[4 * m + 3] << 6);
                        else if (BPS_P == 1) {
                            for (int b = 0; b < 8; b++) v |= (unsigned)decw[8 * m + b] << b;
                        } else {
                            for (int b = 0; b < 8; b++) {
                                const int bi = m * 8 + b;
                                const int si = bi / BPS_P, sb = bi - si * BPS_P;
                                v |= ((unsigned)(decw[si] >> sb) & 1u) << b;
                            }
                        }
                        orow[gb] = (uint8_t)v ^ __ldg(&p.keystream[gb]);
                    }
                }
            }
            __syncwarp();
        }
        if (dead) {
            if (half == 0) fw_flush(fs, spec + j, lane);
            continue;
        }
        fp_pair_sync(pair);                   // the whole packet is in global memory (both warps wrote their bytes)
        if (half == 0) {
            bool crc_ok = true;
            if (p.crc_mode) {
                const int nbytes = fs->nbytes;
                if (nbytes < 4) crc_ok = false;
                else {
                    const uint8_t *pk = bytes_out + (unsigned long long)(unsigned)j * (unsigned long long)byte_stride;
                    const uint32_t c = (words_ok && nbytes >= 8) ? crc32_warp_words(pk, nbytes - 4, s_tab, s_pow, p.crc_pow8, x_2048, lane)
                                                : crc32_warp(pk, nbytes - 4, s_tab, s_pow, x_2048, lane);
                    const uint32_t got = (uint32_t)pk[nbytes - 4] | ((uint32_t)pk[nbytes - 3] << 8)
                                         | ((uint32_t)pk[nbytes - 2] << 16) | ((uint32_t)pk[nbytes - 1] << 24);
                    crc_ok = (c == got);
                }
            }
            if (crc_ok && lane == 0) fs->rec.flags = fs->rec.flags | OFDMX_F_CRC_OK;
            fw_flush(fs, spec + j, lane);
        }
    }
}

static inline size_t framep_smem_bytes(int n_occ_u, int nh, int hsz)
{
    auto al16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t nh8 = (size_t)((nh + 7) & ~7);
    const size_t shared_bytes = 1024 * 8 + 64 * 8 + 256 * 4 + 32 * 4 + 64 + 6 * nh8 * 2;
    const size_t decn = (size_t)((n_occ_u + 15) & ~15);
    const size_t pair_bytes = 2 * decn + 64 + 64;
    const size_t warp_bytes = (size_t)F1K_SLOT * 8 + (size_t)hsz * 8 + sizeof(FwState);
    return al16(shared_bytes) + (size_t)FP_PAIRS * (pair_bytes + 2 * warp_bytes) + 16;
}

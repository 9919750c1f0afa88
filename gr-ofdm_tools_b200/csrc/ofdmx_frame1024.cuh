// ofdmx_frame1024.cuh -- K1+K3+K4 fast path for fft_len = 1024: one CTA per trigger, one WARP per
// OFDM symbol.
//
//   * each warp pulls its symbol's 1024 samples straight from HBM into registers (32 per lane,
//     coalesced 256-byte rows), derotates them with a phasor recurrence (the NCO of
//     frequency_modulator_fc), and runs the 1024-point forward FFT entirely in registers as
//     32 x 32: a generated straight-line 32-point FFT per lane, twiddles from a lane-contiguous
//     shared table, one conflict-free shared-memory transpose, a second 32-point FFT;
//   * all symbols of a frame are transformed concurrently by different warps (no barriers inside the
//     FFT, only __syncwarp);
//   * channel estimation, the decision-directed equaliser, demapping and serialisation then run
//     carrier-parallel: a thread walks its carriers through every symbol of the round without any
//     block barrier (carriers are independent in ofdm_equalizer_simpledfe);
//   * bit packing, descrambling and the CRC-32 check finish the frame; only payload bytes and the
//     32-byte record go back to HBM.
#pragma once
#include "ofdmx_kernels.cuh"
#include "fft32_gen.cuh"

#define F1K_MAXW 9
#define F1K_ROW 34                      // float2 per transpose row (272 B: 16-byte aligned, conflict-free)
#define F1K_SLOT (32 * F1K_ROW)         // float2 per warp buffer (>= 1024)

// exact NCO phasor at item i (piecewise accumulation over the raw triggers); kept out of line: it is only
// needed for the few samples that follow another trigger inside a frame
static __device__ __noinline__ float2 f1k_exact_phasor(long long i, int j, int jend, const long long *__restrict__ trig,
                                                const float *__restrict__ cfo)
{
    double turns = nco_turns(i, j, jend, trig, cfo, 1024);
    turns -= rint(turns);
    float s2, c2;
    sincospif(2.0f * (float)turns, &s2, &c2);
    return make_float2(c2, s2);
}

// phasor of 32 samples of NCO advance (the recurrence step of f1k_symbol); constant over a frame
__device__ __forceinline__ float2 f1k_step_phasor(double kappa)
{
    double tsd = kappa * 32.0;
    tsd -= rint(tsd);
    float ssn, scs;
    sincospif(2.0f * (float)tsd, &ssn, &scs);
    return make_float2(scs, ssn);
}

// streaming sample load: read-only path, no L1 allocation (the 24 KB of L1 left beside the 232 KB of shared
// memory are kept for the small tables the frame loop reads from global memory)
__device__ __forceinline__ float2 f1k_ld_stream(const float2 *p)
{
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

// One symbol: load + derotate + 1024-point FFT.  Result: Tw[k] = X[k], natural order, k < 1024.
__device__ __forceinline__ void f1k_symbol(const KP &p, const float2 *__restrict__ r, long long n, long long i0,
                                           long long t, double kappa, float2 st, bool slow, int j, int jend,
                                           const long long *__restrict__ trig, const float *__restrict__ cfo,
                                           float2 *__restrict__ Tw, const float2 *__restrict__ tws, int lane)
{
    float2 v[32];
    const long long sbase = i0 - p.D + lane;          // stream index of this lane's first sample
    if (sbase - lane >= 0 && sbase - lane + 1024 <= n) {
#pragma unroll
        for (int a = 0; a < 32; a++) v[a] = f1k_ld_stream(&r[sbase + 32 * a]);
    } else {
#pragma unroll
        for (int a = 0; a < 32; a++) {
            const long long s = sbase + 32 * a;
            v[a] = (s >= 0 && s < n) ? __ldg(&r[s]) : make_float2(0.f, 0.f);
        }
    }
    {
        // phase(i) = 2 pi kappa (i - t + 1); phasor recurrence over a (step = 32 samples)
        double tb = kappa * (double)(i0 + lane - t + 1);
        tb -= rint(tb);
        float sn, cs;
        sincospif(2.0f * (float)tb, &sn, &cs);
        float2 ph = make_float2(cs, sn);
        if (!slow) {
#pragma unroll
            for (int a = 0; a < 32; a++) {
                v[a] = cmul(v[a], ph);
                ph = cmul(ph, st);
            }
        } else {
            // another raw trigger falls inside (or before) this symbol: the sample-and-hold value changes
            // there.  Samples before it keep the recurrence; the others get their phase from the piecewise
            // accumulation (nco_turns).
            const long long tnx = (j + 1 < jend) ? trig[j + 1] : 0x7fffffffffffffffLL;
#pragma unroll
            for (int a = 0; a < 32; a++) {
                const long long i = i0 + lane + 32 * a;
                float2 pa = ph;
                if (i >= tnx) pa = f1k_exact_phasor(i, j, jend, trig, cfo);
                v[a] = cmul(v[a], pa);
                ph = cmul(ph, st);
            }
        }
    }
    // n = 32 a + b (b = lane):  y_b[k1] = sum_a x[32 a + b] W32^(a k1)
    fft32_fwd(v);
    // z_b[k1] = y_b[k1] W1024^(b k1), stored transposed: row k1, column b
#pragma unroll
    for (int q = 0; q < 32; q++) {
        const int k1 = brev5(q);
        Tw[k1 * F1K_ROW + lane] = cmul(v[q], tws[k1 * 32 + lane]);
    }
    __syncwarp();
    // lane = k1 now owns row k1:  X[k1 + 32 k2] = sum_b z_b[k1] W32^(b k2)
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

// fft_len 2048 as two interleaved 1024-point transforms (decimation in time): the even samples of the symbol go
// through the same 32x32 register FFT into Ta, the odd samples into Tb; a bin is read back as
//     X[k] = Ta[k mod 1024] + W_2048^k Tb[k mod 1024]
// by the equaliser (rx_framew_kernel's bin accessor), so the combined spectrum is never materialised.
// st = phasor of 64 samples of NCO advance (the recurrence step between a lane's consecutive inputs).
__device__ __forceinline__ void f2k_symbol(const KP &p, const float2 *__restrict__ r, long long n, long long i0,
                                           long long t, double kappa, float2 st, bool slow, int j, int jend,
                                           const long long *__restrict__ trig, const float *__restrict__ cfo,
                                           float2 *__restrict__ Ta, float2 *__restrict__ Tb,
                                           const float2 *__restrict__ tws, int lane)
{
    const long long tnx = (slow && j + 1 < jend) ? trig[j + 1] : 0x7fffffffffffffffLL;
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
        float2 *Tw = half ? Tb : Ta;
        float2 v[32];
        const long long sbase = i0 - p.D + 2 * lane + half;      // stream index of this lane's first sample
#pragma unroll
        for (int a = 0; a < 32; a++) {
            const long long s = sbase + 64 * a;
            v[a] = (s >= 0 && s < n) ? __ldg(&r[s]) : make_float2(0.f, 0.f);
        }
        {
            double tb = kappa * (double)(i0 + 2 * lane + half - t + 1);
            tb -= rint(tb);
            float sn, cs;
            sincospif(2.0f * (float)tb, &sn, &cs);
            float2 ph = make_float2(cs, sn);
#pragma unroll
            for (int a = 0; a < 32; a++) {
                const long long i = i0 + 2 * lane + half + 64 * a;
                float2 pa = ph;
                if (i >= tnx) {      // sample-and-hold value changed inside the frame: exact piecewise phase
                    double turns = nco_turns(i, j, jend, trig, cfo, 2048);
                    turns -= rint(turns);
                    float s2, c2;
                    sincospif(2.0f * (float)turns, &s2, &c2);
                    pa = make_float2(c2, s2);
                }
                v[a] = cmul(v[a], pa);
                ph = cmul(ph, st);
            }
        }
        fft32_fwd(v);
        __syncwarp();
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
}

// decision_maker with the modulation known at compile time
template <int BPS>
__device__ __forceinline__ int f1k_decide(float re, float im, const uint8_t *__restrict__ lut, float inv_w)
{
    if (BPS == 1) return re > 0.f;
    if (BPS == 2) return 2 * (im > 0.f) + (re > 0.f);
    if (BPS == 3) {
        int r = (fabsf(re) <= fabsf(im)) ? 4 : 0;
        if (re <= 0.f) r |= 1;
        if (im <= 0.f) r |= 2;
        return r;
    }
    constexpr int side = (BPS == 4) ? 4 : 8;
    constexpr float half = 0.5f * (float)side;
    // float -> unsigned conversion saturates at 0 (negative, NaN) by itself: only the upper clamp is an instruction
    const unsigned rs = min(__float2uint_rz(fmaf(re, inv_w, half)), (unsigned)(side - 1));
    const unsigned is = min(__float2uint_rz(fmaf(im, inv_w, half)), (unsigned)(side - 1));
    return lut[rs * side + is];
}

__device__ __forceinline__ float f1k_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

#define F1K_THREADS 320

// SIMPLE: one occupied-carrier set, at most one pilot set, no pilot inside the occupied set (every
// reference surface); WANT_Z: the pre-decision symbols are tapped to z_out (debug / parity tests).
template <int BPS_P, bool SIMPLE, bool WANT_Z>
__global__ void __launch_bounds__(F1K_THREADS, 2)
rx_frame1024_kernel(const KP p, const int W, const float2 *__restrict__ samples, long long n, long long stride,
                    const long long *__restrict__ trig, const int *__restrict__ trig_stream,
                    const float *__restrict__ cfo, const int *__restrict__ stream_start,
                    const int *__restrict__ n_trig_dev, ofdmx_frame *__restrict__ spec,
                    uint8_t *__restrict__ bytes_out, long long byte_stride, float2 *__restrict__ z_out,
                    long long z_stride)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, NT = blockDim.x, NW = blockDim.x >> 5;
    float2 *T = reinterpret_cast<float2 *>(smem_raw);          // W warp buffers
    float2 *tws = T + W * F1K_SLOT;                            // [k1*32 + b]
    float2 *Hs = tws + 1024;                                   // [n_occ_u]
    float2 *ipts = Hs + p.n_occ_u;                             // [64] 1 / payload constellation point
    uint32_t *scratch = reinterpret_cast<uint32_t *>(ipts + 64);
    uint8_t *lut = reinterpret_cast<uint8_t *>(scratch + 16);  // [64] payload sector LUT
    uint8_t *hb = lut + 64;
    uint8_t *syms = hb + ((p.hl + 15) & ~15);
    uint8_t *pk = syms + ((p.max_pkt_syms + 15) & ~15);
    uint8_t *ks = pk + ((p.max_pkt_bytes + 15) & ~15);         // scrambler keystream
    __shared__ uint32_t s_crc_tab[256], s_crc_pow[256];
    __shared__ float2 wacc[F1K_THREADS / 32][4];
    __shared__ float wbest[F1K_THREADS / 32];
    __shared__ int wbestg[F1K_THREADS / 32];
    __shared__ int s_off, s_ok, s_plen, s_pnum, s_psyms, s_fsyms;
    __shared__ float2 pcs[F1K_MAXW];

    for (int i = tid; i < 1024; i += NT) {
        const int k1 = i >> 5, b = i & 31;
        float sn, cs;
        sincospif(-(float)(b * k1) * (1.0f / 512.0f), &sn, &cs);
        tws[i] = make_float2(cs, sn);
    }
    for (int i = tid; i < p.max_pkt_bytes; i += NT) ks[i] = p.keystream[i];
    if (tid < 256) { s_crc_tab[tid] = p.crc_tab[tid]; s_crc_pow[tid] = p.crc_pow[tid]; }
    if (tid < 64) {
        lut[tid] = p.lut_p[tid];
        ipts[tid] = (tid < (1 << BPS_P)) ? p.inv_ppts[tid] : make_float2(0.f, 0.f);
    }
    const int nt = *n_trig_dev;
    const int N = 1024, D = p.D;
    constexpr bool want_z = WANT_Z;
    const float al = p.alpha, oma = 1.0f - p.alpha;
    const bool one_set = SIMPLE || (p.n_occ_sets == 1);
    const bool pil_occ = !SIMPLE && (p.pil_in_occ != 0);
    const int size0 = p.occ_size[0];

    for (int j = blockIdx.x; j < nt; j += gridDim.x) {
        __syncthreads();
        const int st = trig_stream[j];
        const long long t = trig[j];
        const float2 *r = samples + (long long)st * stride;
        const int jend = stream_start[st + 1];
        ofdmx_frame rec;
        rec.trigger = t; rec.cfo = cfo[j]; rec.stream = st; rec.flags = 0; rec.pkt_len = 0; rec.pkt_num = 0;
        rec.frame_syms = 0; rec.carr_offset = 0; rec.slot = (uint32_t)j;
        if (t + 3LL * D > n) {
            if (tid == 0) spec[j] = rec;
            continue;
        }
        const long long tnext = (j + 1 < jend) ? trig[j + 1] : 0x7fffffffffffffffLL;
        const double kappa = (double)rec.cfo * (-2.0 / 1024.0) * (1.0 / TWO_PI_D);
        const float2 kstep = f1k_step_phasor(kappa);
        const long long rem = n - t;
        // ---- round 0: symbols 0 .. W-1 (those that fit in the buffer), one per warp
        if (wid < W && (long long)(wid + 1) * D <= rem) {
            const long long i0 = t + (long long)wid * D + p.cp;
            f1k_symbol(p, r, n, i0, t, kappa, kstep, tnext <= i0 + 1023, j, jend, trig, cfo, T + wid * F1K_SLOT, tws, lane);
        }
        __syncthreads();
        const float2 *Y1 = T, *Y2 = T + F1K_SLOT, *Y3 = T + 2 * F1K_SLOT;
        // ---- ofdm_chanest_vcvc: integer carrier offset, B(g) = |sum_k conj(Y1[k+g]) conj(cv[k]) Y2[k+g]|
        const int ng = (p.gpos - p.gneg) / 2 + 1;
        int off_local = 0;
        if (ng <= 4) {
            // few candidates (max_carr_offset given): every thread takes one cv term for all candidates
            float2 acc[4];
#pragma unroll
            for (int gi = 0; gi < 4; gi++) acc[gi] = make_float2(0.f, 0.f);
            for (int c = tid; c < p.n_cv; c += NT) {
                const int kc = p.cv_k[c];
                const float2 cvc = p.cv_conj[c];
#pragma unroll
                for (int gi = 0; gi < 4; gi++)
                    if (gi < ng) {
                        const int k = kc + p.gneg + 2 * gi;
                        acc[gi] = cadd(acc[gi], cmul(cmul_conj(ysh(Y2, k, N), ysh(Y1, k, N)), cvc));
                    }
            }
#pragma unroll
            for (int gi = 0; gi < 4; gi++)
                for (int o = 16; o > 0; o >>= 1) {
                    acc[gi].x += __shfl_xor_sync(0xffffffffu, acc[gi].x, o);
                    acc[gi].y += __shfl_xor_sync(0xffffffffu, acc[gi].y, o);
                }
            if (lane == 0) {
#pragma unroll
                for (int gi = 0; gi < 4; gi++) wacc[wid][gi] = acc[gi];
            }
            __syncthreads();
            if (wid == 0) {
                // warp 0 folds the per-warp partial sums (lane = source warp) and picks the first maximum
                float b = 0.f;
                int g = 0;
#pragma unroll
                for (int gi = 0; gi < 4; gi++) {
                    float2 sacc = (gi < ng && lane < NW) ? wacc[lane][gi] : make_float2(0.f, 0.f);
                    for (int o = 16; o > 0; o >>= 1) {
                        sacc.x += __shfl_xor_sync(0xffffffffu, sacc.x, o);
                        sacc.y += __shfl_xor_sync(0xffffffffu, sacc.y, o);
                    }
                    const float v = sacc.x * sacc.x + sacc.y * sacc.y;
                    if (gi < ng && v > b) { b = v; g = p.gneg + 2 * gi; }
                }
                if (lane == 0) s_off = g;
            }
            __syncthreads();
            off_local = s_off;
        } else {
            float best = 0.f;
            int bestg = 0;
            for (int gi = wid; gi < ng; gi += NW) {
                const int g = p.gneg + 2 * gi;
                float2 acc = make_float2(0.f, 0.f);
                for (int c = lane; c < p.n_cv; c += 32) {
                    const int k = p.cv_k[c] + g;
                    const float2 t1 = cmul_conj(ysh(Y2, k, N), ysh(Y1, k, N));
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
                for (int w = 0; w < NW; w++) {
                    const float v = wbest[w];
                    if (v > b || (v == b && v > 0.f && wbestg[w] < g)) { b = v; g = wbestg[w]; }
                }
                s_off = g;
            }
            __syncthreads();
            off_local = s_off;
        }
        const int off = off_local;
        // ---- taps + header symbol (ofdm_frame_equalizer_vcvc + simpledfe, header constellation, symbol 0)
        {
            float2 pc = make_float2(1.f, 0.f), rot = make_float2(1.f, 0.f);
            if (off != 0) {
                float sn, cs;
                sincosf((float)(-TWO_PI_D * off * p.cp / N * 1), &sn, &cs);
                pc = make_float2(cs, sn);
                sincosf((float)(TWO_PI_D * off * p.cp / N * 1), &sn, &cs);
                rot = make_float2(cs, sn);               // channel state handed to the payload equaliser
            }
            for (int u = tid; u < p.n_occ_u; u += NT) {
                const int k = p.occ_u[u];
                const int src = k + off;
                float2 Hk = make_float2(0.f, 0.f), y = make_float2(0.f, 0.f);
                if (src >= 0 && src < N) {
                    Hk = cmul(ysh(Y2, src, N), p.inv_sw2[k]);
                    y = cmul(ysh(Y3, src, N), pc);
                }
                int d;
                float2 z = make_float2(0.f, 0.f);
                if (pil_occ && p.pil_flag[k]) {
                    const float2 pv = p.pil_val[k];
                    const float2 q = cdivf(y, pv);
                    Hk = make_float2(al * Hk.x + oma * q.x, al * Hk.y + oma * q.y);
                    d = ofdm_decide(p.bps_h, pv.x, pv.y, p.lut_h, p.qiw_h);
                } else {
                    z = cdivf(y, Hk);
                    d = ofdm_decide(p.bps_h, z.x, z.y, p.lut_h, p.qiw_h);
                    const float2 q = cmul(y, p.inv_hpts[d]);
                    Hk = make_float2(al * Hk.x + oma * q.x, al * Hk.y + oma * q.y);
                }
                const int pos = p.pos_su[u];                 // carrier set 0
                if (pos >= 0) {
                    hb[pos] = (uint8_t)d ^ p.hdr_mask[pos];
                    if (want_z) z_out[(long long)j * z_stride + pos] = z;
                }
                Hs[u] = cmul(Hk, rot);
            }
        }
        __syncthreads();
        if (p.bps_h == 1 && p.hl >= 32) {
            // BPSK header (every reference surface): one bit per item; warp 0 gathers the 32 header bits
            // with a ballot and checks the CRC-8 by linearity (per-bit contributions XOR-reduced)
            if (wid == 0) {
                const unsigned bits = __ballot_sync(0xffffffffu, hb[lane] & 1);
                const unsigned len = bits & 0xFFFu, num = (bits >> 12) & 0xFFFu, crc_rx = bits >> 24;
                unsigned c8 = (lane < 24 && ((bits >> lane) & 1u)) ? (unsigned)p.crc8_bit[lane] : 0u;
                for (int o = 16; o > 0; o >>= 1) c8 ^= __shfl_xor_sync(0xffffffffu, c8, o);
                c8 ^= p.crc8_zero;
                if (lane == 0) {
                    int ps = (int)len * 8 / BPS_P;
                    if (((int)len * 8) % BPS_P) ps++;
                    int fl = 0, acc = 0, s2 = 0;
                    if (one_set) fl = (ps + size0 - 1) / size0;
                    else
                        while (acc < ps) { fl++; acc += p.occ_size[s2]; s2 = (s2 + 1 == p.n_occ_sets) ? 0 : s2 + 1; }
                    s_ok = (c8 == crc_rx); s_plen = (int)len; s_pnum = (int)num; s_psyms = ps; s_fsyms = fl;
                }
            }
        } else if (tid == 0) {
            const int bpb = p.bps_h, msk = (1 << bpb) - 1;
            unsigned len = 0, num = 0;
            int k = 0, ok = 1;
            for (int i = 0; i < 12 && k < p.hl; i += bpb, k++) len |= ((unsigned)(hb[k] & msk)) << i;
            if (k < p.hl) {
                for (int i = 0; i < 12 && k < p.hl; i += bpb, k++) num |= ((unsigned)(hb[k] & msk)) << i;
                if (k < p.hl) {
                    const uint8_t crc = crc8_hdr(len, num);
                    for (int i = 0; i < 8 && k < p.hl; i += bpb, k++)
                        if ((hb[k] & msk) != ((crc >> i) & msk)) ok = 0;
                }
            }
            int ps = (int)len * 8 / BPS_P;
            if (((int)len * 8) % BPS_P) ps++;
            int fl = 0, acc = 0, s = 0;
            if (one_set) fl = (ps + size0 - 1) / size0;
            else
                while (acc < ps) { fl++; acc += p.occ_size[s]; s = (s + 1 == p.n_occ_sets) ? 0 : s + 1; }
            s_ok = ok; s_plen = (int)len; s_pnum = (int)num; s_psyms = ps; s_fsyms = fl;
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
        if ((long long)(3 + fsyms) * D > rem) {
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
        {   // pull the next frame's trigger record towards L1 while this frame is equalised
            const int jn = j + gridDim.x;
            if (tid == 0 && jn < nt) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(trig + jn));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(cfo + jn));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(trig_stream + jn));
            }
        }

        // ---- payload: rounds of up to W symbols; round 0 already holds payload symbols 0 .. W-4
        int cbase = 0;                                // serialised symbols before payload symbol `first`
        int first = 0;                                // first payload symbol of this round
        int slot0 = 3;                                // warp buffer holding payload symbol `first`
        int in_round = min(fsyms, W - 3);
        int set_first = (p.n_occ_sets > 1) ? 1 : 0;   // carrier set / pilot set of payload symbol `first`
        int pset_first = (p.n_pil_sets > 1) ? 1 : 0;
        for (;;) {
            if (off != 0) {   // per-symbol phase fix exp(-j 2 pi off cp / N (i+1)) of ofdm_frame_equalizer_vcvc
                if (tid < in_round) {
                    float sn, cs;
                    sincosf((float)(-TWO_PI_D * off * p.cp / N * (first + tid + 1)), &sn, &cs);
                    pcs[tid] = make_float2(cs, sn);
                }
                __syncthreads();
            }
            // carrier-parallel DFE over the symbols of this round (no barriers: carriers are independent)
            for (int u = tid; u < p.n_occ_u; u += NT) {
                const int k = p.occ_u[u];
                const int src = k + off;
                const bool inr = (src >= 0 && src < N);
                const int ysrc = (src & (N - 1)) ^ (N >> 1);
                float2 Hk = Hs[u];
                int cb = cbase, set = set_first, pset = pset_first;
                const float2 *Y = T + slot0 * F1K_SLOT;
                int pos = one_set ? p.pos_su[u] : 0;
                for (int ii = 0; ii < in_round; ii++, Y += F1K_SLOT) {
                    float2 y = make_float2(0.f, 0.f);
                    if (inr) {
                        y = Y[ysrc];
                        if (off != 0) y = cmul(y, pcs[ii]);
                    }
                    int d;
                    float2 z = make_float2(0.f, 0.f);
                    if (pil_occ && p.pil_flag[pset * N + k]) {
                        const float2 pv = p.pil_val[pset * N + k];
                        const float2 q = cdivf(y, pv);
                        Hk = make_float2(al * Hk.x + oma * q.x, al * Hk.y + oma * q.y);
                        d = f1k_decide<BPS_P>(pv.x, pv.y, lut, p.qiw_p);
                    } else {
                        const float rinv = f1k_rcp(fmaf(Hk.x, Hk.x, Hk.y * Hk.y));
                        const float2 nn = cmul_conj(y, Hk);
                        z = make_float2(nn.x * rinv, nn.y * rinv);
                        d = f1k_decide<BPS_P>(z.x, z.y, lut, p.qiw_p);
                        const float2 q = cmul(y, ipts[d]);
                        Hk = make_float2(fmaf(al, Hk.x, oma * q.x), fmaf(al, Hk.y, oma * q.y));
                    }
                    if (!SIMPLE && !one_set) pos = p.pos_su[set * p.n_occ_u + u];
                    if (pos >= 0) {
                        const int idx = cb + pos;
                        if (idx < psyms) {
                            syms[idx] = (uint8_t)d;
                            if (want_z && p.hl + idx < z_stride) z_out[(long long)j * z_stride + p.hl + idx] = z;
                        }
                    }
                    if (SIMPLE || one_set) cb += size0;
                    else {
                        cb += p.occ_size[set];
                        set = (set + 1 == p.n_occ_sets) ? 0 : set + 1;
                    }
                    if (!SIMPLE && p.n_pil_sets > 1) pset = (pset + 1 == p.n_pil_sets) ? 0 : pset + 1;
                }
                Hs[u] = Hk;
            }
            for (int ii = 0; ii < in_round; ii++) {
                cbase += p.occ_size[set_first];
                set_first = (set_first + 1 >= p.n_occ_sets) ? 0 : set_first + 1;
                if (p.n_pil_sets > 1) pset_first = (pset_first + 1 == p.n_pil_sets) ? 0 : pset_first + 1;
            }
            first += in_round;
            if (first >= fsyms) break;
            // next round: transform the next W payload symbols
            __syncthreads();
            in_round = min(fsyms - first, W);
            slot0 = 0;
            if (wid < in_round) {
                const long long i0 = t + (long long)(3 + first + wid) * D + p.cp;
                f1k_symbol(p, r, n, i0, t, kappa, kstep, tnext <= i0 + 1023, j, jend, trig, cfo, T + wid * F1K_SLOT, tws, lane);
            }
            __syncthreads();
        }
        __syncthreads();
        const int cnt = min(cbase, psyms);
        // ---- repack_bits_bb(bps, 8, key, True) + additive_scrambler_bb
        const int nbytes = min(cnt * BPS_P / 8, p.max_pkt_bytes);
        for (int mb = tid; mb < nbytes; mb += NT) {
            unsigned v = 0;
            if (BPS_P == 4) {
                v = (unsigned)syms[2 * mb] | ((unsigned)syms[2 * mb + 1] << 4);
            } else if (BPS_P == 2) {
                v = (unsigned)syms[4 * mb] | ((unsigned)syms[4 * mb + 1] << 2) | ((unsigned)syms[4 * mb + 2] << 4)
                    | ((unsigned)syms[4 * mb + 3] << 6);
            } else {
                for (int b = 0; b < 8; b++) {
                    const int bi = mb * 8 + b;
                    const int si = bi / BPS_P, sb = bi - si * BPS_P;
                    v |= ((unsigned)(syms[si] >> sb) & 1u) << b;
                }
            }
            const uint8_t o = (uint8_t)v ^ ks[mb];
            pk[mb] = o;
            bytes_out[(long long)j * byte_stride + mb] = o;
        }
        __syncthreads();
        bool crc_ok = true;
        if (p.crc_mode) {
            if (nbytes < 4) crc_ok = false;
            else {
                const uint32_t c = crc32_block(pk, nbytes - 4, s_crc_tab, s_crc_pow, scratch);
                const uint32_t got = (uint32_t)pk[nbytes - 4] | ((uint32_t)pk[nbytes - 3] << 8)
                                     | ((uint32_t)pk[nbytes - 2] << 16) | ((uint32_t)pk[nbytes - 1] << 24);
                crc_ok = (c == got);
            }
        }
        if (crc_ok) rec.flags |= OFDMX_F_CRC_OK;
        if (tid == 0) spec[j] = rec;
    }
}

static inline size_t frame1024_smem_bytes(int W, int n_occ_u, int hl, int max_pkt_syms, int max_pkt_bytes)
{
    auto al16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    return (size_t)W * F1K_SLOT * 8 + 1024 * 8 + (size_t)n_occ_u * 8 + 64 * 8 + 64 + 64 + al16(hl) + al16(max_pkt_syms)
           + 2 * al16(max_pkt_bytes) + 32;
}

// ofdmx_frame1024w.cuh -- K1+K3+K4 with ONE WARP PER FRAME: rx_framew_kernel<fft_len, bps, want_z>,
// instantiated for fft_len 1024 (the benchmark path) and for fft_len 64 / 128 (the carrier plans of
// ofdm_tx_rx_hier and ofdm_radio_hier).
//
// The CTA-per-frame kernels (ofdmx_frame1024.cuh, rx_frame_kernel) spend a large part of their instructions on
// redundancy: every warp runs the per-frame scalar code, the header/CRC/pack phases and waits at ~10 block
// barriers per frame.  Here a frame belongs to a single warp, which streams its OFDM symbols one at a time:
//     load + derotate + FFT (f1k_symbol: 32x32 in registers; fsmall_symbol: registers + lane shuffles)
//     ->  equalise/demap the symbol with the lanes spread over the carriers  ->  pack + descramble its bytes
// keeping only one symbol, the channel state and 80 bytes of frame state in shared memory.  There is no
// block-level synchronisation after the prologue; the warps of an SM sit in different phases (FP32-heavy FFT,
// LDS-heavy equaliser, integer CRC) and fill each other's stalls.  Two measured facts shaped the code
// (profiles/r1_final_summary.md): the steady-state loop has to fit the SM's 32 KB instruction cache, and
// throughput scales with the number of resident warps (16 at fft_len 1024, where the register file is full).
//
// Preconditions checked by the host (otherwise a CTA-per-frame kernel runs): one carrier set, no pilot inside the
// occupied set, BPSK header with >= 32 items; at fft_len 1024 also <= 4 integer-offset candidates
// (max_carr_offset given) and bits per OFDM symbol a multiple of 8.
#pragma once
#include "ofdmx_frame1024.cuh"
#include "ofdmx_symbol_small.cuh"
#include <type_traits>

#define FW_WARPS 16
#define FW_THREADS (FW_WARPS * 32)

// Per-warp frame state kept in shared memory: these values live across the register-hungry FFT of every
// symbol, where the compiler would otherwise spill them to local memory (whose reloads miss the small L1
// that the sample stream keeps flushing).
struct __align__(16) FwState {
    ofdmx_frame rec;        // record under construction (written out once per frame)
    double kappa;           // NCO turns per sample
    long long tnext;        // next raw trigger of the stream
    int nsym;               // symbols of this frame: 3, then 3 + frame_syms once the header is decoded
    int nbytes;             // packet bytes to produce
    float stx, sty;         // phasor of 32 samples of NCO advance
};

// zlib CRC-32 of msg[0..len) by one warp: 64-byte chunks per lane inside 2048-byte super-chunks (leading
// zero padding, init folded into the first 4 bytes), terms shifted by x^(512*(31-lane)) and XOR-reduced.
__device__ __forceinline__ uint32_t crc32_warp(const uint8_t *msg, int len, const uint32_t *tab,
                                               const uint32_t *pow64, uint32_t x_2048, int lane)
{
    if (len < 4) {
        uint32_t c = 0xFFFFFFFFu;
        for (int i = 0; i < len; i++) c = tab[(c ^ msg[i]) & 0xFF] ^ (c >> 8);
        return c ^ 0xFFFFFFFFu;
    }
    uint32_t total = 0;
    int done = 0;
    int first = len & 2047;
    if (first == 0) first = 2048;
    while (done < len) {
        const int clen = done ? 2048 : first;
        const int pad = 2048 - clen;
        uint32_t reg = 0;
        for (int q = 0; q < 64; q++) {
            const int vp = lane * 64 + q - pad;
            if (vp < 0) continue;
            const int gi = done + vp;
            uint8_t b = msg[gi];
            if (gi < 4) b ^= 0xFF;
            reg = tab[(reg ^ b) & 0xFF] ^ (reg >> 8);
        }
        uint32_t term = reg ? gf2_mul(reg, pow64[lane]) : 0u;
        for (int o = 16; o > 0; o >>= 1) term ^= __shfl_xor_sync(0xffffffffu, term, o);
        total = (total ? gf2_mul(total, x_2048) : 0u) ^ term;
        done += clen;
    }
    return total ^ 0xFFFFFFFFu;
}

// Same CRC for a 16-byte aligned message of >= 4 bytes, read by 16-byte loads (one 256-entry table: four
// dependent steps per word; shared memory is better spent on a 16th resident warp than on slicing tables).
// Lane L of a 2048-byte super-chunk owns message bytes [64L, 64L+64); with R bytes left at the start of the
// last super-chunk, R = 64*nq + tail (1 <= tail <= 64): lanes < nq are shifted by x^(512*(nq-1-L)), the
// running total by x^(512*nq), the XOR of those by x^(8*tail), and lane nq (the tail) is added unshifted.
__device__ __forceinline__ uint32_t crc32_warp_words(const uint8_t *msg, int len, const uint32_t *tab,
                                                     const uint32_t *pow64, const uint32_t *__restrict__ pow8,
                                                     uint32_t x_2048, int lane)
{
    uint32_t total = 0;
    for (int base = 0; base < len; base += 2048) {
        const int R = len - base;
        const int mb = min(64, R - 64 * lane);                  // bytes this lane owns (<= 0: none)
        uint32_t reg = 0;
        if (mb > 0) {
            const uint4 *wp = reinterpret_cast<const uint4 *>(msg + base + 64 * lane);
            uint32_t w[16];
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (16 * q < mb) {
                    const uint4 v = wp[q];
                    w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
                }
            if (base == 0 && lane == 0) w[0] ^= 0xFFFFFFFFu;    // CRC init folded into the first 4 bytes
            const int nw = mb >> 2;                             // whole words; mb & 3 bytes follow in word nw
            uint32_t x = 0;
#pragma unroll
            for (int q = 0; q < 16; q++) {
                if (q < nw) {
                    reg ^= w[q];
                    reg = tab[reg & 0xFF] ^ (reg >> 8);
                    reg = tab[reg & 0xFF] ^ (reg >> 8);
                    reg = tab[reg & 0xFF] ^ (reg >> 8);
                    reg = tab[reg & 0xFF] ^ (reg >> 8);
                }
                if (q == nw) x = w[q];
            }
#pragma unroll 1
            for (int b = mb & 3; b > 0; b--, x >>= 8) reg = tab[(reg ^ x) & 0xFF] ^ (reg >> 8);
        }
        if (R > 2048) {
            uint32_t term = reg ? gf2_mul(reg, pow64[lane]) : 0u;
            for (int o = 16; o > 0; o >>= 1) term ^= __shfl_xor_sync(0xffffffffu, term, o);
            total = (total ? gf2_mul(total, x_2048) : 0u) ^ term;
        } else {
            const int nq = (R - 1) >> 6, tail = R - 64 * nq;
            const uint32_t a = (lane < nq) ? reg : (lane == nq ? total : 0u);
            uint32_t term = a ? gf2_mul(a, pow64[lane < nq ? 32 - nq + lane : 31 - nq]) : 0u;
            for (int o = 16; o > 0; o >>= 1) term ^= __shfl_xor_sync(0xffffffffu, term, o);
            const uint32_t last = __shfl_sync(0xffffffffu, reg, nq);
            total = (term ? gf2_mul(term, pow8[tail]) : 0u) ^ last;
        }
    }
    return total ^ 0xFFFFFFFFu;
}

// lane 0 copies the finished record to global memory (two 16-byte moves)
__device__ __forceinline__ void fw_flush(volatile FwState *fs, ofdmx_frame *dst, int lane)
{
    __syncwarp();
    if (lane == 0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(const_cast<const FwState *>(fs));
        const uint4 a = s4[0], b = s4[1];
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        d4[0] = a;
        d4[1] = b;
    }
    __syncwarp();
}

template <int NFFT, int BPS_P, bool WANT_Z>
__global__ void __launch_bounds__(FW_THREADS, NFFT >= 1024 ? 1 : 2)
rx_framew_kernel(const KP p, const float2 *__restrict__ samples, long long n, long long stride,
                     const long long *__restrict__ trig, const int *__restrict__ trig_stream,
                     const float *__restrict__ cfo, const int *__restrict__ stream_start,
                     const int *__restrict__ n_trig_dev, ofdmx_frame *__restrict__ spec,
                     uint8_t *__restrict__ bytes_out, long long byte_stride, float2 *__restrict__ z_out,
                     long long z_stride, uint32_t x_2048, int dec_off, int dec_all_arg)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, NTH = blockDim.x, NWARP = blockDim.x >> 5;
    const int nu = p.n_occ_u;
    const int hsz = (max(nu, p.y1_span) + 1) & ~1;              // H area also parks the Y1 bins chanest needs
    // ---- CTA-shared tables
    constexpr int TWN = NFFT;                                     // twiddle entries (fft_len 2048: 32x32 table + W_2048^k, k < 1024)
    float2 *tws = reinterpret_cast<float2 *>(smem_raw);           // [TWN]
    float2 *ipts = tws + TWN;                                    // [64]
    uint32_t *s_tab = reinterpret_cast<uint32_t *>(ipts + 64);    // [256] CRC-32 table
    uint32_t *s_pow = s_tab + 256;                                // [32]
    // carriers are walked in SERIALISER order (position p in occupied_carriers[0]): the decisions of a symbol land at
    // dec[p] without a position look-up, and the channel state Hs[p] is read lane-contiguously
    uint16_t *s_bin = reinterpret_cast<uint16_t *>(s_pow + 32);   // [nu] shifted bin of the carrier at position p
    uint16_t *s_nat = s_bin + ((nu + 7) & ~7);                    // [nu] its natural-order FFT bin when the carrier offset is 0
    uint8_t *lut = reinterpret_cast<uint8_t *>(s_nat + ((nu + 7) & ~7));   // [64]
    // ---- per-warp buffers
    constexpr int YSLOT = (NFFT == 2048) ? 2 * F1K_SLOT : (NFFT == 1024) ? F1K_SLOT : NFFT;      // float2 per symbol buffer
    // fft_len 1024 keeps its steady-state code minimal: configurations whose bits per OFDM symbol are not a byte
    // multiple (dec_all > 0) go to the CTA-per-frame kernel there
    const int dec_all = (NFFT == 1024) ? 0 : dec_all_arg;
    const size_t shared_bytes = (size_t)TWN * 8 + 64 * 8 + 256 * 4 + 32 * 4 + 2 * (size_t)((nu + 7) & ~7) * 2 + 64;
    const size_t dec_bytes = (dec_off >= 0) ? 0 : (size_t)(((dec_all > 0 ? dec_all : nu) + 15) & ~15);
    const size_t per_warp = (size_t)YSLOT * 8 + (size_t)hsz * 8 + dec_bytes + 64 + sizeof(FwState);
    unsigned char *wbase = smem_raw + ((shared_bytes + 15) & ~(size_t)15) + (size_t)wid * per_warp;
    float2 *Y = reinterpret_cast<float2 *>(wbase);                // F1K_SLOT
    float2 *Hs = Y + YSLOT;
    // FFT output bin in natural order (fft_len 2048: recombined from the even- and odd-sample transforms)
    auto ybin = [&](int k) -> float2 {
        if (NFFT == 2048) {
            const int kk = k & 1023;
            const float2 a = Y[kk], b = cmul(tws[1024 + kk], Y[F1K_SLOT + kk]);
            return (k & 1024) ? make_float2(a.x - b.x, a.y - b.y) : make_float2(a.x + b.x, a.y + b.y);
        }
        return Y[k];
    };                                    // hsz
    // decisions of the current symbol: in the guard band of the symbol buffer when the carrier plan leaves one
    // (bins no equaliser read touches; dead before the next FFT overwrites them), else in their own array
    uint8_t *dec = (dec_off >= 0) ? reinterpret_cast<uint8_t *>(Y + dec_off) : reinterpret_cast<uint8_t *>(Hs + hsz);
    uint8_t *hb = reinterpret_cast<uint8_t *>(Hs + hsz) + dec_bytes;   // 64 header items
    volatile FwState *fs = reinterpret_cast<volatile FwState *>(hb + 64);

    for (int i = tid; i < (NFFT == 2048 ? 1024 : NFFT); i += NTH) {
        const int k1 = i >> 5, b = i & 31;
        float sn, cs;
        sincospif(-(float)(b * k1) * (2.0f / (NFFT == 2048 ? 1024 : NFFT)), &sn, &cs);
        tws[i] = make_float2(cs, sn);
    }
    if (NFFT == 2048)
        for (int i = tid; i < 1024; i += NTH) {      // W_2048^k for the even/odd recombination
            float sn, cs;
            sincospif(-(float)i * (1.0f / 1024.0f), &sn, &cs);
            tws[1024 + i] = make_float2(cs, sn);
        }
    for (int i = tid; i < nu; i += NTH) { const int b = p.occ_bins[i]; s_bin[i] = (uint16_t)b; s_nat[i] = (uint16_t)(b ^ (NFFT / 2)); }
    // (strided loops: the CTA may have fewer than 256 threads when the per-warp buffers are large)
    for (int i = tid; i < 256; i += NTH) s_tab[i] = p.crc_tab[i];
    for (int i = tid; i < 32; i += NTH) s_pow[i] = p.crc_pow64[i];
    for (int i = tid; i < 64; i += NTH) {
        lut[i] = p.lut_p[i];
        // (1 - alpha) / constellation point: the decision-directed update is H <- alpha H + (1 - alpha) y / s
        const float2 ip = (i < (1 << BPS_P)) ? p.inv_ppts[i] : make_float2(0.f, 0.f);
        ipts[i] = make_float2((1.0f - p.alpha) * ip.x, (1.0f - p.alpha) * ip.y);
    }
    __syncthreads();

    const int nt = *n_trig_dev;
    constexpr int N = NFFT, HALF = NFFT / 2;
    const int D = p.D;
    LaneTw ltw = {};
    if (NFFT < 1024) ltw = lane_twiddles(lane);      // lane-FFT twiddles
    const float al = p.alpha, oma = 1.0f - p.alpha, qiw = p.qiw_p;
    const int size0 = p.occ_size[0];
    const int sym_bytes = size0 * BPS_P / 8;
    const int ng = (p.gpos - p.gneg) / 2 + 1;
    const int y1_lo = p.y1_lo;                                    // first shifted bin parked from Y1
    const unsigned hmask32 = __ballot_sync(0xffffffffu, p.hdr_mask[lane] & 1);   // header scrambler, bits 0..31
    const bool words_ok = ((reinterpret_cast<uintptr_t>(bytes_out) | (uintptr_t)byte_stride) & 15) == 0;

    for (int j = blockIdx.x * NWARP + wid; j < nt; j += gridDim.x * NWARP) {
        const int st = trig_stream[j];
        const long long t = trig[j];
        const float2 *r = samples + (long long)st * stride;
        const int jend = stream_start[st + 1];
        const long long rem = n - t;
        if (lane == 0) {
            const float cf = cfo[j];
            fs->rec.trigger = t; fs->rec.cfo = cf; fs->rec.stream = st; fs->rec.flags = 0; fs->rec.pkt_len = 0;
            fs->rec.pkt_num = 0; fs->rec.frame_syms = 0; fs->rec.carr_offset = 0; fs->rec.slot = (uint32_t)j;
            const double kap = (double)cf * (-2.0 / NFFT) * (1.0 / TWO_PI_D);
            const float2 kst = f1k_step_phasor(NFFT == 2048 ? 2.0 * kap : kap);
            fs->kappa = kap; fs->stx = kst.x; fs->sty = kst.y;
            fs->tnext = (j + 1 < jend) ? trig[j + 1] : 0x7fffffffffffffffLL;
            fs->nsym = 3;
            fs->nbytes = 0;
        }
        __syncwarp();
        if (3LL * D > rem) {
            fw_flush(fs, spec + j, lane);
            continue;
        }

        // One loop over the frame's OFDM symbols with a SINGLE call site of the (large, straight-line)
        // FFT code, so that all warps of the SM share one copy in the instruction cache.
        int off = 0, psyms = 0;
        bool dead = false;
        float2 phs = make_float2(1.f, 0.f), sD = make_float2(1.f, 0.f);   // fft_len < 1024: carried NCO phasor and its per-symbol step
        if constexpr (NFFT < 1024) {
            double td = fs->kappa * (double)D;
            td -= rint(td);
            float sn, cs;
            sincospif(2.0f * (float)td, &sn, &cs);
            sD = make_float2(cs, sn);
        }
        for (int sidx = 0; sidx < fs->nsym; sidx++) {
            const long long i0 = t + (long long)sidx * D + p.cp;
            if constexpr (NFFT == 2048) {
                f2k_symbol(p, r, n, i0, t, fs->kappa, make_float2(fs->stx, fs->sty), fs->tnext <= i0 + 2047, j, jend, trig, cfo, Y, Y + F1K_SLOT, tws, lane);
            } else if constexpr (NFFT == 1024) {
                f1k_symbol(p, r, n, i0, t, fs->kappa, make_float2(fs->stx, fs->sty), fs->tnext <= i0 + 1023, j, jend, trig, cfo, Y, tws, lane);
                // pull the next symbol's 8 KB towards L2 while this one is processed
                const long long sn = i0 + D - p.D + lane * 32;
                if (sn >= 0 && sn + 32 <= n) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r + sn));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r + sn + 16));
                }
            } else {
                fsmall_symbol<NFFT>(p, r, n, i0, t, fs->kappa, make_float2(fs->stx, fs->sty), fs->tnext <= i0 + NFFT - 1, j, jend,
                                    trig, cfo, Y, tws, lane, ltw, phs, sD, (sidx & 15) == 0);
            }
            __syncwarp();
            if (sidx == 0) {
                // sync word 1: park the bins chanest needs in the (still unused) H area
                for (int q = lane; q < p.y1_span; q += 32) Hs[q] = ybin((y1_lo + q) ^ HALF);
            } else if (sidx == 1) {
                // sync word 2: integer carrier offset (ofdm_chanest_vcvc), then the channel taps
                float b = 0.f;
                for (int g0 = 0; g0 < ng && (NFFT != 1024 || g0 == 0); g0 += 4) {   // candidates in groups of four (one group at fft_len 1024)
                    float2 acc[4];
#pragma unroll
                    for (int gi = 0; gi < 4; gi++) acc[gi] = make_float2(0.f, 0.f);
                    for (int c0 = lane; c0 < p.n_cv; c0 += 128) {
                        // four table entries per lane per trip, loaded together (global, L2-resident)
                        int kc[4];
                        float2 cvc[4];
#pragma unroll
                        for (int w4 = 0; w4 < 4; w4++) {
                            const int c = c0 + 32 * w4;
                            kc[w4] = (c < p.n_cv) ? p.cv_k[c] : -1;
                            cvc[w4] = (c < p.n_cv) ? p.cv_conj[c] : make_float2(0.f, 0.f);
                        }
#pragma unroll
                        for (int w4 = 0; w4 < 4; w4++) {
                            if (kc[w4] < 0) continue;
#pragma unroll
                            for (int gi = 0; gi < 4; gi++)
                                if (g0 + gi < ng) {
                                    const int k = kc[w4] + p.gneg + 2 * (g0 + gi);
                                    acc[gi] = cadd(acc[gi], cmul(cmul_conj(ybin(k ^ HALF), Hs[k - y1_lo]), cvc[w4]));
                                }
                        }
                    }
#pragma unroll
                    for (int gi = 0; gi < 4; gi++) {
                        for (int o = 16; o > 0; o >>= 1) {
                            acc[gi].x += __shfl_xor_sync(0xffffffffu, acc[gi].x, o);
                            acc[gi].y += __shfl_xor_sync(0xffffffffu, acc[gi].y, o);
                        }
                        const float v = acc[gi].x * acc[gi].x + acc[gi].y * acc[gi].y;
                        if (g0 + gi < ng && v > b) { b = v; off = p.gneg + 2 * (g0 + gi); }
                    }
                }
                __syncwarp();
                // H[k] = Y2[k+off] / sw2[k] (overwrites the parked Y1 bins)
                for (int u = lane; u < nu; u += 32) {
                    const int k = s_bin[u];
                    const int src = k + off;
                    float2 Hk = make_float2(0.f, 0.f);
                    if (src >= 0 && src < N) Hk = cmul(ybin(src ^ HALF), p.inv_sw2[k]);
                    Hs[u] = Hk;
                }
                if (WANT_Z && p.h_taps) {     // debug tap (taps of the occupied carriers; this kernel computes no others)
                    __syncwarp();
                    for (int u = lane; u < nu; u += 32) p.h_taps[(long long)j * p.h_stride + s_bin[u]] = Hs[u];
                }
            } else if (sidx == 2) {
                // header symbol: frame equaliser (offset shift + phase fix) + simpledfe with the BPSK header
                float2 pc = make_float2(1.f, 0.f), rot = make_float2(1.f, 0.f);
                if (off != 0) {
                    // exp(-+ j 2 pi off cp / N): the turn count is reduced exactly in integers
                    float sn, cs;
                    sincospif((float)((off * p.cp) & (N - 1)) * (2.0f / N), &sn, &cs);
                    pc = make_float2(cs, -sn);
                    rot = make_float2(cs, sn);
                }
                // two carriers per lane and trip, straight-line (clamped index, predicated stores): the two dependent
                // chains interleave
                auto hdr_eq = [&](auto off0) {
                    constexpr bool OFF0 = decltype(off0)::value;
                    for (int p0 = lane; p0 < nu; p0 += 64) {
#pragma unroll
                    for (int jj = 0; jj < 2; jj++) {
                        // lanes past the end redo a carrier of their own (position `lane`): nothing is stored for them
                        const int pp = p0 + 32 * jj, pc_ = (pp < nu) ? pp : lane;
                        float2 y;
                        if (OFF0) y = ybin(s_nat[pc_]);
                        else {
                            const int src = (int)s_bin[pc_] + off;
                            y = (src >= 0 && src < N) ? cmul(ybin(src ^ HALF), pc) : make_float2(0.f, 0.f);
                        }
                        float2 Hk = Hs[pc_];
                        const float2 hn = cmul_conj(y, Hk);
                        const int d = hn.x > 0.f;                    // sign of re(y / H): |H|^2 > 0 does not change it
                        const float2 q = make_float2(d ? y.x : -y.x, d ? y.y : -y.y);       // y / (+-1)
                        Hk = make_float2(fmaf(al, Hk.x, oma * q.x), fmaf(al, Hk.y, oma * q.y));
                        if (!OFF0) Hk = cmul(Hk, rot);
                        if (pp < nu) {
                            if (pp < 64) hb[pp] = (uint8_t)d;        // descrambled after the ballot (hmask32)
                            if (WANT_Z) {
                                const float2 H0 = Hs[pc_];
                                const float hinv = f1k_rcp(fmaf(H0.x, H0.x, H0.y * H0.y));
                                z_out[(unsigned long long)(unsigned)j * (unsigned long long)z_stride + pp] = make_float2(hn.x * hinv, hn.y * hinv);
                            }
                            Hs[pp] = Hk;
                        }
                    }
                    }
                };
                if (off == 0) hdr_eq(std::true_type{}); else hdr_eq(std::false_type{});
                __syncwarp();
                const unsigned bits = __ballot_sync(0xffffffffu, hb[lane] & 1) ^ hmask32;
                const int plen = (int)(bits & 0xFFFu);
                const int pnum = (int)((bits >> 12) & 0xFFFu);
                unsigned c8 = (lane < 24 && ((bits >> lane) & 1u)) ? (unsigned)p.crc8_bit[lane] : 0u;
                for (int o = 16; o > 0; o >>= 1) c8 ^= __shfl_xor_sync(0xffffffffu, c8, o);
                const bool ok = ((c8 ^ p.crc8_zero) == (bits >> 24));
                psyms = (plen * 8 + BPS_P - 1) / BPS_P;
                const int fsyms = (psyms + size0 - 1) / size0;
                // oversize: the samples are there, but the length field exceeds the slot capacity the caller configured
                // (max_pkt_bytes): the demux still consumes the declared payload, the packet itself is not decoded
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
                // payload symbol i: equalise + demap with the lanes over the carriers, then pack its bytes
                const int i = sidx - 3;
                float2 pc = make_float2(1.f, 0.f);
                if (off != 0) {
                    float sn, cs;
                    sincospif((float)((off * p.cp * (i + 1)) & (N - 1)) * (2.0f / N), &sn, &cs);
                    pc = make_float2(cs, -sn);
                }
                const int cb = i * size0;
                // bits per OFDM symbol not a multiple of 8 (dec_all > 0): keep the decisions of the whole packet, pack at the end
                uint8_t *decw = dec + (dec_all > 0 ? cb : 0);
                // two carriers per lane and trip in serialiser order, straight-line (clamped index, predicated stores) so
                // that the two dependent chains (loads -> 1/|H|^2 -> decision -> table -> update) interleave
                auto pay_eq = [&](auto off0) {
                    constexpr bool OFF0 = decltype(off0)::value;
                    for (int p0 = lane; p0 < nu; p0 += 64) {
                        // phase 1: all loads of both carriers (lanes past the end redo a carrier of their own, position
                        // `lane`; nothing is stored for them), phase 2: arithmetic, phase 3: stores -- no store sits
                        // between the loads, so the two chains are free to interleave
                        float2 y[2], Hk[2];
#pragma unroll
                        for (int jj = 0; jj < 2; jj++) {
                            const int pp = p0 + 32 * jj, pc_ = (pp < nu) ? pp : lane;
                            if (OFF0) y[jj] = ybin(s_nat[pc_]);
                            else {
                                const int src = (int)s_bin[pc_] + off;
                                y[jj] = (src >= 0 && src < N) ? cmul(ybin(src ^ HALF), pc) : make_float2(0.f, 0.f);
                            }
                            Hk[jj] = Hs[pc_];
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
                                d[jj] = f1k_decide<BPS_P>(nn.x, nn.y, lut, rinv * qiw);      // sector of nn / |H|^2
                            }
                            const float2 q = cmul(y[jj], ipts[d[jj]]);          // (1 - alpha) * y / decided point
                            hq[jj] = make_float2(fmaf(al, Hk[jj].x, q.x), fmaf(al, Hk[jj].y, q.y));
                        }
#pragma unroll
                        for (int jj = 0; jj < 2; jj++) {
                            const int pp = p0 + 32 * jj;
                            if (pp < nu) {
                                Hs[pp] = hq[jj];
                                decw[pp] = (uint8_t)d[jj];
                                if (WANT_Z) {
                                    const int idx = cb + pp;
                                    if (idx < psyms && p.hl + idx < z_stride) z_out[(unsigned long long)(unsigned)j * (unsigned long long)z_stride + p.hl + idx] = z[jj];
                                }
                            }
                        }
                    }
                };
                if (off == 0) pay_eq(std::true_type{}); else pay_eq(std::false_type{});
                __syncwarp();
                // repack_bits_bb(bps, 8) + additive_scrambler_bb for the bytes this OFDM symbol completes
                const int b0 = i * sym_bytes;
                const int nbytes = (dec_all > 0) ? 0 : fs->nbytes;
                if ((BPS_P == 4 || BPS_P == 2) && words_ok && (sym_bytes & 3) == 0) {
                    // four bytes per lane: decisions read as words, nibbles / bit pairs squeezed together
                    const uint32_t *dw = reinterpret_cast<const uint32_t *>(dec);
                    const uint32_t *kw = reinterpret_cast<const uint32_t *>(p.keystream + b0);   // L1-resident
                    uint8_t *orow = bytes_out + (unsigned long long)(unsigned)j * (unsigned long long)byte_stride + b0;
                    for (int m = lane; 4 * m < sym_bytes; m += 32) {
                        const int gb = b0 + 4 * m;
                        if (gb >= nbytes) break;
                        uint32_t v;
                        if (BPS_P == 4) {
                            uint32_t lo = dw[2 * m], hi = dw[2 * m + 1];
                            lo = (lo | (lo >> 4)) & 0x00FF00FFu; lo = (lo | (lo >> 8)) & 0xFFFFu;
                            hi = (hi | (hi >> 4)) & 0x00FF00FFu; hi = (hi | (hi >> 8)) & 0xFFFFu;
                            v = lo | (hi << 16);
                        } else {
                            v = 0;
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                uint32_t x = dw[4 * m + q];
                                x = (x | (x >> 6)) & 0x000F000Fu; x = (x | (x >> 12)) & 0xFFu;
                                v |= x << (8 * q);
                            }
                        }
                        v ^= __ldg(&kw[m]);
                        if (gb + 4 <= nbytes) *reinterpret_cast<uint32_t *>(orow + 4 * m) = v;
                        else
                            for (int b = 0; gb + b < nbytes; b++) orow[4 * m + b] = (uint8_t)(v >> (8 * b));
                    }
                } else
                for (int m = lane; m < sym_bytes; m += 32) {
                    const int gb = b0 + m;
                    if (gb >= nbytes) break;
                    unsigned v = 0;
                    if (BPS_P == 4) v = (unsigned)dec[2 * m] | ((unsigned)dec[2 * m + 1] << 4);
                    else if (BPS_P == 2)
                        v = (unsigned)dec[4 * m] | ((unsigned)dec[4 * m + 1] << 2) | ((unsigned)dec[4 * m + 2] << 4) | ((unsigned)dec[4 * m + 3] << 6);
                    else if (BPS_P == 1) {
                        for (int b = 0; b < 8; b++) v |= (unsigned)dec[8 * m + b] << b;
                    } else {
                        for (int b = 0; b < 8; b++) {
                            const int bi = m * 8 + b;
                            const int si = bi / BPS_P, sb = bi - si * BPS_P;
                            v |= ((unsigned)(dec[si] >> sb) & 1u) << b;
                        }
                    }
                    bytes_out[(unsigned long long)(unsigned)j * (unsigned long long)byte_stride + gb] = (uint8_t)v ^ __ldg(&p.keystream[gb]);
                }
            }
            __syncwarp();
        }
        if (dead) {
            fw_flush(fs, spec + j, lane);
            continue;
        }
        if (dec_all > 0) {
            // repack_bits_bb(bps, 8) + additive_scrambler_bb over the stored decisions of the whole packet
            const int nb = fs->nbytes;
            __syncwarp();
            for (int m = lane; m < nb; m += 32) {
                unsigned v = 0;
                for (int b = 0; b < 8; b++) {
                    const int bi = m * 8 + b;
                    const int si = bi / BPS_P, sb = bi - si * BPS_P;
                    v |= ((unsigned)(dec[si] >> sb) & 1u) << b;
                }
                bytes_out[(unsigned long long)(unsigned)j * (unsigned long long)byte_stride + m] = (uint8_t)v ^ __ldg(&p.keystream[m]);
            }
        }
        bool crc_ok = true;
        if (p.crc_mode) {
            const int nbytes = fs->nbytes;
            if (nbytes < 4) crc_ok = false;
            else {
                // the packet bytes were written by this warp: read them back (L2) for the CRC
                __syncwarp();
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

static inline size_t framew_smem_bytes(int nfft, int n_occ_u, int y1_span, int warps, bool dec_in_guard, int dec_all)
{
    auto al16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t nu8 = (size_t)((n_occ_u + 7) & ~7);
    const size_t yslot = (nfft == 2048) ? 2 * (size_t)F1K_SLOT : (nfft == 1024) ? (size_t)F1K_SLOT : (size_t)nfft;
    const size_t shared_bytes = (size_t)nfft * 8 + 64 * 8 + 256 * 4 + 32 * 4 + 2 * nu8 * 2 + 64;
    const size_t hsz = (size_t)((std::max(n_occ_u, y1_span) + 1) & ~1);
    const size_t per_warp = yslot * 8 + hsz * 8 + (dec_in_guard ? 0 : al16(dec_all > 0 ? dec_all : n_occ_u)) + 64 + sizeof(FwState);
    return al16(shared_bytes) + (size_t)warps * per_warp + 16;
}

// ofdmx_tx1024w.cuh -- K5+K1: the TX chain with ONE WARP PER PACKET, tx_framew_kernel<fft_len, bps> for fft_len 1024
// (register IFFT 32 x 32) and fft_len 64 / 128 (register FFT + lane-shuffle FFT on conjugated data).
//
// Mirror image of the warp-per-frame receiver (ofdmx_frame1024w.cuh).  The generic tx_frame_kernel gives a
// 256-thread CTA to every packet and runs a radix-4 shared-memory IFFT with a block barrier per stage; here a
// packet belongs to one warp, which
//   * pulls the payload into shared memory, appends the CRC-32 (crc32_bb) and applies the additive scrambler;
//   * copies the two sync symbols, whose time-domain samples (cyclic prefix, x tx_scale and clipper included) are a
//     constant of the configuration computed once at context creation;
//   * for the header and every payload symbol fills the IFFT input DIRECTLY IN REGISTERS from a per-bin map
//     (data position / pilot / empty): repack_bits_bb + chunks_to_symbols + ofdm_carrier_allocator_cvc become 32
//     table look-ups per lane, no zero-fill and no scatter pass;
//   * runs the 1024-point inverse FFT as 32 x 32 in registers (same generated butterflies as the receiver, one
//     transpose through shared memory) and stores the samples with the cyclic prefix, x tx_scale and the clipper
//     straight from the registers in coalesced 256-byte rows.
// Preconditions checked by the host (otherwise the generic kernel runs): fft_len 1024, one carrier set, at most
// one pilot set (its symbols may cycle over several pilot-symbol sets), no pilot inside the occupied set, BPSK header.
#pragma once
#include "ofdmx_frame1024w.cuh"

#define TXW_WARPS 16
#define TXW_EMPTY 0xFFFFu
#define TXW_PILOT 0x8000u

// inverse FFT of the lane-distributed spectrum v (bin 32 a + lane) + ofdm_cyclic_prefixer(rolloff 0) +
// multiply_const(tx_scale) + clipper, stored to dst[0 .. fft_len + cp)
// With rolloff (nfl = rolloff_len - 1 > 0; ofdm_cyclic_prefixer, python/ofdm_txrx_modules.py:247-253): the first nfl samples
// of the prefixed symbol are x * up + tail, where tail (per warp, shared memory) holds the first nfl body samples of the
// previous symbol times the down flank; this symbol's own first body samples replace it afterwards.  Same unfused float
// operations as the CTA-per-packet kernel.
template <int NFFT, bool ROLL>
__device__ __forceinline__ void tx_ifft_store(float2 (&v)[NFFT / 32], float2 *__restrict__ Tw, const float2 *__restrict__ tws,
                                              const LaneTw &ltw, float2 *__restrict__ dst, int cp, float sc, float clip, int lane,
                                              int nfl_arg, const float *__restrict__ flank, float2 *__restrict__ tail)
{
    const int nfl = ROLL ? nfl_arg : 0;      // compile-time zero without rolloff: the flank code disappears
    auto finish = [&](float2 x) {
        float2 w = make_float2(x.x * sc, x.y * sc);
        if (clip > 0.f) {
            w.x = w.x < -clip ? -clip : (w.x > clip ? clip : w.x);
            w.y = w.y < -clip ? -clip : (w.y > clip ? clip : w.y);
        }
        return w;
    };
    if constexpr (NFFT == 1024) {
        // x[k1 + 32 k2] = sum_b [ (sum_a X[32 a + b] W32^(-a k1)) W1024^(-b k1) ] W32^(-b k2)
        fft32_inv(v);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 32; q++) {
            const int k1 = brev5(q);
            Tw[k1 * F1K_ROW + lane] = cmul_conj(v[q], tws[k1 * 32 + lane]);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const float4 t4 = *reinterpret_cast<const float4 *>(&Tw[lane * F1K_ROW + 2 * q]);
            v[2 * q] = make_float2(t4.x, t4.y);
            v[2 * q + 1] = make_float2(t4.z, t4.w);
        }
        fft32_inv(v);
        // straight from the registers: lane = k1, v[q] = x[k1 + 32 brev5(q)]
        if (!ROLL) {                  // rectangular prefix: the hot loop carries nothing of the flank logic
#pragma unroll
            for (int q = 0; q < 32; q++) {
                const int t = lane + 32 * brev5(q);
                const float2 w = finish(v[q]);
                dst[cp + t] = w;
                if (t >= NFFT - cp) dst[t - (NFFT - cp)] = w;
            }
        } else {
            // The flank arithmetic stays out of the unrolled store loop (32 copies of it made the hot code larger still:
            // 228 -> 258 Gsamples/s; running ONE copy of the 32-point transform twice on top of that measured slower, 251): the loop only parks the raw samples the
            // flanks need in the (now free) transpose buffer -- Tw[m] = x[fft_len - cp + m] for the up flank,
            // Tw[512 + t] = x[t] for the delay line, m, t < nfl -- and two short rolled loops do the arithmetic.
            __syncwarp();             // every lane has read its row of Tw
#pragma unroll
            for (int q = 0; q < 32; q++) {
                const int t = lane + 32 * brev5(q);
                const float2 w = finish(v[q]);
                dst[cp + t] = w;
                const int m = t - (NFFT - cp);
                if (m >= 0) {
                    if (m < nfl) Tw[m] = v[q];
                    else dst[m] = w;
                }
                if (t < nfl) Tw[512 + t] = v[q];
            }
            __syncwarp();
            for (int m = lane; m < nfl; m += 32) {            // up flank + the delay line of the previous symbol
                const float2 x = Tw[m];
                const float up = flank[m];
                const float2 tl = tail[m];
                dst[m] = finish(make_float2(__fadd_rn(__fmul_rn(x.x, up), tl.x), __fadd_rn(__fmul_rn(x.y, up), tl.y)));
            }
            __syncwarp();             // every old tail value has been consumed
            for (int t = lane; t < nfl; t += 32) {
                const float2 x = Tw[512 + t];
                const float dn = flank[nfl + t];
                tail[t] = make_float2(__fmul_rn(x.x, dn), __fmul_rn(x.y, dn));
            }
            __syncwarp();
        }
    } else {
        // inverse transform = conj(forward(conj X)): fft_len/32-point FFT in registers, twiddles, 32-point FFT across
        // the lanes; time samples staged in shared memory so that the global stores are contiguous
        constexpr int R = NFFT / 32;
#pragma unroll
        for (int a = 0; a < R; a++) v[a].y = -v[a].y;
        fftR_fwd<R>(v);
        const int k2 = brev5(lane);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < R; q++) {
            const int k1 = brevR<R>(q);
            const float2 z = (k1 == 0) ? v[q] : cmul(v[q], tws[k1 * 32 + lane]);
            const float2 x = lane_fft32(z, lane, ltw);
            Tw[k1 + R * k2] = make_float2(x.x, -x.y);
        }
        __syncwarp();
        for (int m = lane; m < NFFT + cp; m += 32) {
            float2 x = Tw[(m - cp + NFFT) & (NFFT - 1)];
            if (m < nfl) {
                const float up = flank[m];
                const float2 tl = tail[m];
                x = make_float2(__fadd_rn(__fmul_rn(x.x, up), tl.x), __fadd_rn(__fmul_rn(x.y, up), tl.y));
            }
            dst[m] = finish(x);
        }
        if (nfl > 0) {
            __syncwarp();
            for (int m = lane; m < nfl; m += 32) { const float dn = flank[nfl + m]; const float2 h = Tw[m]; tail[m] = make_float2(__fmul_rn(h.x, dn), __fmul_rn(h.y, dn)); }
            __syncwarp();
        }
    }
}

template <int NFFT, int BPS_P, bool ROLL>
__global__ void __launch_bounds__(TXW_WARPS * 32, 1)
tx_framew_kernel(const KP p, const uint8_t *__restrict__ payload, const long long *__restrict__ pkt_off,
                     long long n_pkts, int first_num, float2 *__restrict__ out, long long cap,
                     const long long *__restrict__ sample_off, const uint16_t *__restrict__ tx_map,
                     const float2 *__restrict__ sync_td, uint32_t x_2048, int pb_bytes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, NTH = blockDim.x;
    // ---- CTA-shared tables
    float2 *tws = reinterpret_cast<float2 *>(smem_raw);           // [1024] W1024^(b k1), forward sign
    float2 *pts = tws + NFFT;                                     // [64] payload constellation
    uint32_t *s_tab = reinterpret_cast<uint32_t *>(pts + 64);     // [256] CRC-32 table
    uint32_t *s_pow = s_tab + 256;                                // [32]
    uint16_t *map = reinterpret_cast<uint16_t *>(s_pow + 32);     // [1024] natural bin -> data position / pilot / empty
    uint8_t *hmask = reinterpret_cast<uint8_t *>(map + NFFT);     // [NFFT] header scrambling mask (zero padded)
    constexpr int TSLOT = (NFFT == 1024) ? F1K_SLOT : NFFT;       // float2 per warp buffer
    const size_t shared_bytes = NFFT * 8 + 64 * 8 + 256 * 4 + 32 * 4 + NFFT * 2 + NFFT;
    const int nfl = (ROLL && p.roll) ? p.roll - 1 : 0;            // flank samples of the cyclic prefixer (0: rectangular)
    const size_t tail_bytes = ((size_t)nfl * 8 + 15) & ~(size_t)15;
    unsigned char *wbase = smem_raw + shared_bytes + (size_t)wid * ((size_t)TSLOT * 8 + pb_bytes + tail_bytes);
    float2 *Tw = reinterpret_cast<float2 *>(wbase);               // TSLOT
    uint8_t *pb = reinterpret_cast<uint8_t *>(Tw + TSLOT);        // packet bytes (+ CRC), 16-byte aligned
    float2 *tail = reinterpret_cast<float2 *>(pb + pb_bytes);     // [nfl] delay line of the prefixer

    for (int i = tid; i < NFFT; i += NTH) {
        const int k1 = i >> 5, b = i & 31;
        float sn, cs;
        sincospif(-(float)(b * k1) * (2.0f / NFFT), &sn, &cs);
        tws[i] = make_float2(cs, sn);
        map[i] = tx_map[i];
        hmask[i] = (i < p.hl) ? p.hdr_mask[i] : 0;
    }
    for (int i = tid; i < 256; i += NTH) s_tab[i] = p.crc_tab[i];
    for (int i = tid; i < 32; i += NTH) s_pow[i] = p.crc_pow64[i];
    for (int i = tid; i < 64; i += NTH) pts[i] = (i < (1 << BPS_P)) ? p.ppts[i] : make_float2(0.f, 0.f);
    __syncthreads();

    constexpr int N = NFFT;
    const int D = p.D, cp = p.cp;
    const LaneTw ltw = lane_twiddles(lane);          // lane-FFT twiddles (fft_len < 1024 only)
    const int size0 = p.occ_size[0];
    const float2 h0 = p.hpts[0], h1 = p.hpts[1];
    const float sc = p.tx_scale, clip = p.tx_clip;
    const int gw = blockIdx.x * (NTH >> 5) + wid, nw = gridDim.x * (NTH >> 5);

    for (long long pk = gw; pk < n_pkts; pk += nw) {
        const long long o0 = pkt_off[pk];
        const int len = (int)(pkt_off[pk + 1] - o0);
        const int lp = len + (p.crc_mode ? 4 : 0);
        if (lp > p.max_pkt_bytes) continue;                       // rejected on the host as well
        __syncwarp();
        // ---- payload -> shared memory
        const uint8_t *src = payload + o0;
        {
            const int full = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) ? (len & ~15) : 0;   // whole 16-byte units
            for (int m = lane * 16; m < full; m += 512) *reinterpret_cast<uint4 *>(pb + m) = __ldg(reinterpret_cast<const uint4 *>(src + m));
            for (int m = full + lane; m < len; m += 32) pb[m] = __ldg(src + m);
        }
        __syncwarp();
        if (p.crc_mode) {   // crc32_bb(check=False): append little-endian CRC
            const uint32_t c = (len >= 4) ? crc32_warp_words(pb, len, s_tab, s_pow, p.crc_pow8, x_2048, lane)
                                          : crc32_warp(pb, len, s_tab, s_pow, x_2048, lane);
            __syncwarp();
            if (lane < 4) pb[len + lane] = (uint8_t)(c >> (8 * lane));
        }
        __syncwarp();
        // ---- additive_scrambler_bb; the two bytes after the packet read as zero (bit windows of the last chunk)
        for (int m = lane; m < lp; m += 32) pb[m] ^= __ldg(&p.keystream[m]);
        if (lane < 2) pb[lp + lane] = 0;
        __syncwarp();
        // ---- packet_header_default::header_formatter (BPSK: one bit per item)
        const unsigned plen = (unsigned)lp & 0xFFFu, pnum = (unsigned)(first_num + (int)pk) & 0xFFFu;
        const unsigned hbits = plen | (pnum << 12) | ((unsigned)crc8_hdr(plen, pnum) << 24);
        const int ns = (lp * 8 + BPS_P - 1) / BPS_P;                   // repack_bits_bb(8, bps, key, False)
        const int n_ofdm = 3 + (ns + size0 - 1) / size0;
        const long long base = sample_off[pk];
        if (base + (long long)n_ofdm * D + nfl > cap) continue;
        // ---- sync symbols: constants of the configuration (flanks applied; the delay line they leave follows them)
        for (int m = lane; m < 2 * D; m += 32) out[base + m] = __ldg(&sync_td[m]);
        for (int m = lane; m < nfl; m += 32) tail[m] = __ldg(&sync_td[2 * D + m]);
        __syncwarp();
        // ---- header and payload symbols
        for (int o = 2; o < n_ofdm; o++) {
            constexpr int R = NFFT / 32;                       // IFFT inputs per lane: bins 32 a + lane
            float2 v[R];
            const int sbase = (o - 3) * size0;
            const int pil0 = (p.n_pil_sym_sets > 1) ? p.pil_sym_base[(o - 2) % p.n_pil_sym_sets] : 0;   // pilot symbols cycle per OFDM symbol
#pragma unroll
            for (int a = 0; a < R; a++) {
                const unsigned code = map[32 * a + lane];
                float2 val = make_float2(0.f, 0.f);
                if (code < TXW_PILOT) {
                    if (o == 2) {
                        const unsigned bit = ((code < 32 ? (hbits >> code) : 0u) ^ hmask[code]) & 1u;
                        val = bit ? h1 : h0;
                    } else {
                        const int si = sbase + (int)code;
                        if (si < ns) {
                            const int bi = si * BPS_P;
                            unsigned c;
                            if (BPS_P == 1 || BPS_P == 2 || BPS_P == 4) c = ((unsigned)pb[bi >> 3] >> (bi & 7)) & ((1u << BPS_P) - 1u);
                            else {
                                const unsigned w = (unsigned)pb[bi >> 3] | ((unsigned)pb[(bi >> 3) + 1] << 8);
                                c = (w >> (bi & 7)) & ((1u << BPS_P) - 1u);
                            }
                            val = pts[c];
                        }
                    }
                } else if (code != TXW_EMPTY) {
                    val = __ldg(&p.pil_sym[pil0 + (code & 0x7FFFu)]);
                }
                v[a] = val;
            }
            float2 *dst = out + base + (long long)o * D;
            tx_ifft_store<NFFT, ROLL>(v, Tw, tws, ltw, dst, cp, sc, clip, lane, nfl, p.roll_flank, tail);
        }
        // packet mode: the delay line is flushed behind the last symbol
        for (int m = lane; m < nfl; m += 32) {
            const float2 tl = tail[m];
            float2 w = make_float2(tl.x * sc, tl.y * sc);
            if (clip > 0.f) {
                w.x = w.x < -clip ? -clip : (w.x > clip ? clip : w.x);
                w.y = w.y < -clip ? -clip : (w.y > clip ? clip : w.y);
            }
            out[base + (long long)n_ofdm * D + m] = w;
        }
        __syncwarp();
    }
}

static inline size_t tx1024w_pb_bytes(int max_pkt_bytes) { return ((size_t)max_pkt_bytes + 8 + 15) & ~(size_t)15; }
static inline size_t txw_smem_bytes(int nfft, int max_pkt_bytes, int warps, int roll)
{
    const size_t tslot = (nfft == 1024) ? (size_t)F1K_SLOT : (size_t)nfft;
    const size_t tail = roll ? (((size_t)(roll - 1) * 8 + 15) & ~(size_t)15) : 0;
    return (size_t)(nfft * 8 + 64 * 8 + 256 * 4 + 32 * 4 + nfft * 2 + nfft) + (size_t)warps * (tslot * 8 + tx1024w_pb_bytes(max_pkt_bytes) + tail);
}
